#!/usr/bin/env python
"""Mean cycles per phase of the -DBUILD_TIMING lines of bvh_build_kernel, grouped by set size (stdin)."""
import collections
import re
import sys

agg = collections.defaultdict(list)
for line in sys.stdin:
    if not line.startswith("BT n"):
        continue
    n = int(re.search(r"BT n (\d+)", line).group(1))
    vals = {k: int(v) for k, v in re.findall(r"(\S+) (\d+)", line.split(":", 1)[1])}
    agg[n // 500 * 500].append(vals)
for n in sorted(agg):
    rows = agg[n]
    keys = rows[0].keys()
    print(f"n~{n} ({len(rows)} launches): " + "  ".join(f"{k} {sum(r[k] for r in rows) / len(rows):.0f}" for k in keys))
