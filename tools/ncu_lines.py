#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
       python tools/ncu_lines.py src.csv [min_pct]
Prints, for each source line holding >= min_pct of the stall samples: share of samples, share of executed
warp instructions, average active threads per instruction, and the top stall reasons.
"""
import csv
import sys
from collections import defaultdict


def _i(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main(path, min_pct=0.5):
    cur_file, hdr = None, None
    agg = defaultdict(lambda: {"s": 0, "i": 0, "t": 0, "src": "", "st": defaultdict(int)})
    key = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            cs, ci, ct = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index(
                "Thread Instructions Executed")
            stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if r[0] != "":
            key = (cur_file, int(r[0]))
            agg[key]["src"] = r[1].strip()[:110]
            continue
        if key is None or hdr is None or len(r) <= ct:
            continue
        a = agg[key]
        a["s"] += _i(r[cs])
        a["i"] += _i(r[ci])
        a["t"] += _i(r[ct])
        for j, h in stall_cols:
            if j < len(r) and r[j]:
                a["st"][h] += _i(r[j])
    ts = sum(a["s"] for a in agg.values()) or 1
    ti = sum(a["i"] for a in agg.values()) or 1
    tt = sum(a["t"] for a in agg.values())
    print(f"total samples {ts}  warp-inst {ti}  thread-inst {tt}  avg threads/inst {tt / ti:.2f}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["s"]):
        if 100.0 * a["s"] / ts < min_pct:
            continue
        top = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{h[6:]}:{100 * v / max(a['s'], 1):.0f}%" for h, v in top)
        print(f"{f}:{ln:<5d} samp {100 * a['s'] / ts:5.1f}%  inst {100 * a['i'] / ti:5.1f}%  thr/inst "
              f"{(a['t'] / a['i'] if a['i'] else 0):5.1f}  [{tops}]  | {a['src']}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
