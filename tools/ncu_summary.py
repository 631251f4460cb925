#!/usr/bin/env python
"""Turn ncu output into the tables committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv>
      <launches.csv> = `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...` of bench.py.
      Prints per kernel: launches, total ms, share of the summed kernel time.

  ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py full raw.csv
      Prints one row per captured launch of an `ncu --set full` report: duration, DRAM bytes read / written
      (bench.py's roofline.traffic), achieved DRAM GB/s, SM / fp64-pipe / warp-occupancy percentages, lane use.
"""
import collections
import csv
import re
import sys

OURS = ("extract_ring_kernel", "pack_features_kernel", "bvh_build_smem_kernel", "bvh_build_kernel", "assoc_knn_smem_kernel",
        "assoc_knn_kernel", "widen_kernel", "assoc_fit_kernel",
        "lm_kernel", "compact_active_kernel", "init_pairs_kernel", "finish_pairs_kernel", "knn_kernel",
        "transform_points_kernel", "big_bbox_kernel", "big_morton_kernel", "big_hist_kernel", "big_scan_kernel",
        "big_scatter_kernel", "big_gather_kernel", "big_topology_kernel", "big_boxes_kernel", "big_header_kernel")


def short(name):
    for k in OURS:
        if re.search(r"\b" + k + r"\b", name):
            return k
    return "other: " + name[:60]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6) * v
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share | first grid | block |\n|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % | {v[2]} | {v[3]} |")
    print(f"| **sum** | {sum(v[0] for v in agg.values())} | {tot:.3f} | 100 % | | |")


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(d, name, unit_scale=None):
        if name not in ix:
            return float("nan")
        v = num(d[ix[name]])
        if unit_scale:
            v *= unit_scale.get(units[ix[name]], 1.0)
        return v

    byte_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    time_scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
    print("| # | kernel | grid | ms | DRAM read MB | DRAM write MB | DRAM GB/s | SM % | fp64 pipe % | warps active % | "
          "threads/inst | regs | L1 hit % | L2 hit % |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n, d in enumerate(data):
        ms = get(d, "gpu__time_duration.sum", time_scale)
        rd = get(d, "dram__bytes_read.sum", byte_scale)
        wr = get(d, "dram__bytes_write.sum", byte_scale)
        print(f"| {n} | {short(d[ix['Kernel Name']])} | {d[ix['Grid Size']]} | {ms:.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{(rd + wr) / 1e9 / (ms * 1e-3):.0f} | {get(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{get(d, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{get(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{get(d, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | "
              f"{get(d, 'launch__registers_per_thread'):.0f} | {get(d, 'l1tex__t_sector_hit_rate.pct'):.0f} | "
              f"{get(d, 'lts__t_sector_hit_rate.pct'):.0f} |")


CLASS_OF = {"extract_ring_kernel": ("extract", "scan"), "pack_features_kernel": ("pack", "scan"),
            "bvh_build_smem_kernel": ("nn_build", "scan"), "bvh_build_kernel": ("nn_build", "scan"),
            "assoc_knn_smem_kernel": ("knn", "pair"), "assoc_knn_kernel": ("knn", "pair"),
            "assoc_fit_kernel": ("fit", "pair"), "lm_kernel": ("lm", "pair")}


def traffic(path, rings, cols, scans, commit):
    """profiles/traffic.json from the raw page of the `ncu --set full` capture of ONE step of `scans` scans: DRAM bytes
    (read + written) of every kernel class per unit of work (scan / registered pair), all launches of the class summed."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    byte_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = collections.OrderedDict()
    for d in data:
        k = short(d[ix["Kernel Name"]])
        if k not in CLASS_OF:
            continue
        b = sum(num(d[ix[m]]) * byte_scale.get(units[ix[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tot[CLASS_OF[k][0]] = tot.get(CLASS_OF[k][0], 0.0) + b
    unit = {c: u for c, u in CLASS_OF.values()}
    n_units = {"scan": int(scans), "pair": int(scans) - 1}
    out = {"_note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel class PER UNIT of work (a scan "
                    "for extract / pack / nn_build, a registered pair for knn / fit / lm: all launches of the class that "
                    "touch the unit), from one `ncu --set full --clock-control none` capture of one step of the default "
                    "`bench.py` command. bench.py turns them into roofline.traffic per launch — only for the shape named "
                    "here.", "shape": [int(rings), int(cols)], "scans_per_step": int(scans), "commit": commit,
           "per_unit": {c: tot[c] / n_units[unit[c]] for c in tot}, "unit": {c: unit[c] for c in tot}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:7])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
