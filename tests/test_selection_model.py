"""CPU model of the extraction kernel's selection (loam_b200/csrc/extract.cu, phase D) against the C
restatement of the reference's greedy walks.

The kernel does not sort a sector and walk it: it decides every candidate by the rule "picked iff no candidate of the
same walk that precedes it within +-(N-1) columns is picked", resolves that recursion in rounds, ranks the picks by
(curvature, column) and lets only the first max+1 invalidate their neighbours.  This test runs exactly that
formulation in Python — bit words and windows left out, they are an encoding — on the fuzz problems and compares the
index lists with the oracle's (which is itself checked against the real reference code): the equivalence the kernel
relies on holds on the CPU alone, odd shapes, ties and zero caps included."""
import numpy as np

import fuzz_cases as FZ
from oracle.pyoracle import FeParams, LidarParams


def model_extract(curv, valid, R, P, fe):
    N, S = int(fe.neighbor_points), int(fe.number_sectors)
    reach = N - 1
    edges, planars = [], []
    for r in range(R):
        c = curv[r * P:(r + 1) * P]
        ok = valid[r * P:(r + 1) * P].copy()
        pps = P // S
        for sec in range(S):
            b, e = sec * pps, (P if sec == S - 1 else sec * pps + pps)
            for planar in (False, True):
                cap = int(fe.max_planar_feats_per_sector if planar else fe.max_edge_feats_per_sector)
                thr = fe.planar_feat_threshold if planar else fe.edge_feat_threshold
                cand = [j for j in range(b, e) if ok[j] and (c[j] < thr if planar else c[j] > thr)]
                if planar:
                    before = lambda t, j: c[t] < c[j] or (c[t] == c[j] and t < j)  # noqa: E731
                else:
                    before = lambda t, j: c[t] > c[j] or (c[t] == c[j] and t > j)  # noqa: E731
                cset = set(cand)
                higher = {j: [t for t in range(max(j - reach, b), min(j + reach, e - 1) + 1)
                              if t != j and t in cset and before(t, j)] for j in cand}
                state = {j: "open" for j in cand}
                while any(s == "open" for s in state.values()):
                    snap = dict(state)  # a round decides on what the previous round left (any interleaving works)
                    for j in cand:
                        if snap[j] != "open":
                            continue
                        if any(snap[t] == "picked" for t in higher[j]):
                            state[j] = "dropped"
                        elif not any(snap[t] == "open" for t in higher[j]):
                            state[j] = "picked"
                picks = [j for j in cand if state[j] == "picked"]
                for j in picks:
                    rank = sum(1 for o in picks if o != j and before(o, j))
                    if rank > cap:
                        continue
                    (planars if planar else edges).append((r, sec, rank, r * P + j))
                    ok[max(j - reach, 0):min(j + reach, P - 1) + 1] = False
    key = lambda t: t[:3]  # noqa: E731  (ring, sector, rank): the reference's output order
    return [t[3] for t in sorted(edges, key=key)], [t[3] for t in sorted(planars, key=key)]


def test_round_formulation_equals_the_greedy_walks(oracle):
    checked = 0
    for seed in FZ.SEEDS[:120]:
        pts, lp_t, fe_t = FZ.make_case(seed)
        lp, fe = LidarParams(*lp_t), FeParams(*fe_t)
        R, P = int(lp.scan_lines), int(lp.points_per_line)
        if int(fe.neighbor_points) >= P:  # CHECK 1 leaves no valid column: nothing to select
            continue
        eo, po = oracle.extract(pts, lp, fe)
        em, pm = model_extract(oracle.curvature(pts, lp, fe), oracle.valid_mask(pts, lp, fe), R, P, fe)
        assert em == list(eo) and pm == list(po), seed
        checked += 1
    assert checked > 80
