#!/usr/bin/env python
"""CPU simulation of the k-NN traversal over the radix tree the NN-build kernels produce, to size the
next step named in §10c before writing it: how many DEPENDENT node loads, box tests, leaf scans and stack operations a
query costs with the binary tree the kernel walks today and with its 4-wide collapse (a node's record holds the boxes
of its grandchildren, so one load decides two levels).  Pure numpy / Python, no GPU, nothing of the product is used:

    python tools/sim_bvh.py [n_queries]

Prints one JSON line per (set, bound) case.  "hint" = the search starts with the exact k-th distance as its bound, which
is what the second outer iteration of a registration has (previous neighbours under a millimetre-level pose change)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loam_b200 import synth  # noqa: E402
from oracle.pyoracle import FeParams, LidarParams, Oracle  # noqa: E402

LEAF = 8
K = 5


def clz32(x):
    return 32 - int(x).bit_length()


def build(pts):
    """Morton order + Karras topology as in bvh_build_kernel: returns sorted points, per-node (first, last, split)."""
    n = len(pts)
    lo, hi = pts.min(0), pts.max(0)
    scale = 1023.999 / max((hi - lo).max(), 1e-30)
    q = np.clip(((pts - lo) * scale).astype(np.int64), 0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    bits_axis = 6
    while bits_axis < 10 and (1 << (2 * (bits_axis - 1))) < n:
        bits_axis += 1
    code >>= 3 * (10 - bits_axis)
    order = np.argsort(code, kind="stable")
    code, spts = code[order], pts[order]

    def delta(i, j):
        if j < 0 or j >= n:
            return -1
        ci, cj = int(code[i]), int(code[j])
        return clz32(ci ^ cj) if ci != cj else 32 + clz32(i ^ j)

    first = np.zeros(n - 1, dtype=np.int64)
    last = np.zeros(n - 1, dtype=np.int64)
    split = np.zeros(n - 1, dtype=np.int64)
    for i in range(n - 1):
        d = 1 if delta(i, i + 1) - delta(i, i - 1) >= 0 else -1
        dmin = delta(i, i - d)
        lmax = 2
        while delta(i, i + lmax * d) > dmin:
            lmax <<= 1
        ln, st = 0, lmax >> 1
        while st >= 1:
            if delta(i, i + (ln + st) * d) > dmin:
                ln += st
            st >>= 1
        j = i + ln * d
        dn = delta(i, j)
        sp, div = 0, 2
        while True:
            st = (ln + div - 1) // div
            if delta(i, i + (sp + st) * d) > dn:
                sp += st
            if st <= 1:
                break
            div <<= 1
        split[i] = i + sp * d + min(d, 0)
        first[i], last[i] = min(i, j), max(i, j)
    return spts, order, first, last, split


class Tree:
    def __init__(self, pts):
        self.pts, self.order, self.first, self.last, self.split = build(pts)
        self.n = len(pts)
        self._box = {}

    def box(self, f, l):
        key = (f, l)
        b = self._box.get(key)
        if b is None:
            seg = self.pts[f:l + 1]
            b = self._box[key] = (seg.min(0), seg.max(0))
        return b

    def lb(self, f, l, q):
        lo, hi = self.box(f, l)
        g = np.maximum(np.maximum(lo - q, q - hi), 0.0)
        return float(g @ g)

    def children(self, f, l):
        """(first, last) ranges of the two children of the internal node covering [f, l]."""
        # the node covering [f, l] is node f or node l (Karras): the one whose range matches
        i = f if (f < self.n - 1 and self.first[f] == f and self.last[f] == l) else l
        s = int(self.split[i])
        return (f, s), (s + 1, l)


def scan_leaf(t, f, l, q, best, c):
    seg = t.pts[f:l + 1]
    d2 = ((seg - q) ** 2).sum(1)
    c["points"] += l - f + 1
    c["leaves"] += 1
    best.extend(d2.tolist())
    best.sort()
    del best[K:]


def knn_binary(t, q, bound0):
    c = dict(steps=0, tests=0, leaves=0, points=0, push=0, pop=0)
    best, stack = [], []
    bound = bound0
    cur = (0, t.n - 1)
    if t.n > LEAF:
        c["steps"] += 1
        c["tests"] += 1
        if t.lb(0, t.n - 1, q) > bound:
            return c
    while True:
        if cur is None:
            if not stack:
                return c
            lbv, f, l = stack.pop()
            c["pop"] += 1
            if lbv <= bound:
                cur = (f, l)
            continue
        f, l = cur
        if l - f < LEAF:
            scan_leaf(t, f, l, q, best, c)
            if len(best) == K:
                bound = min(bound, best[-1])
            cur = None
            continue
        a, b = t.children(f, l)
        c["steps"] += 1  # one dependent load (the two children are adjacent records)
        da = 0.0 if a[0] == a[1] else t.lb(*a, q)
        db = 0.0 if b[0] == b[1] else t.lb(*b, q)
        c["tests"] += (a[0] != a[1]) + (b[0] != b[1])
        near, far, dn, df = (b, a, db, da) if db < da else (a, b, da, db)
        if df <= bound:
            stack.append((df, *far))
            c["push"] += 1
        cur = near if dn <= bound else None


def knn_wide(t, q, bound0, levels=2):
    """2^levels-wide collapse: the record of a node holds the boxes of its descendants `levels` levels down (a
    descendant that is already a leaf stays as it is), so one dependent load tests up to 2^levels subtrees."""
    c = dict(steps=0, tests=0, leaves=0, points=0, push=0, pop=0)
    best, stack = [], []
    bound = bound0
    cur = (0, t.n - 1)
    while True:
        if cur is None:
            if not stack:
                return c
            lbv, f, l = stack.pop()
            c["pop"] += 1
            if lbv <= bound:
                cur = (f, l)
            continue
        f, l = cur
        if l - f < LEAF:
            scan_leaf(t, f, l, q, best, c)
            if len(best) == K:
                bound = min(bound, best[-1])
            cur = None
            continue
        kids = [(f, l)]
        for _ in range(levels):
            nxt = []
            for ch in kids:
                if ch[1] - ch[0] < LEAF:
                    nxt.append(ch)
                else:
                    nxt.extend(t.children(*ch))
            kids = nxt
        c["steps"] += 1
        c["tests"] += len(kids)
        scored = sorted(((0.0 if k[0] == k[1] else t.lb(*k, q)), k) for k in kids)
        keep = [(d, k) for d, k in scored if d <= bound]
        for d, k in reversed(keep[1:]):
            stack.append((d, *k))
            c["push"] += 1
        cur = keep[0][1] if keep else None


def morton_order(pts):
    lo, hi = pts.min(0), pts.max(0)
    q = np.clip(((pts - lo) * (1023.999 / max((hi - lo).max(), 1e-30))).astype(np.int64), 0, 1023)
    code = np.zeros(len(pts), dtype=np.int64)
    for b in range(10):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(code, kind="stable")


def knn_packet(t, qs, bound0):
    """One walk for a whole warp of (Morton-adjacent) queries: a node is entered when ANY lane's conservative bound
    allows it, every lane tests the same record (uniform loads, all lanes busy in the node steps), lanes the subtree
    cannot help are masked in its leaf scans.  Each lane still sees every leaf it needs, so results are unchanged.
    Counts per WARP: node steps, pops (each re-tests the popped subtree's box), leaf scans, and lane-leaf scans."""
    nl = len(qs)
    c = dict(steps=0, pop=0, leaves=0, lane_leaves=0, push=0)
    best = [[] for _ in range(nl)]
    bound = np.full(nl, bound0)
    stack = []

    def lbs(f, l):
        lo, hi = t.box(f, l)
        g = np.maximum(np.maximum(lo - qs, qs - hi), 0.0)
        return (g * g).sum(1)

    def leaf(f, l, want):
        c["leaves"] += 1
        c["lane_leaves"] += int(want.sum())
        seg = t.pts[f:l + 1]
        for i in np.nonzero(want)[0]:
            d2 = ((seg - qs[i]) ** 2).sum(1)
            b = best[i]
            b.extend(d2.tolist())
            b.sort()
            del b[K:]
            if len(b) == K:
                bound[i] = min(bound[i], b[-1])

    cur = (0, t.n - 1, lbs(0, t.n - 1))
    while True:
        if cur is None:
            if not stack:
                return c
            f, l = stack.pop()
            c["pop"] += 1
            d = lbs(f, l)
            if (d <= bound).any():
                cur = (f, l, d)
            continue
        f, l, d = cur
        if l - f < LEAF:
            leaf(f, l, d <= bound)
            cur = None
            continue
        a, b = t.children(f, l)
        c["steps"] += 1
        da = np.zeros(nl) if a[0] == a[1] else lbs(*a)
        db = np.zeros(nl) if b[0] == b[1] else lbs(*b)
        wa, wb = (da <= bound), (db <= bound)
        # the side most lanes would enter first goes first
        a_first = (da <= db).sum() * 2 >= nl
        (n1, d1, w1), (n2, d2, w2) = ((a, da, wa), (b, db, wb)) if a_first else ((b, db, wb), (a, da, wa))
        if w2.any():
            stack.append(n2)
            c["push"] += 1
        cur = (n1[0], n1[1], d1) if w1.any() else None


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    R, P = 64, 1024
    lp, fe = LidarParams(R, P, 1.0, 120.0), FeParams.default()
    orc = Oracle()
    s0 = synth.make_scan(R, P, k=40)[:, :3].astype(np.float64)
    s1 = synth.make_scan(R, P, k=41)[:, :3].astype(np.float64)
    e0, p0 = orc.extract(s0, lp, fe)
    e1, p1 = orc.extract(s1, lp, fe)
    rng = np.random.RandomState(0)
    for name, T, Q, r in (("planar", s0[p0], s1[p1], 2.0), ("edge", s0[e0], s1[e1], 2.0)):
        t = Tree(T)
        qs = Q[rng.choice(len(Q), size=min(nq, len(Q)), replace=False)]
        for mode in ("cold", "hint"):
            tot = {}
            for kind, fn in (("binary", knn_binary), ("wide4", knn_wide),
                             ("wide8", lambda t_, q_, b_: knn_wide(t_, q_, b_, 3))):
                acc = dict(steps=0, tests=0, leaves=0, points=0, push=0, pop=0)
                for q in qs:
                    b0 = r * r
                    if mode == "hint":
                        d2 = np.sort(((T - q) ** 2).sum(1))[K - 1]
                        b0 = min(b0, float(d2))
                    c = fn(t, q, b0)
                    for k_ in acc:
                        acc[k_] += c[k_]
                tot[kind] = {k_: round(v / len(qs), 2) for k_, v in acc.items()}
            print(json.dumps({"set": name, "n": len(T), "queries": len(qs), "bound": mode, **tot}))
        # warp packets over the source set's own Morton order (what the kernel's lanes hold)
        order = morton_order(Q)
        n_warps = max(1, min(nq // 32, len(Q) // 32))
        starts = rng.choice(len(Q) // 32, size=n_warps, replace=False) * 32
        for mode in ("cold", "hint"):
            acc = dict(steps=0, pop=0, leaves=0, lane_leaves=0, push=0)
            solo = dict(steps=0, leaves=0)
            for s0 in starts:
                qw = Q[order[s0:s0 + 32]]
                b0 = r * r
                if mode == "hint":  # (a per-lane hint would be tighter; the packet uses each lane's own below)
                    pass
                cw = knn_packet(t, qw, b0)
                for k_ in acc:
                    acc[k_] += cw[k_]
                for q in qw:
                    cs = knn_binary(t, q, b0)
                    solo["steps"] += cs["steps"]
                    solo["leaves"] += cs["leaves"]
            if mode == "hint":
                continue
            print(json.dumps({"set": name, "packet_per_warp": {k_: round(v / n_warps, 1) for k_, v in acc.items()},
                              "solo_sum_per_warp": {k_: round(v / n_warps, 1) for k_, v in solo.items()},
                              "solo_mean_per_lane": {k_: round(v / n_warps / 32, 2) for k_, v in solo.items()},
                              "warps": int(n_warps)}))


if __name__ == "__main__":
    main()
