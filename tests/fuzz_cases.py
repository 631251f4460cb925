"""Deterministic random feature-extraction problems for the fuzz tests: small organised scans of odd shapes (1 x 1 up
to 6 x 300, sectors larger than the ring, neighbourhoods wider than the ring, zero caps), smooth / jumpy / quantised
(tie-heavy) / no-return-riddled ranges, random thresholds.  neighbor_points == 0 and number_sectors == 0 are excluded:
the reference indexes out of range / divides by zero there, the C-ABI rejects them (tests cover that separately)."""
import numpy as np


def make_case(seed):
    rng = np.random.RandomState(seed)
    R = int(rng.randint(1, 7))
    P = int(rng.choice([1, 2, 3, 5, 7, 11, 16, 33, 64, 100, 173, 300]))
    N = int(rng.choice([1, 2, 3, 4, 5, 8]))
    S = int(rng.choice([1, 2, 3, 6, 7, 12, 40]))
    fe = (N, S, int(rng.randint(0, 7)), int(rng.randint(0, 21)), float(rng.choice([0.5, 5.0, 50.0, 100.0])),
          float(rng.choice([0.01, 0.5, 1.0, 5.0])), float(rng.choice([0.1, 0.5, 1.0])), float(rng.choice([0.02, 0.5, 1.0])))
    lp = (R, P, float(rng.choice([0.1, 1.0, 3.0])), float(rng.choice([8.0, 30.0, 120.0])))
    style = seed % 4
    az = 2 * np.pi * np.arange(P) / max(P, 1)
    pts = np.zeros((R, P, 3))
    for r in range(R):
        el = np.deg2rad(-10 + 5 * r)
        rngs = 4.0 + 2.0 * np.sin(3 * az + r) + rng.normal(0, 0.02, P)
        if style >= 1:  # depth jumps (edges, occlusions)
            for _ in range(int(rng.randint(0, 5))):
                a, b = sorted(rng.randint(0, P + 1, 2))
                rngs[a:b] += rng.uniform(-3, 6)
        rngs = np.abs(rngs)
        pts[r, :, 0] = rngs * np.cos(el) * np.cos(az)
        pts[r, :, 1] = rngs * np.cos(el) * np.sin(az)
        pts[r, :, 2] = rngs * np.sin(el)
    if style == 2:  # quantised coordinates: many exactly equal curvatures (documented tie-break on both sides)
        pts = np.round(pts * 4) / 4
    if style == 3:  # no-return points
        pts[rng.uniform(size=(R, P)) < 0.08] = 0.0
    pts = pts.reshape(-1, 3)
    if seed % 3 == 0:  # float32-representable input, as a sensor driver delivers it
        pts = pts.astype(np.float32).astype(np.float64)
    return pts, lp, fe


SEEDS = list(range(240))
