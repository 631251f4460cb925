"""CPU model of the compact-record box test (loam_b200/csrc/bvh.cuh: query_grid / rec_lower_bounds, common.cuh:
rec_pack_boxes, the grid of bvh_build_smem_kernel): the integer lower bound, in cells^2, never exceeds the true squared
distance from the query to any point inside the box — for queries inside, outside and far outside the set's bounding
box — and the int16 arithmetic of the 16x2 SIMD instructions cannot overflow.  This is the proof obligation behind
"a subtree is skipped only when its true distance is strictly above the bound" (DESIGN.md §5d); the kernel itself is
checked against brute force on the GPU (tests/test_fuzz_registration.py, tools/soak_gpu.py)."""
import numpy as np

CELL_MAX, CELLS = 32767, 32766.0


def grid_of(points):
    lo, hi = points.min(0), points.max(0)
    emax = float((hi - lo).max())
    amax = max(emax, float(np.abs(np.r_[lo, hi]).max()))
    margin = 1e-6 * amax + 1e-9
    inv = CELLS / (emax + 2.0 * margin)
    return lo - margin, inv, inv * inv * (1.0 + 1e-12)


def quantise_box(blo, bhi, org, inv):
    f32 = np.float32
    lo32 = np.nextafter(blo.astype(f32), f32(-np.inf)).astype(np.float64)  # at least as low as __double2float_rd
    hi32 = np.nextafter(bhi.astype(f32), f32(np.inf)).astype(np.float64)
    ql = np.clip(np.floor((lo32 - org) * inv - 1e-6), 0, CELL_MAX).astype(np.int64)
    qh = np.clip(np.ceil((hi32 - org) * inv + 1e-6), 0, CELL_MAX).astype(np.int64)
    return ql, qh


def query_cells(q, org, inv):
    t = (q - org) * inv
    e = 1e-6 + np.abs(t) * 1e-12
    return (np.clip(np.floor(t - e), 0, CELL_MAX).astype(np.int64), np.clip(np.ceil(t + e), 0, CELL_MAX).astype(np.int64))


def test_integer_bound_is_a_lower_bound_and_fits_int16():
    rng = np.random.RandomState(0)
    for trial in range(300):
        scale = 10.0 ** rng.uniform(-2, 3)
        pts = rng.normal(size=(64, 3)) * scale * rng.uniform(0.01, 1, 3) + rng.normal(size=3) * scale * 5
        org, inv, inv2 = grid_of(pts)
        sub = pts[rng.choice(64, rng.randint(1, 9), replace=False)]
        ql, qh = quantise_box(sub.min(0), sub.max(0), org, inv)
        spread = rng.choice([0.1, 1.0, 10.0, 1000.0])
        for q in np.r_[pts[:4] + rng.normal(size=(4, 3)) * scale * 0.01, pts.mean(0) + rng.normal(size=(8, 3)) * scale * spread]:
            qlo, qhi = query_cells(q, org, inv)
            a, b = ql - qhi, qlo - qh  # the two 16x2 additions: lo + (-qhi), (-hi) + qlo
            assert np.abs(a).max() <= 32767 and np.abs(b).max() <= 32767
            gap = np.maximum(np.maximum(a, b), 0)
            s = int((gap * gap).sum())
            assert s < 2 ** 32
            d2 = ((sub - q) ** 2).sum(1).min()  # true squared distance to the nearest point inside the box
            # pruning compares s with ceil(bound * inv2): s > that must imply d2 > bound, i.e. s <= ceil(d2 * inv2)
            assert s <= np.ceil(d2 * inv2), (trial, s, d2 * inv2)
