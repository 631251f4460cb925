"""GPU: what ONE evaluation of the on-device LM solve produces (J^T J, J^T r, cost) against torch autograd of the
reference's literal functors (tests/ceres_model.py) — through the streamed evaluation and through the moment sums of
the inlier planes (register.cu: plane_moments_pass / moment_contrib), at non-identity iterates.

VERDICT round 1, next #1a: "... and vs the kernel's B_j p maps through a small debug C-ABI hook"."""
import numpy as np
import pytest

import ceres_model as CM
import helpers as H
import residual_cases as RC

pytestmark = pytest.mark.gpu


def small_iterate(rng, rot=0.004, trans=0.05):
    """A unit-quaternion iterate a scan-to-scan LM candidate looks like: inside the moment path's validity region."""
    ax = rng.normal(size=3)
    return np.r_[H.axis_angle(rng.uniform(0.2, 1.0) * rot, ax), rng.normal(0, trans, 3)]


def model_sums(kinds, P, A, B, x):
    cost, r, J, g = CM.Problem(kinds, P, A, B).evaluate(np.asarray(x, dtype=np.float64), True)
    return J.T @ J, g, cost


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("mode", [0, 1])
def test_one_evaluation_vs_autograd(ctx, seed, mode):
    rng = np.random.RandomState(300 + seed)
    kinds, P, A, B = RC.random_blocks(20 + seed, 700, 3000, noise=0.08, outliers=0.1)
    for x in ([0, 0, 0, 1, 0, 0, 0.0], small_iterate(rng), small_iterate(rng)):
        Hm, gm, cm = model_sums(kinds, P, A, B, x)
        Hk, gk, ck, used, n_stream = ctx.debug_problem_eval(kinds, P, A, B, x, mode=mode)
        assert used == (mode == 1)
        if mode == 1:
            assert 0 < n_stream < 0.5 * (kinds != 0).sum()  # the Huber outliers (and far points) stay streamed
        assert abs(ck - cm) <= 1e-11 * cm
        np.testing.assert_allclose(gk, gm, rtol=1e-9, atol=1e-11 * np.abs(gm).max())
        np.testing.assert_allclose(Hk, Hm, rtol=1e-9, atol=1e-11 * np.abs(Hm).max())


def test_moment_path_falls_back_outside_its_validity_region(ctx):
    """A large candidate step (|M - I|_F or |t| beyond the bounds the core planes were classified with) must stream
    every record: same sums as mode 0, bit for bit."""
    kinds, P, A, B = RC.random_blocks(31, 200, 1500, noise=0.05, outliers=0.05)
    x = np.r_[H.axis_angle(0.2, [1, 3, 1]), [0.5, -0.3, 0.2]]
    a = ctx.debug_problem_eval(kinds, P, A, B, x, mode=0)
    b = ctx.debug_problem_eval(kinds, P, A, B, x, mode=1)
    assert not b[3]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    Hm, gm, cm = model_sums(kinds, P, A, B, x)
    assert abs(a[2] - cm) <= 1e-11 * cm
    np.testing.assert_allclose(a[0], Hm, rtol=1e-9, atol=1e-11 * np.abs(Hm).max())


def test_planes_only_edges_only_and_uncoverable_planes(ctx):
    rng = np.random.RandomState(5)
    x = small_iterate(rng)
    for ne, npl, noise, outl in ((0, 2500, 0.05, 0.0), (900, 0, 0.05, 0.1), (10, 1200, 0.9, 0.5)):
        kinds, P, A, B = RC.random_blocks(40 + ne, ne, npl, noise=noise, outliers=outl)
        Hm, gm, cm = model_sums(kinds, P, A, B, x)
        Hk, gk, ck, used, n_stream = ctx.debug_problem_eval(kinds, P, A, B, x, mode=1)
        assert abs(ck - cm) <= 1e-11 * cm
        np.testing.assert_allclose(gk, gm, rtol=1e-9, atol=1e-11 * np.abs(gm).max())
        np.testing.assert_allclose(Hk, Hm, rtol=1e-9, atol=1e-11 * np.abs(Hm).max())
    # more uncovered planes than the compact list holds: the solve streams everything
    kinds, P, A, B = RC.random_blocks(77, 0, 6000, noise=2.0, outliers=0.9)
    out = ctx.debug_problem_eval(kinds, P, A, B, x, mode=1)
    ref = ctx.debug_problem_eval(kinds, P, A, B, x, mode=0)
    assert not out[3] and np.array_equal(out[0], ref[0]) and out[2] == ref[2]
