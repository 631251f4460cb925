"""Fuzz: the C restatement against the real reference feature code on random odd-shaped problems (CPU, needs
oracle/_ref), and the CUDA path against the oracle on the same problems (-m gpu)."""
import numpy as np
import pytest

import fuzz_cases as FZ
import helpers as H
from oracle.pyoracle import FeParams, LidarParams


def test_oracle_matches_real_reference_on_fuzz_cases(oracle, reflib):
    compared = 0
    for seed in FZ.SEEDS:
        pts, lp, fe = FZ.make_case(seed)
        lp, fe = LidarParams(*lp), FeParams(*fe)
        assert np.array_equal(oracle.curvature(pts, lp, fe), reflib.curvature(pts, lp, fe)[0]), seed
        assert np.array_equal(oracle.valid_mask(pts, lp, fe), reflib.valid_mask(pts, lp, fe)), seed
        eo, po, ties = oracle.extract(pts, lp, fe, return_ties=True)
        if ties == 0:  # the reference's std::sort leaves the order of equal curvatures unpinned
            er, pr = reflib.extract(pts, lp, fe)
            assert np.array_equal(eo, er) and np.array_equal(po, pr), seed
            compared += 1
    assert compared > len(FZ.SEEDS) // 3


@pytest.mark.gpu
def test_cuda_matches_oracle_on_fuzz_cases(ctx, oracle):
    for seed in FZ.SEEDS:
        pts, lp, fe = FZ.make_case(seed)
        lp, fe = LidarParams(*lp), FeParams(*fe)
        clp, cfe = H.to_capi(lp), H.to_capi(fe)
        assert np.array_equal(ctx.curvature(pts, clp, cfe), oracle.curvature(pts, lp, fe)), seed
        assert np.array_equal(ctx.valid_mask(pts, clp, cfe), oracle.valid_mask(pts, lp, fe)), seed
        eo, po = oracle.extract(pts, lp, fe)
        e, p = ctx.extract(pts, clp, cfe)
        assert np.array_equal(e, eo) and np.array_equal(p, po), seed
        if seed % 3 == 0:  # the float4 (TMA-staged) layout of the same values
            f4 = np.zeros((len(pts), 4), dtype=np.float32)
            f4[:, :3] = pts
            e4, p4 = ctx.extract(f4, clp, cfe)
            assert np.array_equal(e4, eo) and np.array_equal(p4, po), seed
