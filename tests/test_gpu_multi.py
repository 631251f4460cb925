"""GPU: the product-level multi-GPU call (loamgpu_multi_*): contiguous pair blocks, one host thread +
context per device entry, halo scan per block, no collective — results identical to one device whatever the split."""
import threading

import numpy as np
import pytest

from loam_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("slots", [1, 2, 3, 5])
@pytest.mark.parametrize("pack", [4, 3])
def test_split_over_device_slots_equals_one_device(ctx, slots, pack):
    """A device may be listed more than once: on a one-GPU box the split, the halo scans, the per-slot threads and
    the result placement are exercised with `slots` contexts on device 0; with >= 2 GPUs the slots alternate."""
    R, P, n = 16, 512, 12
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])[:, :, :pack]
    ref = ctx.odometry_host(scans, lp, fe, rp)
    devs = [i % max(1, n_devices()) for i in range(slots)]
    m = _capi.MultiContext(devs)
    got = m.odometry_host(scans, lp, fe, rp)
    m.close()
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_more_slots_than_pairs_and_tiny_sequences(ctx):
    R, P = 8, 256
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    m = _capi.MultiContext([0, 0, 0, 0])
    for n in (1, 2, 3):
        scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
        ref = ctx.odometry_host(scans, lp, fe, rp)
        got = m.odometry_host(scans, lp, fe, rp)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)
    m.close()


def test_python_surface_and_errors(ctx):
    import loam_b200 as loam
    R, P, n = 16, 512, 6
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    one = loam.odometry(scans, loam.LidarParams(R, P, 1.0, 120.0))
    two = loam.odometry(scans, loam.LidarParams(R, P, 1.0, 120.0), devices=[0, 0])
    for a, b in zip(one, two):
        assert np.array_equal(a, b)
    with pytest.raises(_capi.LoamGpuError):
        _capi.MultiContext([0, 99])
    with pytest.raises(RuntimeError, match="does not match provided lidar parameters"):
        loam.odometry(scans, loam.LidarParams(R, P + 1, 1.0, 120.0), devices=[0, 0])


def test_contexts_of_finished_threads_are_released():
    """per-thread contexts live in threading.local() and close when their thread ends."""
    import gc

    import torch

    import loam_b200 as loam
    scan = synth.make_scan(16, 512, k=0)
    lp = loam.LidarParams(16, 512, 1.0, 120.0)
    loam.extractFeatureIndices(scan, lp)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]

    def work():
        loam.extractFeatureIndices(scan, lp)

    for _ in range(4):
        t = threading.Thread(target=work)
        t.start()
        t.join()
    gc.collect()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20, (free0, free1)  # four leaked contexts would hold far more
