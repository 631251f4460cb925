"""GPU parity (through the C-ABI): the batched extract + register sequence path (features stay on the device)."""
import numpy as np
import pytest

import helpers as H
from loam_b200 import _capi, synth
from oracle.pyoracle import FeParams, LidarParams, RegParams

pytestmark = pytest.mark.gpu


def oracle_sequence(oracle, scans, lp, fe, rp):
    feats = []
    for s in scans:
        xyz = s[:, :3].astype(np.float64)
        e, p = oracle.extract(xyz, lp, fe)
        feats.append((xyz[e], xyz[p], len(e), len(p)))
    out = []
    for k in range(len(scans) - 1):
        pose, det = oracle.register(feats[k + 1][0], feats[k + 1][1], feats[k][0], feats[k][1], None, rp, want_detail=True)
        out.append((pose, det))
    return feats, out


@pytest.mark.parametrize("shape,n", [((64, 1024), 6), ((16, 1800), 5), ((128, 2048), 3)])
def test_sequence_matches_oracle(ctx, oracle, shape, n):
    R, P = shape
    scans = np.stack([synth.make_scan(R, P, k=40 + k) for k in range(n)])
    lp, fe, rp = LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()
    poses, term, its, ne, npl = ctx.odometry_host(scans, H.to_capi(lp), H.to_capi(fe), H.to_capi(rp))
    feats, ref = oracle_sequence(oracle, scans, lp, fe, rp)
    assert [f[2] for f in feats] == list(ne) and [f[3] for f in feats] == list(npl)
    for k, (po, do) in enumerate(ref):
        assert term[k] == do.termination and its[k] == do.n_iters
        assert H.angular_distance(po[:4], poses[k][:4]) < H.POSE_TOL_RAD
        assert np.abs(po[4:] - poses[k][4:]).max() < H.POSE_TOL_M
        gt = synth.relative_pose(40 + k, 41 + k)
        assert H.angular_distance(gt[:4], poses[k][:4]) < 5e-3 and np.abs(gt[4:] - poses[k][4:]).max() < 5e-2


def test_chunking_and_residency_do_not_change_results(ctx):
    """Same sequence through (a) one chunk, (b) chunks of 2 pairs with the halo scan kept in its feature slot,
    (c) the device-resident entry point: bit-identical poses."""
    import torch
    R, P, n = 64, 1024, 8
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    lp, fe, rp = (H.to_capi(x) for x in (LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()))
    a = ctx.odometry_host(scans, lp, fe, rp)
    ctx.set_chunk_pairs(2)
    b = ctx.odometry_host(scans, lp, fe, rp)
    ctx.set_chunk_pairs(3)
    d_scans = torch.from_numpy(scans).cuda()
    d_pose = torch.zeros((n - 1, 7), dtype=torch.float64, device="cuda")
    d_term = torch.zeros(n - 1, dtype=torch.int32, device="cuda")
    d_it = torch.zeros(n - 1, dtype=torch.int32, device="cuda")
    d_ne = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_np = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, d_pose.data_ptr(), d_term.data_ptr(), d_it.data_ptr(),
                            d_ne.data_ptr(), d_np.data_ptr())
    torch.cuda.synchronize()
    ctx.set_stream(None)
    ctx.set_chunk_pairs(0)  # back to automatic
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(a[0], d_pose.cpu().numpy())
    assert np.array_equal(a[1], d_term.cpu().numpy()) and np.array_equal(a[3], d_ne.cpu().numpy().astype(np.uint32))


def test_determinism_and_identity_round_trip(ctx):
    """Properties that need no oracle: repeated runs are bit-identical; a scan registered onto itself stays at identity."""
    R, P = 64, 1024
    s = synth.make_scan(R, P, k=3)
    scans = np.stack([s, s, synth.make_scan(R, P, k=4)])
    lp, fe, rp = (H.to_capi(x) for x in (LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()))
    a = ctx.odometry_host(scans, lp, fe, rp)
    b = ctx.odometry_host(scans, lp, fe, rp)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[1][0] == 0 and a[2][0] == 1  # converged in one iteration
    # (not exactly identity: each point's fitted plane/line comes from its noisy neighbours, not from itself)
    assert H.angular_distance(a[0][0][:4], np.array([0, 0, 0, 1.0])) < 1e-3 and np.abs(a[0][0][4:]).max() < 1e-2


def test_single_scan_and_launch_counter(ctx):
    R, P = 16, 256
    lp, fe, rp = (H.to_capi(x) for x in (LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()))
    before = ctx.launch_count
    poses, term, its, ne, npl = ctx.odometry_host(synth.make_scan(R, P, k=0)[None], lp, fe, rp)
    assert poses.shape == (0, 7) and ne[0] > 0 and npl[0] > 0
    assert ctx.launch_count == before + 2  # extract + pack


def test_async_host_calls_pipeline_without_changing_results(ctx):
    """loamgpu_odometry_host_async: three different sequences enqueued back to back (page-locked buffers, staging
    memory shared between the calls, copies of call k+1 overlapping the kernels of call k), one wait at the end;
    then an unrelated entry point that uses the same staging memory.  Everything equals the synchronous calls."""
    import torch
    R, P, n = 64, 1024, 7
    lp, fe, rp = (H.to_capi(x) for x in (LidarParams(R, P, 1.0, 120.0), FeParams.default(), RegParams.default()))
    seqs = [np.stack([synth.make_scan(R, P, k=100 * j + k) for k in range(n)]) for j in range(3)]
    ref = [ctx.odometry_host(s, lp, fe, rp) for s in seqs]
    ctx.set_chunk_pairs(2)  # several chunks per call: both staging buffers are in use when the next call starts
    try:
        h_scans = [torch.from_numpy(s).pin_memory() for s in seqs]
        outs = [(torch.zeros((n - 1, 7), dtype=torch.float64).pin_memory(), torch.zeros(n - 1, dtype=torch.int32).pin_memory(),
                 torch.zeros(n - 1, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory(),
                 torch.zeros(n, dtype=torch.int32).pin_memory()) for _ in seqs]
        for rep in range(2):
            for hs, o in zip(h_scans, outs):
                ctx.odometry_host_async_ptr(hs.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in o))
            ctx.synchronize()
            for r, o in zip(ref, outs):
                assert np.array_equal(r[0], o[0].numpy()) and np.array_equal(r[1], o[1].numpy())
                assert np.array_equal(r[2], o[2].numpy().astype(np.uint32))
                assert np.array_equal(r[3], o[3].numpy().astype(np.uint32)) and np.array_equal(r[4], o[4].numpy().astype(np.uint32))
        # an async call still in flight, then single-scan extraction through the same staging memory
        ctx.odometry_host_async_ptr(h_scans[1].data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in outs[0]))
        e, p = ctx.extract(seqs[2][0], lp, fe)
        ctx.synchronize()
        assert len(e) == ref[2][3][0] and len(p) == ref[2][4][0]
        assert np.array_equal(ref[1][0], outs[0][0].numpy())
    finally:
        ctx.set_chunk_pairs(0)  # back to automatic


@pytest.mark.parametrize("shape", [(16, 512), (8, 250), (5, 333)])
def test_packed_xyz_records_give_identical_results(ctx, shape):
    """12-byte {x,y,z} float records (loamgpu_odometry_*_strided): host, asynchronous host and device-resident calls
    and the single-scan entry point all equal the float4 path bit for bit — also for ring lengths whose 12-byte rows
    are not 16-byte multiples (no bulk copy) and odd ring lengths (staging buffer alignment)."""
    import torch
    R, P = shape
    n = 5
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    scans4 = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    scans3 = np.ascontiguousarray(scans4[:, :, :3])
    ref = ctx.odometry_host(scans4, lp, fe, rp)
    got = ctx.odometry_host(scans3, lp, fe, rp)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    e4, p4 = ctx.extract(scans4[1], lp, fe)
    e3, p3 = ctx.extract(scans3[1], lp, fe)
    assert np.array_equal(e4, e3) and np.array_equal(p4, p3) and len(p4) == ref[4][1]
    # device-resident and asynchronous host variants
    dev = torch.device("cuda", 0)
    d_scans = torch.from_numpy(scans3).to(dev)
    outs = [torch.zeros((n - 1, 7), dtype=torch.float64, device=dev), torch.zeros(n - 1, dtype=torch.int32, device=dev),
            torch.zeros(n - 1, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev),
            torch.zeros(n, dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()
    ctx.odometry_device_ptr(d_scans.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in outs), stride=12)
    ctx.synchronize()
    assert np.array_equal(outs[0].cpu().numpy(), ref[0]) and np.array_equal(outs[3].cpu().numpy(), ref[3].astype(np.int32))
    h_scans = torch.from_numpy(scans3).pin_memory()
    h_outs = [torch.zeros((n - 1, 7), dtype=torch.float64).pin_memory(), torch.zeros(n - 1, dtype=torch.int32).pin_memory(),
              torch.zeros(n - 1, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory(),
              torch.zeros(n, dtype=torch.int32).pin_memory()]
    ctx.odometry_host_async_ptr(h_scans.data_ptr(), n, lp, fe, rp, *(t.data_ptr() for t in h_outs), stride=12)
    ctx.synchronize()
    assert np.array_equal(h_outs[0].numpy(), ref[0]) and np.array_equal(h_outs[1].numpy(), ref[1])


def test_strided_calls_reject_other_record_sizes(ctx):
    lp, fe, rp = _capi.CLidarParams(4, 64, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    buf = np.zeros((2, 256, 5), dtype=np.float32)
    with pytest.raises(_capi.LoamGpuError) as ei:
        ctx.odometry_host_ptr(buf.ctypes.data, 2, lp, fe, rp, None, None, None, None, None, stride=20)
    assert ei.value.code == _capi.ERR_UNSUPPORTED


def test_large_max_iterations_stops_enqueuing_once_every_pair_is_done(ctx):
    """max_iterations as "run until converged" (valid in the reference) must not enqueue thousands of
    empty launches: beyond 16 iterations the host checks the active count at iterations 8, 16, 32, ..."""
    R, P, n = 16, 512, 6
    lp, fe, rp = _capi.CLidarParams(R, P, 1.0, 120.0), _capi.default_fe_params(), _capi.default_reg_params()
    scans = np.stack([synth.make_scan(R, P, k=k) for k in range(n)])
    ref = ctx.odometry_host(scans, lp, fe, rp)
    rp.max_iterations = 100000
    before = ctx.launch_count
    got = ctx.odometry_host(scans, lp, fe, rp)
    launched = ctx.launch_count - before
    assert launched < 80, launched  # 8 iterations x 5 launches + extraction / build / bookkeeping
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
