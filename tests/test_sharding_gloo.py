"""N>1 host logic on CPU: world_size-2 gloo.  Shards are disjoint, cover every pair once, carry one halo scan, and
per-pair results gathered rank by rank reassemble in sequence order; the max-over-ranks timing reduce works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loam_b200.sharding import shard_sequence


def test_shards_cover_pairs_once_with_one_halo_scan():
    for n_scans in (1, 2, 3, 10, 257, 10000):
        for world in (1, 2, 3, 4, 8):
            shards = [shard_sequence(n_scans, world, r) for r in range(world)]
            pairs = [p for s in shards for p in range(s.pair_lo, s.pair_hi)]
            assert pairs == list(range(max(n_scans - 1, 0)))
            for s in shards:
                if s.n_pairs:
                    assert (s.scan_lo, s.scan_hi) == (s.pair_lo, s.pair_hi + 1)   # own block + one halo scan
                    assert s.scan_hi <= n_scans
                else:
                    assert s.n_scans == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_scans, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = shard_sequence(n_scans, world, rank)
    # stand-in for the per-pair result of the hot path: a pose row that encodes the pair index
    local = torch.tensor([[float(p)] * 7 for p in range(sh.pair_lo, sh.pair_hi)], dtype=torch.float64).reshape(-1, 7)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]]))
    cap = int(max(s.item() for s in sizes))
    padded = torch.zeros((cap, 7), dtype=torch.float64)
    padded[:local.shape[0]] = local
    bufs = [torch.zeros((cap, 7), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(bufs, padded)
    whole = torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])
    t = torch.tensor([10.0 + rank], dtype=torch.float64)  # "device time" of this rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "whole.npy"), whole.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather_and_max_timing(tmp_path):
    world, n_scans = 2, 11
    mp.spawn(_worker, args=(world, _free_port(), n_scans, str(tmp_path)), nprocs=world, join=True)
    whole = np.load(tmp_path / "whole.npy")
    assert whole.shape == (n_scans - 1, 7)
    assert np.array_equal(whole[:, 0], np.arange(n_scans - 1, dtype=np.float64))
    assert np.load(tmp_path / "tmax.npy")[0] == 11.0
