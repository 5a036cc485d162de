"""CPU tests of the N>1 host logic with a world_size-2 gloo group: contiguous stream sharding + the final PCM gather
(the only exchange of the path).  The per-rank "engine" here is the CPU oracle — allowed in tests — so that the
gathered result can be compared with a single-process run of the whole batch."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import helpers as H
from fixtures import make_feature_batch
from lpcnet_b200.sharding import shard_range, gather_pcm


def test_shard_range_partitions_everything():
    for n in (1, 2, 7, 4096, 4097, 32768):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, T, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    feats = make_feature_batch(range(lo, hi), T)                  # stream ids are global: shards are position-independent
    pcm = torch.from_numpy(H.oracle_synth(feats, "int8", nthreads=2))
    dist.barrier()
    full = gather_pcm(pcm, n_total, dist, dst=0)
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5])
def test_two_rank_gloo_gather_equals_single_process(tmp_path, n_total):
    T = 5
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_total, T, out), nprocs=2, join=True)
    got = np.load(out)
    want = H.oracle_synth(make_feature_batch(range(n_total), T), "int8")
    np.testing.assert_array_equal(got, want)
