"""N>1 host logic on CPU: hypothesis sharding + packed (count,id) MAX all-reduce over gloo,
world_size 2 and 3, against the single-range argmax.  The per-rank counts come from the oracle
(this is a test; the product computes them on the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geometric_mapping_b200 import distributed as D
from geometric_mapping_b200 import synth


def test_key_packing_round_trip_and_ordering():
    assert D.unpack_key(D.pack_key(10, 7)) == (10, 7)
    assert D.unpack_key(D.pack_key(-1, 3)) == (-1, -1)
    assert D.pack_key(5, 9) > D.pack_key(4, 0)            # higher count wins
    assert D.pack_key(5, 2) > D.pack_key(5, 3)            # tie -> lowest id
    assert D.pack_key(-1, 0) < D.pack_key(0, 4000)        # degenerate never beats a valid hypothesis
    assert D.pack_key((1 << 31) - 2, 0) < (1 << 63)       # fits a signed int64 (NCCL/gloo int64 max)
    assert D.best_key_of_counts([3, 9, 9, -1]) == D.pack_key(9, 1)
    assert D.best_key_of_counts([], 0) == D.INVALID_KEY


@pytest.mark.parametrize("H,W", [(4096, 8), (1000, 3), (5, 8), (0, 2)])
def test_shard_ranges_tile_the_hypotheses(H, W):
    r = [D.shard_range(H, k, W) for k in range(W)]
    assert r[0][0] == 0 and r[-1][1] == H
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
    assert sorted(f for k in range(W) for f in D.frames_of_rank(64, k, W)) == list(range(64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    cloud = synth.curved_tunnel(20000, seed=2)            # every rank holds the same scan
    samples = synth.sample_indices(len(cloud), H, 3, seed=3)
    coef, valid = O.plane_hypotheses(cloud, samples)
    lo, hi = D.shard_range(H, rank, world)
    counts = O.count_plane(cloud, coef[lo:hi], valid[lo:hi], 0.05, nthreads=1)
    key = torch.tensor([D.best_key_of_counts(counts, lo)], dtype=torch.int64)
    D.allreduce_best_key(key)
    out[rank] = int(key.item())
    if rank == 0:
        full = O.count_plane(cloud, coef, valid, 0.05, nthreads=1)
        out["ref"] = (int(full[O.argmax(full)]), O.argmax(full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ransac_allreduce_matches_single_rank(world):
    H = 257
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), H, out), nprocs=world, join=True)
        keys = [out[r] for r in range(world)]
        assert len(set(keys)) == 1                        # every rank ends with the same winner
        assert D.unpack_key(keys[0]) == tuple(out["ref"])
