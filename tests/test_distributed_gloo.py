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


# ---- map slabs (config C4): cuts on voxel faces, halo selection, bbox all-reduce -------------------
@pytest.mark.parametrize("leaf", [0.1, 0.25, 0.07])
def test_voxel_cut_is_the_exact_voxel_face(leaf):
    f = np.float32
    inv = f(1.0) / f(leaf)
    for K in (-700, -37, -1, 0, 1, 3, 251, 4999):
        c = f(D.voxel_cut(K, leaf))
        below = np.nextafter(c, f(-np.inf))
        assert int(np.floor(c * inv)) >= K > int(np.floor(below * inv))


@pytest.mark.parametrize("world", [1, 2, 8])
def test_slab_cuts_tile_the_axis_on_voxel_faces_and_own_every_point_once(world):
    leaf, radius = 0.1, 0.05
    pts = synth.curved_tunnel(50_000, seed=9, arc_length=60.0, arc_radius=200.0, bound=40.0)
    pts[5, 0] = np.nan
    cuts = D.slab_cuts(pts[:, 0], world, leaf)
    assert len(cuts) == world and cuts[0][0] < -1e38 and cuts[-1][1] > 1e38
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    x = pts[:, 0]
    owned = np.zeros(len(pts), np.int32)
    inv = np.float32(1.0) / np.float32(leaf)
    vox = np.floor(x * inv)
    for lo, hi in cuts:
        m = (x >= np.float32(lo)) & (x < np.float32(hi))
        owned += m
        if m.any() and hi < 1e38:
            # a voxel never straddles a cut: every owned point's voxel index is below every index right of the cut
            assert vox[m].max() < vox[np.isfinite(x) & (x >= np.float32(hi))].min()
        sel = D.slab_select(pts, 0, lo, hi, halo=1.01 * radius)
        assert m.sum() <= len(sel) <= len(pts)
        # the halo holds every point within `radius` (along x) of an owned point
        if m.any():
            near = np.isfinite(x) & (x >= x[m].min() - np.float32(radius)) & (x <= x[m].max() + np.float32(radius))
            assert near.sum() <= len(sel)
    assert np.array_equal(owned[np.isfinite(x)], np.ones(np.isfinite(x).sum(), np.int32)) and owned[5] == 0
    if world > 1:
        sizes = [int(((x >= np.float32(a)) & (x < np.float32(b))).sum()) for a, b in cuts]
        assert max(sizes) < 1.2 * len(pts) / world        # balanced


def _bbox_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = synth.curved_tunnel(30_000, seed=9, arc_length=40.0, arc_radius=200.0, bound=30.0)
    lo, hi = D.slab_cuts(pts[:, 0], world, 0.1)[rank]
    mine = pts[(pts[:, 0] >= np.float32(lo)) & (pts[:, 0] < np.float32(hi)), :3]
    mn, mx = D.allreduce_bbox(mine.min(0), mine.max(0))
    out[rank] = (mn.tolist(), mx.tolist())
    if rank == 0:
        out["ref"] = (pts[:, :3].min(0).tolist(), pts[:, :3].max(0).tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_bbox_allreduce_gives_every_rank_the_whole_map_box():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_bbox_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert out[0] == out[1] == out["ref"]


# ---- map slabs: the voxels of all slabs gathered into the whole-map VoxelGrid -----------------
class _FakeSlabCtx:
    """Stands in for a capi.Context that has run gm_voxel on one slab: the oracle's VoxelGrid on the GLOBAL lattice,
    restricted to the slab's owned points (that is what gm_set_owned_range + the all-reduced box produce on the GPU)."""

    def __init__(self, keys, centroids, counts):
        self._v = {"keys": keys, "centroids": centroids, "counts": counts}

    def download_voxels(self, with_nn=False):
        return self._v


def _slab_voxels(full_vox, pts, lo, hi):
    own = (pts[:, 0] >= np.float32(lo)) & (pts[:, 0] < np.float32(hi))
    ids = np.unique(full_vox["assign"][own])
    return full_vox["voxel_keys"][ids], full_vox["centroids_fx"][ids], full_vox["voxel_counts"][ids]


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    pts = synth.curved_tunnel(30000, seed=4, outlier_frac=0.0)
    leaf = 0.2
    full = O.voxel(pts, leaf)
    lo, hi = D.slab_cuts(pts[:, 0], world, leaf)[rank]
    keys, cen, cnt = D.gather_slab_voxels(_FakeSlabCtx(*_slab_voxels(full, pts, lo, hi)))
    out[rank] = bool(np.array_equal(keys, full["voxel_keys"]) and np.array_equal(cnt, full["voxel_counts"])
                     and cen.tobytes() == full["centroids_fx"].tobytes())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gathered_slab_voxels_are_the_whole_map_voxelgrid(world):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def test_merge_rejects_slabs_that_share_a_voxel():
    k = np.array([1, 5, 9], np.int32)
    c = np.zeros((3, 4), np.float32)
    n = np.ones(3, np.int32)
    keys, _, _ = D.merge_slab_voxels([(k, c, n), (np.array([2, 7], np.int32), c[:2], n[:2])])
    assert list(keys) == [1, 2, 5, 7, 9]
    with pytest.raises(ValueError):
        D.merge_slab_voxels([(k, c, n), (np.array([5], np.int32), c[:1], n[:1])])
