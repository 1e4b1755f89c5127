"""Multi-GPU plumbing for the two places the path shards.

* RANSAC hypotheses are independent: rank r evaluates ids [shard_range(H, r, W)) on its replica of
  the scan, and the best model is one all-reduce(MAX) of the packed 8-byte key
      key = (uint64(count + 1) << 32) | (0xFFFFFFFF - id)
  (max count, ties -> lowest id: RandomSampleConsensus keeps a model only on a strictly larger
  count).  The key fits a signed int64 because count + 1 < 2^31.
* Multi-frame batches are split by frame (frame f -> rank f mod W); no collective on the data path.

torch.distributed is the PLUMBING (rendezvous, barriers, the one-off exchange of the mailbox handles, gloo on CPU for
the tests).  The data-path collectives themselves -- the (count,id) MAX of a sharded RANSAC round, the bounding-box
MIN/MAX and the scatter-matrix SUM of the map slabs -- are kernels of the library over NVLink peer memory
(csrc/gm_comm.cuh, `connect_peers` + `Context.ransac_sharded / allreduce_voxel_bbox / allreduce_frame`); the
torch.distributed forms of the same collectives (`sharded_ransac_pair`, `allreduce_bbox`) remain as the portable path and
as the cross-check.
"""
from __future__ import annotations

from typing import List, Tuple

INVALID_KEY = 0


def pack_key(count: int, hyp_id: int) -> int:
    """count = -1 for a degenerate hypothesis."""
    if count < -1 or count >= (1 << 31) - 1 or hyp_id < 0 or hyp_id > 0xFFFFFFFF:
        raise ValueError("count/id out of range")
    return ((count + 1) << 32) | (0xFFFFFFFF - hyp_id)


def unpack_key(key: int) -> Tuple[int, int]:
    """-> (count, id); (-1, -1) when no valid hypothesis contributed."""
    key &= 0xFFFFFFFFFFFFFFFF
    count = (key >> 32) - 1
    if count < 0:
        return -1, -1
    return count, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def shard_range(H: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of hypothesis ids for `rank`; ranges tile [0, H)."""
    if world < 1 or not (0 <= rank < world) or H < 0:
        raise ValueError("bad shard arguments")
    per = (H + world - 1) // world
    return min(rank * per, H), min((rank + 1) * per, H)


def frames_of_rank(n_frames: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_frames, world))


def best_key_of_counts(counts, h_begin: int = 0) -> int:
    """Host restatement of the device argmax (csrc/gm_ransac.cuh:k_argmax) for tests."""
    best = INVALID_KEY
    for i, c in enumerate(counts):
        k = pack_key(int(c), h_begin + i)
        if k > best:
            best = k
    return best


def allreduce_best_key(key_tensor, group=None):
    """In-place MAX all-reduce of an int64 key tensor (device tensor under NCCL, CPU under gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key_tensor, op=dist.ReduceOp.MAX, group=group)
    return key_tensor


def sharded_ransac(ctx, kind: int, samples, rank: int, world: int, key_tensor, group=None):
    """One hypothesis-sharded RANSAC round on `ctx` (all ranks hold the same scan and sample list):
    count this rank's shard, all-reduce(MAX) the packed key through `key_tensor` (a 1-element int64
    tensor on the ctx's device and stream), refit the global winner on every rank."""
    lo, hi = shard_range(len(samples), rank, world)
    ctx.ransac(kind, samples, lo, hi)
    if world > 1:
        ctx.ransac_export_key(kind, key_tensor.data_ptr())
        allreduce_best_key(key_tensor, group)
        ctx.ransac_import_key(kind, key_tensor.data_ptr())
    ctx.ransac_select(kind)


def sharded_ransac_pair(ctx, plane_samples, cyl_samples, rank: int, world: int, key_tensor2, group=None):
    """One hypothesis-sharded RANSAC step for both primitives: plane and cylinder shards are counted side by side,
    ONE all-reduce(MAX) carries both packed keys (`key_tensor2`: 2-element int64 tensor on the ctx's device and
    stream), then both winners are refitted side by side on every rank."""
    ctx.ransac_pair(plane_samples, cyl_samples, shard_range(len(plane_samples), rank, world), shard_range(len(cyl_samples), rank, world))
    if world > 1:
        ctx.ransac_export_keys(key_tensor2.data_ptr())
        allreduce_best_key(key_tensor2, group)
        ctx.ransac_import_keys(key_tensor2.data_ptr())
    ctx.ransac_select_pair()


# ---------------------------------------------------------------------------------------------
# Map slabs: one large map cut along an axis, one slab per rank.
#
# A slab owns the points whose VOXEL index along the axis lies in [K_lo, K_hi): cuts sit on voxel faces
# of the pcl::VoxelGrid lattice (floor(p * inv_leaf) is independent of the lattice origin), so no
# voxel is shared by two ranks and, with one global bounding box (all-reduce MIN/MAX of 6 floats, the
# only collective on this path), the per-slab voxel keys and centroids are exactly those of the
# whole map.  Each rank is also handed a halo of >= neighborRadius on both sides: halo points take part
# in the radius search and are then dropped (gm_set_owned_range), so an owned point sees its complete
# neighbourhood and gets the normal a single context would compute.
def voxel_cut(K: int, leaf: float):
    """Smallest float32 c with floor(fl32(c * inv_leaf)) >= K, inv_leaf = fl32(1 / fl32(leaf)): the
    coordinate test `x >= c` is then exactly the voxel-index test `floor(x * inv_leaf) >= K`."""
    import numpy as np

    f = np.float32
    inv = f(1.0) / f(leaf)

    def idx(c):
        return int(np.floor(f(c) * inv))

    c = f(K) / inv
    for _ in range(64):
        if idx(c) < K:
            break
        c = np.nextafter(c, f(-np.inf))
    for _ in range(128):
        if idx(c) >= K:
            return float(c)
        c = np.nextafter(c, f(np.inf))
    raise ValueError("voxel_cut did not converge")


def slab_cuts(coord, world: int, leaf: float):
    """Balanced cuts of the finite values of `coord` (the points' coordinate along the split axis) into
    `world` slabs, snapped to voxel faces.  -> list of (lo, hi) float pairs, lo/hi = -/+3e38 at the ends;
    slab g owns lo <= x < hi.  Slabs tile the axis; a slab may be empty when the data are narrow."""
    import numpy as np

    if world < 1:
        raise ValueError("world < 1")
    x = np.asarray(coord, np.float32)
    x = x[np.isfinite(x)]
    big = 3.0e38
    if world == 1 or x.size == 0:
        return [(-big, big)] + [(big, big)] * (world - 1)
    inv = np.float32(1.0) / np.float32(leaf)
    qs = np.quantile(x.astype(np.float64), np.linspace(0.0, 1.0, world + 1)[1:-1])
    ks = [int(np.floor(np.float32(q) * inv)) for q in qs]
    for i in range(1, len(ks)):
        ks[i] = max(ks[i], ks[i - 1])
    edges = [-big] + [voxel_cut(k, leaf) for k in ks] + [big]
    return [(edges[g], edges[g + 1]) for g in range(world)]


def slab_select(points, axis: int, lo: float, hi: float, halo: float):
    """Rows of `points` (n x 4 float32) a rank is handed: its slab plus the halo (order preserved)."""
    import numpy as np

    v = points[:, axis]
    keep = (v >= np.float32(lo) - np.float32(halo)) & (v < np.float32(hi) + np.float32(halo))
    return np.ascontiguousarray(points[keep])


def robust_box(xyz, q: float = 0.02, margin: float = 1.0):
    """Box holding all but a fraction ~2q per axis of the points, grown by `margin` metres: the region to hand to
    gm_set_grid_box when a few far outliers would otherwise stretch the neighbour grid over the whole crop cube."""
    import numpy as np

    step = max(1, len(xyz) // 200_000)            # quantiles of a subsample are plenty for a box
    sub = np.asarray(xyz[::step], np.float64)
    lo = np.quantile(sub, q, axis=0) - margin
    hi = np.quantile(sub, 1.0 - q, axis=0) + margin
    return lo.astype(np.float32), hi.astype(np.float32)


def allreduce_bbox(mn, mx, group=None):
    """Global bounding box over ranks (MIN / MAX all-reduce of 3 + 3 floats; identity when not distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return mn, mx
    import numpy as np

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    a = torch.tensor(np.asarray(mn, np.float32), device=dev)
    b = torch.tensor(np.asarray(mx, np.float32), device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
    return a.cpu().numpy(), b.cpu().numpy()


def process_map_slab(ctx, slab_points, axis: int, lo: float, hi: float, plane_samples=None, cyl_samples=None, group=None,
                     grid_box=None, compress: bool = True):
    """One rank's share of a large map: normals with halo, owned-point compaction, VoxelGrid on the global
    lattice, local frame, RANSAC primitives of THIS slab (multi-primitive approximation: one plane + one
    cylinder per slab), labels, polyline, compression.  `slab_points` = slab_select(...) of the map."""
    ctx.set_owned_range(axis, lo, hi)
    if grid_box is not None:
        ctx.set_grid_box(grid_box[0], grid_box[1])
    ctx.upload_scan(slab_points)
    ctx.crop()
    ctx.normals()
    mn, mx = ctx.voxel_bbox()
    gmn, gmx = allreduce_bbox(mn, mx, group)
    ctx.set_voxel_bbox(gmn, gmx)
    ctx.voxel()
    ctx.local_frame()
    if plane_samples is not None and len(plane_samples):
        ctx.ransac(0, plane_samples)
        ctx.ransac_select(0)
    if cyl_samples is not None and len(cyl_samples):
        ctx.ransac(1, cyl_samples)
        ctx.ransac_select(1)
    ctx.label()
    ctx.axis_polyline()
    if compress:
        ctx.compress()
    return gmn, gmx


# ---------------------------------------------------------------------------------------------
# Peer-memory collectives: connect the mailboxes of all ranks once, then the library does the rest on the GPU.
def connect_peers(group=None):
    """-> capi.PeerComm connected to every rank of `group` (one process per GPU on one node).  The only use of
    torch.distributed is one all_gather of the 64-byte IPC handles."""
    import torch
    import torch.distributed as dist

    from . import capi

    if not (dist.is_available() and dist.is_initialized()):
        return capi.PeerComm(0, 1)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    comm = capi.PeerComm(rank, world)
    if world == 1:
        return comm
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(comm.handle()), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    comm.connect(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))
    dist.barrier(group=group)  # nobody sends before everybody has mapped everybody
    return comm


def merge_slab_voxels(parts):
    """Whole-map VoxelGrid from the per-slab results.  `parts` = list of (keys int32, centroids n x 4, counts int32) in any
    order; slabs cut at voxel faces on one lattice own disjoint keys, so the map is their union in ascending key order
    (= the output order of pcl::VoxelGrid over the whole cloud)."""
    import numpy as np

    keys = np.concatenate([np.asarray(p[0], np.int32) for p in parts]) if parts else np.empty(0, np.int32)
    cen = np.concatenate([np.asarray(p[1], np.float32).reshape(-1, 4) for p in parts]) if parts else np.empty((0, 4), np.float32)
    cnt = np.concatenate([np.asarray(p[2], np.int32) for p in parts]) if parts else np.empty(0, np.int32)
    order = np.argsort(keys.view(np.uint32), kind="stable")
    keys, cen, cnt = keys[order], cen[order], cnt[order]
    if len(keys) > 1 and (keys[1:] == keys[:-1]).any():
        raise ValueError("slabs share a voxel: they were not cut at voxel faces of one lattice")
    return keys, cen, cnt


def gather_slab_voxels(ctx, group=None):
    """All-gather of the slabs' voxels (keys, centroids, counts) -> the whole-map VoxelGrid on every rank.
    Variable sizes: one all_gather of the counts, then padded all_gathers."""
    import numpy as np
    import torch
    import torch.distributed as dist

    vox = ctx.download_voxels(with_nn=False)
    mine = (vox["keys"], vox["centroids"], vox["counts"])
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return merge_slab_voxels([mine])
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([len(mine[0])], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    sizes = [int(t.item()) for t in ns]
    cap = max(max(sizes), 1)
    buf = np.zeros((cap, 6), np.float32)  # key and count travel as bit patterns beside the centroid
    k = len(mine[0])
    buf[:k, 0:4] = mine[1]
    buf[:k, 4] = mine[0].view(np.float32)
    buf[:k, 5] = mine[2].view(np.float32)
    t = torch.from_numpy(buf).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    parts = []
    for r in range(world):
        a = out[r].cpu().numpy()[: sizes[r]]
        parts.append((np.ascontiguousarray(a[:, 4]).view(np.int32), np.ascontiguousarray(a[:, 0:4]), np.ascontiguousarray(a[:, 5]).view(np.int32)))
    return merge_slab_voxels(parts)


def process_map_slab_peer(ctx, slab_points, axis: int, lo: float, hi: float, map_box, plane_samples=None, cyl_samples=None, grid_box=None,
                          compress: bool = True):
    """process_map_slab with the collectives on the device (ctx has a PeerComm attached): the bounding box is
    all-reduced by one kernel between gm_normals and gm_voxel (no host round trip: `map_box` = (min3, max3) only promises
    where the whole map lies, for the sizing of the VoxelGrid tables), and the local frame is that of the WHOLE map
    (sum of all slabs' scatter matrices)."""
    ctx.set_owned_range(axis, lo, hi)
    if grid_box is not None:
        ctx.set_grid_box(grid_box[0], grid_box[1])
    ctx.set_voxel_bbox_hint(map_box[0], map_box[1])
    ctx.upload_scan(slab_points)
    ctx.crop()
    ctx.normals()
    ctx.allreduce_voxel_bbox()
    ctx.voxel()
    ctx.local_frame()
    ctx.allreduce_frame()
    if plane_samples is not None and len(plane_samples):
        ctx.ransac(0, plane_samples)
        ctx.ransac_select(0)
    if cyl_samples is not None and len(cyl_samples):
        ctx.ransac(1, cyl_samples)
        ctx.ransac_select(1)
    ctx.label()
    ctx.axis_polyline()
    if compress:
        ctx.compress()
