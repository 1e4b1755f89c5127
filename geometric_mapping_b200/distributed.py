"""Multi-GPU plumbing for the two places the path shards.

* RANSAC hypotheses are independent: rank r evaluates ids [shard_range(H, r, W)) on its replica of
  the scan, and the best model is one all-reduce(MAX) of the packed 8-byte key
      key = (uint64(count + 1) << 32) | (0xFFFFFFFF - id)
  (max count, ties -> lowest id: RandomSampleConsensus keeps a model only on a strictly larger
  count).  The key fits a signed int64 because count + 1 < 2^31.
* Multi-frame batches are split by frame (frame f -> rank f mod W); no collective on the data path.

torch.distributed is the transport (NCCL over NVLink on GPUs, gloo on CPU for the tests).
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
