// gm_device.cuh — device-side building blocks shared by the stage kernels (sm_100a).
//   * ordered-int float min/max atomics
//   * warp / block reductions (deterministic order)
//   * single-pass stable stream compaction with decoupled look-back
//   (the LSD radix sort lives in gm_sort.cuh)
// Everything here is hand-written; no CUB / Thrust.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gm {

constexpr unsigned FULL = 0xFFFFFFFFu;

// Spin bound for every inter-block wait.  A predecessor tile is always resident or finished (tile
// indices are handed out by an atomic ticket, see TileCtl: whoever holds ticket t started after the
// holders of 0..t-1), so the bound is never reached in a correct run; if it is, the kernel raises
// device_error instead of hanging the GPU.
constexpr int SPIN_BOUND = 1 << 22;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- packed f32x2 arithmetic (sm_100: SASS FFMA2 / FADD2) --------------------------------------
// Each half is an IEEE round-to-nearest operation, bit-identical to the scalar fmaf / add.
typedef unsigned long long u64;
__device__ __forceinline__ u64 d_pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void d_unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 d_fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 d_add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 d_sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- float <-> order-preserving int (for atomicMin/atomicMax on floats) -------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return (i >= 0) ? i : (i ^ 0x7FFFFFFF);
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float((i >= 0) ? i : (i ^ 0x7FFFFFFF));
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Sum NV doubles per thread across a block of BLOCK threads and store the NV block totals to
// out[0..NV) (global or shared).  Warp shuffle tree, then thread k adds the W warp totals of value k
// in warp order: fixed order -> bitwise reproducible, and no single-thread serial tail.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum_store(const double (&v)[NV], double* smem /* NV * BLOCK/32 */, double* out) {
  constexpr int W = BLOCK / 32;
  static_assert(NV <= BLOCK, "one thread per value");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (l == 0) smem[k * W + w] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < W; ++i) s += smem[threadIdx.x * W + i];
    out[threadIdx.x] = s;
    __threadfence();  // the block totals are consumed by other blocks (last-block / grid-barrier reductions)
  }
}

// ---- single-launch grid reductions -----------------------------------------------------------
// Returns true (block-uniformly) in the last block of the grid to get here, after every block's
// earlier global writes are visible.  The counter is reset for the next launch.  Combined with a
// fixed-order sum over per-block partials this gives a single-launch, bitwise reproducible grid
// reduction (the result does not depend on which block happens to be last).
__device__ __forceinline__ bool d_last_block(unsigned* counter, unsigned nblocks) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    s_last = (t == nblocks - 1);
    if (s_last) { *counter = 0u; __threadfence(); }
  }
  __syncthreads();
  return s_last;
}

// fin[k] = sum over blocks b of partials[b*NV + k], k < NV, for the whole calling block.
// Warp w owns columns w, w+W, ...; lane l adds blocks l, l+32, ... in order, then a fixed shuffle
// tree: the loads of a column are independent (one L2 latency, not nblocks of them) and the
// summation order depends only on (nblocks, blockDim), never on scheduling.
template <int NV>
__device__ __forceinline__ void d_reduce_partials(const double* partials, int nblocks, double* fin) {
  const int W = blockDim.x >> 5, w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int k = w; k < NV; k += W) {
    double s = 0.0;
    for (int b = l; b < nblocks; b += 32) s += __ldcg(partials + (size_t)b * NV + k);
    s = warp_sum(s);
    if (l == 0) fin[k] = s;
  }
  __syncthreads();
}

// ---- decoupled look-back tile state ---------------------------------------------------------
// One 64-bit word per tile: epoch (30 bits) | status (2 bits) | value (32 bits), read and written
// with single 8-byte accesses so all three are always observed together.
//  * The EPOCH is a per-launch number chosen by the host: a word whose epoch differs from the
//    launch's is "empty", so the state arrays never have to be cleared (a cudaMemsetAsync node costs
//    ~5 us, and a scan would need 13 of them).
//  * Words are PUBLISHED with atomicExch, not a plain store: measured on B200
//    (tools/microbench/compact_variants.cu) a relaxed store takes ~1 us to become visible to the
//    spinning readers of other SMs, an atomic goes straight to L2; this alone halves the kernel.
//  * Readers use ld.relaxed.gpu (served by L2).  ld.cg must NOT be used to spin: it never observed
//    the update in the same test.
enum : unsigned { TS_EMPTY = 0u, TS_AGG = 1u, TS_PREFIX = 2u };
constexpr unsigned TS_EPOCH_MASK = 0x3FFFFFFFu;

__device__ __forceinline__ unsigned long long ts_pack(unsigned epoch, unsigned status, unsigned value) {
  return ((unsigned long long)((epoch << 2) | status) << 32) | value;
}
__device__ __forceinline__ void ts_store(unsigned long long* p, unsigned epoch, unsigned status, unsigned value) {
  atomicExch(p, ts_pack(epoch, status, value));
}
// -> status (TS_EMPTY if the word belongs to another epoch) and value
__device__ __forceinline__ unsigned ts_load(const unsigned long long* p, unsigned epoch, unsigned& value) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  value = (unsigned)w;
  unsigned hi = (unsigned)(w >> 32);
  return ((hi >> 2) == (epoch & TS_EPOCH_MASK)) ? (hi & 3u) : TS_EMPTY;
}

// Launch control of the look-back kernels, one per tile-state array, resident in device memory so that a launch needs
// no per-launch host argument (the whole scan is replayed as a CUDA graph):
//  * TILE INDICES come from an atomic ticket, not from blockIdx.x: CUDA does not promise that blocks are dispatched in
//    index order, and a tile that waits for a predecessor which has not been scheduled yet could spin forever.  With
//    tickets every predecessor of tile t is, by construction, already running or done.
//  * The EPOCH of the launch is read from here by every block when it takes its ticket; the last block to FINISH
//    (all tickets taken, all reads done) resets ticket/done and advances the epoch for the next launch on the stream.
struct TileCtl { unsigned ticket, done, epoch, pad_; };

// Every block of a look-back kernel: first statement / before every return.  gridDim.x blocks take part.
__device__ __forceinline__ int tile_begin(TileCtl* c, unsigned& epoch) {
  __shared__ int s_tile;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) {
    s_tile = (int)atomicAdd(&c->ticket, 1u);
    s_epoch = *(volatile unsigned*)&c->epoch;
  }
  __syncthreads();
  epoch = s_epoch;
  return s_tile;
}
__device__ __forceinline__ void tile_end(TileCtl* c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&c->done, 1u) == gridDim.x - 1u) {
      c->done = 0u; c->ticket = 0u;
      c->epoch = (*(volatile unsigned*)&c->epoch + 1u) & 0x3FFFFFFFu;
      __threadfence();
    }
  }
}

// Warp-cooperative look-back: returns the exclusive prefix of tile `tile` (sum of aggregates of
// all earlier tiles).  Call with all 32 lanes of one warp.  `err` is raised on spin-bound.
// Each round inspects a window of 128 predecessor tiles: every lane issues 4 independent loads
// (L2 latency is paid once per round, not once per tile), then the four 32-tile groups are
// folded in order, nearest first.  In a single-wave launch nobody but tile 0 holds a full prefix
// early on, so the walk length, not bandwidth, is what bounds these kernels.
__device__ __forceinline__ unsigned lookback_exclusive(const unsigned long long* state, unsigned epoch, int tile, int* err) {
  unsigned exclusive = 0;
  int base = tile - 1;
  while (base >= 0) {
    unsigned st4[4], val4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int t = base - q * 32 - lane_id();
      val4[q] = 0u;
      st4[q] = (t >= 0) ? ts_load(state + t, epoch, val4[q]) : (unsigned)TS_PREFIX;  // tiles before 0 act as a zero prefix
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int t = base - q * 32 - lane_id();
      unsigned status = st4[q], value = val4[q];
      if (t >= 0 && status == TS_EMPTY) {
        int spins = 0;
        do { status = ts_load(state + t, epoch, value); } while (status == TS_EMPTY && ++spins < SPIN_BOUND);
        if (status == TS_EMPTY) { atomicExch(err, 1); status = TS_PREFIX; value = 0u; }
      }
      unsigned pref_mask = __ballot_sync(FULL, status == TS_PREFIX);
      int first = __ffs(pref_mask) - 1;  // nearest predecessor in this group that holds a full prefix
      unsigned contrib = (first < 0 || lane_id() <= first) ? value : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(FULL, contrib, o);
      exclusive += contrib;
      if (first >= 0) return exclusive;
    }
    base -= 128;
  }
  return exclusive;
}

// ---- stable stream compaction -------------------------------------------------------------
// Tile = BLOCK threads x IPT items, striped (item j of thread t is tile_base + j*BLOCK + t) so
// that global loads are coalesced.  flags[j] in, ranks[j] out (exclusive rank among the flagged
// items of the whole input, i.e. the output position), plus the running total.
// `epoch` must differ from the epoch of the previous launch that used `state` and `tile` comes from tile_begin().
template <int BLOCK, int IPT>
struct CompactSmem {
  unsigned warp_cnt[IPT * (BLOCK / 32)];
  unsigned tile_excl;
  unsigned tile_total;
};

template <int BLOCK, int IPT>
__device__ __forceinline__ void tile_compact_ranks(const bool (&flags)[IPT], unsigned (&ranks)[IPT],
                                                   unsigned& total_inclusive, unsigned long long* state, unsigned epoch,
                                                   int tile, int* err, CompactSmem<BLOCK, IPT>& sm) {
  constexpr int W = BLOCK / 32;
  const int w = threadIdx.x >> 5;
  unsigned ball[IPT];
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    ball[j] = __ballot_sync(FULL, flags[j]);
    if (lane_id() == 0) sm.warp_cnt[j * W + w] = __popc(ball[j]);
  }
  __syncthreads();
  if (w == 0) {
    // exclusive scan over IPT*W (<= 32*k) counts in (j, w) order
    unsigned carry = 0;
    for (int b = 0; b < IPT * W; b += 32) {
      int i = b + lane_id();
      unsigned v = (i < IPT * W) ? sm.warp_cnt[i] : 0u;
      unsigned inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned n = __shfl_up_sync(FULL, inc, o);
        if (lane_id() >= o) inc += n;
      }
      if (i < IPT * W) sm.warp_cnt[i] = carry + inc - v;
      carry += __shfl_sync(FULL, inc, 31);
    }
    unsigned tile_total = carry;
    if (lane_id() == 0) ts_store(state + tile, epoch, tile == 0 ? TS_PREFIX : TS_AGG, tile_total);
    unsigned excl = lookback_exclusive(state, epoch, tile, err);
    if (lane_id() == 0) {
      if (tile != 0) ts_store(state + tile, epoch, TS_PREFIX, excl + tile_total);
      sm.tile_excl = excl;
      sm.tile_total = tile_total;
    }
  }
  __syncthreads();
  const unsigned excl = sm.tile_excl;
#pragma unroll
  for (int j = 0; j < IPT; ++j) ranks[j] = excl + sm.warp_cnt[j * W + w] + __popc(ball[j] & lanemask_lt());
  total_inclusive = excl + sm.tile_total;
}

}  // namespace gm
