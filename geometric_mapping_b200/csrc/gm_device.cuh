// gm_device.cuh — device-side building blocks shared by the stage kernels (sm_100a).
//   * ordered-int float min/max atomics
//   * warp / block reductions (deterministic order)
//   * single-pass stable stream compaction with decoupled look-back
//   * onesweep LSD radix sort of (key,value) u32 pairs
// Everything here is hand-written; no CUB / Thrust.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gm {

constexpr unsigned FULL = 0xFFFFFFFFu;

// Spin bound for every inter-block wait.  A predecessor tile is always resident or finished
// (tiles are taken in launch order), so the bound is never reached in a correct run; if it is,
// the kernel raises device_error instead of hanging the GPU.
constexpr int SPIN_BOUND = 1 << 22;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

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
// `epoch` must differ from the epoch of the previous launch that used `state`; tiles are blockIdx.x.
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

// ---- onesweep radix sort ------------------------------------------------------------------
constexpr int RS_BLOCK = 256;
constexpr int RS_IPT = 16;  // 4096 keys per tile: half as many tiles -> half as long look-back walks in a single-wave launch
constexpr int RS_TILE = RS_BLOCK * RS_IPT;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_MAX_PASSES = 4;

// One read of the keys -> digit histograms of all passes.  hist[pass*256 + d] must be zero on entry:
// the kernel that PRODUCES the keys clears it (d_zero_hist), so no memset node is needed.
__device__ __forceinline__ void d_zero_hist(unsigned* hist) {
  if (blockIdx.x == 0) for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += blockDim.x) hist[i] = 0u;
}

__global__ void __launch_bounds__(RS_BLOCK) k_radix_hist(const unsigned* __restrict__ keys, const int* __restrict__ n_ptr,
                                                         int passes, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[RS_MAX_PASSES * 256];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += RS_BLOCK) sh[i] = 0;
  __syncthreads();
  const int n = *n_ptr;
  for (int i = blockIdx.x * RS_BLOCK + threadIdx.x; i < n; i += gridDim.x * RS_BLOCK) {
    unsigned k = keys[i];
    for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * 256 + ((k >> (8 * p)) & 255u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += RS_BLOCK) {
    unsigned v = sh[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// One LSD pass.  state: [tiles][256] epoch-tagged words per pass (never cleared); the ticket counter
// is reset by whichever block draws the last ticket.
__global__ void __launch_bounds__(RS_BLOCK)
k_radix_onesweep(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                 unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                 const int* __restrict__ n_ptr, int pass, const unsigned* __restrict__ hist,
                 unsigned long long* __restrict__ state, unsigned epoch, unsigned* __restrict__ ticket, int* __restrict__ err) {
  __shared__ unsigned s_warp_hist[RS_WARPS][257];
  __shared__ unsigned s_keys[RS_TILE];
  __shared__ unsigned s_vals[RS_TILE];
  __shared__ unsigned s_local_base[256];   // position of digit d in the tile-sorted order
  __shared__ unsigned s_global_base[256];  // output position of the first key of digit d of this tile
  __shared__ unsigned s_scan[RS_WARPS];
  __shared__ int s_tile;

  const int n = *n_ptr;
  const int ntiles = (n + RS_TILE - 1) / RS_TILE;
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) *ticket = 0u;  // everybody has drawn: ready for the next launch
    s_tile = (int)t;
  }
  for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_BLOCK) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int tile = s_tile;
  if (tile >= ntiles) return;

  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int shift = 8 * pass;
  const int chunk = tile * RS_TILE + w * (32 * RS_IPT);

  unsigned key[RS_IPT], val[RS_IPT], rank[RS_IPT];
  int dig[RS_IPT];
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    int g = chunk + i * 32 + l;
    bool ok = g < n;
    key[i] = ok ? keys_in[g] : 0xFFFFFFFFu;
    val[i] = ok ? vals_in[g] : 0u;
    dig[i] = ok ? (int)((key[i] >> shift) & 255u) : 256;
  }
  // stable ranks inside the warp chunk: order = (i, lane)
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    unsigned m = __match_any_sync(FULL, dig[i]);
    unsigned before = s_warp_hist[w][dig[i]];
    __syncwarp();
    rank[i] = before + __popc(m & lanemask_lt());
    if ((m & lanemask_lt()) == 0) s_warp_hist[w][dig[i]] = before + __popc(m);
    __syncwarp();
  }
  __syncthreads();

  // per digit (thread d): exclusive offsets across warps, tile count, look-back
  {
    const int d = threadIdx.x;
    unsigned run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      unsigned c = s_warp_hist[ww][d];
      s_warp_hist[ww][d] = run;
      run += c;
    }
    const unsigned tile_cnt = run;
    unsigned long long* st = state + (size_t)tile * 256 + d;
    ts_store(st, epoch, tile == 0 ? TS_PREFIX : TS_AGG, tile_cnt);

    // global exclusive digit offset = sum(hist[pass][0..d))  (block scan over 256 digits)
    unsigned hv = hist[pass * 256 + d];
    unsigned inc = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(FULL, inc, o);
      if (l >= o) inc += t;
    }
    if (l == 31) s_scan[w] = inc;
    // tile-local digit base (scan of tile_cnt over d)
    unsigned linc = tile_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(FULL, linc, o);
      if (l >= o) linc += t;
    }
    __syncthreads();
    unsigned hbase = 0;
    for (int ww = 0; ww < w; ++ww) hbase += s_scan[ww];
    const unsigned digit_global = hbase + inc - hv;
    __syncthreads();
    if (l == 31) s_scan[w] = linc;
    __syncthreads();
    unsigned lbase = 0;
    for (int ww = 0; ww < w; ++ww) lbase += s_scan[ww];
    s_local_base[d] = lbase + linc - tile_cnt;

    // look back over earlier tiles for this digit: 16 independent loads per step (one L2 round
    // trip per 16 tiles), folded nearest first until a tile with a full prefix is met
    unsigned excl = 0;
    bool done = false;
    for (int t0 = tile - 1; t0 >= 0 && !done; t0 -= 16) {
      unsigned wst[16], wval[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        wval[q] = 0u;
        wst[q] = (t0 - q >= 0) ? ts_load(state + (size_t)(t0 - q) * 256 + d, epoch, wval[q]) : (unsigned)TS_PREFIX;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (done) break;
        if (wst[q] == TS_EMPTY) {
          const unsigned long long* ps = state + (size_t)(t0 - q) * 256 + d;
          int spins = 0;
          do { wst[q] = ts_load(ps, epoch, wval[q]); } while (wst[q] == TS_EMPTY && ++spins < SPIN_BOUND);
          if (wst[q] == TS_EMPTY) { atomicExch(err, 2); done = true; break; }
        }
        excl += wval[q];
        if (wst[q] == TS_PREFIX) done = true;
      }
    }
    if (tile != 0) ts_store(st, epoch, TS_PREFIX, excl + tile_cnt);
    s_global_base[d] = digit_global + excl;
  }
  __syncthreads();

  // scatter into tile-sorted order in shared memory
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    if (dig[i] < 256) {
      unsigned pos = s_local_base[dig[i]] + s_warp_hist[w][dig[i]] + rank[i];
      s_keys[pos] = key[i];
      s_vals[pos] = val[i];
    }
  }
  __syncthreads();
  const int tile_n = min(RS_TILE, n - tile * RS_TILE);
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    int idx = i * RS_BLOCK + threadIdx.x;
    if (idx < tile_n) {
      unsigned k = s_keys[idx];
      int d = (int)((k >> shift) & 255u);
      unsigned g = s_global_base[d] + (unsigned)idx - s_local_base[d];
      keys_out[g] = k;
      vals_out[g] = s_vals[idx];
    }
  }
}

}  // namespace gm
