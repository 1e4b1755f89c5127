// sort_v1.cuh -- round 1's tile-form LSD radix sort (three kernels per 8-bit pass), kept OUT of the product library as the
// baseline of tools/microbench/sort_v2.cu.  The product sort is geometric_mapping_b200/csrc/gm_sort.cuh.
#pragma once
#include "gm_device.cuh"

namespace gm {

// ---- LSD radix sort of (key, value) u32 pairs, 8-bit digits ----------------------------------
// Three wait-free kernels per pass (no inter-block spinning at all):
//   k_rs_upsweep    tile t -> 256-bin digit histogram, stored digit-major  hist[d][t]
//   k_rs_scan       block d: exclusive scan of row d over the tiles (in place) + row total
//   k_rs_downsweep  tile t: stable in-tile ranking, then scatter to
//                   digit_base[d] (scan of the 256 totals) + row_prefix[d][t] + rank-in-tile
// History (round-1 microbenchmark, B200): a single-pass "onesweep" with decoupled look-back
// was measured at 17 us per 4096-key tile = 6.4 us of __match_any_sync ranking (~760 cycles per
// round with many distinct digits) + 4-5 us of look-back (resident tiles advance in lockstep, so
// every tile walks ~300 predecessors x 256 digits = 600 KB of L2 reads) + 3.3 us scatter, i.e.
// 1.4 TB/s.  Ranking here uses 8 ballots per key (fixed latency) instead of match_any.
// Keys per thread (RS_IPT) is a template parameter: 8 (2048-key tiles) is the default; 4 was tried for ~1M-key
// inputs (more, shorter blocks) and measured slightly slower (see gm_capi.cu:radix_sort).
constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_MAX_PASSES = 4;

template <int RS_IPT>
__global__ void __launch_bounds__(RS_BLOCK)
k_rs_upsweep(const unsigned* __restrict__ keys, const int* __restrict__ n_ptr, int pass, int tile_stride,
             unsigned* __restrict__ hist /* [256][tile_stride] */) {
  constexpr int RS_TILE = RS_BLOCK * RS_IPT;
  __shared__ unsigned sh[256];
  const int n = *n_ptr;
  const int tile = blockIdx.x, base = tile * RS_TILE;
  if (base >= n) return;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 8 * pass;
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    int g = base + i * RS_BLOCK + threadIdx.x;
    if (g < n) atomicAdd(&sh[(keys[g] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * tile_stride + tile] = sh[threadIdx.x];
}

// grid = 256 blocks (one per digit)
__global__ void __launch_bounds__(RS_BLOCK)
k_rs_scan(unsigned* __restrict__ hist, const int* __restrict__ n_ptr, int tile_keys, int tile_stride, unsigned* __restrict__ totals /* [256] */) {
  __shared__ unsigned s_warp[RS_WARPS];
  __shared__ unsigned s_carry;
  const int n = *n_ptr;
  const int ntiles = (n + tile_keys - 1) / tile_keys;
  unsigned* row = hist + (size_t)blockIdx.x * tile_stride;
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += RS_BLOCK) {
    const int t = base + threadIdx.x;
    const unsigned v = (t < ntiles) ? row[t] : 0u;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned x = __shfl_up_sync(FULL, inc, o); if (l >= o) inc += x; }
    if (l == 31) s_warp[w] = inc;
    __syncthreads();
    unsigned off = s_carry;
    for (int ww = 0; ww < w; ++ww) off += s_warp[ww];
    if (t < ntiles) row[t] = off + inc - v;
    __syncthreads();
    if (threadIdx.x == RS_BLOCK - 1) s_carry = off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

template <int RS_IPT>
__global__ void __launch_bounds__(RS_BLOCK)
k_rs_downsweep(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in, unsigned* __restrict__ keys_out,
               unsigned* __restrict__ vals_out, const int* __restrict__ n_ptr, int pass, int tile_stride,
               const unsigned* __restrict__ hist /* row prefixes */, const unsigned* __restrict__ totals) {
  constexpr int RS_TILE = RS_BLOCK * RS_IPT;
  __shared__ unsigned s_warp_hist[RS_WARPS][257];
  __shared__ unsigned s_keys[RS_TILE];
  __shared__ unsigned s_vals[RS_TILE];
  __shared__ unsigned s_local_base[256];   // position of digit d in the tile-sorted order
  __shared__ unsigned s_global_base[256];  // output position of the first key of digit d of this tile
  __shared__ unsigned s_scan[RS_WARPS];

  const int n = *n_ptr;
  const int tile = blockIdx.x;
  if (tile * RS_TILE >= n) return;
  for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_BLOCK) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();

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
  // stable ranks inside the warp chunk, order = (i, lane).  peers = lanes holding the same digit,
  // found with one ballot per digit bit (invalid tail items form their own group via bit 8).
  // The lowest peer adds the group size to the warp's digit counter with a shared-memory atomic and
  // broadcasts the old value: rounds have no register dependence on one another, and same-warp
  // atomics on one address execute in program order, so all RS_IPT rounds pipeline (the former
  // read -> __syncwarp -> write chain cost ~350 cycles per round).
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    unsigned peers = FULL;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      const bool bit = (dig[i] >> b) & 1;
      const unsigned vote = __ballot_sync(FULL, bit);
      peers &= bit ? vote : ~vote;
    }
    const int leader = __ffs(peers) - 1;
    unsigned before = 0;
    if (l == leader) before = atomicAdd(&s_warp_hist[w][dig[i]], (unsigned)__popc(peers));
    before = __shfl_sync(FULL, before, leader);
    rank[i] = before + __popc(peers & lanemask_lt());
  }
  __syncthreads();

  // per digit (thread d): exclusive offsets across warps, tile count, bases
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
    const unsigned row_prefix = hist[(size_t)d * tile_stride + tile];
    // digit base = exclusive scan of the 256 row totals; tile-local digit base = scan of tile_cnt
    unsigned hv = totals[d];
    unsigned inc = hv, linc = tile_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(FULL, inc, o), t2 = __shfl_up_sync(FULL, linc, o);
      if (l >= o) { inc += t; linc += t2; }
    }
    if (l == 31) s_scan[w] = inc;
    __syncthreads();
    unsigned hbase = 0;
    for (int ww = 0; ww < w; ++ww) hbase += s_scan[ww];
    __syncthreads();
    if (l == 31) s_scan[w] = linc;
    __syncthreads();
    unsigned lbase = 0;
    for (int ww = 0; ww < w; ++ww) lbase += s_scan[ww];
    s_local_base[d] = lbase + linc - tile_cnt;
    s_global_base[d] = hbase + inc - hv + row_prefix;
  }
  __syncthreads();

  // scatter into tile-sorted order in shared memory, then coalesced-run writes
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
