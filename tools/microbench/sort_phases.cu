// Phase timing of the LSD downsweep (clone of gm::k_rs_downsweep with %globaltimer stamps) plus a
// correctness check of the product sort kernels against std::stable_sort.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I../../geometric_mapping_b200/csrc -o sort_phases sort_phases.cu
#include "gm_device.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace gm;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(RS_BLOCK)
k_down_timed(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in, unsigned* __restrict__ keys_out,
               unsigned* __restrict__ vals_out, const int* __restrict__ n_ptr, int pass, int tile_stride,
               const unsigned* __restrict__ hist, const unsigned* __restrict__ totals, unsigned long long* ts) {
  unsigned long long T0 = gtime();
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
  unsigned long long T1 = gtime();
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
  unsigned long long T2 = gtime();
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

  unsigned long long T3 = gtime();
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
  unsigned long long T4 = gtime();
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
  unsigned long long T5 = gtime();
  if (threadIdx.x == 0) { unsigned long long* o = ts + (size_t)tile * 8; o[0]=T0; o[1]=T1; o[2]=T2; o[3]=T3; o[4]=T4; o[5]=T5; }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  const int bits = argc > 2 ? atoi(argv[2]) : 23;
  const int passes = (bits + 7) / 8;
  std::vector<unsigned> hk(n), hv(n);
  srand(7);
  for (int i = 0; i < n; ++i) { hk[i] = ((unsigned)rand() * 2654435761u) & ((1u << bits) - 1); hv[i] = i; }
  const int ntiles = (n + RS_TILE - 1) / RS_TILE;
  unsigned *dk[2], *dv[2], *d_hist, *d_tot; unsigned long long* d_ts; int* d_n;
  for (int b = 0; b < 2; ++b) { cudaMalloc(&dk[b], n * 4); cudaMalloc(&dv[b], n * 4); }
  cudaMalloc(&d_hist, (size_t)ntiles * 256 * 4); cudaMalloc(&d_tot, 1024); cudaMalloc(&d_ts, (size_t)ntiles * 64); cudaMalloc(&d_n, 4);
  cudaMemcpy(d_n, &n, 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dk[0], hk.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dv[0], hv.data(), n * 4, cudaMemcpyHostToDevice);
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    k_rs_upsweep<<<ntiles, RS_BLOCK>>>(dk[cur], d_n, p, ntiles, d_hist);
    k_rs_scan<<<256, RS_BLOCK>>>(d_hist, d_n, ntiles, d_tot);
    k_rs_downsweep<<<ntiles, RS_BLOCK>>>(dk[cur], dv[cur], dk[cur ^ 1], dv[cur ^ 1], d_n, p, ntiles, d_hist, d_tot);
    cur ^= 1;
  }
  std::vector<unsigned> ok(n), ov(n);
  cudaMemcpy(ok.data(), dk[cur], n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ov.data(), dv[cur], n * 4, cudaMemcpyDeviceToHost);
  std::vector<unsigned> idx(n); for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return hk[a] < hk[b]; });
  int bad = 0; for (int i = 0; i < n; ++i) bad += (ov[i] != idx[i]) || (ok[i] != hk[idx[i]]);
  printf("n=%d bits=%d passes=%d tiles=%d mismatches=%d (%s)\n", n, bits, passes, ntiles, bad, cudaGetErrorString(cudaGetLastError()));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto&& f) { float best = 1e9; for (int r = 0; r < 6; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); } printf("%-10s %.1f us\n", name, best * 1e3); };
  cudaMemcpy(dk[0], hk.data(), n * 4, cudaMemcpyHostToDevice);
  timeit("upsweep", [&] { k_rs_upsweep<<<ntiles, RS_BLOCK>>>(dk[0], d_n, 0, ntiles, d_hist); });
  timeit("scan", [&] { k_rs_scan<<<256, RS_BLOCK>>>(d_hist, d_n, ntiles, d_tot); });
  k_rs_upsweep<<<ntiles, RS_BLOCK>>>(dk[0], d_n, 0, ntiles, d_hist); k_rs_scan<<<256, RS_BLOCK>>>(d_hist, d_n, ntiles, d_tot);
  timeit("downsweep", [&] { k_rs_downsweep<<<ntiles, RS_BLOCK>>>(dk[0], dv[0], dk[1], dv[1], d_n, 0, ntiles, d_hist, d_tot); });
  k_down_timed<<<ntiles, RS_BLOCK>>>(dk[0], dv[0], dk[1], dv[1], d_n, 0, ntiles, d_hist, d_tot, d_ts);
  cudaDeviceSynchronize();
  std::vector<unsigned long long> ts((size_t)ntiles * 8);
  cudaMemcpy(ts.data(), d_ts, ts.size() * 8, cudaMemcpyDeviceToHost);
  double ph[5] = {0}; unsigned long long tmin = ~0ull, tmax = 0;
  for (int t = 0; t < ntiles; ++t) { for (int k = 0; k < 5; ++k) ph[k] += (double)(ts[t * 8 + k + 1] - ts[t * 8 + k]); tmin = std::min(tmin, ts[t * 8]); tmax = std::max(tmax, ts[t * 8 + 5]); }
  const char* names[5] = {"init+load", "rank(ballot)", "digit bases", "smem scatter", "global store"};
  printf("downsweep wall %.1f us; mean per-tile phase times (us):", (tmax - tmin) / 1e3);
  for (int k = 0; k < 5; ++k) printf("  %s %.2f", names[k], ph[k] / ntiles / 1e3);
  printf("\n");
  return 0;
}
