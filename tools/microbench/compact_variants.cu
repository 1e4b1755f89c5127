// Where do the ~20 us of the 1M-point stream-compaction kernels go?  (Result, B200: publishing the tile
// state with a plain st.relaxed.gpu made the look-back ~1 us per round; atomicExch publishes cut the
// kernel from 18 to 12 us, and every cudaMemsetAsync of the state array cost another ~4.8 us, hence
// the epoch-tagged states of gm_device.cuh.  ld.cg spinning never observed the update at all.)  Times k_crop (the product
// kernel, decoupled look-back) against (a) a plain float4 copy with the same tiling and (b) the
// same kernel with the look-back replaced by one atomicAdd per tile (unordered output; timing only).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../geometric_mapping_b200/csrc -o compact_variants compact_variants.cu
#include "gm_stages.cuh"
#include <cstdio>
#include <vector>
using namespace gm;

template <int IPT>
__global__ void __launch_bounds__(CP_BLOCK) k_copy(const float4* __restrict__ in, int n, float4* __restrict__ out) {
  const int base = blockIdx.x * CP_BLOCK * IPT;
  float4 p[IPT];
#pragma unroll
  for (int j = 0; j < IPT; ++j) { int i = base + j * CP_BLOCK + threadIdx.x; if (i < n) p[j] = in[i]; }
#pragma unroll
  for (int j = 0; j < IPT; ++j) { int i = base + j * CP_BLOCK + threadIdx.x; if (i < n) out[i] = p[j]; }
}

template <int IPT>
__global__ void __launch_bounds__(CP_BLOCK) k_crop_atomic(const float4* __restrict__ in, int n, float lo, float hi, float4* __restrict__ out, unsigned* counter) {
  __shared__ unsigned s_cnt[IPT * (CP_BLOCK / 32)];
  __shared__ unsigned s_base;
  const int base = blockIdx.x * CP_BLOCK * IPT, w = threadIdx.x >> 5;
  float4 p[IPT]; bool f[IPT]; unsigned ball[IPT];
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    int i = base + j * CP_BLOCK + threadIdx.x;
    f[j] = false;
    if (i < n) { p[j] = in[i]; f[j] = !((p[j].x < lo) || (p[j].y < lo) || (p[j].z < lo) || (p[j].x > hi) || (p[j].y > hi) || (p[j].z > hi)); }
    ball[j] = __ballot_sync(FULL, f[j]);
    if (lane_id() == 0) s_cnt[j * (CP_BLOCK / 32) + w] = __popc(ball[j]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int i = 0; i < IPT * (CP_BLOCK / 32); ++i) { unsigned c = s_cnt[i]; s_cnt[i] = run; run += c; }
    s_base = atomicAdd(counter, run);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < IPT; ++j) if (f[j]) out[s_base + s_cnt[j * (CP_BLOCK / 32) + w] + __popc(ball[j] & lanemask_lt())] = p[j];
}

// instrumented clone of k_crop: timestamps (ns, %globaltimer) per tile: start, loaded, published, lookback done, end
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(CP_BLOCK, 4)
k_crop_timed(const float4* __restrict__ in, int n, float lo, float hi, float4* __restrict__ out, unsigned long long* state, DevState* st,
             unsigned epoch, unsigned long long* ts /* 6 per tile */) {
  __shared__ CompactSmem<CP_BLOCK, CPL_IPT> sm;
  const int tile = blockIdx.x, base = tile * CPL_TILE;
  unsigned long long t0 = gtime();
  float4 p[CPL_IPT]; bool f[CPL_IPT];
#pragma unroll
  for (int j = 0; j < CPL_IPT; ++j) {
    int i = base + j * CP_BLOCK + threadIdx.x;
    f[j] = false;
    if (i < n) { p[j] = in[i]; f[j] = !((p[j].x < lo) || (p[j].y < lo) || (p[j].z < lo) || (p[j].x > hi) || (p[j].y > hi) || (p[j].z > hi)); }
  }
  constexpr int W = CP_BLOCK / 32;
  const int w = threadIdx.x >> 5;
  unsigned ball[CPL_IPT];
#pragma unroll
  for (int j = 0; j < CPL_IPT; ++j) { ball[j] = __ballot_sync(FULL, f[j]); if (lane_id() == 0) sm.warp_cnt[j * W + w] = __popc(ball[j]); }
  __syncthreads();
  unsigned long long t1 = gtime(), t2 = 0, t3 = 0;
  int rounds = 0;
  if (w == 0) {
    unsigned carry = 0;
    for (int b = 0; b < CPL_IPT * W; b += 32) {
      int i = b + lane_id();
      unsigned v = (i < CPL_IPT * W) ? sm.warp_cnt[i] : 0u, inc = v;
      for (int o = 1; o < 32; o <<= 1) { unsigned nn = __shfl_up_sync(FULL, inc, o); if (lane_id() >= o) inc += nn; }
      if (i < CPL_IPT * W) sm.warp_cnt[i] = carry + inc - v;
      carry += __shfl_sync(FULL, inc, 31);
    }
    unsigned tile_total = carry;
    if (lane_id() == 0) ts_store(state + tile, epoch, tile == 0 ? TS_PREFIX : TS_AGG, tile_total);
    t2 = gtime();
    unsigned excl = lookback_exclusive(state, epoch, tile, &st->error);
    t3 = gtime();
    if (lane_id() == 0) { if (tile != 0) ts_store(state + tile, epoch, TS_PREFIX, excl + tile_total); sm.tile_excl = excl; sm.tile_total = tile_total; }
  }
  __syncthreads();
  const unsigned excl = sm.tile_excl;
#pragma unroll
  for (int j = 0; j < CPL_IPT; ++j) if (f[j]) out[excl + sm.warp_cnt[j * W + w] + __popc(ball[j] & lanemask_lt())] = p[j];
  unsigned long long t4 = gtime();
  if (threadIdx.x == 0) { ts[tile * 6 + 0] = t0; ts[tile * 6 + 1] = t1; ts[tile * 6 + 2] = t2; ts[tile * 6 + 3] = t3; ts[tile * 6 + 4] = t4; unsigned sid; asm("mov.u32 %0, %%smid;" : "=r"(sid)); ts[tile * 6 + 5] = sid; }
  (void)rounds;
}

int main() {
  const int n = 1000000;
  std::vector<float4> h(n);
  srand(1);
  for (auto& p : h) p = make_float4(rand() % 1100 / 100.f - 5.5f, rand() % 500 / 100.f - 2.5f, rand() % 500 / 100.f - 2.5f, 1.f);
  float4 *d_in, *d_out; unsigned long long* d_state; DevState* d_st; unsigned* d_cnt;
  cudaMalloc(&d_in, n * 16); cudaMalloc(&d_out, n * 16); cudaMalloc(&d_state, 8 * 4096); cudaMemset(d_state, 0, 8 * 4096); cudaMalloc(&d_st, sizeof(DevState)); cudaMalloc(&d_cnt, 4);
  cudaMemcpy(d_in, h.data(), n * 16, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, auto&& launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e9, sum = 0;
    for (int i = 0; i < 20; ++i) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); sum += ms; }
    printf("%-44s best %.2f us  mean %.2f us  (%s)\n", name, best * 1e3, sum / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  time("copy IPT=4 (977 blocks)", [&] { k_copy<4><<<(n + 1023) / 1024, 256>>>(d_in, n, d_out); });
  time("copy IPT=8 (489 blocks)", [&] { k_copy<8><<<(n + 2047) / 2048, 256>>>(d_in, n, d_out); });
  time("memset state only", [&] { cudaMemsetAsync(d_state, 0, 8 * 1024); });
  unsigned epoch = 0;
  time("k_crop look-back (product, epoch-tagged)", [&] { k_crop<<<(n + CPL_TILE - 1) / CPL_TILE, CP_BLOCK>>>(d_in, n, -5.f, 5.f, 1, d_out, d_state, ++epoch, d_st); });
  time("k_crop atomic IPT=8 (memset + kernel)", [&] { cudaMemsetAsync(d_cnt, 0, 4); k_crop_atomic<8><<<(n + 2047) / 2048, 256>>>(d_in, n, -5.f, 5.f, d_out, d_cnt); });
  time("k_crop atomic IPT=4 (memset + kernel)", [&] { cudaMemsetAsync(d_cnt, 0, 4); k_crop_atomic<4><<<(n + 1023) / 1024, 256>>>(d_in, n, -5.f, 5.f, d_out, d_cnt); });
  {
    int tiles = (n + CPL_TILE - 1) / CPL_TILE;
    unsigned long long* d_ts; cudaMalloc(&d_ts, tiles * 6 * 8);
    for (int rep = 0; rep < 3; ++rep) k_crop_timed<<<tiles, CP_BLOCK>>>(d_in, n, -5.f, 5.f, d_out, d_state, d_st, ++epoch, d_ts);
    cudaDeviceSynchronize();
    std::vector<unsigned long long> ts(tiles * 6);
    cudaMemcpy(ts.data(), d_ts, tiles * 6 * 8, cudaMemcpyDeviceToHost);
    unsigned long long tmin = ~0ull; for (int t = 0; t < tiles; ++t) tmin = ts[t * 6] < tmin ? ts[t * 6] : tmin;
    printf("tile: start loaded published lookback_done end (us since first start) smid\n");
    for (int t = 0; t < tiles; t += (t < 16 ? 1 : 12)) printf("%4d: %6.2f %6.2f %6.2f %6.2f %6.2f  sm%llu\n", t, (ts[t*6]-tmin)/1e3, (ts[t*6+1]-tmin)/1e3, (ts[t*6+2]-tmin)/1e3, (ts[t*6+3]-tmin)/1e3, (ts[t*6+4]-tmin)/1e3, ts[t*6+5]);
    double mx = 0; for (int t = 0; t < tiles; ++t) mx = fmax(mx, (ts[t*6+4]-tmin)/1e3); printf("last end %.2f us\n", mx);
  }
  DevState st; cudaMemcpy(&st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
  printf("n_crop=%d err=%d\n", st.n_crop, st.error);
  return 0;
}
