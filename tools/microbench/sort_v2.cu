// Correctness (against std::stable_sort) and timing of the segment-form radix sort (gm_sort.cuh) beside the round-1
// tile-form kernels (gm_device.cuh: k_rs_upsweep / k_rs_scan / k_rs_downsweep).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -lineinfo -I../../geometric_mapping_b200/csrc -I. -o sort_v2 sort_v2.cu
//   run:   ./sort_v2            (prints one line per case; exit code 1 on any mismatch)
#include "gm_sort.cuh"
#include "sort_v1.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace gm;

static unsigned *dk[2], *dv[2], *d_hist1, *d_hist2, *d_tot, *d_flush;
static int* d_n;
static Rs2Aux* d_aux;
static size_t cap;

static int sort_v1(int n, int bits) {
  constexpr int TILE = RS_BLOCK * 8;
  const int ntiles = (n + TILE - 1) / TILE, passes = std::max(1, (bits + 7) / 8);
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    k_rs_upsweep<8><<<ntiles, RS_BLOCK>>>(dk[cur], d_n, p, ntiles, d_hist1);
    k_rs_scan<<<256, RS_BLOCK>>>(d_hist1, d_n, TILE, ntiles, d_tot);
    k_rs_downsweep<8><<<ntiles, RS_BLOCK>>>(dk[cur], dv[cur], dk[cur ^ 1], dv[cur ^ 1], d_n, p, ntiles, d_hist1, d_tot);
    cur ^= 1;
  }
  return cur;
}
static int sort_v2(int n, int bits) { return rs2_sort(0, dk, dv, d_n, (size_t)n, bits, d_hist2, d_aux, nullptr); }

static int check(int n, int bits, unsigned seed, int n_cap_extra) {
  std::vector<unsigned> hk(n), hv(n);
  unsigned s = seed;
  const unsigned mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; hk[i] = (s >> 3) & mask; hv[i] = i; }
  if (seed & 1) for (int i = 0; i < n; ++i) hk[i] = (hk[i] & 0xFF00FFu) & mask;  // many equal keys: stability
  cudaMemcpy(d_n, &n, 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dk[0], hk.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dv[0], hv.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  // n_cap (grid sizing) may exceed n (kernels read the true size from device memory)
  const int cur = rs2_sort(0, dk, dv, d_n, (size_t)n + n_cap_extra, bits, d_hist2, d_aux, nullptr);
  cudaDeviceSynchronize();
  std::vector<unsigned> ok(n), ov(n);
  cudaMemcpy(ok.data(), dk[cur], (size_t)n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ov.data(), dv[cur], (size_t)n * 4, cudaMemcpyDeviceToHost);
  std::vector<unsigned> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return hk[a] < hk[b]; });
  long long bad = 0;
  for (int i = 0; i < n; ++i) bad += (ov[i] != idx[i]) || (ok[i] != hk[idx[i]]);
  const Rs2Plan p = rs2_plan((size_t)n + n_cap_extra, bits);
  printf("check n=%d cap=+%d bits=%d passes=%d digit_bits=%d segments=%d q=%d r=%d mismatches=%lld (%s)\n", n, n_cap_extra, bits, p.passes,
         p.bits, p.segments, p.seg.q, p.seg.r, bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}

template <class F>
static float time_us(F&& f, bool flush) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, sum = 0;
  const int reps = 8;
  for (int r = 0; r < reps; ++r) {
    if (flush) cudaMemsetAsync(d_flush, r, 256u << 20);
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = fminf(best, ms); if (r) sum += ms;
  }
  (void)sum;
  return best * 1e3f;
}

static void bench(int n, int bits) {
  std::vector<unsigned> hk(n), hv(n);
  unsigned s = 12345u;
  const unsigned mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; hk[i] = (s >> 3) & mask; hv[i] = i; }
  cudaMemcpy(d_n, &n, 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dk[0], hk.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dv[0], hv.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  const Rs2Plan p = rs2_plan((size_t)n, bits);
  const int passes1 = std::max(1, (bits + 7) / 8);
  constexpr int TILE = RS_BLOCK * 8;
  const int ntiles = (n + TILE - 1) / TILE;
  for (int flush = 0; flush < 2; ++flush) {
    // (sorting an already sorted / permuted buffer again costs the same: LSD passes are data-oblivious up to digit skew)
    const float t1 = time_us([&] { sort_v1(n, bits); }, flush);
    const float t2 = time_us([&] { sort_v2(n, bits); }, flush);
    const double bytes1 = 20.0 * passes1 * n, bytes2 = 20.0 * p.passes * n;
    printf("bench n=%d bits=%d flush=%d | v1 %d passes %.1f us (%.0f GB/s on 20PM) | v2 %d passes x %d bits %.1f us (%.0f GB/s on 20PM of v1's passes: %.0f)\n", n, bits,
           flush, passes1, t1, bytes1 / t1 * 1e-3, p.passes, p.bits, t2, bytes2 / t2 * 1e-3, bytes1 / t2 * 1e-3);
  }
  cudaMemcpy(dk[0], hk.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  const unsigned m = (1u << p.bits) - 1u;
  printf("  kernels n=%d (L2 warm): v1 upsweep %.1f scan %.1f downsweep %.1f | v2 hist+memset %.1f (scan %.0f: none) hist+down %.1f us\n", n,
         time_us([&] { k_rs_upsweep<8><<<ntiles, RS_BLOCK>>>(dk[0], d_n, 0, ntiles, d_hist1); }, false),
         time_us([&] { k_rs_upsweep<8><<<ntiles, RS_BLOCK>>>(dk[0], d_n, 0, ntiles, d_hist1); k_rs_scan<<<256, RS_BLOCK>>>(d_hist1, d_n, TILE, ntiles, d_tot); }, false) ,
         time_us([&] { k_rs_downsweep<8><<<ntiles, RS_BLOCK>>>(dk[0], dv[0], dk[1], dv[1], d_n, 0, ntiles, d_hist1, d_tot); }, false),
         time_us([&] { k_rs2_hist<<<p.segments, RS2_HIST_THREADS>>>(dk[0], d_n, 0, m, p.seg, d_hist2, d_aux); cudaMemsetAsync(d_aux, 0, sizeof(Rs2Aux)); }, false),
         0.0f,
         time_us([&] { k_rs2_hist<<<p.segments, RS2_HIST_THREADS>>>(dk[0], d_n, 0, m, p.seg, d_hist2, d_aux); rs2_launch_down<8>(0, p, dk[0], dv[0], dk[1], dv[1], d_n, 0, d_hist2, d_aux); }, false));
  printf("  (the two 'scan' columns time histogram + scan together: subtract the first column)\n");
}

int main(int argc, char** argv) {
  cap = 20u << 20;
  for (int b = 0; b < 2; ++b) { cudaMalloc(&dk[b], cap * 4); cudaMalloc(&dv[b], cap * 4); }
  cudaMalloc(&d_hist1, (cap / 2048 + 1) * 256 * 4);
  cudaMalloc(&d_hist2, RS2_HIST_WORDS * 4);
  cudaMalloc(&d_tot, 1024); cudaMalloc(&d_aux, sizeof(Rs2Aux)); cudaMemset(d_aux, 0, sizeof(Rs2Aux)); cudaMalloc(&d_n, 4); cudaMalloc(&d_flush, 256u << 20);
  if (argc > 1 && atoi(argv[1]) == 2) {  // profiling mode: one 10M sort (27-bit keys) and nothing else
    check(10000000, 27, 7u, 0);
    return 0;
  }
  if (argc > 1 && atoi(argv[1]) == 3) {  // profiling mode: uniformly random 32-bit keys, 8-bit digits
    check(10000000, 32, 8u, 0);
    return 0;
  }
  int bad = 0;
  const int sizes[] = {1, 31, 2047, 2048, 2049, 100003, 1000000, 1212416 + 5, 10000000};
  const int bitsv[] = {3, 8, 9, 17, 24, 27, 32};
  unsigned seed = 1;
  for (int n : sizes)
    for (int bits : bitsv) {
      if (n >= 1000000 && (bits == 3 || bits == 9)) continue;
      bad += check(n, bits, seed++, (n % 3 == 0) ? 40000 : 0);
    }
  printf("total failing cases: %d\n", bad);
  if (argc > 1 && atoi(argv[1]) == 0) return bad != 0;
  bench(1000000, 24);
  bench(10000000, 27);
  bench(10000000, 32);
  return bad != 0;
}
