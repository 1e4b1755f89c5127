// Host-only check of the bookkeeping of the segment-form radix sort (geometric_mapping_b200/csrc/gm_sort.cuh): the plan
// (passes, digit width, segments), the segment geometry, and the position arithmetic of one pass (segment histograms ->
// group sums -> first output position per (segment, digit) -> stable scatter) replayed on the CPU with the library's own
// rs2_plan / rs2_segment_of and constants, against std::stable_sort.  No kernel is launched: runs without a GPU.
//   nvcc -std=c++17 -I geometric_mapping_b200/csrc -o sort_plan_check tests/cpp/sort_plan_check.cu
#include "gm_sort.cuh"
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <vector>
using namespace gm;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { ++fails; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static void check_plan(size_t n_cap, int n, int key_bits) {
  const Rs2Plan p = rs2_plan(n_cap, key_bits);
  CHECK(p.passes >= 1 && p.passes <= 4, "passes %d", p.passes);
  CHECK(p.bits >= 4 && p.bits <= 8, "bits %d", p.bits);
  CHECK(p.bits * p.passes >= std::min(std::max(key_bits, 1), 32), "digits cover %d bits: %d x %d", key_bits, p.passes, p.bits);
  CHECK(p.segments >= 1 && p.segments <= RS2_MAX_SEGMENTS, "segments %d", p.segments);
  CHECK((size_t)p.segments * 256 <= RS2_HIST_WORDS, "hist words");
  // the segments tile [0, n) without gaps or overlap, and no segment holds more than one chunk more than another
  int expect = 0, lo = 1 << 30, hi = 0;
  for (int b = 0; b < p.segments; ++b) {
    int beg, end;
    rs2_segment_of(p.seg, b, n, beg, end);
    CHECK(beg == std::min(expect, n), "n_cap=%zu n=%d segment %d begins at %d, expected %d", n_cap, n, b, beg, expect);
    CHECK(end >= beg && end <= n, "segment %d end %d", b, end);
    CHECK(beg >= n || beg % RS2_CHUNK == 0, "an active segment starts on a chunk boundary (16-byte aligned copies)");
    const int chunks = p.seg.q + (b < p.seg.r ? 1 : 0);
    lo = std::min(lo, chunks); hi = std::max(hi, chunks);
    expect = (int)std::min<long long>((long long)expect + (long long)chunks * RS2_CHUNK, 1ll << 30);
  }
  CHECK(expect >= n, "n_cap=%zu: the segments end at %d < n=%d", n_cap, expect, n);
  CHECK(hi - lo <= 1, "uneven segments: %d..%d chunks", lo, hi);
}

// one LSD sort replayed with the kernels' arithmetic
static void check_sort(int n, size_t n_cap, int key_bits, unsigned seed) {
  std::vector<unsigned> key(n), val(n);
  const unsigned kmask = key_bits >= 32 ? 0xFFFFFFFFu : ((1u << key_bits) - 1u);
  unsigned s = seed;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; key[i] = (s >> 5) & kmask; val[i] = (unsigned)i; }
  std::vector<unsigned> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return key[a] < key[b]; });
  const Rs2Plan p = rs2_plan(n_cap, key_bits);
  const unsigned mask = (1u << p.bits) - 1u;
  std::vector<unsigned> k2(n), v2(n);
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.bits;
    std::vector<unsigned> hist((size_t)p.segments * 256, 0u), group((size_t)RS2_GROUPS * 256, 0u);
    for (int b = 0; b < p.segments; ++b) {  // k_rs2_hist
      int beg, end;
      rs2_segment_of(p.seg, b, n, beg, end);
      for (int g = beg; g < end; ++g) ++hist[(size_t)b * 256 + ((key[g] >> shift) & mask)];
      for (int d = 0; d < 256; ++d) group[(size_t)(b / RS2_GROUP) * 256 + d] += hist[(size_t)b * 256 + d];
    }
    const int ngrp = (p.segments + RS2_GROUP - 1) / RS2_GROUP;
    CHECK(ngrp <= RS2_GROUPS, "groups");
    for (int b = 0; b < p.segments; ++b) {  // k_rs2_down
      int beg, end;
      rs2_segment_of(p.seg, b, n, beg, end);
      if (beg >= n) continue;
      unsigned gbase[256], run = 0;
      for (int d = 0; d < 256; ++d) {
        unsigned total = 0, pre = 0;
        for (int g = 0; g < ngrp; ++g) { const unsigned v = group[(size_t)g * 256 + d]; total += v; if (g < b / RS2_GROUP) pre += v; }
        for (int bb = (b / RS2_GROUP) * RS2_GROUP; bb < b; ++bb) pre += hist[(size_t)bb * 256 + d];
        gbase[d] = run + pre;
        run += total;
      }
      for (int g = beg; g < end; ++g) {  // chunk by chunk in the kernel; the running offset per digit is the same thing
        const unsigned d = (key[g] >> shift) & mask;
        const unsigned o = gbase[d]++;
        if (o >= (unsigned)n) { CHECK(false, "position %u out of range", o); continue; }
        k2[o] = key[g]; v2[o] = val[g];
      }
    }
    key.swap(k2); val.swap(v2);
  }
  long long bad = 0;
  for (int i = 0; i < n; ++i) bad += val[i] != idx[i];
  CHECK(bad == 0, "n=%d n_cap=%zu bits=%d: %lld keys out of place", n, n_cap, key_bits, bad);
}

int main() {
  const int sizes[] = {0, 1, 2047, 2048, 2049, 32768, 1000000, 1212416, 1212416 + 1, 9855675, 10000000, 100000000, (1 << 30) - 1};
  for (int n : sizes)
    for (int bits : {1, 4, 8, 9, 16, 17, 24, 25, 27, 32}) {
      check_plan((size_t)n, n, bits);
      check_plan(((size_t)n + 32767) / 32768 * 32768, n, bits);  // grids are sized by the 32768-point bucket above n
    }
  check_sort(1, 32768, 8, 1u);
  check_sort(2049, 2049, 9, 2u);
  check_sort(70001, 98304, 17, 3u);
  check_sort(1300000, 1310720, 24, 4u);  // 635 chunks: more than 592 -> uneven segments, 19 groups
  check_sort(1300000, 1310720, 27, 5u);
  check_sort(300000, 327680, 32, 6u);
  std::printf("%s (%d failures)\n", fails ? "FAILED" : "ok", fails);
  return fails != 0;
}
