// gm_sort.cuh — LSD radix sort of (key, value) u32 pairs, segment form (round 2).
//
// Replaces `std::sort` of the (cell key, point index) pairs in pcl::VoxelGrid (/root/reference
// src/tunnel_processing.cpp:215-220) and the kd-tree build of pcl::NormalEstimation (:58-70): both
// become one stable sort of 32-bit keys with their 32-bit payload.
//
// The input is cut into G <= 148*4 contiguous SEGMENTS (a whole number of 256*IPT-key chunks each, dealt out evenly).
// One pass = TWO wait-free kernels, no inter-block spinning, no memsets:
//   k_rs2_hist  block b: digit histogram of segment b               -> hist[b][256]
//               per-thread BYTE counters in shared memory: thread t owns byte (d & 3) of word
//               [(d >> 2) * 128 + t] -- bank = lane, so the read-modify-write needs no atomics and has no
//               bank conflicts (the shared-memory atomics of the round-1 histogram ran at ~0.8 cycles per
//               key and SM: 28 us for 10M keys); columns are summed with packed 16-bit adds.
//               + the block's counts added (global atomics, 256 per block) to the sums of its GROUP of 32 segments
//               (Rs2Aux)
//   k_rs2_down  block b first sums, per digit, the groups before its own and the segments before it inside its group
//               (<= 18 + 31 coalesced L2 loads per thread, all independent: this replaces a separate scan kernel, which
//               cost 7 us per pass), then walks its segment chunk by chunk: stable ranks inside the warp by one ballot per
//               digit bit, chunk-sorted order in shared memory, runs written to
//               digit_base[d] + hist[b][d] + (keys of digit d in the earlier chunks of the segment).
//               The running per-digit offset lives in a register of thread d: no global read per chunk;
//               the next chunk arrives by warp-private cp.async while the current one is ranked and written.
// The last block to finish that prologue clears Rs2Aux for the next pass (self-cleaning: the buffer is zero between sorts).
// Digits have BITS = ceil(key_bits / passes) <= 8 bits (27-bit keys: four 7-bit passes, 7 ballots per key).
#pragma once
#include "gm_device.cuh"

namespace gm {

constexpr int RS2_HIST_THREADS = 128;
constexpr int RS2_DOWN_THREADS = 256;
constexpr int RS2_DOWN_WARPS = RS2_DOWN_THREADS / 32;
constexpr int RS2_MAX_SEGMENTS = 148 * 4;
constexpr int RS2_GROUP = 32;  // segments per group
constexpr int RS2_GROUPS = (RS2_MAX_SEGMENTS + RS2_GROUP - 1) / RS2_GROUP;
constexpr int RS2_IPT = 8;
constexpr int RS2_CHUNK = RS2_DOWN_THREADS * RS2_IPT;

// Segment b covers chunks [b * q + min(b, r), ... + q + (b < r)): the chunks are dealt out as evenly as they go (the first r
// segments hold one more), so that with 592 segments every SM -- blocks are handed to the SMs round robin -- gets the same
// number of chunks to within one.  (Equal segments of ceil(chunks / 592) chunks left 49 of the 148 SMs with 3 blocks
// instead of 4 and idle for the last quarter of the kernel: 35 % of the warp slots active instead of 46 %.)
struct Rs2Seg {
  int q, r;
};
__host__ __device__ __forceinline__ void rs2_segment_of(const Rs2Seg sg, long long b, int n, int& beg, int& end) {
  const long long first = b * sg.q + (b < sg.r ? b : sg.r);
  const long long nch = sg.q + (b < sg.r ? 1 : 0);
  const long long bb = first * RS2_CHUNK, ee = bb + nch * RS2_CHUNK;
  beg = (int)(bb < n ? bb : n);
  end = (int)(ee < n ? ee : n);
}
__device__ __forceinline__ void rs2_segment(const Rs2Seg sg, int n, int& beg, int& end) { rs2_segment_of(sg, blockIdx.x, n, beg, end); }

// Device-side sums of one pass; all zero between passes and between sorts (cleared by k_rs2_down).
struct Rs2Aux {
  unsigned group[RS2_GROUPS][256];  // per digit: keys in the 32 segments of group g
  unsigned done;                    // blocks of k_rs2_down that have read what they need
};

// ---- segment histogram --------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS2_HIST_THREADS)
k_rs2_hist(const unsigned* __restrict__ keys, const int* __restrict__ n_ptr, int shift, unsigned mask, Rs2Seg seg,
           unsigned* __restrict__ hist /* [G][256] */, Rs2Aux* __restrict__ aux) {
  constexpr int T = RS2_HIST_THREADS;
  constexpr int ROUND = T * 240;  // a byte counter holds 255 keys of one thread; 240 = 15 whole batches of 16
  __shared__ unsigned s_cnt[64 * T];
  unsigned char* s_bytes = reinterpret_cast<unsigned char*>(s_cnt);
  const int n = *n_ptr;
  const int t = threadIdx.x;
  int beg, end;
  rs2_segment(seg, n, beg, end);
  const int q = t >> 1, half = t & 1;
  unsigned tot0 = 0, tot1 = 0, tot2 = 0, tot3 = 0;  // digits 4q .. 4q+3, this thread's half of the columns
  for (int base = beg; base < end; base += ROUND) {
#pragma unroll
    for (int i = 0; i < 64; ++i) s_cnt[i * T + t] = 0;  // own column only: no barrier needed before counting
    const int rend = min(end, base + ROUND);
    // whole batches of 16 * T keys as four 16-byte loads per thread, no bounds checks (every segment but the last is a
    // whole number of batches; segment starts are 16-byte aligned), the next batch's loads in flight while this one is
    // counted: 32 keys per thread in flight -- with 4-byte loads the kernel sat on the load latency (ncu: long scoreboard).
    // The order in which a segment's keys are counted does not matter.  The ragged rest of the last segment: one by one.
    constexpr int BATCH = T * 16;
    const int nfull = (rend - base) / BATCH;
    const uint4* kp = reinterpret_cast<const uint4*>(keys + base) + t;
    uint4 k[4], kn[4];
    if (nfull > 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) kn[u] = kp[u * T];
    }
    for (int bt = 0; bt < nfull; ++bt) {
#pragma unroll
      for (int u = 0; u < 4; ++u) k[u] = kn[u];
      kp += BATCH / 4;
      if (bt + 1 < nfull) {
#pragma unroll
        for (int u = 0; u < 4; ++u) kn[u] = kp[u * T];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned kk[4] = {k[u].x, k[u].y, k[u].z, k[u].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const unsigned d = (kk[c] >> shift) & mask;
          s_bytes[((d >> 2) * T + t) * 4 + (d & 3u)] += 1;
        }
      }
    }
    for (int g = base + nfull * BATCH + t; g < rend; g += T) {
      const unsigned d = (keys[g] >> shift) & mask;
      s_bytes[((d >> 2) * T + t) * 4 + (d & 3u)] += 1;
    }
    __syncthreads();
    unsigned lo = 0, hi = 0;  // packed 16-bit sums of bytes (0,2) and (1,3): at most 64 * 255 each
#pragma unroll 8
    for (int i = 0; i < T / 2; ++i) {
      const unsigned w = s_cnt[q * T + half * (T / 2) + ((i + t) & (T / 2 - 1))];
      lo += w & 0x00FF00FFu;
      hi += (w >> 8) & 0x00FF00FFu;
    }
    tot0 += lo & 0xFFFFu; tot2 += lo >> 16;
    tot1 += hi & 0xFFFFu; tot3 += hi >> 16;
    __syncthreads();
  }
  tot0 += __shfl_xor_sync(FULL, tot0, 1);
  tot1 += __shfl_xor_sync(FULL, tot1, 1);
  tot2 += __shfl_xor_sync(FULL, tot2, 1);
  tot3 += __shfl_xor_sync(FULL, tot3, 1);
  if (half == 0) {
    reinterpret_cast<uint4*>(hist + (size_t)blockIdx.x * 256)[q] = make_uint4(tot0, tot1, tot2, tot3);
    const unsigned tot[4] = {tot0, tot1, tot2, tot3};
    unsigned* grp = aux->group[blockIdx.x / RS2_GROUP];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (tot[c]) atomicAdd(&grp[4 * q + c], tot[c]);  // 32 blocks per address (one sum of all blocks per digit, 592 blocks per
                                                       // address, was measured: +20 us per pass at 1M keys)
    }
  }
}

// peers &= lanes whose digit agrees with mine in the bit `bitmask`: 4 instructions (LOP3 -> predicate, VOTE, SEL, one
// three-input LOP3 = peers & ~(vote ^ m), m = all ones where my bit is set); the C++ form
// `peers &= bit ? vote : ~vote` compiles to 6-7, a pair of predicated LOP3 to 5.
__device__ __forceinline__ unsigned rs2_match_bit(unsigned peers, unsigned dig, unsigned bitmask) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 v, m;\n\t"
      "and.b32 v, %1, %2;\n\t"
      "setp.ne.u32 p, v, 0;\n\t"
      "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
      "selp.b32 m, 0xffffffff, 0, p;\n\t"
      "lop3.b32 %0, %0, v, m, 0x90;\n\t"
      "}"
      : "+r"(peers)
      : "r"(dig), "r"(bitmask));
  return peers;
}

// ---- stable scatter of a segment ---------------------------------------------------------------------------
__device__ __forceinline__ void rs2_cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void rs2_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared memory of one block (41 KB: four blocks per SM)
struct Rs2DownSmem {
  unsigned warp_hist[RS2_DOWN_WARPS][256];  // row w: digit counters of warp w while ranking, then the position of the warp's
                                            // first key of each digit in the chunk-sorted order
  uint2 kv[RS2_CHUNK + RS2_CHUNK / 16];     // the chunk in sorted order, one pad element per 16 (rs2_kv_slot): with evenly
                                            // filled digits the lanes of a warp write positions run_length * digit + (about the
                                            // same offset) -- unpadded that is 4-8 banks for 32 lanes (measured: 8 wavefronts
                                            // per scatter STS)
  unsigned in_keys[RS2_CHUNK];              // the NEXT chunk, landing here (cp.async) while this one is ranked and written:
  unsigned in_vals[RS2_CHUNK];              //   warp w owns [w * 32 * IPT, (w + 1) * 32 * IPT) -- no block barrier involved
  unsigned global_base[256];
  unsigned scan[RS2_DOWN_WARPS];
};

__device__ __forceinline__ unsigned rs2_kv_slot(unsigned pos) { return pos + (pos >> 4); }

// warp-private prefetch of the warp's 32 * IPT keys (or values) of the chunk at `cb` (16-byte copies, zero-filled past `end`)
__device__ __forceinline__ void rs2_prefetch(unsigned* s_dst, const unsigned* __restrict__ src_arr, int cb, int end) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < RS2_IPT / 4; ++j) {
    const int off = w * (32 * RS2_IPT) + (j * 32 + l) * 4;  // first of the 4 elements this copy moves
    const int left = end - (cb + off);
    const int bytes = left >= 4 ? 16 : (left > 0 ? left * 4 : 0);
    const int src = bytes ? cb + off : cb;  // a copy of 0 bytes reads nothing; keep its address inside the array anyway
    rs2_cp_async16(&s_dst[off], src_arr + src, bytes);
  }
}

template <int BITS, bool FULL_CHUNK>
__device__ __forceinline__ void rs2_chunk(Rs2DownSmem& sm, const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                          unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out, int cb, int n_valid,
                                          int end, int n, int shift, unsigned& gbase) {
  constexpr unsigned MASK = (1u << BITS) - 1u;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  unsigned key[RS2_IPT], val[RS2_IPT], rank[RS2_IPT];
  const int wbase = w * (32 * RS2_IPT) + l;  // position in the chunk of this lane's first key
  rs2_cp_async_wait();
  __syncwarp();
#pragma unroll
  for (int i = 0; i < RS2_IPT; ++i) {
    const int p = wbase + i * 32;
    key[i] = sm.in_keys[p];
    val[i] = sm.in_vals[p];
    if (!FULL_CHUNK && p >= n_valid) key[i] = 0xFFFFFFFFu;  // tail: the highest digit, after every valid key -> never written out
  }
  __syncwarp();
  if (FULL_CHUNK && cb + RS2_CHUNK < end) {  // (leaving the values in shared memory until the scatter -- 48 registers, 5 blocks
    rs2_prefetch(sm.in_keys, keys_in, cb + RS2_CHUNK, end);  //  per SM -- was measured 3 % slower)
    rs2_prefetch(sm.in_vals, vals_in, cb + RS2_CHUNK, end);
  }

  // stable ranks inside the warp's 32*IPT keys, order = (i, lane).  peers = lanes holding the same digit: one ballot per
  // digit bit.  The warp's digit counter is read by every peer (one broadcast LDS) and rewritten by the HIGHEST peer, whose
  // new value is its own rank + 1: plain LDS + STS, no atomics (measured equal to one ATOMS by the lowest peer + a shuffle).
#pragma unroll
  for (int i = 0; i < RS2_IPT; ++i) {
    const unsigned dig = (key[i] >> shift) & MASK;
    unsigned peers = FULL;
#pragma unroll
    for (int b = 0; b < BITS; ++b) peers = rs2_match_bit(peers, dig, 1u << b);
    const unsigned before = sm.warp_hist[w][dig];
    __syncwarp();
    rank[i] = before + __popc(peers & lanemask_lt());
    if ((peers >> l) == 1u) sm.warp_hist[w][dig] = rank[i] + 1u;
    __syncwarp();
  }
  __syncthreads();

  // thread d: the digit's keys per warp -> offsets of the warps inside the digit's run, the run's place in the
  // chunk-sorted order (exclusive scan over the digits) and in the output
  {
    const int d = threadIdx.x;
    unsigned c[RS2_DOWN_WARPS];
    unsigned cnt = 0;
#pragma unroll
    for (int ww = 0; ww < RS2_DOWN_WARPS; ++ww) { c[ww] = sm.warp_hist[ww][d]; cnt += c[ww]; }
    unsigned inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(FULL, inc, o); if (l >= o) inc += x; }
    if (l == 31) sm.scan[w] = inc;
    __syncthreads();
    unsigned lbase = inc - cnt;
#pragma unroll
    for (int ww = 0; ww < RS2_DOWN_WARPS; ++ww) lbase += ww < w ? sm.scan[ww] : 0u;
    unsigned run = lbase;
#pragma unroll
    for (int ww = 0; ww < RS2_DOWN_WARPS; ++ww) { sm.warp_hist[ww][d] = run; run += c[ww]; }
    sm.global_base[d] = gbase - lbase;  // output index of chunk-sorted position p (digit d) = global_base[d] + p
    gbase += cnt;
  }
  __syncthreads();

#pragma unroll
  for (int i = 0; i < RS2_IPT; ++i) {
    const unsigned dig = (key[i] >> shift) & MASK;
    sm.kv[rs2_kv_slot(sm.warp_hist[w][dig] + rank[i])] = make_uint2(key[i], val[i]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) sm.warp_hist[w][j * 32 + l] = 0;  // own row, for the next chunk's ranking (nobody else touches it before)
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS2_IPT; ++i) {
    const int p = i * RS2_DOWN_THREADS + threadIdx.x;
    if (FULL_CHUNK || p < n_valid) {
      const uint2 kv = sm.kv[rs2_kv_slot((unsigned)p)];
      const unsigned g = sm.global_base[(kv.x >> shift) & MASK] + (unsigned)p;
      if (g < (unsigned)n) {  // always true when Rs2Aux was clean on entry; keeps a broken invariant from writing out of bounds
        keys_out[g] = kv.x;
        vals_out[g] = kv.y;
      }
    }
  }
  // no barrier here: the next chunk writes sm.scan / sm.global_base / sm.kv only after its own first block barrier, which
  // no thread passes before every thread has left this loop
}

#ifndef RS2_DOWN_MIN_BLOCKS
#define RS2_DOWN_MIN_BLOCKS 4
#endif
template <int BITS>
__global__ void __launch_bounds__(RS2_DOWN_THREADS, RS2_DOWN_MIN_BLOCKS)
k_rs2_down(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in, unsigned* __restrict__ keys_out,
           unsigned* __restrict__ vals_out, const int* __restrict__ n_ptr, int shift, Rs2Seg seg,
           const unsigned* __restrict__ hist /* [G][256] segment histograms */, Rs2Aux* aux) {
  __shared__ __align__(16) Rs2DownSmem sm;
  __shared__ bool s_last;
  const int n = *n_ptr;
  int beg, end;
  rs2_segment(seg, n, beg, end);
  const bool active = beg < n;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  unsigned gbase = 0;
  if (active) {
    rs2_prefetch(sm.in_keys, keys_in, beg, end);
    rs2_prefetch(sm.in_vals, vals_in, beg, end);
#pragma unroll
    for (int j = 0; j < 8; ++j) sm.warp_hist[w][j * 32 + l] = 0;
    // thread d: first output position of digit d for this segment = (keys of lower digits) + (keys of digit d in the
    // earlier segments = whole groups before mine + the segments before me in my group)
    const int grp = blockIdx.x / RS2_GROUP, ngrp = (gridDim.x + RS2_GROUP - 1) / RS2_GROUP;
    unsigned pre = 0, hv = 0;  // hv: keys of digit d in the whole input
#pragma unroll 4
    for (int g = 0; g < ngrp; ++g) {
      const unsigned v = __ldcg(&aux->group[g][threadIdx.x]);
      hv += v;
      pre += g < grp ? v : 0u;
    }
#pragma unroll 8
    for (int b = grp * RS2_GROUP; b < (int)blockIdx.x; ++b) pre += hist[(size_t)b * 256 + threadIdx.x];
    unsigned inc = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(FULL, inc, o); if (l >= o) inc += x; }
    if (l == 31) sm.scan[w] = inc;
    __syncthreads();
    gbase = inc - hv + pre;
#pragma unroll
    for (int ww = 0; ww < RS2_DOWN_WARPS; ++ww) gbase += ww < w ? sm.scan[ww] : 0u;
  }
  // every block (also the ones past the end of the input) reports that it has read the sums; the last one clears them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&aux->done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < RS2_GROUPS * 256; i += RS2_DOWN_THREADS) (&aux->group[0][0])[i] = 0;
    if (threadIdx.x == 0) aux->done = 0;
  }
  if (!active) return;
  for (int cb = beg; cb < end; cb += RS2_CHUNK) {
    if (cb + RS2_CHUNK <= end)
      rs2_chunk<BITS, true>(sm, keys_in, vals_in, keys_out, vals_out, cb, RS2_CHUNK, end, n, shift, gbase);
    else
      rs2_chunk<BITS, false>(sm, keys_in, vals_in, keys_out, vals_out, cb, end - cb, end, n, shift, gbase);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
struct Rs2Plan {
  int passes, bits, segments;
  Rs2Seg seg;
};
inline Rs2Plan rs2_plan(size_t n_cap, int key_bits) {
  Rs2Plan p;
  key_bits = key_bits < 1 ? 1 : (key_bits > 32 ? 32 : key_bits);
  p.passes = (key_bits + 7) / 8;
  p.bits = (key_bits + p.passes - 1) / p.passes;
  if (p.bits < 4) p.bits = 4;
  long long chunks = ((long long)n_cap + RS2_CHUNK - 1) / RS2_CHUNK;
  if (chunks < 1) chunks = 1;
  p.segments = (int)(chunks < RS2_MAX_SEGMENTS ? chunks : RS2_MAX_SEGMENTS);
  p.seg.q = (int)(chunks / p.segments);
  p.seg.r = (int)(chunks % p.segments);
  return p;
}
constexpr size_t RS2_HIST_WORDS = (size_t)RS2_MAX_SEGMENTS * 256;

template <int BITS>
inline void rs2_launch_down(cudaStream_t st, const Rs2Plan& p, const unsigned* ki, const unsigned* vi, unsigned* ko, unsigned* vo,
                            const int* n_ptr, int shift, const unsigned* hist, Rs2Aux* aux) {
  k_rs2_down<BITS><<<p.segments, RS2_DOWN_THREADS, 0, st>>>(ki, vi, ko, vo, n_ptr, shift, p.seg, hist, aux);
}

// Sorts (keys[0], vals[0]) using (keys[1], vals[1]) as the other half of the ping-pong; returns the index of the buffer
// holding the result.  `hist` holds RS2_HIST_WORDS words; `aux` must be zero on entry (cudaMemset once after allocating it)
// and is zero again when the sort has run.  2 launches per pass.
inline int rs2_sort(cudaStream_t st, unsigned* const keys[2], unsigned* const vals[2], const int* n_ptr, size_t n_cap, int key_bits,
                    unsigned* hist, Rs2Aux* aux, int* launches) {
  const Rs2Plan p = rs2_plan(n_cap, key_bits);
  const unsigned mask = (1u << p.bits) - 1u;
  int cur = 0;
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.bits;
    k_rs2_hist<<<p.segments, RS2_HIST_THREADS, 0, st>>>(keys[cur], n_ptr, shift, mask, p.seg, hist, aux);
    switch (p.bits) {
      case 4: rs2_launch_down<4>(st, p, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_ptr, shift, hist, aux); break;
      case 5: rs2_launch_down<5>(st, p, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_ptr, shift, hist, aux); break;
      case 6: rs2_launch_down<6>(st, p, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_ptr, shift, hist, aux); break;
      case 7: rs2_launch_down<7>(st, p, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_ptr, shift, hist, aux); break;
      default: rs2_launch_down<8>(st, p, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_ptr, shift, hist, aux); break;
    }
    cur ^= 1;
  }
  if (launches) *launches += 2 * p.passes;
  return cur;
}

}  // namespace gm
