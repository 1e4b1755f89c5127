// gm_comm.cuh — the data-path collectives of the multi-GPU legs, written directly over
// NVLink peer memory instead of a library call:
//   * hypothesis-sharded RANSAC: all-reduce(MAX) of the packed (count, id) key of each primitive, carrying the winner's
//     coefficients with it (so a rank only ever generates its own share of the hypotheses);
//   * map slabs: MIN/MAX of the 6-float bounding box (one VoxelGrid lattice for all slabs) and SUM of the 6 scatter sums
//     of getLocalFrame (one frame for the whole map).
// Every rank owns a MAILBOX in its own HBM (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by the peers with
// cudaIpcOpenMemHandle; NVSwitch gives every GPU a direct path to every peer).  A collective is: every rank STORES its
// record into its slot of every peer's mailbox (plain st.global on the mapped peer pointer, then a release store of
// the sequence number), and POLLS ITS OWN mailbox (local HBM/L2 reads) until the records of all ranks carry the
// expected sequence number.  Payloads are tens of bytes, so the cost is one NVLink store latency (~1 us), not bandwidth;
// the send is issued from inside the kernel that produces the value (the last block of the inlier-counting kernel), so
// the transfer overlaps the tail of the compute and no host code, library call or extra copy sits between them.
// Sequence numbers live in device memory (per channel), which keeps the launches free of per-call host arguments.
// Reductions visit the ranks in rank order: sums are bitwise identical on every rank.
#pragma once
#include "gm_device.cuh"

namespace gm {

constexpr int COMM_MAX_RANKS = 16;
constexpr int COMM_SLOTS = 4;      // a rank can be at most one collective ahead of a peer on a channel: 2 would do
constexpr int COMM_CHANNELS = 4;   // 0 plane key, 1 cylinder key, 2 bounding box, 3 sums
constexpr int COMM_WORDS = 31;

struct __align__(16) CommRecord { unsigned long long seq; unsigned long long w[COMM_WORDS]; };  // 256 bytes
struct CommBox {
  CommRecord rec[COMM_CHANNELS][COMM_SLOTS][COMM_MAX_RANKS];
  unsigned long long seq[COMM_CHANNELS];  // collectives completed on each channel (same on every rank between collectives)
};
struct CommDev {
  CommBox* local;
  CommBox* peer[COMM_MAX_RANKS];  // peer[rank] == local
  int rank, world;
};

__device__ __forceinline__ unsigned long long d_comm_next_seq(const CommDev& c, int ch) {
  return *(volatile unsigned long long*)&c.local->seq[ch] + 1ull;
}
// one thread sends `nw` words to peer `r` (all peers in parallel: call with r = lane, r < world)
__device__ __forceinline__ void d_comm_send(const CommDev& c, int ch, unsigned long long seq, int r, const unsigned long long* words, int nw) {
  CommRecord* dst = &c.peer[r]->rec[ch][seq % COMM_SLOTS][c.rank];
  for (int k = 0; k < nw; ++k) ((volatile unsigned long long*)dst->w)[k] = words[k];
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(seq) : "memory");
}
// one thread waits for the record of `sender`; returns its address in the local mailbox
__device__ __forceinline__ const CommRecord* d_comm_wait(const CommDev& c, int ch, unsigned long long seq, int sender, int* err) {
  const CommRecord* src = &c.local->rec[ch][seq % COMM_SLOTS][sender];
  unsigned long long got = 0ull;
  int spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(&src->seq) : "memory");
  } while (got != seq && ++spins < SPIN_BOUND * 4);
  if (got != seq) atomicExch(err, 5);  // a peer never arrived: raise device_error instead of hanging
  return src;
}
__device__ __forceinline__ void d_comm_done(const CommDev& c, int ch, unsigned long long seq) {
  *(volatile unsigned long long*)&c.local->seq[ch] = seq;
  __threadfence();
}

// ---- RANSAC best model -----------------------------------------------------------------------------------------
// record = { key, coefficient words }: plane 4 floats (2 words), cylinder model7 + test12 = 19 floats (10 words)
constexpr int COMM_PLANE_WORDS = 1 + 2;
constexpr int COMM_CYL_WORDS = 1 + 10;

// Called by ONE warp of the block that has just reduced this rank's local best key (the last block of the counting
// kernel): lane r stores (key, coefficients of the local winner) into peer r's mailbox.
__device__ __forceinline__ void d_comm_send_model(const CommDev& c, int kind, unsigned long long key, const float4* __restrict__ plane_coef,
                                                  const float* __restrict__ model7, const float* __restrict__ test12, int H) {
  const int lane = threadIdx.x & 31;
  if (lane >= c.world) return;
  unsigned long long w[COMM_CYL_WORDS];
  float f[20];
  for (int k = 0; k < 20; ++k) f[k] = 0.f;
  const int count = (int)(unsigned)(key >> 32) - 1;
  const int id = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
  if (count >= 0 && id >= 0 && id < H) {
    if (kind == 0) { const float4 q = plane_coef[id]; f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w; }
    else {
      for (int k = 0; k < 7; ++k) f[k] = model7[(size_t)id * 7 + k];
      for (int k = 0; k < 12; ++k) f[7 + k] = test12[(size_t)id * 12 + k];
    }
  }
  w[0] = key;
  for (int k = 0; k < 10; ++k) w[1 + k] = (unsigned long long)__float_as_uint(f[2 * k]) | ((unsigned long long)__float_as_uint(f[2 * k + 1]) << 32);
  d_comm_send(c, kind, d_comm_next_seq(c, kind), lane, w, kind == 0 ? COMM_PLANE_WORDS : COMM_CYL_WORDS);
}

// One warp: waits for the records of all ranks, takes the maximum key (ids are globally unique, so there are no ties
// between ranks), installs the winner's coefficients in this rank's hypothesis tables and the key in *key_out.
__global__ void k_comm_recv_model(CommDev c, int kind, int H, unsigned long long* key_out, float4* plane_coef, float* model7, float* test12,
                                  int* counts, int* err) {
  const int lane = threadIdx.x & 31;
  const unsigned long long seq = d_comm_next_seq(c, kind);
  unsigned long long key = 0ull;
  const CommRecord* rec = nullptr;
  if (lane < c.world) { rec = d_comm_wait(c, kind, seq, lane, err); key = ((volatile const unsigned long long*)rec->w)[0]; }
  unsigned long long best = key;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(FULL, best, o); best = t > best ? t : best; }
  const unsigned who = __ballot_sync(FULL, lane < c.world && key == best);
  if (lane == __ffs(who) - 1) {
    const int count = (int)(unsigned)(best >> 32) - 1;
    const int id = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
    if (count >= 0 && id >= 0 && id < H) {
      float f[20];
      for (int k = 0; k < 10; ++k) {
        const unsigned long long v = ((volatile const unsigned long long*)rec->w)[1 + k];
        f[2 * k] = __uint_as_float((unsigned)v); f[2 * k + 1] = __uint_as_float((unsigned)(v >> 32));
      }
      if (kind == 0) plane_coef[id] = make_float4(f[0], f[1], f[2], f[3]);
      else {
        for (int k = 0; k < 7; ++k) model7[(size_t)id * 7 + k] = f[k];
        for (int k = 0; k < 12; ++k) test12[(size_t)id * 12 + k] = f[7 + k];
      }
      counts[id] = count;
    }
    *key_out = best;
  }
  __syncwarp();
  if (lane == 0) d_comm_done(c, kind, seq);
}

// ---- bounding box (map slabs: one VoxelGrid lattice for all ranks) -------------------------------------------------
// VoxState keeps the box as order-preserving ints; MIN of the mins, MAX of the maxes, in place.
__global__ void k_comm_bbox(CommDev c, int* bbox_min3, int* bbox_max3, int* err) {
  const int lane = threadIdx.x & 31;
  const int ch = 2;
  const unsigned long long seq = d_comm_next_seq(c, ch);
  unsigned long long w[3];
  for (int a = 0; a < 3; ++a) w[a] = (unsigned long long)(unsigned)bbox_min3[a] | ((unsigned long long)(unsigned)bbox_max3[a] << 32);
  if (lane < c.world) d_comm_send(c, ch, seq, lane, w, 3);
  int mn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  if (lane < c.world) {
    const CommRecord* rec = d_comm_wait(c, ch, seq, lane, err);
    for (int a = 0; a < 3; ++a) {
      const unsigned long long v = ((volatile const unsigned long long*)rec->w)[a];
      mn[a] = (int)(unsigned)v; mx[a] = (int)(unsigned)(v >> 32);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn[a] = min(mn[a], __shfl_xor_sync(FULL, mn[a], o)); mx[a] = max(mx[a], __shfl_xor_sync(FULL, mx[a], o)); }
  }
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) { bbox_min3[a] = mn[a]; bbox_max3[a] = mx[a]; }
    d_comm_done(c, ch, seq);
  }
}

// ---- sums (whole-map frame): nw <= 31 doubles, summed in rank order -> bitwise identical on every rank --------------
__global__ void k_comm_sum(CommDev c, double* data, int nw, int* err) {
  const int lane = threadIdx.x & 31;
  const int ch = 3;
  const unsigned long long seq = d_comm_next_seq(c, ch);
  if (lane < c.world) {
    unsigned long long w[COMM_WORDS];
    for (int k = 0; k < nw; ++k) w[k] = (unsigned long long)__double_as_longlong(data[k]);
    d_comm_send(c, ch, seq, lane, w, nw);
    d_comm_wait(c, ch, seq, lane, err);
  }
  __syncwarp();
  if (lane < nw) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += __longlong_as_double((long long)((volatile const unsigned long long*)c.local->rec[ch][seq % COMM_SLOTS][r].w)[lane]);
    data[lane] = s;
  }
  __syncwarp();
  if (lane == 0) d_comm_done(c, ch, seq);
}

}  // namespace gm
