// gm_ransac.cuh — builder-defined stage a8: batched RANSAC plane / cylinder hypotheses from
// injected sample indices, inlier counting, argmax, least-squares refit, labels.
// The reference only stubs this stage (src/tunnel_processing.cpp:149-154); the definitions follow
// pcl::SampleConsensusModelPlane / Cylinder / RandomSampleConsensus.
// Compiled with -fmad=false: hypothesis generation is the plain unfused expression order that the
// oracle uses (bit-identical coefficients); the inlier tests are explicit fmaf chains.
#pragma once
#include "gm_stages.cuh"
#include "gm_comm.cuh"

namespace gm {

// ---- hypothesis generation ------------------------------------------------------------------
// plane: samples H x 3 -> coef float4 {a,b,c,d}, valid
__global__ void k_plane_hypotheses(const float4* __restrict__ pts, const int* __restrict__ n_ptr,
                                   const int* __restrict__ samples, int H, float4* __restrict__ coef,
                                   int* __restrict__ valid, int* __restrict__ counts, int h_begin, int h_end, int gen_all) {
  const int n = *n_ptr;
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
    const bool mine = h >= h_begin && h < h_end;
    // gen_all = 0 (peer-memory sharded round): another rank's hypothesis is not generated, its samples are not even here
    if (!mine && !gen_all) { counts[h] = -1; valid[h] = 0; continue; }
    int i0 = samples[h * 3], i1 = samples[h * 3 + 1], i2 = samples[h * 3 + 2];
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    int ok = 0;
    if (i0 >= 0 && i1 >= 0 && i2 >= 0 && i0 < n && i1 < n && i2 < n && i0 != i1 && i0 != i2 && i1 != i2) {
      float4 p0 = pts[i0], p1 = pts[i1], p2 = pts[i2];
      float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
      float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
      float nx = ay * bz - az * by;
      float ny = az * bx - ax * bz;
      float nz = ax * by - ay * bx;
      float len2 = (nx * nx + ny * ny) + nz * nz;
      if (len2 > 0.0f && isfinite(len2)) {
        float len = sqrtf(len2);
        nx = nx / len; ny = ny / len; nz = nz / len;
        float d = -((nx * p0.x + ny * p0.y) + nz * p0.z);
        if (isfinite(d)) { c = make_float4(nx, ny, nz, d); ok = 1; }
      }
    }
    coef[h] = c;
    valid[h] = ok;
    counts[h] = (mine && ok) ? 0 : -1;  // -1: degenerate or not this rank's id
  }
}

// canonical orthonormal basis perpendicular to unit dir (matches oracle perp_basis)
__device__ __forceinline__ void d_perp_basis(const float dir[3], float u[3], float w[3]) {
  float ax = fabsf(dir[0]), ay = fabsf(dir[1]), az = fabsf(dir[2]);
  int k = 0; float m = ax;
  if (ay < m) { k = 1; m = ay; }
  if (az < m) { k = 2; m = az; }
  float c[3];
  if (k == 0) { c[0] = 0.0f; c[1] = dir[2]; c[2] = -dir[1]; }
  else if (k == 1) { c[0] = -dir[2]; c[1] = 0.0f; c[2] = dir[0]; }
  else { c[0] = dir[1]; c[1] = -dir[0]; c[2] = 0.0f; }
  float l = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
  u[0] = c[0] / l; u[1] = c[1] / l; u[2] = c[2] / l;
  w[0] = dir[1] * u[2] - dir[2] * u[1];
  w[1] = dir[2] * u[0] - dir[0] * u[2];
  w[2] = dir[0] * u[1] - dir[1] * u[0];
}

// model7 {q,dir,r} -> test12 {u,du, w,dw, mid,half,0,0}; returns false if any value is non-finite
__device__ __forceinline__ bool d_cyl_test_params(const float* m7, float tau, float* t12) {
  float u[3], w[3];
  d_perp_basis(m7 + 3, u, w);
  float du = -((u[0] * m7[0] + u[1] * m7[1]) + u[2] * m7[2]);
  float dw = -((w[0] * m7[0] + w[1] * m7[1]) + w[2] * m7[2]);
  float r = m7[6];
  float hi = r + tau, lo = r - tau;
  float hi2 = hi * hi, lo2 = lo * lo;
  float mid, half;
  if (lo >= 0.0f) { mid = 0.5f * (hi2 + lo2); half = 0.5f * (hi2 - lo2); }
  else { mid = 0.0f; half = hi2; }
  t12[0] = u[0]; t12[1] = u[1]; t12[2] = u[2]; t12[3] = du;
  t12[4] = w[0]; t12[5] = w[1]; t12[6] = w[2]; t12[7] = dw;
  t12[8] = mid; t12[9] = half; t12[10] = 0.0f; t12[11] = 0.0f;
  bool fin = true;
  for (int k = 0; k < 10; ++k) fin = fin && isfinite(t12[k]);
  return fin;
}

// cylinder: samples H x 2 (+ normals) -> model7, test12, valid
__global__ void k_cyl_hypotheses(const float4* __restrict__ pts, const float4* __restrict__ normals,
                                 const int* __restrict__ n_ptr, const int* __restrict__ samples, int H,
                                 float rmin, float rmax, float tau, float* __restrict__ model7,
                                 float* __restrict__ test12, int* __restrict__ valid, int* __restrict__ counts,
                                 int h_begin, int h_end, int gen_all) {
  const int n = *n_ptr;
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
    const bool mine = h >= h_begin && h < h_end;
    if (!mine && !gen_all) { counts[h] = -1; valid[h] = 0; continue; }  // another rank's hypothesis
    float m[7] = {0, 0, 0, 0, 0, 0, 0};
    float t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int ok = 0;
    int i1 = samples[h * 2], i2 = samples[h * 2 + 1];
    if (i1 >= 0 && i2 >= 0 && i1 < n && i2 < n && i1 != i2) {
      float4 P1 = pts[i1], P2 = pts[i2];
      float4 N1 = normals[2 * (size_t)i1], N2 = normals[2 * (size_t)i2];
      float p1[3] = {P1.x, P1.y, P1.z}, p2[3] = {P2.x, P2.y, P2.z};
      float n1[3] = {N1.x, N1.y, N1.z}, n2[3] = {N2.x, N2.y, N2.z};
      const float eps = 1.1920928955078125e-07f;
      bool coincident = fabsf(p1[0] - p2[0]) <= eps && fabsf(p1[1] - p2[1]) <= eps && fabsf(p1[2] - p2[2]) <= eps;
      if (!coincident) {
        float w[3] = {(n1[0] + p1[0]) - p2[0], (n1[1] + p1[1]) - p2[1], (n1[2] + p1[2]) - p2[2]};
        float a = (n1[0] * n1[0] + n1[1] * n1[1]) + n1[2] * n1[2];
        float b = (n1[0] * n2[0] + n1[1] * n2[1]) + n1[2] * n2[2];
        float c = (n2[0] * n2[0] + n2[1] * n2[1]) + n2[2] * n2[2];
        float d = (n1[0] * w[0] + n1[1] * w[1]) + n1[2] * w[2];
        float e = (n2[0] * w[0] + n2[1] * w[1]) + n2[2] * w[2];
        float den = a * c - b * b;
        float sc, tc;
        if (den < 1e-8f) { sc = 0.0f; tc = (b > c) ? (d / b) : (e / c); }
        else { sc = (b * e - c * d) / den; tc = (a * e - b * d) / den; }
        float q[3], dir[3];
        for (int k = 0; k < 3; ++k) q[k] = (p1[k] + n1[k]) + sc * n1[k];
        for (int k = 0; k < 3; ++k) dir[k] = (p2[k] + tc * n2[k]) - q[k];
        float dl2 = (dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2];
        if (dl2 > 0.0f && isfinite(dl2)) {
          float dl = sqrtf(dl2);
          dir[0] = dir[0] / dl; dir[1] = dir[1] / dl; dir[2] = dir[2] / dl;
          float v[3] = {q[0] - p1[0], q[1] - p1[1], q[2] - p1[2]};
          float cx = dir[1] * v[2] - dir[2] * v[1];
          float cy = dir[2] * v[0] - dir[0] * v[2];
          float cz = dir[0] * v[1] - dir[1] * v[0];
          float c2 = (cx * cx + cy * cy) + cz * cz;
          float d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2];
          float r = sqrtf(c2 / d2);
          if (isfinite(r) && !(r > rmax) && !(r < rmin)) {
            m[0] = q[0]; m[1] = q[1]; m[2] = q[2]; m[3] = dir[0]; m[4] = dir[1]; m[5] = dir[2]; m[6] = r;
            if (d_cyl_test_params(m, tau, t)) ok = 1;
            else { for (int k = 0; k < 12; ++k) t[k] = 0.0f; for (int k = 0; k < 7; ++k) m[k] = 0.0f; }
          }
        }
      }
    }
    for (int k = 0; k < 7; ++k) model7[(size_t)h * 7 + k] = m[k];
    for (int k = 0; k < 12; ++k) test12[(size_t)h * 12 + k] = t[k];
    valid[h] = ok;
    counts[h] = (mine && ok) ? 0 : -1;
  }
}

// ---- canonical inlier tests -----------------------------------------------------------------
__device__ __forceinline__ bool d_plane_inlier(const float4 c, const float4 p, float tau) {
  float d = fmaf(c.x, p.x, fmaf(c.y, p.y, fmaf(c.z, p.z, c.w)));
  return fabsf(d) < tau;
}
struct CylTest { float4 u, w; float mid, half; };
__device__ __forceinline__ bool d_cyl_inlier(const CylTest& t, const float4 p) {
  float A = fmaf(t.u.x, p.x, fmaf(t.u.y, p.y, fmaf(t.u.z, p.z, t.u.w)));
  float B = fmaf(t.w.x, p.x, fmaf(t.w.y, p.y, fmaf(t.w.z, p.z, t.w.w)));
  float v = fmaf(A, A, fmaf(B, B, -t.mid));
  return fabsf(v) < t.half;
}
__device__ __forceinline__ CylTest d_load_cyl_test(const float* __restrict__ t12) {
  CylTest t;
  t.u = make_float4(t12[0], t12[1], t12[2], t12[3]);
  t.w = make_float4(t12[4], t12[5], t12[6], t12[7]);
  t.mid = t12[8]; t.half = t12[9];
  return t;
}

// ---- inlier counting ------------------------------------------------------------------------
// Each thread owns K hypotheses in registers; the block streams its slice of the cloud through
// shared memory (transposed to SoA while staging) and every thread reads 4 points per LDS.128
// triple as warp-wide broadcasts.  Two points are evaluated per instruction with the packed
// fma.rn.f32x2 (SASS FFMA2, coefficient broadcast as the scalar operand); each half is an IEEE
// fma, so the result is bit-identical to the scalar chain of the oracle.  The |d| < tau outcome
// is accumulated with setp + predicated add (2 non-FMA issue slots per test; measured best of the
// idioms in tools/microbench/count_variants.cu, profiles/r01_microbench_count.md).
// Counts stay in registers for the whole slice; one integer atomicAdd per (thread, hypothesis).
// grid = (point slices, hypothesis groups).
// cnt += (v < thr)   [v is already |.|; NaN compares false]
__device__ __forceinline__ void d_count_lt(float v, float thr, int& cnt) {
  asm("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt) : "f"(v), "f"(thr));
}

// best key = max over h of ((count+1) << 32) | (0xFFFFFFFF - h): max count, ties -> lowest id
// (RandomSampleConsensus keeps a model only on a strictly larger count)
__device__ __forceinline__ unsigned long long d_key_of(int count, int h) {
  return ((unsigned long long)(unsigned)(count + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
}
// block-wide argmax over counts[h_begin..h_end) (read through L2: other blocks produced them with atomics)
template <int BLOCK>
__device__ __forceinline__ void d_block_argmax(const int* counts, int h_begin, int h_end, unsigned long long* key_out) {
  __shared__ unsigned long long s_best[BLOCK / 32];
  unsigned long long best = 0ull;
  for (int h = h_begin + threadIdx.x; h < h_end; h += BLOCK) {
    unsigned long long k = d_key_of(__ldcg(counts + h), h);
    best = (k > best) ? k : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(FULL, best, o);
    best = (t > best) ? t : best;
  }
  if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK / 32; ++w) best = (s_best[w] > best) ? s_best[w] : best;
    *key_out = best;
    s_best[0] = best;
  }
  __syncthreads();
}
// ... and, in a hypothesis-sharded round, straight into the mailboxes of all peers (gm_comm.cuh): the block that
// reduces the local best model also ships it, with the winner's coefficients
template <int BLOCK>
__device__ __forceinline__ void d_block_argmax_send(const int* counts, int h_begin, int h_end, unsigned long long* key_out, const CommDev* comm, int kind,
                                                    const float4* __restrict__ plane_coef, const float* __restrict__ model7,
                                                    const float* __restrict__ test12, int H) {
  d_block_argmax<BLOCK>(counts, h_begin, h_end, key_out);
  if (comm != nullptr && threadIdx.x < 32) {
    const unsigned long long key = *(volatile unsigned long long*)key_out;
    d_comm_send_model(*comm, kind, key, plane_coef, model7, test12, H);
  }
}

constexpr int RC_BLOCK = 64;
constexpr int RC_TILE = 512;   // points staged per iteration (6 KB of shared memory)
constexpr int RC_KP = 8;       // plane hypotheses per thread
constexpr int RC_KC = 4;       // cylinder hypotheses per thread

// Stage points [base, base+m) into SoA shared arrays; entries m..m4 are NaN (never inliers).
__device__ __forceinline__ void d_stage_tile(const float4* __restrict__ pts, int base, int m, int m4,
                                             float* sx, float* sy, float* sz) {
  for (int i = threadIdx.x; i < m4; i += RC_BLOCK) {
    float4 p = (i < m) ? pts[base + i] : make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f);
    sx[i] = p.x; sy[i] = p.y; sz[i] = p.z;
  }
}

template <int K>
__global__ void __launch_bounds__(RC_BLOCK)
k_count_plane(const float4* __restrict__ pts, const int* __restrict__ n_ptr, const float4* __restrict__ coef,
              const int* __restrict__ valid, int h_begin, int h_end, float tau, int* __restrict__ counts,
              int H, unsigned long long* key_out, unsigned* ticket, const CommDev* comm) {
  __shared__ __align__(16) float sx[RC_TILE], sy[RC_TILE], sz[RC_TILE];
  const int n = *n_ptr;
  const int per = (((n + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
  const int p_begin = blockIdx.x * per, p_end = min(n, p_begin + per);
  const int hbase = h_begin + blockIdx.y * (RC_BLOCK * K) + threadIdx.x;
  float4 c[K];
  int cnt[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int h = hbase + k * RC_BLOCK;
    // degenerate hypotheses (all-zero coefficients) must never match: d = +inf
    c[k] = (h < h_end && valid[h]) ? coef[h] : make_float4(0.f, 0.f, 0.f, CUDART_INF_F);
    cnt[k] = 0;
  }
  for (int base = p_begin; base < p_end; base += RC_TILE) {
    const int m = min(RC_TILE, p_end - base), m4 = (m + 3) & ~3;
    __syncthreads();
    d_stage_tile(pts, base, m, m4, sx, sy, sz);
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < m4; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(&sx[i]);
      const float4 y = *reinterpret_cast<const float4*>(&sy[i]);
      const float4 z = *reinterpret_cast<const float4*>(&sz[i]);
      const u64 x01 = d_pack2(x.x, x.y), x23 = d_pack2(x.z, x.w), y01 = d_pack2(y.x, y.y), y23 = d_pack2(y.z, y.w);
      const u64 z01 = d_pack2(z.x, z.y), z23 = d_pack2(z.z, z.w);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const u64 A = d_pack2(c[k].x, c[k].x), B = d_pack2(c[k].y, c[k].y), Cz = d_pack2(c[k].z, c[k].z), D = d_pack2(c[k].w, c[k].w);
        const u64 t0 = d_fma2(A, x01, d_fma2(B, y01, d_fma2(Cz, z01, D)));
        const u64 t1 = d_fma2(A, x23, d_fma2(B, y23, d_fma2(Cz, z23, D)));
        float d0, d1, d2, d3;
        d_unpack2(t0, d0, d1); d_unpack2(t1, d2, d3);
        d_count_lt(fabsf(d0), tau, cnt[k]); d_count_lt(fabsf(d1), tau, cnt[k]);
        d_count_lt(fabsf(d2), tau, cnt[k]); d_count_lt(fabsf(d3), tau, cnt[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int h = hbase + k * RC_BLOCK;
    if (h < h_end && cnt[k]) atomicAdd(&counts[h], cnt[k]);
  }
  // the last block to finish reduces the local best model (saves a launch)
  if (d_last_block(ticket, gridDim.x * gridDim.y))
    d_block_argmax_send<RC_BLOCK>(counts, h_begin, h_end, key_out, comm, 0, coef, nullptr, nullptr, H);
}

template <int K>
__global__ void __launch_bounds__(RC_BLOCK)
k_count_cyl(const float4* __restrict__ pts, const int* __restrict__ n_ptr, const float* __restrict__ test12,
            int h_begin, int h_end, int* __restrict__ counts, int H, unsigned long long* key_out, unsigned* ticket,
            const CommDev* comm, const float* __restrict__ model7) {
  __shared__ __align__(16) float sx[RC_TILE], sy[RC_TILE], sz[RC_TILE];
  const int n = *n_ptr;
  const int per = (((n + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
  const int p_begin = blockIdx.x * per, p_end = min(n, p_begin + per);
  const int hbase = h_begin + blockIdx.y * (RC_BLOCK * K) + threadIdx.x;
  float4 u[K], w[K];
  float negmid[K], half[K];
  int cnt[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int h = hbase + k * RC_BLOCK;
    // invalid hypotheses have an all-zero test (half = 0): |v| < 0 is never true
    const float* t = test12 + (size_t)min(h, h_end - 1) * 12;
    const bool ok = h < h_end;
    u[k] = ok ? make_float4(t[0], t[1], t[2], t[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    w[k] = ok ? make_float4(t[4], t[5], t[6], t[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
    negmid[k] = ok ? -t[8] : 0.f;
    half[k] = ok ? t[9] : 0.f;
    cnt[k] = 0;
  }
  for (int base = p_begin; base < p_end; base += RC_TILE) {
    const int m = min(RC_TILE, p_end - base), m4 = (m + 3) & ~3;
    __syncthreads();
    d_stage_tile(pts, base, m, m4, sx, sy, sz);
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < m4; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(&sx[i]);
      const float4 y = *reinterpret_cast<const float4*>(&sy[i]);
      const float4 z = *reinterpret_cast<const float4*>(&sz[i]);
      const u64 x01 = d_pack2(x.x, x.y), x23 = d_pack2(x.z, x.w), y01 = d_pack2(y.x, y.y), y23 = d_pack2(y.z, y.w);
      const u64 z01 = d_pack2(z.x, z.y), z23 = d_pack2(z.z, z.w);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const u64 UX = d_pack2(u[k].x, u[k].x), UY = d_pack2(u[k].y, u[k].y), UZ = d_pack2(u[k].z, u[k].z), UD = d_pack2(u[k].w, u[k].w);
        const u64 WX = d_pack2(w[k].x, w[k].x), WY = d_pack2(w[k].y, w[k].y), WZ = d_pack2(w[k].z, w[k].z), WD = d_pack2(w[k].w, w[k].w);
        const u64 NM = d_pack2(negmid[k], negmid[k]);
        const u64 A0 = d_fma2(UX, x01, d_fma2(UY, y01, d_fma2(UZ, z01, UD)));
        const u64 B0 = d_fma2(WX, x01, d_fma2(WY, y01, d_fma2(WZ, z01, WD)));
        const u64 A1 = d_fma2(UX, x23, d_fma2(UY, y23, d_fma2(UZ, z23, UD)));
        const u64 B1 = d_fma2(WX, x23, d_fma2(WY, y23, d_fma2(WZ, z23, WD)));
        const u64 v0 = d_fma2(A0, A0, d_fma2(B0, B0, NM));
        const u64 v1 = d_fma2(A1, A1, d_fma2(B1, B1, NM));
        float e0, e1, e2, e3;
        d_unpack2(v0, e0, e1); d_unpack2(v1, e2, e3);
        d_count_lt(fabsf(e0), half[k], cnt[k]); d_count_lt(fabsf(e1), half[k], cnt[k]);
        d_count_lt(fabsf(e2), half[k], cnt[k]); d_count_lt(fabsf(e3), half[k], cnt[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int h = hbase + k * RC_BLOCK;
    if (h < h_end && cnt[k]) atomicAdd(&counts[h], cnt[k]);
  }
  // the last block to finish reduces the local best model (saves a launch)
  if (d_last_block(ticket, gridDim.x * gridDim.y))
    d_block_argmax_send<RC_BLOCK>(counts, h_begin, h_end, key_out, comm, 1, nullptr, model7, test12, H);
}

// ---- tile-culled inlier counting ---------------------------------------------------------------
// Same counts as the brute-force kernels above, bit for bit, with most of the work skipped.
// Counting is order independent, so it runs over the CELL-SORTED cloud (k_normals writes a copy in
// which the points dropped by the NaN compaction are blanked to NaN): consecutive points are
// spatial neighbours, and a 32-point LEAF (one warp of k_normals, which also records the leaf's
// bounding box) is a small box.  A hypothesis whose inlier band provably misses that box cannot
// have an inlier in it:
//   plane     |n.c + d| - sum_k |n_k| h_k  >  tau + margin
//   cylinder  rho(c) - e > (r + tau) + margin   or   rho(c) + e < (r - tau) - margin,
//             rho(c) = |(A(c), B(c))| distance of the box centre to the axis, e = |(|u|.h, |w|.h)|
// (c, h = centre / half extent of the box).  `margin` covers the float rounding of the exact test
// and of the bound itself (1e-5 relative to the magnitudes involved + 1e-3 tau), so the cull is
// conservative: every (hypothesis, point) pair that the exact fma test would accept is still
// evaluated by the exact fma test.  Survival measured on the C1 scan: ~8 % (plane), ~12 % (cylinder).
//
// One block per SUPERTILE of 16 leaves (512 points, staged once in shared memory as SoA):
//   L1  every hypothesis of the chunk against the supertile box      -> s_l1
//   L2  the L1 survivors against the 16 leaf boxes                   -> one list per leaf, s_l2
//   count  warps take (leaf, 32 hypotheses) rounds: a lane owns one hypothesis and streams the
//          leaf's 32 points as broadcast LDS.128, two points per FFMA2 exactly as above.
constexpr int TC_BLOCK = 128;
constexpr int TC_LEAF = 32;
constexpr int TC_LEAVES = 16;
constexpr int TC_SUPER = TC_LEAF * TC_LEAVES;
constexpr int TC_HC = 512;  // hypotheses per chunk (list entries are u16)

struct TileBox { float cx, cy, cz, hx, hy, hz; };

// Cull tests.  The rounding margin is evaluated ONCE per (hypothesis, supertile) from magnitudes
// that bound every point of the supertile box (centre + half extent), so it is valid for every leaf
// inside it and the per-leaf test stays a handful of FMAs.
struct PlaneCull { float4 c; float ax, ay, az, thr; };  // thr = tau + margin
__device__ __forceinline__ PlaneCull d_plane_cull_setup(const float4 c, const TileBox& sb, float tau) {
  PlaneCull p;
  p.c = c; p.ax = fabsf(c.x); p.ay = fabsf(c.y); p.az = fabsf(c.z);
  const float mag = fmaf(p.ax, fabsf(sb.cx) + sb.hx, fmaf(p.ay, fabsf(sb.cy) + sb.hy, fmaf(p.az, fabsf(sb.cz) + sb.hz, fabsf(c.w))));
  p.thr = tau + fmaf(2e-5f, mag, 1e-3f * tau);
  return p;
}
__device__ __forceinline__ bool d_plane_may_hit(const PlaneCull& p, const TileBox& b) {
  const float dc = fmaf(p.c.x, b.cx, fmaf(p.c.y, b.cy, fmaf(p.c.z, b.cz, p.c.w)));
  const float ext = fmaf(p.ax, b.hx, fmaf(p.ay, b.hy, p.az * b.hz));
  return !(fabsf(dc) - ext > p.thr);  // NaN anywhere -> "may hit"
}
// cylinder: band [lo, hi] of distances to the axis; e = |u|.h + |w|.h bounds the change of that
// distance across the box (L1 bound of the exact sqrt(eA^2+eB^2): no square root per test)
struct CylCull { float4 u, w; float aux, auy, auz, awx, awy, awz, hi_m, lo_m; };
__device__ __forceinline__ CylCull d_cyl_cull_setup(const CylTest& t, const TileBox& sb, float tau) {
  CylCull c;
  c.u = t.u; c.w = t.w;
  c.aux = fabsf(t.u.x); c.auy = fabsf(t.u.y); c.auz = fabsf(t.u.z);
  c.awx = fabsf(t.w.x); c.awy = fabsf(t.w.y); c.awz = fabsf(t.w.z);
  const float hi = sqrtf(t.mid + t.half);
  const float lo2 = t.mid - t.half;
  const float lo = (lo2 > 0.0f) ? sqrtf(lo2) : -CUDART_INF_F;
  const float X = fabsf(sb.cx) + sb.hx, Y = fabsf(sb.cy) + sb.hy, Z = fabsf(sb.cz) + sb.hz;
  const float magA = fmaf(c.aux, X, fmaf(c.auy, Y, fmaf(c.auz, Z, fabsf(t.u.w))));
  const float magB = fmaf(c.awx, X, fmaf(c.awy, Y, fmaf(c.awz, Z, fabsf(t.w.w))));
  const float margin = fmaf(2e-5f, (magA + magB) + hi, 1e-3f * tau);
  c.hi_m = hi + margin;
  c.lo_m = lo - margin;
  return c;
}
__device__ __forceinline__ bool d_cyl_may_hit(const CylCull& c, const TileBox& b) {
  const float A = fmaf(c.u.x, b.cx, fmaf(c.u.y, b.cy, fmaf(c.u.z, b.cz, c.u.w)));
  const float B = fmaf(c.w.x, b.cx, fmaf(c.w.y, b.cy, fmaf(c.w.z, b.cz, c.w.w)));
  const float e = fmaf(c.aux + c.awx, b.hx, fmaf(c.auy + c.awy, b.hy, (c.auz + c.awz) * b.hz));
  const float rho2 = fmaf(A, A, B * B);
  const float far = c.hi_m + e, near = c.lo_m - e;
  const bool outside = (rho2 > far * far) || (near > 0.0f && rho2 < near * near);
  return !outside;
}
__device__ __forceinline__ CylTest d_load_cyl_test4(const float* __restrict__ t12) {  // 48-byte records, 16-byte aligned
  const float4* q = reinterpret_cast<const float4*>(t12);
  const float4 m = q[2];
  CylTest t;
  t.u = q[0]; t.w = q[1]; t.mid = m.x; t.half = m.y;
  return t;
}

template <int KIND>
__global__ void __launch_bounds__(TC_BLOCK)
k_count_tiles(const float4* __restrict__ sv, const float4* __restrict__ leaf_bounds, const int* __restrict__ n_ptr,
              const float4* __restrict__ plane_coef, const float* __restrict__ test12, const int* __restrict__ valid,
              int h_begin, int h_end, float tau, int* __restrict__ counts, int H, unsigned long long* key_out, unsigned* ticket,
              const CommDev* comm, const float* __restrict__ model7) {
  __shared__ __align__(16) float sx[TC_SUPER], sy[TC_SUPER], sz[TC_SUPER];
  __shared__ TileBox s_leaf[TC_LEAVES];
  __shared__ TileBox s_super;
  __shared__ unsigned short s_l1[TC_HC];
  __shared__ unsigned short s_l2[TC_LEAVES][TC_HC];
  __shared__ int s_n1, s_n2[TC_LEAVES], s_rounds[TC_LEAVES + 1];
  __shared__ unsigned char s_round_leaf[TC_LEAVES * (TC_HC / 32)];
  __shared__ int s_cnt[TC_HC];  // per-hypothesis counts of this block: ONE coalesced global add per chunk
  const int n = *n_ptr;
  const int base = blockIdx.x * TC_SUPER;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool live = base < n;  // block-uniform
  if (live) {
    const float qnan = CUDART_NAN_F;
    for (int i = threadIdx.x; i < TC_SUPER; i += TC_BLOCK) {
      const float4 p = (base + i < n) ? sv[base + i] : make_float4(qnan, qnan, qnan, 0.f);
      sx[i] = p.x; sy[i] = p.y; sz[i] = p.z;
    }
    if (warp == 0) {
      const int nleaf = (n + TC_LEAF - 1) / TC_LEAF, leaf = (base >> 5) + lane;
      float4 lo = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f), hi = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, 0.f);
      if (lane < TC_LEAVES && leaf < nleaf) { lo = leaf_bounds[2 * (size_t)leaf]; hi = leaf_bounds[2 * (size_t)leaf + 1]; }
      if (lane < TC_LEAVES) {
        TileBox b;
        const bool empty = !(lo.x <= hi.x);
        b.cx = empty ? 0.f : 0.5f * (lo.x + hi.x); b.cy = empty ? 0.f : 0.5f * (lo.y + hi.y); b.cz = empty ? 0.f : 0.5f * (lo.z + hi.z);
        b.hx = empty ? -1.f : 0.5f * (hi.x - lo.x); b.hy = empty ? -1.f : 0.5f * (hi.y - lo.y); b.hz = empty ? -1.f : 0.5f * (hi.z - lo.z);
        s_leaf[lane] = b;
      }
      const float slx = warp_min(lo.x), sly = warp_min(lo.y), slz = warp_min(lo.z);
      const float shx = warp_max(hi.x), shy = warp_max(hi.y), shz = warp_max(hi.z);
      if (lane == 0) {
        TileBox b;
        const bool empty = !(slx <= shx);
        b.cx = empty ? 0.f : 0.5f * (slx + shx); b.cy = empty ? 0.f : 0.5f * (sly + shy); b.cz = empty ? 0.f : 0.5f * (slz + shz);
        b.hx = empty ? -1.f : 0.5f * (shx - slx); b.hy = empty ? -1.f : 0.5f * (shy - sly); b.hz = empty ? -1.f : 0.5f * (shz - slz);
        s_super = b;
      }
    }
  }
  __syncthreads();
  live = live && s_super.hx >= 0.0f;
  if (live) {
    const TileBox sb = s_super;
    for (int chunk = h_begin; chunk < h_end; chunk += TC_HC) {
      const int hc = min(TC_HC, h_end - chunk);
      if (threadIdx.x == 0) s_n1 = 0;
      if (threadIdx.x < TC_LEAVES) s_n2[threadIdx.x] = 0;
      for (int k = threadIdx.x; k < hc; k += TC_BLOCK) s_cnt[k] = 0;
      __syncthreads();
      // L1: chunk hypotheses against the supertile box (warp-aggregated append)
      for (int k0 = 0; k0 < hc; k0 += TC_BLOCK) {
        const int k = k0 + threadIdx.x, h = chunk + k;
        bool keep = false;
        if (k < hc && valid[h]) {
          if (KIND == 0) keep = d_plane_may_hit(d_plane_cull_setup(plane_coef[h], sb, tau), sb);
          else keep = d_cyl_may_hit(d_cyl_cull_setup(d_load_cyl_test4(test12 + (size_t)h * 12), sb, tau), sb);
        }
        const unsigned m = __ballot_sync(FULL, keep);
        int pos = 0;
        if (lane == 0 && m) pos = atomicAdd(&s_n1, __popc(m));
        pos = __shfl_sync(FULL, pos, 0);
        if (keep) s_l1[pos + __popc(m & lanemask_lt())] = (unsigned short)k;
      }
      __syncthreads();
      const int n1 = s_n1;
      // L2: survivors against the leaf boxes
      for (int k = threadIdx.x; k < n1; k += TC_BLOCK) {
        const unsigned short hl = s_l1[k];
        const int h = chunk + hl;
        PlaneCull pc; CylCull cc;
        if (KIND == 0) pc = d_plane_cull_setup(plane_coef[h], sb, tau);
        else cc = d_cyl_cull_setup(d_load_cyl_test4(test12 + (size_t)h * 12), sb, tau);
#pragma unroll 4
        for (int lf = 0; lf < TC_LEAVES; ++lf) {
          const TileBox b = s_leaf[lf];
          if (b.hx < 0.0f) continue;
          const bool keep = (KIND == 0) ? d_plane_may_hit(pc, b) : d_cyl_may_hit(cc, b);
          if (keep) s_l2[lf][atomicAdd(&s_n2[lf], 1)] = hl;
        }
      }
      __syncthreads();
      if (warp == 0) {  // rounds per leaf -> exclusive scan -> round -> leaf table
        const int nr = (lane < TC_LEAVES) ? ((s_n2[lane] + 31) >> 5) : 0;
        int inc = nr;
#pragma unroll
        for (int o = 1; o < TC_LEAVES; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
        if (lane < TC_LEAVES) {
          s_rounds[lane] = inc - nr;
          for (int r = inc - nr; r < inc; ++r) s_round_leaf[r] = (unsigned char)lane;
        }
        if (lane == TC_LEAVES - 1) s_rounds[TC_LEAVES] = inc;
      }
      __syncthreads();
      const int R = s_rounds[TC_LEAVES];
      for (int r = warp; r < R; r += TC_BLOCK / 32) {
        const int lf = s_round_leaf[r];
        const int j = (r - s_rounds[lf]) * 32 + lane;
        const bool act = j < s_n2[lf];
        const int hl = act ? (int)s_l2[lf][j] : 0;
        const int h = chunk + hl;
        const float* px = sx + lf * TC_LEAF; const float* py = sy + lf * TC_LEAF; const float* pz = sz + lf * TC_LEAF;
        int cnt = 0;
        if (KIND == 0) {
          const float4 c = act ? plane_coef[h] : make_float4(0.f, 0.f, 0.f, CUDART_INF_F);
          const u64 A = d_pack2(c.x, c.x), B = d_pack2(c.y, c.y), Cz = d_pack2(c.z, c.z), D = d_pack2(c.w, c.w);
#pragma unroll
          for (int i = 0; i < TC_LEAF; i += 4) {
            const float4 x = *reinterpret_cast<const float4*>(px + i);
            const float4 y = *reinterpret_cast<const float4*>(py + i);
            const float4 z = *reinterpret_cast<const float4*>(pz + i);
            const u64 t0 = d_fma2(A, d_pack2(x.x, x.y), d_fma2(B, d_pack2(y.x, y.y), d_fma2(Cz, d_pack2(z.x, z.y), D)));
            const u64 t1 = d_fma2(A, d_pack2(x.z, x.w), d_fma2(B, d_pack2(y.z, y.w), d_fma2(Cz, d_pack2(z.z, z.w), D)));
            float d0, d1, d2, d3;
            d_unpack2(t0, d0, d1); d_unpack2(t1, d2, d3);
            d_count_lt(fabsf(d0), tau, cnt); d_count_lt(fabsf(d1), tau, cnt);
            d_count_lt(fabsf(d2), tau, cnt); d_count_lt(fabsf(d3), tau, cnt);
          }
        } else {
          CylTest t = d_load_cyl_test4(test12 + (size_t)h * 12);
          if (!act) t.half = 0.f;  // |v| < 0 never holds
          const u64 UX = d_pack2(t.u.x, t.u.x), UY = d_pack2(t.u.y, t.u.y), UZ = d_pack2(t.u.z, t.u.z), UD = d_pack2(t.u.w, t.u.w);
          const u64 WX = d_pack2(t.w.x, t.w.x), WY = d_pack2(t.w.y, t.w.y), WZ = d_pack2(t.w.z, t.w.z), WD = d_pack2(t.w.w, t.w.w);
          const u64 NM = d_pack2(-t.mid, -t.mid);
#pragma unroll
          for (int i = 0; i < TC_LEAF; i += 4) {
            const float4 x = *reinterpret_cast<const float4*>(px + i);
            const float4 y = *reinterpret_cast<const float4*>(py + i);
            const float4 z = *reinterpret_cast<const float4*>(pz + i);
            const u64 x01 = d_pack2(x.x, x.y), x23 = d_pack2(x.z, x.w), y01 = d_pack2(y.x, y.y), y23 = d_pack2(y.z, y.w);
            const u64 z01 = d_pack2(z.x, z.y), z23 = d_pack2(z.z, z.w);
            const u64 A0 = d_fma2(UX, x01, d_fma2(UY, y01, d_fma2(UZ, z01, UD)));
            const u64 B0 = d_fma2(WX, x01, d_fma2(WY, y01, d_fma2(WZ, z01, WD)));
            const u64 A1 = d_fma2(UX, x23, d_fma2(UY, y23, d_fma2(UZ, z23, UD)));
            const u64 B1 = d_fma2(WX, x23, d_fma2(WY, y23, d_fma2(WZ, z23, WD)));
            const u64 v0 = d_fma2(A0, A0, d_fma2(B0, B0, NM));
            const u64 v1 = d_fma2(A1, A1, d_fma2(B1, B1, NM));
            float e0, e1, e2, e3;
            d_unpack2(v0, e0, e1); d_unpack2(v1, e2, e3);
            d_count_lt(fabsf(e0), t.half, cnt); d_count_lt(fabsf(e1), t.half, cnt);
            d_count_lt(fabsf(e2), t.half, cnt); d_count_lt(fabsf(e3), t.half, cnt);
          }
        }
        if (act && cnt) atomicAdd(&s_cnt[hl], cnt);
      }
      __syncthreads();
      for (int k = threadIdx.x; k < hc; k += TC_BLOCK) {
        const int c = s_cnt[k];
        if (c) atomicAdd(&counts[chunk + k], c);
      }
      // (the next chunk's first barrier orders these reads before the lists / counts are rebuilt)
    }
  }
  if (d_last_block(ticket, gridDim.x))
    d_block_argmax_send<TC_BLOCK>(counts, h_begin, h_end, key_out, comm, KIND, plane_coef, model7, test12, H);
}

constexpr int AM_BLOCK = 256;
__global__ void __launch_bounds__(AM_BLOCK) k_argmax(const int* __restrict__ counts, int H, unsigned long long* key_out, const CommDev* comm, int kind,
                                                      const float4* __restrict__ plane_coef, const float* __restrict__ model7,
                                                      const float* __restrict__ test12) {
  d_block_argmax_send<AM_BLOCK>(counts, 0, H, key_out, comm, kind, plane_coef, model7, test12, H);
}

// ---- selection + refit ----------------------------------------------------------------------
// Device-resident model state (mirrors gm_model plus double-precision iterate of the cylinder).
struct ModelState {
  int kind, best_id, best_count, refit_count;
  float hyp[8];
  float coef[8];
  float rms, pad_;
  float test_hyp[12];   // cylinder: inlier test of the winning hypothesis (fixes the refit set)
  float test_coef[12];  // cylinder: inlier test of the refined model (labels)
  double q[3], dir[3], r;  // cylinder iterate
  int n_inl, pad2_;        // size of the compacted inlier array (cylinder)
};

// decode the (possibly all-reduced) key and initialise the model state from the hypothesis table;
// called by ONE thread (the refit kernels fold this in instead of a 1-thread launch)
__device__ void d_select(const unsigned long long* __restrict__ key, int kind, int H,
                         const float4* __restrict__ plane_coef, const float* __restrict__ model7,
                         const float* __restrict__ test12, ModelState* ms) {
  unsigned long long k = *key;
  int count = (int)(unsigned)(k >> 32) - 1;
  int id = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
  ms->kind = kind;
  ms->refit_count = 0;
  ms->rms = 0.f;
  ms->pad_ = 0.f;
  ms->n_inl = 0; ms->pad2_ = 0;
  for (int i = 0; i < 8; ++i) { ms->hyp[i] = 0.f; ms->coef[i] = 0.f; }
  for (int i = 0; i < 12; ++i) { ms->test_hyp[i] = 0.f; ms->test_coef[i] = 0.f; }
  if (count < 0 || id < 0 || id >= H) { ms->best_id = -1; ms->best_count = -1; return; }
  ms->best_id = id;
  ms->best_count = count;
  if (kind == 0) {
    float4 c = plane_coef[id];
    ms->hyp[0] = c.x; ms->hyp[1] = c.y; ms->hyp[2] = c.z; ms->hyp[3] = c.w;
    for (int i = 0; i < 4; ++i) ms->coef[i] = ms->hyp[i];
  } else {
    for (int i = 0; i < 7; ++i) { ms->hyp[i] = model7[(size_t)id * 7 + i]; ms->coef[i] = ms->hyp[i]; }
    for (int i = 0; i < 12; ++i) { ms->test_hyp[i] = test12[(size_t)id * 12 + i]; ms->test_coef[i] = ms->test_hyp[i]; }
    for (int i = 0; i < 3; ++i) { ms->q[i] = ms->hyp[i]; ms->dir[i] = ms->hyp[3 + i]; }
    ms->r = ms->hyp[6];
  }
}

constexpr int RF_BLOCK = 256;

// plane refit: double sums {xx,xy,xz,yy,yz,zz,x,y,z,count} over the hypothesis inliers, then
// (last block) covariance -> Jacobi -> coefficients
__global__ void __launch_bounds__(RF_BLOCK)
k_plane_refit(const float4* __restrict__ pts, const int* __restrict__ n_ptr, const unsigned long long* __restrict__ key, int H,
              const float4* __restrict__ plane_coef, ModelState* ms, float tau, double* __restrict__ partials, unsigned* counter) {
  __shared__ double sm[10 * (RF_BLOCK / 32)];
  __shared__ double fin[10];
  const int n = *n_ptr;
  double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // every block decodes the winner itself (k_select folded in)
  const unsigned long long kk = *key;
  const int best_count = (int)(unsigned)(kk >> 32) - 1;
  const int best_id = (int)(0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull));
  const bool have = best_count >= 0 && best_id >= 0 && best_id < H;
  if (have) {
    const float4 c = plane_coef[best_id];
    for (int i = blockIdx.x * RF_BLOCK + threadIdx.x; i < n; i += gridDim.x * RF_BLOCK) {
      float4 p = pts[i];
      if (d_plane_inlier(c, p, tau)) {
        double x = p.x, y = p.y, z = p.z;
        s[0] += x * x; s[1] += x * y; s[2] += x * z; s[3] += y * y; s[4] += y * z; s[5] += z * z;
        s[6] += x; s[7] += y; s[8] += z; s[9] += 1.0;
      }
    }
  }
  block_sum_store<10, RF_BLOCK>(s, sm, partials + (size_t)blockIdx.x * 10);
  if (!d_last_block(counter, gridDim.x)) return;
  d_reduce_partials<10>(partials, gridDim.x, fin);
  if (threadIdx.x != 0) return;
  d_select(key, 0, H, plane_coef, nullptr, nullptr, ms);
  if (!have) return;
  long long cnt = (long long)(fin[9] + 0.5);
  ms->refit_count = (int)cnt;
  if (cnt <= 3) return;
  double c = (double)cnt, mx = fin[6] / c, my = fin[7] / c, mz = fin[8] / c;
  double C[9] = {fin[0] / c - mx * mx, fin[1] / c - mx * my, fin[2] / c - mx * mz, 0, fin[3] / c - my * my, fin[4] / c - my * mz, 0, 0, fin[5] / c - mz * mz};
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  double vals[3], vecs[9];
  d_jacobi3(C, vals, vecs);
  double nx = vecs[0], ny = vecs[3], nz = vecs[6];
  if (nx * ms->hyp[0] + ny * ms->hyp[1] + nz * ms->hyp[2] < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  ms->coef[0] = (float)nx; ms->coef[1] = (float)ny; ms->coef[2] = (float)nz;
  ms->coef[3] = (float)(-(nx * mx + ny * my + nz * mz));
  ms->rms = (float)sqrt(fmax(vals[0], 0.0));
}

__device__ __forceinline__ void d_perp_basis_d(const double dir[3], double u[3], double w[3]) {
  double ax = fabs(dir[0]), ay = fabs(dir[1]), az = fabs(dir[2]);
  int k = 0; double m = ax;
  if (ay < m) { k = 1; m = ay; }
  if (az < m) { k = 2; m = az; }
  double c[3];
  if (k == 0) { c[0] = 0; c[1] = dir[2]; c[2] = -dir[1]; }
  else if (k == 1) { c[0] = -dir[2]; c[1] = 0; c[2] = dir[0]; }
  else { c[0] = dir[1]; c[1] = -dir[0]; c[2] = 0; }
  double l = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  u[0] = c[0] / l; u[1] = c[1] / l; u[2] = c[2] / l;
  w[0] = dir[1] * u[2] - dir[2] * u[1]; w[1] = dir[2] * u[0] - dir[0] * u[2]; w[2] = dir[0] * u[1] - dir[1] * u[0];
}

constexpr int GN_NV = 22;  // J^T J (15) | J^T r (5) | count | sum r^2

// Gauss-Newton terms of one inlier at the iterate (q, dir, r) with (u, w) the perpendicular basis of dir
__device__ __forceinline__ void d_gn_accumulate(const float4 p, const double q0, const double q1, const double q2, const double* dir,
                                                const double* u, const double* w, const double r, double (&s)[GN_NV]) {
  s[20] += 1.0;
  const double vx = (double)p.x - q0, vy = (double)p.y - q1, vz = (double)p.z - q2;
  const double A = u[0] * vx + u[1] * vy + u[2] * vz;
  const double B = w[0] * vx + w[1] * vy + w[2] * vz;
  const double tt = dir[0] * vx + dir[1] * vy + dir[2] * vz;
  const double dist = sqrt(A * A + B * B);
  if (!(dist > 1e-12)) return;
  const double res = dist - r;
  s[21] += res * res;
  const double J[5] = {-A / dist, -B / dist, -A * tt / dist, -B * tt / dist, -1.0};
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 5; ++a) {
#pragma unroll
    for (int b = a; b < 5; ++b) s[idx++] += J[a] * J[b];
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) s[15 + a] += J[a] * res;
}

// The refit set is FIXED (inliers of the winning hypothesis), so it is compacted once (stable look-back compaction) and
// the Gauss-Newton passes run over the dense array.  (Accumulating the sums of the first Gauss-Newton step here as
// well was measured: the compaction went from 13 to 44 us -- 128 registers, double-precision chains inside a look-back
// kernel -- for 35 us saved in k_cyl_gn_all; the sums are taken by k_cyl_gn_all's own first pass instead.)
__global__ void __launch_bounds__(CP_BLOCK, 4)
k_cyl_inlier_compact(const float4* __restrict__ pts, const int* __restrict__ n_ptr, const unsigned long long* __restrict__ key, int H,
                     const float* __restrict__ model7, const float* __restrict__ test12, ModelState* ms,
                     float4* __restrict__ inl, unsigned long long* state, TileCtl* ctl, int* err) {
  __shared__ CompactSmem<CP_BLOCK, CPL_IPT> sm;
  const int n = *n_ptr;
  unsigned epoch;
  const int tile = tile_begin(ctl, epoch), base = tile * CPL_TILE;
  if (base >= n) {
    // empty cloud: tile 0 still has to publish the (empty) model state
    if (tile == 0 && threadIdx.x == 0) { d_select(key, 1, H, nullptr, model7, test12, ms); ms->n_inl = 0; }
    tile_end(ctl);
    return;
  }
  // every block decodes the winner itself (k_select folded in)
  const unsigned long long kk = *key;
  const int best_count = (int)(unsigned)(kk >> 32) - 1;
  const int best_id = (int)(0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull));
  const bool have = best_count >= 0 && best_id >= 0 && best_id < H;
  const CylTest t = d_load_cyl_test(test12 + (size_t)(have ? best_id : 0) * 12);
  bool f[CPL_IPT];
  float4 p[CPL_IPT];
#pragma unroll
  for (int j = 0; j < CPL_IPT; ++j) {
    int i = base + j * CP_BLOCK + threadIdx.x;
    f[j] = false;
    if (i < n) { p[j] = pts[i]; f[j] = have && d_cyl_inlier(t, p[j]); }
  }
  unsigned ranks[CPL_IPT], total;
  tile_compact_ranks<CP_BLOCK, CPL_IPT>(f, ranks, total, state, epoch, tile, err, sm);
#pragma unroll
  for (int j = 0; j < CPL_IPT; ++j)
    if (f[j]) inl[ranks[j]] = p[j];
  if (base + CPL_TILE >= n && threadIdx.x == 0) { d_select(key, 1, H, nullptr, model7, test12, ms); ms->n_inl = (int)total; }
  tile_end(ctl);
}

__device__ bool d_solve5(double A[5][5], double b[5], double x[5]) {
  for (int c = 0; c < 5; ++c) {
    int piv = c; double best = fabs(A[c][c]);
    for (int r = c + 1; r < 5; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (!(best > 1e-300)) return false;
    if (piv != c) {
      for (int k = 0; k < 5; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < 5; ++r) {
      double f = A[r][c] / A[c][c];
      for (int k = c; k < 5; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 4; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < 5; ++k) s -= A[r][k] * x[k];
    x[r] = s / A[r][r];
  }
  return true;
}

// All Gauss-Newton passes of the cylinder refit in ONE cooperative launch (grid <= co-resident capacity, guaranteed by
// cudaLaunchCooperativeKernel).  Per pass every block reduces its share of the compacted inliers to 22 doubles
// (J^T J: 15, J^T r: 5, count, sum r^2), publishes them, crosses a grid barrier, and then EVERY block sums the per-block
// partials in block order and solves the same 5x5 system (redundant but deterministic: no second barrier, no broadcast).
// The loop ends when the step is below 1e-4 (m / rad; Gauss-Newton converges quadratically here, the next step would be
// ~1e-8: same rule in the oracle) -- the RMS at the new iterate then follows from the sums already at hand through the
// Gauss-Newton model, sum (r + J x)^2 = sum r^2 + 2 x.J^T r + x^T J^T J x, exact to O(|x|^2) -- or when `iters` updates were
// made, in which case the sums of one more pass give the RMS.  Typical scan: TWO passes over the inliers (round 1: six).
__device__ __forceinline__ void d_grid_barrier(unsigned* count, unsigned target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(count, 1u);
    int spins = 0;
    while ((int)(*((volatile unsigned*)count) - target) < 0 && ++spins < SPIN_BOUND) {}  // wrap-safe
    if (spins >= SPIN_BOUND) atomicExch(err, 3);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RF_BLOCK)
k_cyl_gn_all(const float4* __restrict__ inl, ModelState* ms, int iters, float tau, double* __restrict__ partials /* 2 x grid x 22 */,
             unsigned* bars /* [0],[1] barrier counters used alternately, [2] which one this launch uses */, int* err) {
  __shared__ double sm[GN_NV * (RF_BLOCK / 32)];
  __shared__ double fin[GN_NV];
  __shared__ double it_q[3], it_dir[3], it_r, s_rms2;
  __shared__ int s_state;  // 0 = another pass is needed, 1 = finished
  // two barrier counters used alternately by consecutive launches: this launch counts on
  // barrier_count from 0 and clears the other one for the next launch (the number of passes is
  // data dependent, so a cumulative target cannot be pre-computed).  Which one is whose lives in device
  // memory too (bars[2], toggled by block 0 at the end of the launch), so that consecutive launches need no
  // host-side argument: the launch is replayable from a CUDA graph.
  const unsigned par = *(volatile unsigned*)&bars[2] & 1u;
  unsigned* barrier_count = bars + par;
  unsigned* barrier_next = bars + (par ^ 1u);
  const bool have = ms->best_id >= 0;
  const int n = have ? ms->n_inl : 0;
  const int nb = gridDim.x;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) { it_q[k] = ms->q[k]; it_dir[k] = ms->dir[k]; }
    it_r = ms->r;
    s_state = 0; s_rms2 = 0.0;
  }
  __syncthreads();
  long long cnt = 0;
  int pass = 0;  // barriers crossed so far in this launch
  for (int it = -1;; ++it) {
    // it = -1: no sums yet, take the first pass.  Otherwise fin = the sums at the current iterate; every block takes
    // the same decisions from the same numbers
    if (it >= 0 && threadIdx.x == 0) {
      cnt = (long long)(fin[20] + 0.5);
      bool done = true;
      if (cnt > 5) {
        s_rms2 = fin[21] / (double)cnt;
        if (it < iters) {
          double JTJ[5][5], rhs[5], x[5];
          int idx = 0;
          for (int a = 0; a < 5; ++a) for (int b = a; b < 5; ++b) { JTJ[a][b] = fin[idx]; JTJ[b][a] = fin[idx]; ++idx; }
          for (int a = 0; a < 5; ++a) rhs[a] = -fin[15 + a];
          if (d_solve5(JTJ, rhs, x)) {
            double dir[3] = {it_dir[0], it_dir[1], it_dir[2]}, u[3], w[3];
            d_perp_basis_d(dir, u, w);
            for (int k = 0; k < 3; ++k) it_q[k] += x[0] * u[k] + x[1] * w[k];
            for (int k = 0; k < 3; ++k) dir[k] += x[2] * u[k] + x[3] * w[k];
            const double dl = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
            for (int k = 0; k < 3; ++k) it_dir[k] = dir[k] / dl;
            it_r += x[4];
            if (x[0] * x[0] + x[1] * x[1] + x[4] * x[4] < 1e-8 && x[2] * x[2] + x[3] * x[3] < 1e-8) {
              // converged: RMS at the new iterate from the Gauss-Newton model of the residuals
              double lin = 0.0, quad = 0.0;
              idx = 0;
              for (int a = 0; a < 5; ++a) {
                lin += x[a] * fin[15 + a];
                for (int b = a; b < 5; ++b) { quad += (a == b ? 1.0 : 2.0) * x[a] * x[b] * fin[idx]; ++idx; }
              }
              s_rms2 = fmax((fin[21] + 2.0 * lin + quad) / (double)cnt, 0.0);
            } else {
              done = false;  // sums at the new iterate are needed (for the next step, or for the RMS after the last one)
            }
          }
        }
      }
      s_state = done ? 1 : 0;
    }
    __syncthreads();
    if (it >= 0 && s_state) break;
    double s[GN_NV];
#pragma unroll
    for (int k = 0; k < GN_NV; ++k) s[k] = 0.0;
    {
      const double q0 = it_q[0], q1 = it_q[1], q2 = it_q[2], r = it_r;
      double dir[3] = {it_dir[0], it_dir[1], it_dir[2]}, u[3], w[3];
      d_perp_basis_d(dir, u, w);
      for (int i = blockIdx.x * RF_BLOCK + threadIdx.x; i < n; i += nb * RF_BLOCK) d_gn_accumulate(inl[i], q0, q1, q2, dir, u, w, r, s);
    }
    double* buf = partials + (size_t)(pass & 1) * nb * GN_NV;
    block_sum_store<GN_NV, RF_BLOCK>(s, sm, buf + (size_t)blockIdx.x * GN_NV);
    ++pass;
    d_grid_barrier(barrier_count, (unsigned)pass * (unsigned)nb, err);
    d_reduce_partials<GN_NV>(buf, nb, fin);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *barrier_next = 0u;
    bars[2] = par ^ 1u;
    if (have) {
      ms->refit_count = (int)cnt;
      if (cnt > 5) {
        ms->rms = (float)sqrt(s_rms2);
        for (int k = 0; k < 3; ++k) { ms->q[k] = it_q[k]; ms->dir[k] = it_dir[k]; ms->coef[k] = (float)it_q[k]; ms->coef[3 + k] = (float)it_dir[k]; }
        ms->r = it_r;
        ms->coef[6] = (float)it_r;
        float t12[12];
        if (d_cyl_test_params(ms->coef, tau, t12)) for (int k = 0; k < 12; ++k) ms->test_coef[k] = t12[k];
      }
    }
    __threadfence();
  }
}

// labels from the refined models: 1 = plane inlier, else 2 = cylinder inlier, else 0
__global__ void k_label(const float4* __restrict__ pts, const int* __restrict__ n_ptr, const ModelState* __restrict__ plane,
                        const ModelState* __restrict__ cyl, int have_plane, int have_cyl, float tau, unsigned char* __restrict__ labels) {
  const int n = *n_ptr;
  const bool hp = have_plane && plane->best_id >= 0, hc = have_cyl && cyl->best_id >= 0;
  float4 pc = make_float4(0, 0, 0, 0);
  CylTest ct;
  ct.u = make_float4(0, 0, 0, 0); ct.w = ct.u; ct.mid = 0.f; ct.half = 0.f;
  if (hp) pc = make_float4(plane->coef[0], plane->coef[1], plane->coef[2], plane->coef[3]);
  if (hc) ct = d_load_cyl_test(cyl->test_coef);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    unsigned char l = 0;
    if (hp && d_plane_inlier(pc, p, tau)) l = 1;
    else if (hc && d_cyl_inlier(ct, p)) l = 2;
    labels[i] = l;
  }
}

}  // namespace gm
