// gm_compress.cuh — builder-defined "geometric approximation (compression)" of the segmented
// cloud (the reference only carries the idea in its package name).
//   compressed scan = refined plane  {coef, in-plane basis, 2-D bounds, count, rms}
//                   + refined cylinder {coef, extent along its axis, count, rms}
//                   + center-axis polyline slices
//                   + the residual points (label 0) voxel-downsampled at the VoxelGrid leaf size
//   ratio = 16*M' / bytes_out;  reconstruction error = RMS over all M' points of the distance to
//   what represents them (plane / cylinder surface, or the centroid of their residual voxel).
// Pass 1 (k_comp_split) is one stable look-back compaction of the label-0 points that also
// accumulates the per-primitive statistics (double partials per block, fixed-order reduction in the
// last block: reproducible).  The residual cloud then goes through the same voxel kernels as a4.
#pragma once
#include "gm_polyline.cuh"

namespace gm {

struct CompStats {
  int n_points, n_plane, n_cyl, n_residual;
  float plane_coef[4], plane_u[3], plane_v[3], plane_bounds[4], plane_rms;
  float cyl_coef[7], cyl_t_range[2], cyl_rms;
  float residual_rms, total_rms;
  double sq_plane, sq_cyl, sq_res;  // sums of squared distances
};

constexpr int CS_NV = 4;  // sq_plane, n_plane, sq_cyl, n_cyl

// plane in-plane basis (double, canonical perp basis of the normal) and cylinder axis, computed
// identically by every block
struct CompFrame { double pu[3], pv[3], q[3], dir[3], r; float4 pc; CylTest ct; bool hp, hc; };

__device__ __forceinline__ void d_comp_frame(const ModelState* plane, const ModelState* cyl, int have_plane, int have_cyl, CompFrame& f) {
  f.hp = have_plane && plane->best_id >= 0;
  f.hc = have_cyl && cyl->best_id >= 0;
  f.pc = make_float4(0, 0, 0, 0);
  for (int k = 0; k < 3; ++k) { f.pu[k] = f.pv[k] = f.q[k] = f.dir[k] = 0.0; }
  f.r = 0.0;
  f.ct.u = make_float4(0, 0, 0, 0); f.ct.w = f.ct.u; f.ct.mid = 0.f; f.ct.half = 0.f;
  if (f.hp) {
    f.pc = make_float4(plane->coef[0], plane->coef[1], plane->coef[2], plane->coef[3]);
    double nrm[3] = {(double)plane->coef[0], (double)plane->coef[1], (double)plane->coef[2]};
    d_perp_basis_d(nrm, f.pu, f.pv);
  }
  if (f.hc) {
    f.ct = d_load_cyl_test(cyl->test_coef);
    for (int k = 0; k < 3; ++k) { f.q[k] = (double)cyl->coef[k]; f.dir[k] = (double)cyl->coef[3 + k]; }
    f.r = (double)cyl->coef[6];
  }
}

__global__ void __launch_bounds__(CP_BLOCK, 3)
k_comp_split(const float4* __restrict__ pts, const unsigned char* __restrict__ labels, const int* __restrict__ n_ptr,
             const ModelState* __restrict__ plane, const ModelState* __restrict__ cyl, int have_plane, int have_cyl,
             float4* __restrict__ res_pts, VoxState* res_vs, unsigned long long* state, TileCtl* ctl, int* err,
             double* __restrict__ part_sum /* CS_NV x grid */, unsigned long long* __restrict__ part_mm /* 6 x grid */,
             unsigned* ticket, CompStats* out) {
  __shared__ CompactSmem<CP_BLOCK, CP_IPT> sm;
  __shared__ double s_sum[CS_NV * (CP_BLOCK / 32)];
  __shared__ unsigned long long s_mm[6][CP_BLOCK / 32];
  __shared__ float s_red[6][CP_BLOCK / 32];
  __shared__ double fin[CS_NV];
  const int n = *n_ptr;
  unsigned epoch;
  const int tile = tile_begin(ctl, epoch), base = tile * CP_TILE;
  CompFrame F;
  d_comp_frame(plane, cyl, have_plane, have_cyl, F);
  bool f[CP_IPT];
  float4 p[CP_IPT];
  double s[CS_NV] = {0, 0, 0, 0};
  unsigned long long mm[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};  // umin umax vmin vmax tmin tmax (ordered doubles)
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
  for (int j = 0; j < CP_IPT; ++j) {
    int i = base + j * CP_BLOCK + threadIdx.x;
    f[j] = false;
    if (i < n) {
      p[j] = pts[i];
      const unsigned char l = labels[i];
      const double x = p[j].x, y = p[j].y, z = p[j].z;
      if (l == 1) {
        const float d = fmaf(F.pc.x, p[j].x, fmaf(F.pc.y, p[j].y, fmaf(F.pc.z, p[j].z, F.pc.w)));
        s[0] += (double)d * (double)d; s[1] += 1.0;
        unsigned long long ou = d_double_to_ordered(F.pu[0] * x + F.pu[1] * y + F.pu[2] * z);
        unsigned long long ov = d_double_to_ordered(F.pv[0] * x + F.pv[1] * y + F.pv[2] * z);
        mm[0] = ou < mm[0] ? ou : mm[0]; mm[1] = ou > mm[1] ? ou : mm[1];
        mm[2] = ov < mm[2] ? ov : mm[2]; mm[3] = ov > mm[3] ? ov : mm[3];
      } else if (l == 2) {
        const float A = fmaf(F.ct.u.x, p[j].x, fmaf(F.ct.u.y, p[j].y, fmaf(F.ct.u.z, p[j].z, F.ct.u.w)));
        const float B = fmaf(F.ct.w.x, p[j].x, fmaf(F.ct.w.y, p[j].y, fmaf(F.ct.w.z, p[j].z, F.ct.w.w)));
        const double e = sqrt((double)A * (double)A + (double)B * (double)B) - F.r;
        s[2] += e * e; s[3] += 1.0;
        unsigned long long ot = d_double_to_ordered(F.dir[0] * (x - F.q[0]) + F.dir[1] * (y - F.q[1]) + F.dir[2] * (z - F.q[2]));
        mm[4] = ot < mm[4] ? ot : mm[4]; mm[5] = ot > mm[5] ? ot : mm[5];
      } else {
        f[j] = true;
        mn[0] = fminf(mn[0], p[j].x); mn[1] = fminf(mn[1], p[j].y); mn[2] = fminf(mn[2], p[j].z);
        mx[0] = fmaxf(mx[0], p[j].x); mx[1] = fmaxf(mx[1], p[j].y); mx[2] = fmaxf(mx[2], p[j].z);
      }
    }
  }
  if (base < n) {
    unsigned ranks[CP_IPT], total;
    tile_compact_ranks<CP_BLOCK, CP_IPT>(f, ranks, total, state, epoch, tile, err, sm);
#pragma unroll
    for (int j = 0; j < CP_IPT; ++j)
      if (f[j]) res_pts[ranks[j]] = p[j];
    if (base + CP_TILE >= n && threadIdx.x == 0) res_vs->n = (int)total;
  }
  tile_end(ctl);
  // residual bounding box (order independent)
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = warp_min(mn[a]), hi = warp_max(mx[a]);
    if (lane_id() == 0) { s_red[a][w] = lo; s_red[3 + a][w] = hi; }
  }
  // min/max of the projections
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    unsigned long long v = mm[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(FULL, v, o);
      v = (k & 1) ? (t > v ? t : v) : (t < v ? t : v);
    }
    if (lane_id() == 0) s_mm[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float lo = CUDART_INF_F, hi = -CUDART_INF_F;
    for (int ww = 0; ww < CP_BLOCK / 32; ++ww) { lo = fminf(lo, s_red[threadIdx.x][ww]); hi = fmaxf(hi, s_red[3 + threadIdx.x][ww]); }
    if (lo <= hi) { atomicMin(&res_vs->bbox_min[threadIdx.x], float_to_ordered(lo)); atomicMax(&res_vs->bbox_max[threadIdx.x], float_to_ordered(hi)); }
  }
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    unsigned long long v = s_mm[k][0];
    for (int ww = 1; ww < CP_BLOCK / 32; ++ww) { unsigned long long t = s_mm[k][ww]; v = (k & 1) ? (t > v ? t : v) : (t < v ? t : v); }
    part_mm[(size_t)blockIdx.x * 6 + k] = v;
    __threadfence();
  }
  block_sum_store<CS_NV, CP_BLOCK>(s, s_sum, part_sum + (size_t)blockIdx.x * CS_NV);
  if (!d_last_block(ticket, gridDim.x)) return;
  d_reduce_partials<CS_NV>(part_sum, gridDim.x, fin);
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    unsigned long long v = (k & 1) ? 0ull : ~0ull;
    for (int b = 0; b < (int)gridDim.x; ++b) { unsigned long long t = __ldcg(part_mm + (size_t)b * 6 + k); v = (k & 1) ? (t > v ? t : v) : (t < v ? t : v); }
    s_mm[k][0] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  out->n_points = n;
  out->n_plane = (int)(fin[1] + 0.5);
  out->n_cyl = (int)(fin[3] + 0.5);
  out->n_residual = n - out->n_plane - out->n_cyl;
  out->sq_plane = fin[0]; out->sq_cyl = fin[2]; out->sq_res = 0.0;
  for (int k = 0; k < 4; ++k) { out->plane_coef[k] = F.hp ? plane->coef[k] : 0.f; out->plane_bounds[k] = 0.f; }
  for (int k = 0; k < 3; ++k) { out->plane_u[k] = (float)F.pu[k]; out->plane_v[k] = (float)F.pv[k]; }
  for (int k = 0; k < 7; ++k) out->cyl_coef[k] = F.hc ? cyl->coef[k] : 0.f;
  out->cyl_t_range[0] = out->cyl_t_range[1] = 0.f;
  if (out->n_plane > 0) for (int k = 0; k < 4; ++k) out->plane_bounds[k] = (float)d_ordered_to_double(s_mm[k][0]);
  if (out->n_cyl > 0) { out->cyl_t_range[0] = (float)d_ordered_to_double(s_mm[4][0]); out->cyl_t_range[1] = (float)d_ordered_to_double(s_mm[5][0]); }
  out->plane_rms = out->n_plane > 0 ? (float)sqrt(fin[0] / fin[1]) : 0.f;
  out->cyl_rms = out->n_cyl > 0 ? (float)sqrt(fin[2] / fin[3]) : 0.f;
  out->residual_rms = 0.f; out->total_rms = 0.f;
}

// squared distance of every residual point to the centroid of its voxel; last block finishes the stats
constexpr int CR_BLOCK = 256;
__global__ void __launch_bounds__(CR_BLOCK)
k_comp_residual_error(const float4* __restrict__ res_pts, const int* __restrict__ assign, const float4* __restrict__ centroids,
                      const VoxState* __restrict__ vs, double* __restrict__ partials, unsigned* ticket, CompStats* out) {
  __shared__ double sm[CR_BLOCK / 32];
  __shared__ double fin[1];
  const int n = vs->n;
  double s[1] = {0.0};
  for (int i = blockIdx.x * CR_BLOCK + threadIdx.x; i < n; i += gridDim.x * CR_BLOCK) {
    const float4 p = res_pts[i], c = centroids[assign[i]];
    const double dx = (double)p.x - (double)c.x, dy = (double)p.y - (double)c.y, dz = (double)p.z - (double)c.z;
    s[0] += dx * dx + dy * dy + dz * dz;
  }
  block_sum_store<1, CR_BLOCK>(s, sm, partials + blockIdx.x);
  if (!d_last_block(ticket, gridDim.x)) return;
  d_reduce_partials<1>(partials, gridDim.x, fin);
  if (threadIdx.x != 0) return;
  out->sq_res = fin[0];
  out->residual_rms = n > 0 ? (float)sqrt(fin[0] / (double)n) : 0.f;
  const double tot = out->sq_plane + out->sq_cyl + fin[0];
  out->total_rms = out->n_points > 0 ? (float)sqrt(tot / (double)out->n_points) : 0.f;
}

__global__ void k_vox_reset(VoxState* v) { if (threadIdx.x == 0 && blockIdx.x == 0) d_vox_reset(v); }

}  // namespace gm
