// gm_map.cuh — aggregated voxel map across scans (builder-defined: the reference
// processes every scan on its own, src/geometric_mapping.cpp:48-125, and keeps nothing).
// An open-addressing hash table keyed by the GLOBAL voxel index floor(p * inv_leaf) (the VoxelGrid
// cell of pcl::VoxelGrid, independent of any bounding box), holding per voxel the point count and the
// coordinate sums in 2^-20 m fixed point.  Integer atomics are associative, so the map does not depend
// on the order in which scans, blocks or threads arrive: bitwise reproducible, and two maps built from
// the same scans in a different order are identical.
#pragma once
#include "gm_stages.cuh"

namespace gm {

constexpr unsigned long long MAP_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int MAP_COORD_BITS = 21;                 // +-2^20 voxels per axis
constexpr int MAP_COORD_OFF = 1 << 20;
constexpr float MAP_FX = GM_FX20;                  // 2^20: exact scaling of a float (same fixed point as the VoxelGrid centroids)

struct MapState { unsigned long long n_points; int n_voxels; int overflow; int out_of_range; int pad_; };

__host__ __device__ __forceinline__ unsigned long long map_hash(unsigned long long k) {  // splitmix64 finaliser
  k ^= k >> 30; k *= 0xBF58476D1CE4E5B9ull;
  k ^= k >> 27; k *= 0x94D049BB133111EBull;
  k ^= k >> 31;
  return k;
}
__host__ __device__ __forceinline__ bool map_pack(int ix, int iy, int iz, unsigned long long& key) {
  const int lim = MAP_COORD_OFF;
  if (ix < -lim || ix >= lim || iy < -lim || iy >= lim || iz < -lim || iz >= lim) return false;
  key = (unsigned long long)(unsigned)(ix + lim) | ((unsigned long long)(unsigned)(iy + lim) << MAP_COORD_BITS) |
        ((unsigned long long)(unsigned)(iz + lim) << (2 * MAP_COORD_BITS));
  return true;
}
__host__ __device__ __forceinline__ void map_unpack(unsigned long long key, int& ix, int& iy, int& iz) {
  const unsigned long long m = (1ull << MAP_COORD_BITS) - 1ull;
  ix = (int)(key & m) - MAP_COORD_OFF;
  iy = (int)((key >> MAP_COORD_BITS) & m) - MAP_COORD_OFF;
  iz = (int)((key >> (2 * MAP_COORD_BITS)) & m) - MAP_COORD_OFF;
}

struct Pose34 { float r[12]; };  // row-major [R | t]

// Insert the points pts[0..n) (optionally only those with labels[i] == label_filter) after the rigid
// transform p' = R p + t, evaluated as fmaf chains (bit-identical to the oracle's std::fmaf).
__global__ void k_map_insert(const float4* __restrict__ pts, const unsigned char* __restrict__ labels, const int* __restrict__ n_ptr,
                             int label_filter, Pose34 T, int use_pose, float inv_leaf, unsigned long long* __restrict__ keys,
                             int* __restrict__ cnt, long long* __restrict__ sums /* 3 per slot */, unsigned long long slot_mask,
                             MapState* st) {
  const int n = *n_ptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (label_filter >= 0 && labels[i] != (unsigned char)label_filter) continue;
    const float4 p = pts[i];
    float x = p.x, y = p.y, z = p.z;
    if (use_pose) {
      x = fmaf(T.r[0], p.x, fmaf(T.r[1], p.y, fmaf(T.r[2], p.z, T.r[3])));
      y = fmaf(T.r[4], p.x, fmaf(T.r[5], p.y, fmaf(T.r[6], p.z, T.r[7])));
      z = fmaf(T.r[8], p.x, fmaf(T.r[9], p.y, fmaf(T.r[10], p.z, T.r[11])));
    }
    if (!finite3(x, y, z)) { atomicAdd(&st->out_of_range, 1); continue; }
    const float fx = floorf(x * inv_leaf), fy = floorf(y * inv_leaf), fz = floorf(z * inv_leaf);
    unsigned long long key;
    if (!(fabsf(fx) < 2.0e6f && fabsf(fy) < 2.0e6f && fabsf(fz) < 2.0e6f) || !map_pack((int)fx, (int)fy, (int)fz, key)) {
      atomicAdd(&st->out_of_range, 1);
      continue;
    }
    unsigned long long slot = map_hash(key) & slot_mask;
    bool placed = false;
    for (unsigned long long probe = 0; probe <= slot_mask; ++probe) {
      unsigned long long prev = keys[slot];
      if (prev == MAP_EMPTY) {
        prev = atomicCAS(&keys[slot], MAP_EMPTY, key);
        if (prev == MAP_EMPTY) { atomicAdd(&st->n_voxels, 1); prev = key; }
      }
      if (prev == key) { placed = true; break; }
      slot = (slot + 1) & slot_mask;
    }
    if (!placed) { atomicExch(&st->overflow, 1); continue; }
    atomicAdd(&cnt[slot], 1);
    // x * 2^20 is exact in float (power-of-two scaling); llrintf rounds it to the nearest integer
    atomicAdd(reinterpret_cast<unsigned long long*>(sums) + 3 * slot + 0, (unsigned long long)__float2ll_rn(x * MAP_FX));
    atomicAdd(reinterpret_cast<unsigned long long*>(sums) + 3 * slot + 1, (unsigned long long)__float2ll_rn(y * MAP_FX));
    atomicAdd(reinterpret_cast<unsigned long long*>(sums) + 3 * slot + 2, (unsigned long long)__float2ll_rn(z * MAP_FX));
    atomicAdd(&st->n_points, 1ull);
  }
}

// Raw re-insertion of saved voxels (gm_map_load): key, count and fixed-point sums as stored.
__global__ void k_map_restore(const unsigned long long* __restrict__ in_keys, const int* __restrict__ in_cnt,
                              const long long* __restrict__ in_sums, int n, unsigned long long* __restrict__ keys,
                              int* __restrict__ cnt, long long* __restrict__ sums, unsigned long long slot_mask, MapState* st) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long key = in_keys[i];
    unsigned long long slot = map_hash(key) & slot_mask;
    bool placed = false;
    for (unsigned long long probe = 0; probe <= slot_mask; ++probe) {
      unsigned long long prev = atomicCAS(&keys[slot], MAP_EMPTY, key);
      if (prev == MAP_EMPTY) { atomicAdd(&st->n_voxels, 1); prev = key; }
      if (prev == key) { placed = true; break; }
      slot = (slot + 1) & slot_mask;
    }
    if (!placed) { atomicExch(&st->overflow, 1); continue; }
    atomicAdd(&cnt[slot], in_cnt[i]);
    for (int a = 0; a < 3; ++a)
      atomicAdd(reinterpret_cast<unsigned long long*>(sums) + 3 * slot + a, (unsigned long long)in_sums[3 * (size_t)i + a]);
    atomicAdd(&st->n_points, (unsigned long long)in_cnt[i]);
  }
}

// Gather the occupied slots into dense arrays (order = slot order; the host sorts by key).
__global__ void k_map_export(const unsigned long long* __restrict__ keys, const int* __restrict__ cnt, const long long* __restrict__ sums,
                             unsigned long long n_slots, unsigned long long* __restrict__ out_keys, int* __restrict__ out_cnt,
                             long long* __restrict__ out_sums, int capacity, int* cursor) {
  for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < n_slots; s += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[s];
    if (k == MAP_EMPTY) continue;
    const int o = atomicAdd(cursor, 1);
    if (o >= capacity) continue;
    out_keys[o] = k; out_cnt[o] = cnt[s];
    out_sums[3 * (size_t)o] = sums[3 * s]; out_sums[3 * (size_t)o + 1] = sums[3 * s + 1]; out_sums[3 * (size_t)o + 2] = sums[3 * s + 2];
  }
}

}  // namespace gm
