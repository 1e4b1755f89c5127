#include "tunnel_processing.hpp"

#include <algorithm>
#include <cstring>

namespace gmhost {
namespace {
// file-scope state of the shim (the node's globals live at src/geometric_mapping.cpp:35-45)
gm_ctx* g_ctx = nullptr;
size_t g_cap = 0;
gm_params g_base;
bool g_have_base = false;

void ck(gm_status s, const char* where, gm_status also_ok = GM_OK) {
  if (s != GM_OK && s != also_ok) throw GmError(s, where);
}

gm_ctx* ensureContext(size_t points) {
  if (!g_have_base) { gm_params_default(&g_base); g_have_base = true; }
  if (g_ctx && points <= g_cap) return g_ctx;
  gm_params cur = g_base;
  if (g_ctx) { ck(gm_get_params(g_ctx, &cur), "gm_get_params"); gm_destroy(g_ctx); g_ctx = nullptr; }
  g_cap = std::max<size_t>(points + points / 4, 1u << 16);
  ck(gm_create(&cur, g_cap, 16, &g_ctx), "gm_create");
  return g_ctx;
}

// The reference passes these values per call; the context holds them.  Read - modify - write against the CONTEXT's
// parameter block, so everything else in it (quirk switches, RANSAC fields) is preserved.
void setParam(gm_ctx* ctx, double gm_params::*field, double value) {
  gm_params cur;
  ck(gm_get_params(ctx, &cur), "gm_get_params");
  if (cur.*field != value) { cur.*field = value; ck(gm_set_params(ctx, &cur), "gm_set_params"); }
}
}  // namespace

void useParameters(const gm_params& p) {
  g_base = p; g_have_base = true;
  if (g_ctx) ck(gm_set_params(g_ctx, &p), "gm_set_params");
}
gm_ctx* context() { return ensureContext(1); }
void releaseContext() { if (g_ctx) gm_destroy(g_ctx); g_ctx = nullptr; g_cap = 0; }

CloudPtr chopCloud(const double& bound, const CloudPtr& cloud) {
  gm_ctx* ctx = ensureContext(cloud->size());
  setParam(ctx, &gm_params::boxFilterBound, bound);
  ck(gm_upload_scan(ctx, cloud->empty() ? nullptr : &cloud->data()->x, cloud->size(), sizeof(PointXYZ)), "gm_upload_scan");
  ck(gm_crop(ctx), "gm_crop");
  gm_counts c;
  ck(gm_get_counts(ctx, &c), "gm_get_counts");
  auto out = std::make_shared<Cloud>((size_t)c.n_cropped);
  if (!out->empty()) ck(gm_download_cloud(ctx, 0, &out->data()->x, out->size()), "gm_download_cloud");
  return out;
}

NormalsPtr getNormals(const double& neighborRadius, CloudPtr& cloud, SearchPtr& kdtree) {
  gm_ctx* ctx = ensureContext(1);
  kdtree = std::make_shared<SearchHandle>(ctx);  // "set to dynamic memory in function": the device neighbour grid of ctx
  setParam(ctx, &gm_params::neighborRadius, neighborRadius);
  ck(gm_normals(ctx), "gm_normals");
  gm_counts c;
  ck(gm_get_counts(ctx, &c), "gm_get_counts");
  auto normals = std::make_shared<Normals>((size_t)c.n_valid);
  cloud->resize((size_t)c.n_valid);  // in-place compaction of the caller's cloud, as the reference does
  if (c.n_valid > 0) {
    ck(gm_download_normals(ctx, 1, normals->data()->normal, normals->size()), "gm_download_normals");
    ck(gm_download_cloud(ctx, 1, &cloud->data()->x, cloud->size()), "gm_download_cloud");
  }
  return normals;
}

void getLocalFrame(const int& cloudSize, const double& weightingFactor, const NormalsPtr& cloud_normals, Vector3f*& eigenVals,
                   Matrix3f*& eigenVecs) {
  if ((size_t)cloudSize != cloud_normals->size()) throw GmError(GM_ERR_INVALID_ARG, "getLocalFrame: cloudSize");
  gm_ctx* ctx = ensureContext(1);
  setParam(ctx, &gm_params::weightingFactor, weightingFactor);
  ck(gm_local_frame(ctx), "gm_local_frame");
  gm_frame f;
  ck(gm_get_frame(ctx, &f), "gm_get_frame");
  eigenVals = new Vector3f();   // src/tunnel_processing.cpp:131,135: `new`, owned by the caller
  eigenVecs = new Matrix3f();
  std::memcpy(eigenVals->data(), f.vals, sizeof(f.vals));
  std::memcpy(eigenVecs->m, f.vecs, sizeof(f.vecs));
}

Marker* rvizArrow(const Vector3f& start, const Vector3f& end, const Vector3f& scale, const Vector4f& color, const std::string& ns,
                  const int& id, const std::string& frame) {
  Marker* m = new Marker();
  m->frame_id = frame; m->ns = ns; m->id = id;
  for (int k = 0; k < 3; ++k) { m->points[0][k] = start[k]; m->points[1][k] = end[k]; m->scale[k] = scale[k]; }
  m->a = color[0]; m->r = color[1]; m->g = color[2]; m->b = color[3];  // pushed as (a,r,g,b)
  return m;
}

namespace {
void pushArrow(MarkerArray* out, const gm_arrow& a, const char* ns) {
  Marker* m = rvizArrow({a.start[0], a.start[1], a.start[2]}, {a.end[0], a.end[1], a.end[2]}, {a.scale[0], a.scale[1], a.scale[2]},
                        {a.color_argb[0], a.color_argb[1], a.color_argb[2], a.color_argb[3]}, ns, a.id);
  out->push_back(*m);  // the reference copies *rvizArrow(...) and leaks the original (:251, :289); here it is freed
  delete m;
}
}  // namespace

MarkerArray* rvizNormals(const double& leafSize, const CloudPtr& cloud, const SearchPtr& kdtree, const NormalsPtr& normals) {
  (void)cloud; (void)normals;  // both are resident on the device from getNormals
  gm_ctx* ctx = kdtree->ctx;
  setParam(ctx, &gm_params::voxelGridLeafSize, leafSize);
  ck(gm_voxel(ctx), "gm_voxel");
  gm_counts c;
  ck(gm_get_counts(ctx, &c), "gm_get_counts");
  const size_t V = (size_t)c.n_voxels;
  std::vector<float> cen(V * 4), nn(V * 8);
  MarkerArray* out = new MarkerArray();
  if (V == 0) return out;
  gm_status s = gm_download_voxels(ctx, cen.data(), nullptr, nullptr, nullptr, nn.data(), V);
  if (s == GM_ERR_NN_INDEX_RANGE) { delete out; throw GmError(s, "rvizNormals: normals->at(index) out of range (reference quirk B.3)"); }
  ck(s, "gm_download_voxels", GM_WARN_VOXEL_OVERFLOW);
  std::vector<gm_arrow> arrows(V);
  gm_params prm;
  ck(gm_get_params(ctx, &prm), "gm_get_params");
  gm_markers_normals_mode(cen.data(), nn.data(), (int32_t)V, prm.arrow_mode, arrows.data());  // 0 = the reference's arrows (quirk B.4)
  out->reserve(V);
  for (const gm_arrow& a : arrows) pushArrow(out, a, "normals");
  return out;
}

MarkerArray* rvizEigens(const Vector3f& eigenVals, const Matrix3f& eigenVecs) {
  gm_frame f{};
  std::memcpy(f.vals, eigenVals.data(), sizeof(f.vals));
  std::memcpy(f.vecs, eigenVecs.m, sizeof(f.vecs));
  gm_arrow arrows[3];
  gm_markers_eigen(&f, arrows);
  MarkerArray* out = new MarkerArray();
  for (int i = 0; i < 3; ++i) pushArrow(out, arrows[i], "eigenBasis");
  return out;
}

}  // namespace gmhost
