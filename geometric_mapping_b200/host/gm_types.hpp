// gm_types.hpp — layout-compatible stand-ins for the PCL / ROS message types that cross the
// reference's L1 seam, so the host shim compiles where ROS and PCL are absent.
//   gmhost::PointXYZ  == pcl::PointXYZ   (16 bytes: x y z + padding word 1.0f)
//   gmhost::Normal    == pcl::Normal     (32 bytes: nx ny nz 0 | curvature 0 0 0)
//   gmhost::Marker    == the numeric/string payload rvizArrow fills in a visualization_msgs::Marker
// Where PCL exists, `cloud->points.data()` can be passed straight to the C-ABI (same bytes).
#pragma once
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gm_capi.h"

namespace gmhost {

struct alignas(16) PointXYZ { float x, y, z, pad = 1.0f; };
struct alignas(16) Normal { float normal[3]; float pad0 = 0.f; float curvature; float pad1[3] = {0.f, 0.f, 0.f}; };
static_assert(sizeof(PointXYZ) == 16 && sizeof(Normal) == 32, "PCL layout");

using Cloud = std::vector<PointXYZ>;
using CloudPtr = std::shared_ptr<Cloud>;
using Normals = std::vector<Normal>;
using NormalsPtr = std::shared_ptr<Normals>;
using Vector3f = std::array<float, 3>;
using Vector4f = std::array<float, 4>;
struct Matrix3f { float m[9]; float operator()(int r, int c) const { return m[r * 3 + c]; } Vector3f col(int c) const { return {m[c], m[3 + c], m[6 + c]}; } };

struct Marker {
  std::string frame_id, ns;
  int id = 0, type_arrow = 1, action_add = 1, seq = 0;
  std::array<double, 3> points[2];
  std::array<double, 3> scale;
  float a, r, g, b;
};
using MarkerArray = std::vector<Marker>;

// Replaces the pcl::search::KdTree<PointXYZ>::Ptr out-parameter of getNormals: the neighbour
// structure lives on the device inside the gm_ctx.
struct SearchHandle {
  gm_ctx* ctx = nullptr;
  explicit SearchHandle(gm_ctx* c) : ctx(c) {}
};
using SearchPtr = std::shared_ptr<SearchHandle>;

struct GmError : std::runtime_error {
  gm_status status;
  GmError(gm_status s, const std::string& where) : std::runtime_error(where + ": " + gm_status_string(s)), status(s) {}
};

}  // namespace gmhost
