// ROS-free harness with the control flow of the reference node (src/geometric_mapping.cpp): read
// the parameters once, then run the body of cloud_cb (:48-125) on one scan and print what the
// reference logs (eigenvalues / eigenvectors / center axis, src/tunnel_processing.cpp:133-143).
//   geometric_mapping_node [--launch mapping.launch] [--set key=value ...] [--scan file.f32x4 | --synthetic N]
// The scan file is raw little-endian float32 x,y,z,pad records (pcl::PointXYZ memory layout).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <vector>

#include "paramHandler.hpp"
#include "tunnel_processing.hpp"

using namespace gmhost;

static Parameters* params = nullptr;

// the body of cloud_cb (src/geometric_mapping.cpp:48-125), statement for statement, on the mirrored seam
static void cloud_cb(const CloudPtr& cloud) {
  std::fprintf(stderr, "[ INFO] Callback started...\n");
  CloudPtr cloudChopped = chopCloud(params->getBoxFilterBound(), cloud);                       // :57
  std::fprintf(stderr, "[ INFO] Box filter applied...\n");
  SearchPtr kdtree;                                                                             // :62
  NormalsPtr cloudNormals = getNormals(params->getNeighborRadius(), cloudChopped, kdtree);     // :63
  MarkerArray* normalsDisp = rvizNormals(params->getLeafSize(), cloudChopped, kdtree, cloudNormals);  // :70-75
  std::fprintf(stderr, "[ INFO] Surface normals found...\n");
  Vector3f* eigenVals = nullptr;                                                                // :79-80
  Matrix3f* eigenVecs = nullptr;
  getLocalFrame((int)cloudChopped->size(), params->getWeightingFactor(), cloudNormals, eigenVals, eigenVecs);  // :82-88
  Vector3f* centerAxis = new Vector3f(eigenVecs->col(0));  // :91-92 assume that first eigenvec is smallest (ascending order)
  MarkerArray* eigenBasis = rvizEigens(*eigenVals, *eigenVecs);                                 // :96
  std::printf("Weights made of size:\t%zu\n", cloudChopped->size());
  std::printf("Eigenvalues are:\n%g %g %g\n", (*eigenVals)[0], (*eigenVals)[1], (*eigenVals)[2]);
  std::printf("Eigenvectors are:\n");
  for (int r = 0; r < 3; ++r) std::printf("%g %g %g\n", (*eigenVecs)(r, 0), (*eigenVecs)(r, 1), (*eigenVecs)(r, 2));
  std::printf("Center Axis is:\n%g %g %g\n", (*centerAxis)[0], (*centerAxis)[1], (*centerAxis)[2]);
  // :100-117: what is published, as the ROS-free byte payloads of the messages (gm_encode_*)
  if (params->displayCloud()) {
    std::vector<unsigned char> msg(gm_pointcloud2_size(cloudChopped->size(), "/velodyne"));
    size_t bytes = 0;
    gm_encode_pointcloud2(cloudChopped->empty() ? nullptr : &cloudChopped->data()->x, cloudChopped->size(), "/velodyne", 0, 0, 1, msg.data(),
                          msg.size(), &bytes);
    std::printf("publish cloudOutput: %zu points, %zu bytes\n", cloudChopped->size(), bytes);
  }
  if (params->displayNormals()) std::printf("publish normalsOutput: %zu markers\n", normalsDisp->size());
  if (params->displayCenterAxis()) std::printf("publish eigenBasisOutput: %zu markers, E1 scale %g %g %g\n", eigenBasis->size(),
                                               (*eigenBasis)[0].scale[0], (*eigenBasis)[0].scale[1], (*eigenBasis)[0].scale[2]);
  std::fprintf(stderr, "[ INFO] Published...\n");
  delete normalsDisp; delete eigenBasis; delete eigenVals; delete eigenVecs; delete centerAxis;  // (the reference leaks all five)
}

int main(int argc, char** argv) {
  ParamSource src;
  std::string scan_path;
  size_t synthetic = 100000;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--launch" && i + 1 < argc) {
      std::ifstream f(argv[++i]);
      std::stringstream ss; ss << f.rdbuf();
      for (auto& kv : parseLaunchParams(ss.str())) src[kv.first] = kv.second;
    } else if (a == "--set" && i + 1 < argc) {
      std::string kv = argv[++i];
      auto eq = kv.find('=');
      if (eq != std::string::npos) src[kv.substr(0, eq)] = kv.substr(eq + 1);
    } else if (a == "--scan" && i + 1 < argc) scan_path = argv[++i];
    else if (a == "--synthetic" && i + 1 < argc) synthetic = std::strtoull(argv[++i], nullptr, 10);
  }
  Parameters param(src);
  params = &param;
  std::fprintf(stderr, "[ INFO] Launched geometric_mapping_node...\n");

  auto cloud = std::make_shared<Cloud>();
  if (!scan_path.empty()) {
    std::ifstream f(scan_path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", scan_path.c_str()); return 2; }
    size_t bytes = (size_t)f.tellg();
    cloud->resize(bytes / sizeof(PointXYZ));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(cloud->data()), (std::streamsize)(cloud->size() * sizeof(PointXYZ)));
  } else {  // straight cylinder R = 2.5 m along x (config 0 of BASELINE.json)
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> ux(-5.f, 5.f), ut(0.f, 6.2831853f);
    std::normal_distribution<float> nr(0.f, 0.01f);
    cloud->resize(synthetic);
    for (auto& p : *cloud) { float t = ut(rng), r = 2.5f + nr(rng); p.x = ux(rng); p.y = r * std::cos(t); p.z = r * std::sin(t); p.pad = 1.0f; }
  }

  useParameters(param.toGm());  // parameters that are not arguments of the seam functions (quirk switches, is_dense)
  int rc = 0;
  try {
    cloud_cb(cloud);
  } catch (const GmError& e) {
    std::fprintf(stderr, "error: %s (%s)\n", e.what(), gm_last_error(context()));
    rc = 4;
  }
  if (rc == 0) {
    gm_params cur;
    gm_get_params(context(), &cur);
    std::printf("context weight_mode %d arrow_mode %d\n", cur.weight_mode, cur.arrow_mode);
  }
  releaseContext();
  return rc;
}
