// Host mirror of the reference's processing / visualisation seam
// (include/geometric_mapping/tunnel_processing.hpp:38-82): the SAME function names, argument lists, out-parameters and
// return types (raw pointers to heap objects included), with PCL/Eigen/ROS types replaced by the layout-compatible PODs
// of gm_types.hpp, so the body of cloud_cb (src/geometric_mapping.cpp:48-125) calls them unchanged.  The CUDA context the
// functions run on is a file-scope object of the shim, like `Parameters* params` is one of the node
// (src/geometric_mapping.cpp:35): gmhost::useParameters() installs the node's parameters once, the context itself is
// created on first use and grows with the scans.  Every function is a thin call into the CUDA library through
// include/gm_capi.h; nothing is computed on the CPU here except the marker formatting the reference also does on the host.
#pragma once
#include "gm_types.hpp"

namespace gmhost {

// chopCloud(bound, cloud) -> cropped cloud                       (src/tunnel_processing.cpp:39-49)
CloudPtr chopCloud(const double& bound, const CloudPtr& cloud);
// getNormals(neighborRadius, cloud&, kdtree&): cloud is compacted in place (NaN normals removed),
// kdtree is set in the function                                   (src/tunnel_processing.cpp:52-89)
NormalsPtr getNormals(const double& neighborRadius, CloudPtr& cloud, SearchPtr& kdtree);
// getLocalFrame(cloudSize, weightingFactor, normals, eigenVals*&, eigenVecs*&): both set to dynamic memory   (:92-148)
void getLocalFrame(const int& cloudSize, const double& weightingFactor, const NormalsPtr& cloud_normals, Vector3f*& eigenVals,
                   Matrix3f*& eigenVecs);
// rvizArrow(start, end, scale, color, ns, id, frame)                                  (:161-205)
Marker* rvizArrow(const Vector3f& start, const Vector3f& end, const Vector3f& scale, const Vector4f& color, const std::string& ns,
                  const int& id = 0, const std::string& frame = "/velodyne");
// rvizNormals(leafSize, cloud, kdtree, normals) -> one arrow per voxel               (:208-257)
MarkerArray* rvizNormals(const double& leafSize, const CloudPtr& cloud, const SearchPtr& kdtree, const NormalsPtr& normals);
// rvizEigens(eigenVals, eigenVecs) -> 3 arrows                                        (:260-300)
MarkerArray* rvizEigens(const Vector3f& eigenVals, const Matrix3f& eigenVecs);

// ---- not in the reference: what replaces the PCL objects behind the seam ---------------------------------------------
// parameters that are not arguments of the functions above (is_dense, quirk switches, ...): call once before the first scan
void useParameters(const gm_params& p);
gm_ctx* context();       // the shim's CUDA context (created on first use)
void releaseContext();   // destroy it (end of the node)

}  // namespace gmhost
