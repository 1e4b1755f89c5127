// Host mirror of the reference's `class Parameters` (include/geometric_mapping/paramHandler.hpp:9-37):
// same getters, same defaults; the ROS NodeHandle is replaced by a key -> value source so the class
// works without ROS (values parsed from launch/mapping.launch-style <param> lines or set in code).
#pragma once
#include <map>
#include <string>

#include "../../include/gm_capi.h"

namespace gmhost {

using ParamSource = std::map<std::string, std::string>;
// <param name="k" value="v" .../> lines of a ROS launch file -> ParamSource ($(arg x) left verbatim)
ParamSource parseLaunchParams(const std::string& launch_xml_text);

class Parameters {
 public:
  explicit Parameters(const ParamSource& node);
  double getBoxFilterBound() const { return boxFilterBound; }
  double getLeafSize() const { return leafSize; }
  double getNeighborRadius() const { return neighborRadius; }
  double getWeightingFactor() const { return weightingFactor; }
  bool displayCloud() const { return rvizCloud; }
  bool displayNormals() const { return rvizNormals; }
  bool displayCenterAxis() const { return rvizCenterAxis; }
  bool usePCLViz() const { return pclviz; }
  gm_params toGm() const;  // the same values as the C-ABI's parameter block

 private:
  double boxFilterBound = 5.0, leafSize = .1, neighborRadius = .03, weightingFactor = .2;
  bool rvizCloud = true, rvizNormals = true, rvizCenterAxis = true, pclviz = false;
  int weightMode = 0, arrowMode = 0;  // quirk switches (gm_params.weight_mode / arrow_mode); 0 = as the reference
};

}  // namespace gmhost
