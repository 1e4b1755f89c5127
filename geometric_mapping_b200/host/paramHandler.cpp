// See paramHandler.hpp.  Table-driven restatement of src/paramHandler.cpp:12-66: each ROS key is
// looked up once; a missing key keeps the default and logs "ERROR: <field> set to default...".
#include "paramHandler.hpp"

#include <cstdio>
#include <cstdlib>
#include <regex>

namespace gmhost {

ParamSource parseLaunchParams(const std::string& text) {
  ParamSource out;
  static const std::regex re("<param\\s+name=\"([^\"]+)\"\\s+value=\"([^\"]*)\"");
  for (std::sregex_iterator it(text.begin(), text.end(), re), end; it != end; ++it) out[(*it)[1]] = (*it)[2];
  return out;
}

namespace {
bool toBool(const std::string& v) { return v == "true" || v == "True" || v == "1"; }
}  // namespace

Parameters::Parameters(const ParamSource& node) {
  struct D { const char* key; const char* field; double* dst; };
  struct B { const char* key; const char* field; bool* dst; };
  const D doubles[] = {{"boxFilterBound", "boxFilterBound", &boxFilterBound}, {"voxelGridLeafSize", "leafSize", &leafSize},
                       {"neighborRadius", "neighborRadius", &neighborRadius}, {"weightingFactor", "weightingFactor", &weightingFactor}};
  const B bools[] = {{"displayCloud", "displayCloud", &rvizCloud}, {"displayNormals", "displayNormals", &rvizNormals},
                     {"displayCenterAxis", "displayCenterAxis", &rvizCenterAxis}, {"usePCLViz", "pclviz", &pclviz}};
  for (const D& d : doubles) {
    auto it = node.find(d.key);
    if (it != node.end()) { *d.dst = std::strtod(it->second.c_str(), nullptr); std::fprintf(stderr, "[ INFO] %s set to:\t %f\n", d.field, *d.dst); }
    else std::fprintf(stderr, "[ INFO] ERROR: %s set to default...\n", d.field);
  }
  // not parameters of the reference: optional switches for its quirks (absent = reference-faithful, no log line)
  { auto it = node.find("weightMode"); if (it != node.end()) weightMode = std::atoi(it->second.c_str()); }
  { auto it = node.find("arrowMode"); if (it != node.end()) arrowMode = std::atoi(it->second.c_str()); }
  for (const B& b : bools) {
    auto it = node.find(b.key);
    if (it != node.end() && it->second.find("$(") == std::string::npos) { *b.dst = toBool(it->second); std::fprintf(stderr, "[ INFO] %s set to:\t %d\n", b.field, (int)*b.dst); }
    else std::fprintf(stderr, "[ INFO] ERROR: %s set to default...\n", b.field);
  }
}

gm_params Parameters::toGm() const {
  gm_params p;
  gm_params_default(&p);
  p.boxFilterBound = boxFilterBound; p.voxelGridLeafSize = leafSize; p.neighborRadius = neighborRadius; p.weightingFactor = weightingFactor;
  p.displayCloud = rvizCloud; p.displayNormals = rvizNormals; p.displayCenterAxis = rvizCenterAxis; p.usePCLViz = pclviz;
  p.weight_mode = weightMode; p.arrow_mode = arrowMode;
  return p;
}

}  // namespace gmhost
