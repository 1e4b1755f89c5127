"""Host mirror of the reference's `class Parameters` (include/geometric_mapping/paramHandler.hpp:9-37,
src/paramHandler.cpp:12-66): same ROS keys, same defaults, same getter names.  Where ROS exists the
values come from the parameter server; here they come from a dict (e.g. parsed from
launch/mapping.launch) and a missing key keeps its default and logs the reference's message."""
from __future__ import annotations

import logging

from . import capi

log = logging.getLogger("geometric_mapping")

# ROS key -> (field, default)   [src/paramHandler.cpp:13,19,25,31,37,43,49,61]
_KEYS = {
    "boxFilterBound": ("boxFilterBound", 5.0),
    "voxelGridLeafSize": ("leafSize", 0.1),
    "neighborRadius": ("neighborRadius", 0.03),
    "weightingFactor": ("weightingFactor", 0.2),
    "displayCloud": ("rvizCloud", True),
    "displayNormals": ("rvizNormals", True),
    "displayCenterAxis": ("rvizCenterAxis", True),
    "usePCLViz": ("pclviz", False),
}

# values set by launch/mapping.launch:7-17 (displayCylinder is set there but never read)
LAUNCH_FILE_VALUES = {"boxFilterBound": 5.0, "voxelGridLeafSize": 0.5, "neighborRadius": 0.5, "weightingFactor": 0.2,
                      "displayCloud": True, "displayNormals": False, "displayCenterAxis": True, "usePCLViz": False}


class Parameters:
    def __init__(self, node_params: dict | None = None):
        node_params = node_params or {}
        for key, (field, default) in _KEYS.items():
            if key in node_params:
                setattr(self, field, type(default)(node_params[key]))
                log.info("%s set to:\t %s", field if key != "usePCLViz" else "pclviz", getattr(self, field))
            else:
                setattr(self, field, default)
                log.info("ERROR: %s set to default...", field)

    def getBoxFilterBound(self) -> float:
        return self.boxFilterBound

    def getLeafSize(self) -> float:
        return self.leafSize

    def getNeighborRadius(self) -> float:
        return self.neighborRadius

    def getWeightingFactor(self) -> float:
        return self.weightingFactor

    def displayCloud(self) -> bool:
        return self.rvizCloud

    def displayNormals(self) -> bool:
        return self.rvizNormals

    def displayCenterAxis(self) -> bool:
        return self.rvizCenterAxis

    def usePCLViz(self) -> bool:
        return self.pclviz

    def to_gm_params(self, **builder_defined) -> capi.gm_params:
        return capi.default_params(
            boxFilterBound=self.boxFilterBound, voxelGridLeafSize=self.leafSize, neighborRadius=self.neighborRadius,
            weightingFactor=self.weightingFactor, displayCloud=int(self.rvizCloud), displayNormals=int(self.rvizNormals),
            displayCenterAxis=int(self.rvizCenterAxis), usePCLViz=int(self.pclviz), **builder_defined)
