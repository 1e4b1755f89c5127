"""geometric_mapping_b200 — B200-native (sm_100a) per-scan point-cloud hot path of the ROS package
wangqiaoli/geometric_mapping, behind a C-ABI (include/gm_capi.h).

Python here is plumbing only: `capi` binds the shared library with ctypes, `tunnel_processing`
mirrors the reference's free functions (chopCloud / getNormals / rvizNormals / getLocalFrame /
rvizEigens) on top of it, `synth` generates the seeded synthetic scans of SURVEY.md section 8(d).
There is no CPU fallback: importing `capi` without the built CUDA library raises.
"""
__version__ = "0.1.0"
