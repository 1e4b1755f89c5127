"""Host mirror of the reference's L1 seam (include/geometric_mapping/tunnel_processing.hpp:38-82):
the same free functions, argument meaning and order, on numpy arrays laid out like the PCL types
(PointXYZ = n x 4 float32, Normal = n x 8 float32).  Every function forwards to the CUDA library
through the C-ABI; the kd-tree out-parameter of getNormals is replaced by the `Context` that owns
the device-side neighbour grid.  There is no CPU implementation behind these.  The Context argument is optional:
without it the functions run on a module-level context (created on first use, like the file-scope context of the
C++ shim and `Parameters* params` of the node, src/geometric_mapping.cpp:35), so the reference's own argument lists work:
`chopCloud(bound, cloud)`, `getNormals(radius, cloud)`, `getLocalFrame(n, wf, normals)`, `rvizNormals(leaf, cloud, None, normals)`.

    cloudChopped = chopCloud(bound, cloud, ctx)                       # src/tunnel_processing.cpp:39
    normals, cloudChopped = getNormals(radius, cloudChopped, ctx)     # :52  (cloud is compacted, as the reference mutates it)
    markers = rvizNormals(leaf, cloudChopped, ctx, normals)           # :208
    vals, vecs = getLocalFrame(len(cloudChopped), wf, normals, ctx)   # :92
    basis = rvizEigens(vals, vecs)                                    # :260
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

FRAME_ID = "/velodyne"  # include/geometric_mapping/tunnel_processing.hpp:73


_CTX = None


def use_context(ctx):
    """Install the module-level Context (None = drop it; a fresh one is created on next use)."""
    global _CTX
    _CTX = ctx


def context(points: int = 1 << 20) -> capi.Context:
    global _CTX
    if _CTX is None or _CTX.max_points < points:
        prm = _CTX.params if _CTX is not None else None
        if _CTX is not None:
            _CTX.close()
        _CTX = capi.Context(prm, max_points=max(points + points // 4, 1 << 16), max_hypotheses=4096)
    return _CTX


def _same_cloud(ctx: capi.Context, cloud: np.ndarray, which: int) -> bool:
    c = ctx.counts()
    n = c.n_cropped if which == 0 else c.n_valid
    return n == len(cloud)


def chopCloud(bound: float, cloud: np.ndarray, ctx: capi.Context | None = None) -> np.ndarray:
    ctx = ctx if ctx is not None else context(len(cloud))
    if ctx.params.boxFilterBound != bound:
        p = ctx.params
        p.boxFilterBound = bound
        ctx.set_params(p)
    ctx.upload_scan(cloud)
    ctx.crop()
    return ctx.download_cloud(0)


def getNormals(neighborRadius: float, cloud: np.ndarray, ctx: capi.Context | None = None):
    """Returns (cloud_normals, cloud) — both compacted by the NaN-normal removal, like the reference
    which mutates `cloud` in place (src/tunnel_processing.cpp:81-85).  `cloud` must be the result of
    chopCloud on the same ctx (the device copy is used; the argument is only checked for size)."""
    ctx = ctx if ctx is not None else context()
    if ctx.params.neighborRadius != neighborRadius:
        p = ctx.params
        p.neighborRadius = neighborRadius
        ctx.set_params(p)
    if not _same_cloud(ctx, cloud, 0):
        raise ValueError("getNormals: cloud is not the chopCloud result held by this Context")
    ctx.normals()
    return ctx.download_normals(1), ctx.download_cloud(1)


def getLocalFrame(cloudSize: int, weightingFactor: float, cloud_normals: np.ndarray, ctx: capi.Context | None = None):
    """-> (eigenVals[3] ascending, eigenVecs[3,3] with eigenvectors as columns); the center axis is
    eigenVecs[:, 0] (src/geometric_mapping.cpp:91-92)."""
    ctx = ctx if ctx is not None else context()
    if ctx.params.weightingFactor != weightingFactor:
        p = ctx.params
        p.weightingFactor = weightingFactor
        ctx.set_params(p)
    if cloudSize != len(cloud_normals) or not _same_cloud(ctx, cloud_normals, 1):
        raise ValueError("getLocalFrame: normals are not the getNormals result held by this Context")
    ctx.local_frame()
    f = ctx.frame()
    return f["vals"], f["vecs"]


def rvizArrow(start, end, scale, color, ns: str, id: int = 0, frame: str = FRAME_ID) -> dict:
    """Marker payload of one ARROW (src/tunnel_processing.cpp:161-205); color is pushed as (a,r,g,b)."""
    return {"header": {"frame_id": frame, "seq": 0}, "ns": ns, "id": int(id), "type": "ARROW", "action": "ADD",
            "points": [tuple(float(v) for v in start), tuple(float(v) for v in end)],
            "scale": tuple(float(v) for v in scale),
            "color": {"a": float(color[0]), "r": float(color[1]), "g": float(color[2]), "b": float(color[3])}}


def rvizNormals(leafSize: float, cloud: np.ndarray, ctx: capi.Context | None, normals: np.ndarray) -> list:
    ctx = ctx if ctx is not None else context()
    if ctx.params.voxelGridLeafSize != leafSize:
        p = ctx.params
        p.voxelGridLeafSize = leafSize
        ctx.set_params(p)
    if not _same_cloud(ctx, cloud, 1) or len(normals) != len(cloud):
        raise ValueError("rvizNormals: cloud/normals are not the getNormals results held by this Context")
    ctx.voxel()
    vox = ctx.download_voxels()
    arrows = capi.markers_normals(vox["centroids"], vox["nn_normal"], int(ctx.params.arrow_mode))
    return [rvizArrow(a["start"], a["end"], a["scale"], a["color_argb"], "normals", int(a["id"])) for a in arrows]


def rvizEigens(eigenVals: np.ndarray, eigenVecs: np.ndarray) -> list:
    f = capi.gm_frame()
    f.vals = (C.c_float * 3)(*[float(v) for v in eigenVals])
    f.vecs = (C.c_float * 9)(*[float(v) for v in np.asarray(eigenVecs, np.float32).reshape(-1)])
    arrows = capi.markers_eigen(f)
    return [rvizArrow(a["start"], a["end"], a["scale"], a["color_argb"], "eigenBasis", int(a["id"])) for a in arrows]
