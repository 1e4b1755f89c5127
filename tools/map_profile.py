"""Per-segment times of one whole-map pass (configs[4] on one GPU): python tools/map_profile.py [points]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geometric_mapping_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = synth.tunnel_map(n, seed=4, bound=60.0)
fin = np.isfinite(pts[:, :3]).all(1)
p = capi.default_params(boxFilterBound=60.0, neighborRadius=0.05, voxelGridLeafSize=0.1)
ctx = capi.Context(p, max_points=n, max_hypotheses=4096)
from geometric_mapping_b200 import distributed as D
ctx.set_grid_box(*D.robust_box(pts[fin, :3]))
d = torch.from_numpy(pts).cuda()
def go(sp, sc):
    ctx.set_scan_device(d.data_ptr(), n); ctx.crop(); ctx.normals(); ctx.voxel(); ctx.local_frame()
    for k, s in ((0, sp), (1, sc)):
        if s is not None: ctx.ransac(k, s); ctx.ransac_select(k)
    ctx.label(); ctx.axis_polyline(); ctx.compress()
go(None, None)
nv = ctx.counts().n_valid
sp, sc = synth.sample_indices(nv, 512, 3, seed=3), synth.sample_indices(nv, 512, 2, seed=4)
go(sp, sc)
ctx.profile_enable(True)
for _ in range(3): go(sp, sc)
prof = ctx.profile_read()
c = ctx.counts()
print("n_valid", c.n_valid, "cells", c.n_cells, "voxels", c.n_voxels, "err", c.device_error)
tot = 0
for k, (ms, calls) in prof.items():
    if calls: print(f"{k:14s} {ms / 3:8.3f} ms  ({calls // 3} calls/pass)"); tot += ms / 3
print("sum", tot)
print("plane", ctx.model(0)["best_count"], "cyl", ctx.model(1)["best_count"])
