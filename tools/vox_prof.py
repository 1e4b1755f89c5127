import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from geometric_mapping_b200 import capi, synth
n=1_000_000
pts=synth.curved_tunnel(n,seed=2); d=torch.from_numpy(pts).cuda()
ctx=capi.Context(capi.default_params(neighborRadius=0.05,voxelGridLeafSize=0.1),max_points=n,max_hypotheses=512)
ctx.set_scan_device(d.data_ptr(),n); ctx.crop(); ctx.normals()
for _ in range(4): ctx.voxel()
ctx.synchronize(); print(ctx.counts().n_voxels)
