import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from geometric_mapping_b200 import capi, synth
n=10_000_000
pts=synth.curved_tunnel(n,seed=8)
out={}
with capi.Context(capi.default_params(neighborRadius=0.02,voxelGridLeafSize=0.1),max_points=n,max_hypotheses=1024) as ctx:
    for mode in (0,1):
        ctx.set_voxel_mode(mode)
        ctx.upload_scan(pts); ctx.crop(); ctx.normals(); ctx.voxel()
        keys,assign,_=ctx.download_voxel_assignment(); vox=ctx.download_voxels()
        out[mode]=(keys,assign,vox["keys"],vox["counts"],vox["centroids"],vox["nn_index"])
        print("mode",mode,"valid",ctx.counts().n_valid,"voxels",ctx.counts().n_voxels,"err",ctx.counts().device_error, flush=True)
    same=all(np.array_equal(np.asarray(a).view(np.uint8),np.asarray(b).view(np.uint8)) for a,b in zip(out[0],out[1]))
    print("dense == sort at 10M:",same)
    # centroid of each voxel == mean of its members in double (tolerance 2e-6)
    k,a,vk,vc,cen,_=out[0]
    cloud=ctx.download_cloud(1).astype(np.float64)
    s=np.zeros((len(vk),3)); np.add.at(s,a,cloud[:,:3]); m=s/vc[:,None]
    print("max |centroid - double mean|:",np.abs(m-cen[:,:3]).max(), "counts sum ok:", int(vc.sum())==len(cloud))
