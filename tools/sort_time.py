"""Time the radix sorts inside the real pipeline (neighbour-grid sort, sort-based VoxelGrid) on C1-type scans.
    python tools/sort_time.py [points] [radius]
Prints the library's own CUDA-event segment times (gm_profile_*), median over the scans, and a checksum of the outputs
(a stable sort leaves every downstream result bit-identical: compare with the line of the previous build)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from geometric_mapping_b200 import capi, synth

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    radius = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    pts = [synth.curved_tunnel(n, seed=2 + s) for s in range(3)]
    prm = capi.default_params(neighborRadius=radius, voxelGridLeafSize=0.1)
    stream = torch.cuda.Stream()
    out = {}
    with capi.Context(prm, max_points=n, max_hypotheses=64) as ctx:
        ctx.set_stream(stream.cuda_stream)
        for vm in (0, 1):
            ctx.set_voxel_mode(vm)
            seg = {}
            chk = 0
            for it in range(7):
                ctx.upload_scan(pts[it % 3])
                ctx.profile_enable(True)
                ctx.crop()
                ctx.normals()
                ctx.voxel()
                torch.cuda.synchronize()
                pr = ctx.profile_read()
                ctx.profile_enable(False)
                if it:
                    for k, v in pr.items():
                        seg.setdefault(k, []).append(v[0] / max(v[1], 1))
                c = ctx.counts()
                assert c.device_error == 0
                if it == 6:
                    cen = ctx.download_voxels(with_nn=False)["centroids"]
                    nrm = ctx.download_normals()
                    chk = int(np.nan_to_num(nrm).view(np.uint32).astype(np.uint64).sum())
                    chk ^= int(np.ascontiguousarray(cen).view(np.uint32).astype(np.uint64).sum())
            med = {k: float(np.median(v)) for k, v in seg.items()}
            out[vm] = (med, chk, c.n_cropped, c.n_valid, c.n_voxels)
    tag = "segment-form sort"
    for vm, (med, chk, nc, nv, V) in out.items():
        keys = [k for k in ("grid_sort", "voxel_sort", "grid_keys", "grid_heads", "voxel_keys", "voxel_reduce") if k in med]
        print(f"{tag} points={n} r={radius} voxel_mode={vm} cropped={nc} valid={nv} voxels={V} | " +
              " ".join(f"{k}={med[k] * 1e3:.1f}us" for k in keys) + f" | checksum={chk}")


if __name__ == "__main__":
    main()
