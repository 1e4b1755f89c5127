"""Time the k-NN normals kernel on one C1 scan (tunables come from the environment: GM_KNN_BLOCK, GM_KNN_CARVEOUT).
    GM_KNN_BLOCK=128 GM_KNN_CARVEOUT=50 python tools/knn_time.py [k] [cell] [cap]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from geometric_mapping_b200 import capi, synth

    k = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.07
    cap = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
    n = 1_000_000
    pts = synth.curved_tunnel(n, seed=2)
    prm = capi.default_params(neighborRadius=cell, voxelGridLeafSize=0.1)
    stream = torch.cuda.Stream()
    with capi.Context(prm, max_points=n, max_hypotheses=64) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_knn(k, max_radius=cap)
        ctx.upload_scan(pts)
        ctx.crop()
        ctx.profile_enable(True)
        ts, ks = [], []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.normals()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            ks.append(ctx.profile_read()["normals"][0])
        c = ctx.counts()
        assert c.device_error == 0
        nrm = ctx.download_normals()
        chk = int(np.nan_to_num(nrm).view(np.uint32).astype(np.uint64).sum())
    print(f"block={os.environ.get('GM_KNN_BLOCK', 'default')} carveout={os.environ.get('GM_KNN_CARVEOUT', 'default')} k={k} cell={cell} cap={cap} "
          f"kernel ms: min {min(ks[1:]):.3f} median {sorted(ks[1:])[2]:.3f}; grid+normals+compact ms: min {min(ts[1:]):.3f}  n_valid={c.n_valid} checksum={chk}")


if __name__ == "__main__":
    main()
