"""Times gm_ransac (hypothesis generation + inlier counting + argmax) alone: tile-culled (mode 0) vs
brute force (mode 1), plane and cylinder, at a few hypothesis counts.  CUDA events on the ctx stream.
    python tools/count_bench.py [points] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geometric_mapping_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
radius = 0.05 if n <= 2_000_000 else 0.02
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
pts = synth.curved_tunnel(n, seed=2)
d = torch.from_numpy(pts).cuda()
ctx = capi.Context(capi.default_params(neighborRadius=radius, voxelGridLeafSize=0.1), max_points=n, max_hypotheses=8192)
ctx.set_stream(stream.cuda_stream)
ctx.set_scan_device(d.data_ptr(), n); ctx.crop(); ctx.normals(); nv = ctx.counts().n_valid
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for H in (512, 2048):
    smp = {0: synth.sample_indices(nv, H, 3, seed=3), 1: synth.sample_indices(nv, H, 2, seed=4)}
    ref = {}
    for mode in (1, 0):
        ctx.set_count_mode(mode)
        for kind in (0, 1):
            ts = []
            for r in range(reps + 3):
                flush.zero_()
                e0.record(stream); ctx.ransac(kind, smp[kind]); e1.record(stream); torch.cuda.synchronize()
                if r >= 3: ts.append(e0.elapsed_time(e1))
            cnt = ctx.download_hypotheses(kind, H)[2]
            if mode == 1: ref[kind] = cnt
            same = np.array_equal(cnt, ref[kind])
            flop = nv * H * (6 if kind == 0 else 16)
            print(f"n={n} H={H} mode={'tiles' if mode == 0 else 'brute'} kind={'plane' if kind == 0 else 'cyl'}: "
                  f"{1e3 * np.median(ts):8.1f} us (min {1e3 * min(ts):.1f})  {flop / np.median(ts) / 1e9:7.1f} algorithmic TFLOP/s  counts_equal={same}", flush=True)
