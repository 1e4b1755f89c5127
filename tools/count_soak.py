"""Soak: tile-culled vs brute-force inlier counts on many different scans / radii / thresholds (exact equality).
    python tools/count_soak.py [scans]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geometric_mapping_b200 import capi, synth

scans = int(sys.argv[1]) if len(sys.argv) > 1 else 24
g = np.random.Generator(np.random.Philox(77))
bad = 0
for k in range(scans):
    n = int(g.choice([200_000, 500_000, 1_000_000]))
    radius = float(g.choice([0.04, 0.05, 0.08]))
    tau = float(g.choice([0.01, 0.05, 0.15]))
    arc = float(g.choice([30.0, 50.0, 1000.0]))
    pts = synth.curved_tunnel(n, seed=500 + k, arc_radius=arc, noise=float(g.choice([0.0, 0.02, 0.05])), outlier_frac=float(g.choice([0.0, 0.01, 0.1])))
    with capi.Context(capi.default_params(neighborRadius=radius, ransacThreshold=tau), max_points=n, max_hypotheses=4096) as ctx:
        ctx.upload_scan(pts); ctx.crop(); ctx.normals()
        nv = ctx.counts().n_valid
        H = int(g.choice([300, 1024, 2500]))
        ps, cs = synth.sample_indices(nv, H, 3, seed=k), synth.sample_indices(nv, H, 2, seed=1000 + k)
        res = {}
        for mode in (0, 1):
            ctx.set_count_mode(mode)
            ctx.ransac(0, ps); ctx.ransac(1, cs)
            res[mode] = (ctx.download_hypotheses(0, H)[2], ctx.download_hypotheses(1, H)[2])
        ok = np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and ctx.counts().device_error == 0
        bad += 0 if ok else 1
        print(f"scan {k}: n={n} r={radius} tau={tau} arc={arc} H={H} valid={nv} best plane {res[1][0].max()} cyl {res[1][1].max()} equal={ok}", flush=True)
print("done, mismatches:", bad)
