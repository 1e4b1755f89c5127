"""One context, a few whole scans with plain stream launches (no graph, no stream fork): the program ncu is pointed at.

    python tools/profile_scan.py [--points N] [--radius R] [--hyp H] [--scans K] [--voxel-mode 0|1] [--knn K] [--count-mode 0|1]

ncu launch list (every kernel of one steady-state scan, 34 launches):
    ncu --metrics gpu__time_duration.sum --clock-control none -s <34*(K-1)+1> -c 34 --csv --log-file launches.csv python tools/profile_scan.py
ncu full capture of the same launches:
    ncu --set full --clock-control none --import-source on -s <...> -c 34 -o prof python tools/profile_scan.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GM_GRAPH", "0")
os.environ.setdefault("GM_SERIAL", "1")
from geometric_mapping_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--radius", type=float, default=0.05)
ap.add_argument("--leaf", type=float, default=0.1)
ap.add_argument("--hyp", type=int, default=1024)
ap.add_argument("--scans", type=int, default=4)
ap.add_argument("--voxel-mode", type=int, default=0)
ap.add_argument("--count-mode", type=int, default=0)
ap.add_argument("--knn", type=int, default=0)
ap.add_argument("--knn-cap", type=float, default=0.0)
a = ap.parse_args()
pts = synth.curved_tunnel(a.points, seed=2)
prm = capi.default_params(neighborRadius=a.radius, voxelGridLeafSize=a.leaf)
with capi.Context(prm, max_points=a.points, max_hypotheses=max(a.hyp, 16)) as ctx:
    ctx.set_voxel_mode(a.voxel_mode)
    ctx.set_count_mode(a.count_mode)
    if a.knn:
        ctx.set_knn(a.knn, max_radius=a.knn_cap)
    ctx.upload_scan(pts)
    ctx.crop()
    ctx.normals()
    nv = ctx.counts().n_valid
    ps = synth.sample_indices(nv, a.hyp // 2, 3, seed=3)
    cs = synth.sample_indices(nv, a.hyp - a.hyp // 2, 2, seed=4)
    ctx.reset_launch_count()
    for _ in range(a.scans):
        ctx.upload_scan(pts)
        ctx.process_scan(ps, cs)
    c = ctx.counts()
    assert c.device_error == 0
    print(f"scans {a.scans} launches {ctx.launch_count} ({ctx.launch_count // a.scans} per scan incl. begin_scan) n_valid {c.n_valid} voxels {c.n_voxels}")
