import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from geometric_mapping_b200 import capi, synth
n = 1_000_000
pts = synth.curved_tunnel(n, seed=2)
d = torch.from_numpy(pts).cuda()
hp = torch.from_numpy(pts).pin_memory()
params = capi.default_params(neighborRadius=0.05, voxelGridLeafSize=0.1)
ctx = capi.Context(params, max_points=n, max_hypotheses=4096)
ctx.set_scan_device(d.data_ptr(), n); ctx.crop(); ctx.normals(); nv = ctx.counts().n_valid
ps, cs = synth.sample_indices(nv, 512, 3, seed=3), synth.sample_indices(nv, 512, 2, seed=4)
for _ in range(5):
    ctx.set_scan_device(d.data_ptr(), n); ctx.process_scan(ps, cs)
ctx.synchronize()
# host cost: enqueue 20 scans without waiting
t0 = time.perf_counter()
for _ in range(20):
    ctx.set_scan_device(d.data_ptr(), n); ctx.process_scan(ps, cs)
t1 = time.perf_counter()
ctx.synchronize()
t2 = time.perf_counter()
print("host enqueue per scan %.3f ms ; total per scan %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
# H2D / D2H of 16 MB pinned
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = torch.empty_like(hp).pin_memory()
for name, fn in (("H2D", lambda: d.copy_(hp, non_blocking=True)), ("D2H", lambda: out.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); e0.record(); [fn() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    print(name, "16 MB: %.3f ms (%.1f GB/s)" % (e0.elapsed_time(e1) / 10, 16e-3 / (e0.elapsed_time(e1) / 10 * 1e-3)))
# pure host cost: enqueue ONE scan onto an idle GPU (no blocking on earlier work)
ts = []
for _ in range(10):
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.set_scan_device(d.data_ptr(), n); ctx.process_scan(ps, cs)
    ts.append(time.perf_counter() - t0)
ctx.synchronize()
print("host cost of one gm_process_scan on an idle GPU: median %.3f ms, min %.3f ms (%d launches)" % (np.median(ts) * 1e3, min(ts) * 1e3, 42))
