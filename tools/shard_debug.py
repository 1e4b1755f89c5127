import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from geometric_mapping_b200 import capi, synth
from geometric_mapping_b200 import distributed as D
n = 1_000_000
pts = synth.curved_tunnel(n, seed=2)
d = torch.from_numpy(pts).cuda()
ctx = capi.Context(capi.default_params(neighborRadius=0.05, voxelGridLeafSize=0.1), max_points=n, max_hypotheses=4096)
ctx.set_scan_device(d.data_ptr(), n); ctx.crop(); ctx.normals(); nv = ctx.counts().n_valid
H = 2048
sc = synth.sample_indices(nv, H, 2, seed=4)
ctx.ransac(1, sc)
_, _, full = ctx.download_hypotheses(1, H)
print("full argmax", int(np.argmax(full)), int(full.max()))
for W in (2, 8):
    best = 0
    for r in range(W):
        lo, hi = D.shard_range(H, r, W)
        ctx.ransac(1, sc, lo, hi)
        _, _, c = ctx.download_hypotheses(1, H)
        ok = np.array_equal(c[lo:hi], full[lo:hi]) and (c[:lo] == -1).all() and (c[hi:] == -1).all()
        key_t = torch.as_tensor(D.DeviceKey(ctx.ransac_key_device_ptr(1)), device="cuda")
        k = int(key_t.item())
        print("W", W, "rank", r, "range", lo, hi, "counts ok", ok, "local best", D.unpack_key(k), "expected", (int(c.max()), int(np.argmax(c))))
        best = max(best, k)
    print("W", W, "global", D.unpack_key(best))
