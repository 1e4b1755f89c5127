"""A/B of the radius-normals kernels on the C1 scan (and a 200k frame, and r = 0.15): ms per launch from the library's own
per-segment CUDA events.  GM_NORMALS_KERNEL: 0 = direct per-lane loop, 1 = staged (768 candidates / warp), 2 = staged (1024)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geometric_mapping_b200 import capi, synth  # noqa: E402

out = {}
for name, n, radius in (("C1_1M_r0.05", 1_000_000, 0.05), ("frame_200k_r0.05", 200_000, 0.05), ("C1_1M_r0.1", 1_000_000, 0.1)):
    pts = synth.curved_tunnel(n, seed=2)
    ref = None
    for kern in (0, 1, 2):
        os.environ["GM_NORMALS_KERNEL"] = str(kern)
        with capi.Context(capi.default_params(neighborRadius=radius), max_points=n, max_hypotheses=64) as ctx:
            ctx.upload_scan(pts)
            for _ in range(3):
                ctx.crop()
                ctx.normals()
            ctx.profile_enable(True)
            reps = 20
            for _ in range(reps):
                ctx.crop()
                ctx.normals()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            cnt = ctx.download_neighbor_counts()
            nrm = ctx.download_normals(0)
            cand, nbr = ctx.search_stats()
            assert ctx.counts().device_error == 0
        if ref is None:
            ref = (cnt, nrm)
        else:
            assert np.array_equal(cnt, ref[0]), "neighbour counts differ between kernels"
            fin = np.isfinite(nrm[:, 0])
            assert np.array_equal(fin, np.isfinite(ref[1][:, 0]))
            ang = np.arccos(np.clip(np.abs((nrm[fin, :3] * ref[1][fin, :3]).sum(1)), -1, 1))
            out[f"{name}/kernel{kern}_vs_0_angle_p99"] = float(np.percentile(ang, 99))
        out[f"{name}/kernel{kern}_ms"] = prof["normals"][0] / reps
        out[f"{name}/candidates_per_point"] = cand / max(len(cnt), 1)
        out[f"{name}/neighbours_per_point"] = nbr / max(len(cnt), 1)
print(json.dumps(out, indent=1))
