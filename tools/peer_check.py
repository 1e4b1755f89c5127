"""Multi-GPU check of the peer-memory collectives (csrc/gm_comm.cuh).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py

1. hypothesis-sharded RANSAC over the mailboxes (Context.ransac_sharded) against the unsharded run on the same scan:
   identical winners, counts and refined coefficients on every rank; timing against the torch.distributed (NCCL) form.
2. map slabs: bounding-box MIN/MAX on the device against torch.distributed's; the all-reduced frame against the frame a
   single context computes for the whole map (1e-5 |S|); the gathered voxels against the whole-map VoxelGrid (bit-exact).
Rank 0 prints one JSON line; exit code != 0 on any mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from geometric_mapping_b200 import capi, synth
    from geometric_mapping_b200 import distributed as gmd

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    comm = gmd.connect_peers()
    out = {"world": world}
    n = int(os.environ.get("PEER_POINTS", 1_000_000))
    pts = synth.curved_tunnel(n, seed=2)
    prm = capi.default_params(neighborRadius=0.05, voxelGridLeafSize=0.1)
    ok = True
    with capi.Context(prm, max_points=n, max_hypotheses=65536) as ctx, capi.Context(prm, max_points=n, max_hypotheses=65536) as ref:
        for c in (ctx, ref):
            c.set_stream(stream.cuda_stream)
            c.upload_scan(pts)
            c.crop()
            c.normals()
        ctx.set_comm(comm)
        nv = ctx.counts().n_valid
        key_t = torch.zeros(2, dtype=torch.int64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for H in (4096, 65536):
            ps, cs = synth.sample_indices(nv, H // 2, 3, seed=3), synth.sample_indices(nv, H // 2, 2, seed=4)
            ref.ransac_pair(ps, cs)
            ref.ransac_select_pair()
            ctx.ransac_sharded(ps, cs)
            for kind in (0, 1):
                a, b = ctx.model(kind), ref.model(kind)
                same = (a["best_id"] == b["best_id"] and a["best_count"] == b["best_count"] and a["refit_count"] == b["refit_count"]
                        and np.array_equal(a["coef"].view(np.uint32), b["coef"].view(np.uint32)))
                ok &= bool(same)
                out[f"H{H}_kind{kind}"] = {"same": bool(same), "best": [a["best_id"], a["best_count"]], "ref_best": [b["best_id"], b["best_count"]]}
            reps = 20
            times = {}
            for name, fn in (("peer", lambda: ctx.ransac_sharded(ps, cs)),
                             ("nccl", lambda: gmd.sharded_ransac_pair(ctx, ps, cs, rank, world, key_t))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(reps):
                    fn()
                e1.record(stream)
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times[name] = float(t.item())
            out[f"H{H}_ms_per_round"] = times
        assert ctx.counts().device_error == 0

    # ---- map slabs ------------------------------------------------------------------------------------------
    nm = int(os.environ.get("PEER_MAP_POINTS", 2_000_000))
    bound, leaf, rad = 60.0, 0.1, 0.05
    full = synth.tunnel_map(nm, seed=4, bound=bound)
    mprm = capi.default_params(boxFilterBound=bound, neighborRadius=rad, voxelGridLeafSize=leaf)
    fin = np.isfinite(full[:, :3]).all(1)
    map_box = gmd.robust_box(full[fin, :3], q=0.02, margin=2.0)
    lo, hi = gmd.slab_cuts(full[:, 0], world, leaf)[rank]
    slab = gmd.slab_select(full, 0, lo, hi, halo=1.01 * rad)
    with capi.Context(mprm, max_points=max(len(slab), 1), max_hypotheses=64) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_comm(comm)
        ctx.set_owned_range(0, lo, hi)
        ctx.set_grid_box(map_box[0], map_box[1])
        ctx.set_voxel_bbox_hint(map_box[0], map_box[1])
        ctx.upload_scan(slab)
        ctx.crop()
        ctx.normals()
        mn, mx = ctx.voxel_bbox()
        gmn, gmx = gmd.allreduce_bbox(mn, mx)
        ctx.allreduce_voxel_bbox()
        dmn, dmx = ctx.voxel_bbox()
        same_box = bool(np.array_equal(dmn, np.asarray(gmn, np.float32)) and np.array_equal(dmx, np.asarray(gmx, np.float32)))
        ok &= same_box
        ctx.voxel()
        ctx.local_frame()
        ctx.allreduce_frame()
        fr = ctx.frame()
        keys, cen, cnt = gmd.gather_slab_voxels(ctx)
        assert ctx.counts().device_error == 0
    out["bbox_equal_to_torch_allreduce"] = same_box
    if rank == 0:
        with capi.Context(mprm, max_points=nm, max_hypotheses=64) as whole:
            whole.set_stream(stream.cuda_stream)
            whole.set_grid_box(map_box[0], map_box[1])
            whole.upload_scan(full)
            whole.crop()
            whole.normals()
            whole.voxel()
            whole.local_frame()
            wf = whole.frame()
            wv = whole.download_voxels(with_nn=False)
        sn = float(np.abs(wf["scatter"]).max())
        d = float(np.abs(fr["scatter"].astype(np.float64) - wf["scatter"]).max() / sn)
        ax = float(np.arccos(min(1.0, abs(float(np.dot(fr["vecs"][:, 0], wf["vecs"][:, 0]))))))
        vox_same = bool(np.array_equal(keys, wv["keys"]) and np.array_equal(cnt, wv["counts"]) and cen.tobytes() == wv["centroids"].tobytes())
        out["frame_scatter_rel_diff_vs_single_context"] = d
        out["frame_axis_angle_rad"] = ax
        out["gathered_voxels_equal_whole_map"] = vox_same
        out["voxels"] = int(len(keys))
        ok &= d <= 1e-5 and ax <= 1e-4 and vox_same
    okt = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    out["ok"] = bool(okt.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
