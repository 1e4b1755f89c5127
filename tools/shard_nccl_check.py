"""torchrun --nproc-per-node N tools/shard_nccl_check.py : every rank must end with the single-range winner."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from geometric_mapping_b200 import capi, synth
from geometric_mapping_b200 import distributed as D
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
dist.init_process_group("nccl", device_id=dev)
n = 300_000
pts = synth.curved_tunnel(n, seed=2)
d = torch.from_numpy(pts).to(dev)
ctx = capi.Context(capi.default_params(neighborRadius=0.08), max_points=n, max_hypotheses=4096)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_scan_device(d.data_ptr(), n); ctx.crop(); ctx.normals(); nv = ctx.counts().n_valid
key_t = torch.zeros(2, dtype=torch.int64, device=dev)
ok = True
H = 1024
smp = {0: synth.sample_indices(nv, H, 3, seed=3), 1: synth.sample_indices(nv, H, 2, seed=4)}
want = {}
for kind in (0, 1):
    ctx.ransac(kind, smp[kind])
    _, _, full = ctx.download_hypotheses(kind, H)
    want[kind] = (int(np.argmax(full)), int(full.max()))
D.sharded_ransac_pair(ctx, smp[0], smp[1], rank, world, key_t)
for kind in (0, 1):
    m = ctx.model(kind)
    got = (m["best_id"], m["best_count"])
    owner = want[kind][0] // ((H + world - 1) // world)
    print(f"rank {rank} kind {kind} want {want[kind]} (owned by rank {owner}) got {got}", flush=True)
    ok = ok and got == want[kind]
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("ALL OK" if int(t.item()) == 1 else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
