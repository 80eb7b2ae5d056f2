"""Node-partitioned GraphLAM rollout: STRONG scaling of one forecast on N GPUs.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_partition.py --grid 1024 --steps 10

Every rank owns a strip of grid / mesh nodes, exchanges boundary sender rows (NCCL grouped
send/recv) before each of the 6 InteractionNet calls of a step, and the rollout keeps each rank's
own grid rows local.  Prints one JSON line on rank 0 (time = max over ranks, CUDA events).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from neural_lam_b200 import dist as nd, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1024)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
spec = synthetic.make_graph_spec(args.grid, args.grid)
ds = synthetic.SyntheticDatastore(spec)
torch.manual_seed(42)
model = nd.PartitionedGraphLAM(ds, spec, rank, world, hidden_dim=64, processor_layers=4, math="auto").to(dev).eval()
sl = model.own_grid_slice()
G_own = sl.stop - sl.start
B = args.batch
g = torch.Generator().manual_seed(123 + rank)
prev, pprev = torch.randn(B, G_own, 17, generator=g).to(dev), torch.randn(B, G_own, 17, generator=g).to(dev)
forc = torch.randn(B, G_own, 18, generator=g).to(dev)
bmask = model.boundary_mask_local.to(dev)


def step(prev, pprev):
    new, _ = model(prev, pprev, forc)
    new = bmask * prev + (1 - bmask) * new  # boundary rows keep the (synthetic) truth
    return new, prev


with torch.no_grad():
    for _ in range(args.warmup):
        prev, pprev = step(prev, pprev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(args.steps):
        prev, pprev = step(prev, pprev)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    sent, recv = model.halo_bytes_per_step(B)
    print(json.dumps({"metric": "forecast-steps/sec, node-partitioned (strong scaling)", "grid": f"{args.grid}x{args.grid}",
                      "n_gpus": world, "batch": B, "steps": args.steps, "value": B * args.steps / (ms.item() * 1e-3),
                      "ms_per_step": ms.item() / args.steps, "halo_bytes_sent_per_step_rank0": sent,
                      "halo_bytes_recv_per_step_rank0": recv, "finite": bool(torch.isfinite(prev).all())}))
if world > 1:
    dist.destroy_process_group()
