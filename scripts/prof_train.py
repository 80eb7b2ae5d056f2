"""Per-kernel breakdown of one GraphLAM training step (forward + backward) at the MEPS size, B=4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_lam_b200 import models, ops, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
spec = synthetic.make_graph_spec(268, 238)
ds = synthetic.SyntheticDatastore(spec, d_state=17, d_forcing=18, d_static=4, boundary_width=10)
torch.manual_seed(42)
m = models.GraphLAM(ds, spec, hidden_dim=64, processor_layers=4).to(dev)
G = m.num_grid_nodes
prev, pprev, forc, tgt = (torch.randn(B, G, d, device=dev) for d in (17, 17, 18, 17))
def step():
    m.zero_grad(set_to_none=True)
    p, _ = m(prev, pprev, forc)
    ((p - tgt) ** 2).mean().backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
with ops.profile_launches() as prof:
    step()
agg = {}
for n, us, nb in prof.rows:
    a = agg.setdefault(n, [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += nb
tot = sum(v[1] for v in agg.values())
print(f"train step B={B}: {len(prof.rows)} launches of the library, {tot/1e3:.2f} ms in them")
for n, (c, us, nb) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {n:34s} x{c:4d} {us/1e3:8.2f} ms  {nb/us/1e3 if us else 0:8.1f} GB/s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
print("wall per step (ms):", e0.elapsed_time(e1) / 3)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof2:
    step()
    torch.cuda.synchronize()
print(prof2.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
print(prof2.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
