"""Run the MEPS m2m InteractionNet layer a few times (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
math = sys.argv[2] if len(sys.argv) > 2 else "tf32"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
spec = synthetic.make_graph_spec(268, 238)
ei = spec["m2m_edge_index"]
torch.manual_seed(0)
net = nlb.InteractionNet(ei, 64, math=math).to(dev)
mesh = torch.randn(B, 6561, 64, device=dev)
edge = torch.randn(B, ei.shape[1], 64, device=dev)
with torch.no_grad():
    for _ in range(iters):
        r, e = net(mesh, mesh, edge)
    torch.cuda.synchronize()
    # graph-replayed timing (no Python/launch overhead between kernels)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        net(mesh, mesh, edge)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        r, e = net(mesh, mesh, edge)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
print("graph replay ms:", [round(t, 4) for t in ts])
