import os, sys
os.environ["NLAM_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
spec = synthetic.make_graph_spec(268, 238)
ei = spec["g2m_edge_index"]
torch.manual_seed(0)
net = nlb.InteractionNet(ei, 64, update_edges=False, math="tf32").to(dev)
send = torch.randn(B, 268 * 238, 64, device=dev)
rec = torch.randn(6561, 64, device=dev).unsqueeze(0).expand(B, -1, -1)
edge = torch.randn(ei.shape[1], 64, device=dev).unsqueeze(0).expand(B, -1, -1)
with torch.no_grad():
    for i in range(2):
        print("--- call", i, file=sys.stderr)
        net(send, rec, edge)
torch.cuda.synchronize()
