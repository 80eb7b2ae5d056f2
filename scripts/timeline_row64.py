import sys, os
os.environ["NLAM_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import ops, _lib
B = 8
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
G = 268 * 238
m = nlb.make_mlp([64 * ns, 64, 64], layer_norm=True).to(dev)
ds = [torch.randn(B, G, 64, device=dev) for _ in range(ns)]
with torch.no_grad():
    for i in range(2):
        print("--- call", i, file=sys.stderr)
        ops.rowmlp(m, ds, res=ds[0], flags=_lib.MATH_TF32)
torch.cuda.synchronize()
