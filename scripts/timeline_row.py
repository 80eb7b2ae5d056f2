import sys, os
os.environ["NLAM_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N = 8, 63784
enc = nlb.make_mlp([64, 64, 64]).to(dev)
x = torch.randn(B, N, 64, device=dev)
emb = nlb.make_mlp([56, 64, 64]).to(dev)
srcs = [torch.randn(B, N, 17, device=dev), torch.randn(B, N, 17, device=dev), torch.randn(B, N, 18, device=dev), torch.randn(N, 4, device=dev)]
om = nlb.make_mlp([64, 64, 17], layer_norm=False).to(dev)
with torch.no_grad():
    for name, fn in (("encoding_grid_mlp", lambda: ops.rowmlp(enc, [x], res=x, flags=_lib.MATH_TF32)),
                     ("grid_embedder", lambda: ops.rowmlp(emb, srcs, flags=_lib.MATH_TF32)),
                     ("output_map", lambda: ops.rowmlp(om, [x], flags=_lib.MATH_TF32))):
        for i in range(2):
            print("---", name, i, file=sys.stderr)
            fn()
torch.cuda.synchronize()
