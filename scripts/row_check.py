"""GPU check of the streaming row-MLP kernel (tc4.cu) vs the fp64 oracle + grid-sized timings.
Compare with the previous kernel: NLAM_TC_ROW=v1 python scripts/row_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_lam_b200 as nlb
from neural_lam_b200 import ops, _lib
from oracle import reference_port as rp

dev = torch.device("cuda:0")
torch.manual_seed(0)
TF = _lib.MATH_TF32


def err(a, b):
    return (a.double().cpu() - b).abs().max().item()


def check(name, bp, shapes, res_idx=None):
    m = nlb.make_mlp(bp, layer_norm=True)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    srcs = [torch.randn(*s) for s in shapes]
    B_eff = max([s.shape[0] for s in srcs if s.dim() == 3], default=1)
    cat = torch.cat([s if s.dim() == 3 else s.unsqueeze(0).expand(B_eff, -1, -1) for s in srcs], dim=-1).double()
    params = {f"m.{k}": v.double() for k, v in m.state_dict().items()}
    want = rp.mlp(cat, params, "m", layer_norm=True)
    if res_idx is not None:
        r = srcs[res_idx]
        want = want + (r if r.dim() == 3 else r.unsqueeze(0)).double()
    m = m.to(dev)
    ds = [s.to(dev) for s in srcs]
    with torch.no_grad():
        got = ops.rowmlp(m, ds, res=ds[res_idx] if res_idx is not None else None, flags=TF)
    torch.cuda.synchronize()
    if got.dim() == 2:
        got = got.unsqueeze(0)
    print(f"{name:38s} tf32 err {err(got, want):.3e}  shape {tuple(got.shape)}", flush=True)


check("K=64 res (encoding)", [64, 64, 64], [(2, 1000, 64)], 0)
check("K=64 no res (embed)", [64, 64, 64], [(3, 300, 64)])
check("K=128 res=rec", [128, 64, 64], [(2, 777, 64), (2, 777, 64)], 0)
check("K=128 res=aggr", [128, 64, 64], [(2, 130, 64), (2, 130, 64)], 1)
check("K=128 bcast rec, no res", [128, 64, 64], [(257, 64), (3, 257, 64)])
check("K=128 many tiles per CTA", [128, 64, 64], [(3, 128 * 400 + 3, 64), (3, 128 * 400 + 3, 64)], 0)
check("K=64 many tiles per CTA", [64, 64, 64], [(2, 128 * 700 + 77, 64)], 0)
check("one row", [64, 64, 64], [(1, 1, 64)], 0)

G = 268 * 238
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in (8, 32):
    for name, bp, ns, res in (("encoding K=64 res", [64, 64, 64], 1, 0), ("node K=128 res", [128, 64, 64], 2, 0)):
        m = nlb.make_mlp(bp, layer_norm=True).to(dev)
        ds = [torch.randn(B, G, 64, device=dev) for _ in range(ns)]
        with torch.no_grad():
            for _ in range(3):
                ops.rowmlp(m, ds, res=ds[res], flags=TF)
            ts = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(True), torch.cuda.Event(True)
                s.record(); ops.rowmlp(m, ds, res=ds[res], flags=TF); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        t = sorted(ts)[len(ts) // 2]
        nbytes = B * G * 256 * (ns + 1)
        print(f"{name} B={B}: {t*1e3:.1f} us -> {nbytes/t/1e6:.0f} GB/s ({os.environ.get('NLAM_TC_ROW', 'v2')})", flush=True)
# narrow inputs / narrow outputs
def check_generic(name, bp, shapes, ln):
    m = nlb.make_mlp(bp, layer_norm=ln)
    srcs = [torch.randn(*s) for s in shapes]
    B_eff = max([s.shape[0] for s in srcs if s.dim() == 3], default=1)
    cat = torch.cat([s if s.dim() == 3 else s.unsqueeze(0).expand(B_eff, -1, -1) for s in srcs], dim=-1).double()
    want = rp.mlp(cat, {f"m.{k}": v.double() for k, v in m.state_dict().items()}, "m", layer_norm=ln)
    m = m.to(dev)
    with torch.no_grad():
        got = ops.rowmlp(m, [s.to(dev) for s in srcs], flags=TF)
    torch.cuda.synchronize()
    print(f"{name:38s} tf32 err {err(got, want):.3e}  shape {tuple(got.shape)}", flush=True)


check_generic("embedder 17|17|18|4 (400 rows)", [56, 64, 64], [(2, 400, 17), (2, 400, 17), (2, 400, 18), (400, 4)], True)
check_generic("embedder many tiles", [56, 64, 64], [(3, 128 * 300 + 8, 17), (3, 128 * 300 + 8, 17), (3, 128 * 300 + 8, 18), (128 * 300 + 8, 4)], True)
check_generic("output_map 64->64->17", [64, 64, 17], [(2, 500, 64)], False)
check_generic("output_map many tiles", [64, 64, 17], [(3, 128 * 300 + 5, 64)], False)
check_generic("narrow in, narrow out 20->64->9", [20, 64, 9], [(2, 1000, 13), (2, 1000, 7)], False)

# fused output_map + step epilogue
Bq, Gq, D = 3, 128 * 40 + 12, 17
m = nlb.make_mlp([64, 64, D], layer_norm=False)
x, prev, bnd = torch.randn(Bq, Gq, 64), torch.randn(Bq, Gq, D), torch.randn(Bq, Gq, D)
mask = (torch.rand(Gq) < 0.3).float()
std, mean = torch.rand(D) + 0.5, torch.randn(D)
y = rp.mlp(x.double(), {f"m.{k}": v.double() for k, v in m.state_dict().items()}, "m", layer_norm=False)
pred = prev.double() + (y * std.double() + mean.double())
want_b = mask.double()[None, :, None] * bnd.double() + (1 - mask.double()[None, :, None]) * pred
m = m.to(dev)
with torch.no_grad():
    g1 = ops.rowmlp_step(m, x.to(dev), prev.to(dev), None, None, std.to(dev), mean.to(dev))
    g2 = ops.rowmlp_step(m, x.to(dev), prev.to(dev), bnd.to(dev), mask.to(dev), std.to(dev), mean.to(dev))
torch.cuda.synchronize()
print(f"fused step epilogue: no boundary err {err(g1, pred):.3e}, boundary err {err(g2, want_b):.3e}", flush=True)

for B in (8,):
    m = nlb.make_mlp([56, 64, 64], layer_norm=True).to(dev)
    ds = [torch.randn(B, G, 17, device=dev), torch.randn(B, G, 17, device=dev), torch.randn(B, G, 18, device=dev), torch.randn(G, 4, device=dev)]
    mo = nlb.make_mlp([64, 64, 17], layer_norm=False).to(dev)
    xo, pv = torch.randn(B, G, 64, device=dev), torch.randn(B, G, 17, device=dev)
    sd, mn = torch.rand(17, device=dev), torch.randn(17, device=dev)
    for name, fn in (("grid embedder", lambda: ops.rowmlp(m, ds, flags=TF)), ("output_map", lambda: ops.rowmlp(mo, [xo], flags=TF)),
                     ("output_map + epilogue", lambda: ops.rowmlp_step(mo, xo, pv, None, None, sd, mn))):
        with torch.no_grad():
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(True), torch.cuda.Event(True)
                s.record(); fn(); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        print(f"{name} B={B}: {sorted(ts)[5]*1e3:.1f} us ({os.environ.get('NLAM_TC_ROW', 'v2')})", flush=True)
print("row_check done")
