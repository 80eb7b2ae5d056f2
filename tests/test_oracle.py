"""Pin the CPU oracle (oracle/reference_port.py):
  * against the golden vectors generated from the reference's own source,
  * against the reference source directly when /root/reference is present,
  * and run the reference's own tests/test_gnn_layers.py sections A-H against the
    reference source + PyG stand-in (validates the stand-in itself).
"""
import pytest
import torch

from oracle import load_reference, reference_port as rp
from golden_util import load_golden_cases

CASES = load_golden_cases()


def _run_oracle(case, dtype=torch.float32, requires_grad=False):
    t = case.t
    send = t["send"].to(dtype).clone().requires_grad_(requires_grad)
    rec = send if case.same else t["rec"].to(dtype).clone().requires_grad_(requires_grad)
    edge = t["edge"].to(dtype).clone().requires_grad_(requires_grad)
    params = {k: v.to(dtype).clone().requires_grad_(requires_grad) for k, v in case.params.items()}
    kw = dict(case.kwargs)
    ue = kw.pop("update_edges", True)
    rec_out, edge_out, aggr, msg = rp.interaction_net(
        params, t["edge_index"], send, rec, edge, propagation=case.propagation,
        update_edges=ue, return_internals=True, **kw,
    )
    return dict(send=send, rec=rec, edge=edge, params=params, rec_out=rec_out,
                edge_out=edge_out if ue else None)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_matches_golden_forward(case):
    out = _run_oracle(case)
    # same op sequence, same dtype, CPU: differences only from summation order
    torch.testing.assert_close(out["rec_out"], case.t["rec_out"], rtol=1e-5, atol=1e-5)
    if out["edge_out"] is not None:
        torch.testing.assert_close(out["edge_out"], case.t["edge_out"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_fp64_close_to_golden(case):
    out = _run_oracle(case, torch.float64)
    torch.testing.assert_close(out["rec_out"].float(), case.t["rec_out"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_matches_golden_backward(case):
    out = _run_oracle(case, requires_grad=True)
    loss = (out["rec_out"] * case.t["w_rec"]).sum()
    if out["edge_out"] is not None:
        loss = loss + (out["edge_out"] * case.t["w_edge"]).sum()
    loss.backward()
    torch.testing.assert_close(out["send"].grad, case.t["g_send"] , rtol=1e-4, atol=1e-4)
    if not case.same:
        torch.testing.assert_close(out["rec"].grad, case.t["g_rec"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["edge"].grad, case.t["g_edge"], rtol=1e-4, atol=1e-4)
    for k, g in case.gparams.items():
        torch.testing.assert_close(out["params"][k].grad, g, rtol=1e-4, atol=1e-4, msg=lambda m: f"{k}: {m}")


needs_ref = pytest.mark.skipif(not load_reference.available(), reason="/root/reference not present (GPU box)")


@needs_ref
def test_reference_own_tests_pass_on_standin():
    """Sections A-H of the reference's tests/test_gnn_layers.py, exec'd against the
    reference classes loaded behind the PyG stand-in: validates the stand-in."""
    ref = load_reference.load()
    ns = {"InteractionNet": ref.InteractionNet, "PropagationNet": ref.PropagationNet, "__name__": "ref_tests"}
    exec(compile(load_reference.reference_test_source(), "ref_test_gnn_layers", "exec"), ns)
    ran = 0
    for name, obj in list(ns.items()):
        if isinstance(obj, type) and name.startswith("Test"):
            inst = obj()
            for m in dir(inst):
                if m.startswith("test_"):
                    getattr(inst, m)()
                    ran += 1
    assert ran == 25


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_reference_source_random(seed):
    """Fresh random graphs / weights, reference source vs oracle, fp32 CPU."""
    ref = load_reference.load()
    torch.manual_seed(seed)
    ns, nr, ne, H, B = 11 + seed, 7, 40, 8 * (seed + 1), 2
    ei = torch.stack([torch.randint(0, ns, (ne,)), torch.randint(0, nr, (ne,))])
    ei[1, -1] = nr - 1
    for cls, prop in (("InteractionNet", False), ("PropagationNet", True)):
        for aggr in ("sum", "mean"):
            net = getattr(ref, cls)(ei.clone(), H, aggr=aggr)
            send, rec, edge = torch.randn(B, ns, H), torch.randn(B, nr, H), torch.randn(B, ne, H)
            r_ref, e_ref = net(send, rec, edge)
            r_o, e_o = rp.interaction_net(dict(net.state_dict()), ei, send, rec, edge, aggr=aggr, propagation=prop)
            torch.testing.assert_close(r_o, r_ref, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(e_o, e_ref, rtol=1e-5, atol=1e-5)
