"""GPU parity: backward / train step vs gradients produced by the reference's own classes
(tests/golden) and vs torch.autograd of the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, shipped_state_dict, state_dict_from
from oracle import graph as og
from oracle import models as om

pytestmark = pytest.mark.gpu


def _model(kind, w, L):
    from fesr_b200.models.model import KernelNN, TEECNet
    if kind == "neuralop":
        return KernelNN(w, w, L, in_width=4, out_width=4)
    return TEECNet(4, w, 4, num_layers=L, retrieve_weight=False)


def _grads(model, x, ei, ea, y, precision="fp32"):
    model = model.cuda().train()
    model.precision = precision
    for p in model.parameters():
        p.grad = None
    out = model(torch.from_numpy(x).cuda(), torch.from_numpy(ei).cuda(), torch.from_numpy(ea).cuda())
    loss = torch.nn.functional.mse_loss(out, torch.from_numpy(y).cuda())
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}


@pytest.mark.parametrize("name,kind,w,L", [("kernelnn_w16", "neuralop", 16, 3), ("teecnet_w12", "teecnet", 12, 2),
                                           ("kernelnn_w48", "neuralop", 48, 2)])
def test_grads_vs_reference_vectors(golden, name, kind, w, L):
    m = _model(kind, w, L)
    m.load_state_dict(state_dict_from(golden, name))
    loss, g = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    assert abs(loss - float(golden[name + "_loss"])) <= 1e-5 * abs(float(golden[name + "_loss"]))
    for k, v in g.items():
        ref = golden[f"{name}_grad::{k}"]
        assert v.shape == ref.shape, k
        assert rel_l2(v, ref) < 2e-4, (k, rel_l2(v, ref))


@pytest.mark.parametrize("kind", ["neuralop", "teecnet"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tf32", 5e-3)])
def test_grads_w43_vs_oracle_autograd(shipped, kind, precision, tol):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(5)
    src, dst, ea = og.build_edges(mesh.cells, mesh.pos)
    ei = np.stack([src, dst])
    sd = shipped_state_dict(shipped, kind)
    m = _model(kind, 43, 5)
    m.load_state_dict(sd)
    o = om.make_model(kind, 43, 5).double()
    o.load_state_dict({k: v.double() for k, v in sd.items()})
    out = o(torch.from_numpy(mesh.x).double(), torch.from_numpy(ei), torch.from_numpy(ea).double())
    lo = torch.nn.functional.mse_loss(out, torch.from_numpy(mesh.y).double())
    lo.backward()
    loss, g = _grads(m, mesh.x, ei, ea, mesh.y, precision)
    assert abs(loss - float(lo)) <= max(tol, 1e-5) * abs(float(lo))
    for k, p in o.named_parameters():
        err = rel_l2(g[k], p.grad.numpy())
        assert err < tol, (k, err)


def test_grad_x(golden):
    m = _model("neuralop", 16, 3).cuda()
    sd = state_dict_from(golden, "kernelnn_w16")
    m.load_state_dict(sd)
    o = om.make_model("neuralop", 16, 3)
    o.load_state_dict(sd)
    x = torch.from_numpy(golden["x"]).requires_grad_(True)
    ei, ea = torch.from_numpy(golden["ref_edge_index"]), torch.from_numpy(golden["ref_edge_attr"])
    o(x, ei, ea).square().sum().backward()
    xg = torch.from_numpy(golden["x"]).cuda().requires_grad_(True)
    m(xg, ei.cuda(), ea.cuda()).square().sum().backward()
    assert rel_l2(xg.grad.cpu().numpy(), x.grad.numpy()) < 2e-4


def test_backward_is_deterministic(golden):
    m = _model("neuralop", 16, 3)
    m.load_state_dict(state_dict_from(golden, "kernelnn_w16"))
    _, a = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    _, b = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    for k in a:
        assert np.array_equal(a[k], b[k]), k
