"""Per-parameter gradient error of the tf32 training arm against fp64 oracle autograd (w = 43 shipped weights, 5^3 duct)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import rel_l2
from fesr_b200.dataset.synthetic import make_duct_mesh
from fesr_b200.models.model import KernelNN, TEECNet
from oracle import graph as og, models as om
z = np.load(os.path.join(ROOT, "tests/golden/shipped_w43_weights.npz"))
mesh = make_duct_mesh(int(os.environ.get("MESH_N", "5")))
src, dst, ea = og.build_edges(mesh.cells, mesh.pos)
ei = np.stack([src, dst])
for kind in ("neuralop", "teecnet"):
    sd = {k[len(kind) + 2:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(kind + "::")}
    o = om.make_model(kind, 43, 5).double(); o.load_state_dict({k: v.double() for k, v in sd.items()})
    out = o(torch.from_numpy(mesh.x).double(), torch.from_numpy(ei), torch.from_numpy(ea).double())
    torch.nn.functional.mse_loss(out, torch.from_numpy(mesh.y).double()).backward()
    m = (KernelNN(43, 43, 5, in_width=4, out_width=4) if kind == "neuralop" else TEECNet(4, 43, 4, num_layers=5, retrieve_weight=False))
    m.load_state_dict(sd); m = m.cuda().train(); m.precision = "tf32"
    y = m(torch.from_numpy(mesh.x).cuda(), torch.from_numpy(ei).cuda(), torch.from_numpy(ea).cuda())
    torch.nn.functional.mse_loss(y, torch.from_numpy(mesh.y).cuda()).backward()
    errs = {k: rel_l2(dict(m.named_parameters())[k].grad.cpu().numpy(), p.grad.numpy()) for k, p in o.named_parameters()}
    print(kind, "max", f"{max(errs.values()):.2e}", {k: f"{v:.1e}" for k, v in errs.items()})
