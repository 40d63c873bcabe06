import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_l2
from fesr_b200.dataset.synthetic import make_duct_mesh
from fesr_b200.models.model import KernelNN, TEECNet
from oracle import graph as og, models as om
z = np.load(os.path.join(ROOT, "tests/golden/shipped_w43_weights.npz"))
g = dict(np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz")))
def sd_of(tag): return {k[len(tag)+2:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(tag+"::")}
mesh = make_duct_mesh("50k")
src, dst, ea = og.build_edges(mesh.cells, mesh.pos)
ei = np.stack([src, dst])
torch.set_num_threads(16)
for kind in ("neuralop", "teecnet"):
    sd = sd_of(kind)
    o = om.make_model(kind, 43, 5).double(); o.load_state_dict({k: v.double() for k, v in sd.items()})
    with torch.no_grad(): yo = o(torch.from_numpy(mesh.x).double(), torch.from_numpy(ei), torch.from_numpy(ea).double()).numpy()
    o32 = om.make_model(kind, 43, 5); o32.load_state_dict(sd)
    with torch.no_grad(): y32 = o32(torch.from_numpy(mesh.x), torch.from_numpy(ei), torch.from_numpy(ea)).numpy()
    m = (KernelNN(43,43,5,in_width=4,out_width=4) if kind=="neuralop" else TEECNet(4,43,4,num_layers=5,retrieve_weight=False))
    m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "fp32"
    with torch.no_grad(): y = m(torch.from_numpy(mesh.x).cuda(), torch.from_numpy(ei).cuda(), torch.from_numpy(ea).cuda()).cpu().numpy()
    print(kind, "SIMT" if os.environ.get("FESR_FP32_SIMT") else "3xTF32", "gpu vs fp64 oracle %.3e | cpu fp32 oracle vs fp64 %.3e | gpu vs cpu fp32 %.3e" % (rel_l2(y, yo), rel_l2(y32, yo), rel_l2(y, y32)))
