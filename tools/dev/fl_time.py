"""Mean CUDA-event time of the fused layer kernel at 526 848 cells for the library in FESR_LIB_PATH (same-box A/B)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fesr_b200 import ops, _lib
from fesr_b200.dataset.synthetic import make_duct_mesh
from fesr_b200.models.model import KernelNN
z = np.load(os.path.join(ROOT, "tests/golden/shipped_w43_weights.npz"))
sd = {k[10:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("neuralop::")}
nm = int(os.environ.get("MESH_N", "28"))
mesh = make_duct_mesh(nm)
part, b = ops.assemble(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(), 7 if nm == 28 else 10)
m = KernelNN(43, 43, 5, in_width=4, out_width=4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "f16"
x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
res = []
with torch.no_grad():
    for _ in range(5): y = m(x, b.csr, b.edge_attr)
    torch.cuda.synchronize()
    for rep in range(3):
        _lib.profile_enable(True)
        for _ in range(30): m(x, b.csr, b.edge_attr)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        res.append(prof["layer_fused"][0] / prof["layer_fused"][1])
print(json.dumps({"lib": os.path.basename(os.environ.get("FESR_LIB_PATH", "libfesr.so")), "tile": os.environ.get("FESR_FL_TILE", "16"), "layer_fused_ms": [round(r, 5) for r in res], "ysum": float(y.double().abs().sum())}))
