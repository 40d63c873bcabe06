"""Per-role wait accounting of the fused layer kernel (libfesr_tr.so, built with -DFL_TRACE)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["FESR_LIB_PATH"] = os.path.join(ROOT, "fesr_b200/lib/libfesr_tr.so")
import numpy as np, torch
from fesr_b200 import ops, _lib
from fesr_b200.dataset.synthetic import make_duct_mesh
from fesr_b200.models.model import KernelNN
z = np.load(os.path.join(ROOT, "tests/golden/shipped_w43_weights.npz"))
sd = {k[10:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("neuralop::")}
mesh = make_duct_mesh(28)
part, b = ops.assemble(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(), 7)
m = KernelNN(43, 43, 5, in_width=4, out_width=4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "f16"
x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
with torch.no_grad():
    for _ in range(5): m(x, b.csr, b.edge_attr)
    torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (148 * 24 * 4))()
assert lib.fesr_dev_fl_trace(buf) == 0
a = np.frombuffer(buf, dtype=np.int64).reshape(148, 24, 4).astype(np.float64)
tiles = a[:, :, 3].mean()
roles = {"consumer g0": slice(0, 8), "consumer g1": slice(8, 16), "epilogue": slice(16, 20), "mma": slice(20, 21), "producer": slice(21, 24)}
print(f"exp={os.environ.get('FESR_FL_EXP','0')} tiles per CTA {tiles:.1f}; cycles per tile (mean over CTAs and the role's warps):")
for name, sl in roles.items():
    tot, w0, w1 = a[:, sl, 0].mean() / tiles, a[:, sl, 1].mean() / tiles, a[:, sl, 2].mean() / tiles
    print(f"  {name:12s} total {tot:7.0f}  wait0 {w0:7.0f}  wait1 {w1:7.0f}  busy {tot - w0 - w1:7.0f}")
