"""Per-role wait accounting of the 16-node fused layer kernel (libfesr_tr.so, built by tools/dev/build_trace.py)."""
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
nmesh = int(os.environ.get("MESH_N", "28"))
mesh = make_duct_mesh(nmesh)
part, b = ops.assemble(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(), 7 if nmesh == 28 else 10)
m = KernelNN(43, 43, 5, in_width=4, out_width=4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "f16"
x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
with torch.no_grad():
    for _ in range(5): m(x, b.csr, b.edge_attr)
    torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (148 * 24 * 12))()
assert lib.fesr_dev_fl16_trace(buf) == 0
a = np.frombuffer(buf, dtype=np.int64).reshape(148, 24, 12).astype(np.float64)
tiles = a[:, :, 1].mean()
roles = {"consumer": slice(0, 16), "epilogue": slice(16, 20), "mma": slice(20, 21), "producer": slice(21, 24)}
print(f"tiles per CTA {tiles:.1f}; cycles per 16-node tile (mean over CTAs and the role's warps): total | w0 w1 | m2 m3 m9")
for name, sl in roles.items():
    tot = a[:, sl, 0].mean() / tiles
    v = [a[:, sl, 2 + q].mean() / tiles for q in range(10)]
    print(f"  {name:10s} total {tot:7.0f}  wait0 {v[0]:7.0f}  wait1 {v[1]:7.0f}  mark2 {v[2]:7.0f}  mark3 {v[3]:7.0f}  mark9 {v[9]:7.0f}")
print("per consumer warp (mean over CTAs): warp: total wait0 wait1 m2 m3 m9")
for w in range(24):
    v = a[:, w, :].mean(axis=0) / tiles
    print(f"  w{w:2d} smsp{w%4} total {v[0]:6.0f} w0 {v[2]:6.0f} w1 {v[3]:6.0f} m2 {v[4]:6.0f} m3 {v[5]:6.0f} m9 {v[11]:6.0f}")
tot = a[:, 0, 0]          # consumer warp 0: cycles from the start of the main loop to its end, per CTA
nt = a[:, 0, 1]
print(f"per-CTA span (consumer warp 0): mean {tot.mean():.0f} max {tot.max():.0f} min {tot.min():.0f} cycles; max/mean {tot.max()/tot.mean():.3f}; tiles per CTA min {nt.min():.0f} max {nt.max():.0f}")
import numpy as np
order = np.argsort(tot)
print("slowest CTAs:", [(int(i), int(tot[i])) for i in order[-5:]], "fastest:", [(int(i), int(tot[i])) for i in order[:5]])
