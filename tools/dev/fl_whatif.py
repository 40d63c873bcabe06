"""What-if timing of the fused layer kernel (FESR_FL_EXP bits: 1 no h gathers, 2 no fix-up row, 4 no outer-product
MMAs, 8 one k-block of the contraction only).  Results are WRONG by construction; only the time matters."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
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
        _lib.profile_enable(True)
        for _ in range(30): m(x, b.csr, b.edge_attr)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
    print(json.dumps({"exp": os.environ.get("FESR_FL_EXP", "0"), "layer_fused_ms": prof["layer_fused"][0] / prof["layer_fused"][1]}))
else:
    for e in (0, 1, 2, 4, 8, 3, 6, 12, 15):
        env = dict(os.environ, FESR_FL_EXP=str(e))
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
