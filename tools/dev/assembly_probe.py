import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fesr_b200 import ops
from fesr_b200.dataset.synthetic import make_duct_mesh
for n, lv in ((28, 7), (60, 10)):
    mesh = make_duct_mesh(n)
    pos, cells = torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda()
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        part, batch = ops.assemble(pos, cells, lv)
        e1.record()
        torch.cuda.synchronize()
        print(f"n={n} rep={rep} cells={mesh.num_cells} gpu_ms={e0.elapsed_time(e1):.2f} wall_ms={(time.perf_counter()-t0)*1e3:.2f} n_tot={batch.n_tot} e_tot={batch.e_tot}", flush=True)
