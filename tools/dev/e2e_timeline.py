"""GPU timeline (kernels + copies, start / duration in us) of ONE end-to-end predict step through the reference-facing
API with host inputs, BASELINE config 2 on one GPU: where the e2e step's time beyond the resident step goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
from fesr_b200.dataset.GraphDataset import SyntheticDuctDataset
from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
from fesr_b200.pipeline import MeshPredictor
mesh_n = int(os.environ.get("MESH_N", "28"))
levels = bench.levels_for(mesh_n)
class C: pass
ctx = C(); ctx.dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mesh = bench.make_mesh(mesh_n)
pos = torch.from_numpy(mesh.pos).cuda(); cells = torch.from_numpy(mesh.cells).cuda()
model = bench.make_model(ctx, "neuralop", "f16")
pred = MeshPredictor(model, pos, cells, levels)
ds = SyntheticDuctDataset(mesh_n=mesh_n, num_meshes=1, sub_size=1 << levels, device=ctx.dev)
ds._cache[0] = {"mesh": mesh, "pos": pos, "part": pred.part, "batch": pred.batch, "x": None, "y": None, "occ": pred.occ}
sched = GNNPartitionScheduler("bench", 1, ds, model, train=True); sched.models = [model]
base = ds.get_one_full_sample(0, materialize=False)
gid = pred.batch.global_ids.cpu().numpy()
xh = torch.from_numpy(mesh.x[gid]).pin_memory(); yh = torch.from_numpy(mesh.y[gid]).pin_memory()
sample = base.with_host_inputs(xh, yh)
def step():
    p, r, mi, wl = sched.predict(sample)
    out = ds.reconstruct_from_partition(p, r, 0, mi, wl)
    f = out.field_local
    p.wait()
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print(f"e2e step {1e3 * (time.perf_counter() - t0) / 50:.3f} ms (wall, 50 steps)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last step = events after the last but one 'stitch' kernel
t_first = ev[0].time_range.start
rows = [(e.time_range.start - t_first, e.time_range.end - e.time_range.start, e.name[:60]) for e in ev]
# find step boundaries: gaps > 100 us
steps, cur = [], [rows[0]]
for r in rows[1:]:
    if r[0] - (cur[-1][0] + cur[-1][1]) > 60 and ("edge_hidden" in r[2]):
        steps.append(cur); cur = []
    cur.append(r)
steps.append(cur)
last = steps[-1]
s0 = last[0][0]
print(f"last step: {len(last)} device activities, span {last[-1][0] + last[-1][1] - s0:.0f} us")
for st, du, nm in last:
    print(f"  +{st - s0:8.1f} us  {du:8.1f} us  {nm}")
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and ("cudaLaunch" in e.name or "cudaMemcpy" in e.name or "cudaEvent" in e.name or "cudaStream" in e.name)]
print("host-side CUDA API calls in the profile:", len(cpu))
