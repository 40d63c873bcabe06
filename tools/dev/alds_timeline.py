"""Device timeline of one ALDS predict step (BASELINE config 3 shape on one GPU)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
from fesr_b200.dataset.GraphDataset import SyntheticDuctDataset
from fesr_b200.models import scheduler_gnn as sg
from fesr_b200.models.classifier import KMeansClassifier
from fesr_b200.models.encoder import PCAEncoder
class C: pass
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
ctx = C(); ctx.dev = torch.device("cuda", local); torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=ctx.dev)
    from fesr_b200 import comm
    comm.init_from_torch_distributed()
mesh_n = int(os.environ.get("MESH_N", "44")); k = 4
os.chdir(tempfile.mkdtemp(prefix=f"alds_r{rank}_"))
os.makedirs("logs/models/collection_c", exist_ok=True)
sd = bench.load_weights("neuralop")
for i in range(k):
    s = {kk: v.clone() for kk, v in sd.items()}; s["fc2.bias"] = s["fc2.bias"] + 0.05 * i
    torch.save(s, f"logs/models/collection_c/partition_{i}.pth")
model = bench.make_model(ctx, "neuralop", "f16")
ds = SyntheticDuctDataset(mesh_n=mesh_n, num_meshes=1, device=ctx.dev)
x = ds.get_one_full_sample(0, materialize=False)
enc, clf = PCAEncoder(n_components=2), KMeansClassifier(n_clusters=k)
b = x.batch; ptr = b.node_ptr.cpu().numpy(); xs = x.x_dev.cpu().numpy()
enc.model.fit(np.stack([xs[ptr[s]:ptr[s] + 280].reshape(-1) for s in range(b.n_sub)])); enc._save_model("logs/models/collection_c")
clf.train(enc.get_latent_space(x), save_model=True, path="logs/models/collection_c")
sched = sg.GNNPartitionScheduler("c", k, ds, model, train=False, encoder=enc, classifier=clf)
def step():
    p, r, mi, wl = sched.predict(x)
    return ds.reconstruct_from_partition(p, r, 0, mi, wl)
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
if rank == 0: print(f"alds step {1e3 * (time.perf_counter() - t0) / 20:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
import collections
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
if rank != 0:
    sys.exit(0)
print(f"span {ev[-1].time_range.end - t0:.0f} us, {len(ev)} activities, busy {sum(v[1] for v in agg.values()):.0f} us")
for kname, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {v[1]:9.1f} us  n={v[0]:3d}  {kname}")

print("timeline (start us, dur us, name):")
for e in ev:
    print(f"  +{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:60]}")
