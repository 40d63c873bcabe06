#!/usr/bin/env python
"""bench.py -- super-resolved mesh cells/s of the predict hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference ...                     the reference's algorithm on host cores

Headline workload (config.workload), the same at every N so that the driver's 1 -> 8 curve is one strong-scaling
curve: BASELINE config 5 -- the 5 184 000-cell synthetic duct (n = 60), KernelNN (--model=neuralop) with the shipped
w = 43 checkpoint, 5 layers, 2^10 kd subdomains with a one-cell halo (AssignToAllIntersectingRegions), subdomains
sharded over the N ranks, f16 arm.  One step = one predict pass: model forward over every subdomain of the rank's
shard (one block-diagonal batch) + node weight + [N > 1: ONE in-place all-gather of the predictions, issued by libfesr
on its own NCCL communicator] + overlap stitch onto the fine mesh (N > 1: every rank stitches its slice of the mesh
nodes).  Subdomain assembly is one-time preprocessing in the reference (GraphDataset.get_partition_domain) and is
timed separately (extra.assembly_ms).

value : cells/s with inputs resident in HBM; timed with CUDA events, max over ranks.
e2e   : the same pass through the reference-facing API -- GNNPartitionScheduler.predict(sample) +
        dataset.reconstruct_from_partition(...) -- with HOST (pinned) inputs: H2D of x and y rows, predict, stitch,
        D2H of the stitched field (each rank its slice), the per-subdomain predictions and the node weights.

Further records on the same JSON line (BASELINE's other configs; `--no-extras` skips them):
  extra.config2   config 2 (526 848 cells, 1 GPU): value / e2e / kernels           [N = 1 only]
  arms            config 2 through the fp32 (<= 1e-5), tf32 and f16 (<= 1e-3) arms   [N = 1 only]
  train           config 4: train step (fwd + MSE + bwd + all-reduce + Adam) on the 2 044 416-cell duct, tf32 / fp32
  alds            config 3: PCA + k-means routing to 4 models + predict + stitch on the 2 044 416-cell duct
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEAD_N = 60                      # 24 * 60^3 = 5 184 000 cells (BASELINE config 5)
C2_N = 28                        # 24 * 28^3 =   526 848 cells (config 2)
C34_N = 44                       # 24 * 44^3 = 2 044 416 cells (configs 3 / 4)
METRIC = "super-resolved mesh cells/s (predict pass: forward + node weight + stitch)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fesr", choices=["fesr", "reference"])
    ap.add_argument("--model", default="neuralop", choices=["neuralop", "teecnet"])
    ap.add_argument("--precision", default=os.environ.get("FESR_PRECISION", "f16"),
                    help="arithmetic of the node contraction: f16 | tf32 (rel-L2 <= 1e-3 arms), fp32 (<= 1e-5 arm)")
    ap.add_argument("--mesh-n", type=int, default=HEAD_N)
    ap.add_argument("--levels", type=int, default=-1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no config2 / arms / train / alds)")
    ap.add_argument("--strong", action="store_true", help="(accepted for compatibility: the workload is always strong)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- shared setup
def make_mesh(n):
    from fesr_b200.dataset import synthetic as syn
    return syn.make_duct_mesh(n)


def load_weights(kind):
    import torch
    path = os.path.join(ROOT, "tests", "golden", "shipped_w43_weights.npz")
    z = np.load(path)
    pre = kind + "::"
    return {k[len(pre):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(pre)}


def levels_for(mesh_n, levels=-1):
    if levels >= 0:
        return levels
    from fesr_b200.dataset import synthetic as syn
    return syn.default_kd_levels((mesh_n + 1) ** 2 * (4 * mesh_n + 1))


def workload_config(args, world, levels, cells):
    """Identical in both arms (the driver compares them): nothing measured or arm-specific goes in here."""
    return {"workload": f"BASELINE config 5: ALDS/DS super-resolution predict + overlap stitch, synthetic duct n={args.mesh_n} "
                        f"({cells} cells), --model={args.model} w=43 shipped checkpoint, 5 layers, 2^{levels} kd subdomains + "
                        f"1-cell halo, strong scaling over the ranks",
            "cells": int(cells), "subdomains": 1 << levels, "model": args.model, "width": 43, "layers": 5,
            "parallelism": f"subdomain-sharded x{world}",
            "l2": "inputs + per-layer streams exceed the 126 MB L2 (g alone is 0.2-1.7 GB per layer pass)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.windows = []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def window(self, t0, t1):
        """A timed region (datetime pair): the samples inside the windows are the ones `under load`."""
        self.windows.append((t0, t1))

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm_all, sm_load, mx, reasons = [], [], [], set()
        for r in rows:
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f")
                clk, cmax = float(r[2]), float(r[3])
            except ValueError:
                continue
            sm_all.append(clk)
            mx.append(cmax)
            loaded = any(a <= ts <= b for a, b in self.windows)
            if loaded:
                sm_load.append(clk)
            if loaded or not self.windows:
                for nm, v in zip(names, r[6:10]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
        use = sm_load if sm_load else sm_all
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(use), "samples_total": len(sm_all),
                "note": "median over the samples taken inside the timed regions" if sm_load else
                        "no sample fell inside a timed region: median over the whole run"}


# ----------------------------------------------------------------------------- reference arm
def oracle_subdomain_pass(model, sub, mesh, s_list, home_cells):
    """The reference's per-subdomain predict loop (models/scheduler_gnn.py:217-226) + stitch sample."""
    import torch
    from oracle import graph as og
    from oracle import models as om
    preds, gids = [], []
    cells = 0
    t0 = time.perf_counter()
    for s in s_list:
        nl, nh = sub["node_ptr"][s], sub["node_ptr"][s + 1]
        el, eh = sub["edge_ptr"][s], sub["edge_ptr"][s + 1]
        g = sub["global_ids"][nl:nh]
        x = torch.from_numpy(mesh.x[g])
        y = torch.from_numpy(mesh.y[g])
        ei = torch.from_numpy(np.stack([sub["edge_src"][el:eh] - nl, sub["edge_dst"][el:eh] - nl]))
        ea = torch.from_numpy(sub["edge_attr"][el:eh]).unsqueeze(1)
        with torch.no_grad():
            p = model(x, ei, ea)
            om.compute_node_weight(p, y, ei, ea, x.shape[0])
        preds.append(p.numpy())
        gids.append(g)
        cells += int(home_cells[s])
    og.stitch_mean(np.concatenate(preds), np.concatenate(gids), mesh.num_nodes)
    return time.perf_counter() - t0, cells


def cpu_model(kind):
    import torch
    from oracle import models as om
    torch.set_num_threads(os.cpu_count() or 1)
    model = om.make_model(kind, 43, 5)
    model.load_state_dict(load_weights(kind))
    return model.eval()


def cpu_setup(args, levels):
    from oracle import graph as og
    mesh = make_mesh(args.mesh_n)
    part = og.kd_partition(mesh.pos, mesh.cells, levels)
    sub = og.build_subdomains(mesh.pos, mesh.cells, part["leaf_ptr"], part["leaf_cells"])
    home_cells = np.bincount(part["home_leaf"], minlength=1 << levels)
    return mesh, sub, home_cells, cpu_model(args.model)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = args.gpus
    levels = levels_for(args.mesh_n, args.levels)
    mesh, sub, home_cells, model = cpu_setup(args, levels)
    S = 1 << levels
    # size the per-step sample so that one step is ~1/3 of the budget (first call untimed: thread pool start-up)
    oracle_subdomain_pass(model, sub, mesh, [0], home_cells)
    t1, c1 = oracle_subdomain_pass(model, sub, mesh, [1 % S], home_cells)
    per_step = max(1, min(S, int(max(args.cpu_seconds / 3.0, 1.0) / max(t1, 1e-3))))
    s_list = [int(v) for v in np.linspace(0, S - 1, per_step).round()]       # the same spread-out sample at every N
    for _ in range(min(args.warmup, 1)):
        oracle_subdomain_pass(model, sub, mesh, s_list[:max(1, per_step // 4)], home_cells)
    steps = max(1, min(args.steps, 3))
    times, cells = [], 0
    for _ in range(steps):
        t, c = oracle_subdomain_pass(model, sub, mesh, s_list, home_cells)
        times.append(t)
        cells = c
    best = min(times)
    value = cells / best
    sample = f"{per_step} of {S} subdomains ({cells} cells) per step, best of {steps}, torch CPU + numpy stitch"
    line = {"impl": "reference", "metric": METRIC,
            "value": value, "unit": "cells/s", "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": best * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, levels, mesh.num_cells),
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ----------------------------------------------------------------------------- fesr arm
class Ctx:
    """Process-wide state of the fesr arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
            from fesr_b200 import comm
            comm.init_from_torch_distributed()
        self.sampler = None

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed steps, then EXACTLY `steps` timed ones between barrier + synchronize; CUDA events; max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        w0 = datetime.datetime.now()
        t0, t1 = self.ev(), self.ev()
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        self.barrier()
        if self.sampler is not None:
            self.sampler.window(w0, datetime.datetime.now())
        ms = t0.elapsed_time(t1)
        if self.world > 1:
            t = self.torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def free(self):
        import gc
        from fesr_b200 import ops
        gc.collect()
        ops.workspace.clear()
        self.torch.cuda.empty_cache()


def make_model(ctx, kind, precision):
    from fesr_b200.models.model import KernelNN, TEECNet
    model = (KernelNN(43, 43, 5, in_width=4, out_width=4) if kind == "neuralop" else
             TEECNet(4, 43, 4, num_layers=5, retrieve_weight=False))
    model.load_state_dict(load_weights(kind))
    model = model.to(ctx.dev).eval()
    model.precision = precision
    return model


def peaks():
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": mp["hbm_gbs"], "bf16_tflops": mp["bf16_tflops"], "src": "MEASURED_PEAKS.json"}
    except (OSError, KeyError, ValueError):
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}


def predict_workload(ctx, kind, precision, mesh_n, levels, steps, warmup, want_e2e=True, want_kernels=True):
    """One mesh through the resident step (`value`), an instrumented pass (per-kernel CUDA events) and the e2e API."""
    torch = ctx.torch
    from fesr_b200 import _lib
    from fesr_b200.pipeline import MeshPredictor
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    mesh = make_mesh(mesh_n)
    pos = torch.from_numpy(mesh.pos).to(dev)
    cells = torch.from_numpy(mesh.cells).to(dev)
    model = make_model(ctx, kind, precision)
    a0, a1 = ctx.ev(), ctx.ev()
    torch.cuda.synchronize()
    a0.record()
    pred = MeshPredictor(model, pos, cells, levels, rank=rank, world=world)
    a1.record()
    torch.cuda.synchronize()
    res = {"cells": mesh.num_cells, "nodes": mesh.num_nodes, "assembly_ms": a0.elapsed_time(a1),
           "batch_nodes": pred.batch.n_tot, "batch_edges": pred.batch.e_tot, "precision": precision, "mesh": mesh,
           "pred": pred, "model": model}
    sh = pred.shard
    gid = sh.global_ids
    x_dev = torch.from_numpy(mesh.x).to(dev)[gid]
    y_dev = torch.from_numpy(mesh.y).to(dev)[gid]

    def step_resident():
        pred.step(x_dev, y_dev)

    launches0 = _lib.launch_count()
    ms = ctx.timed(step_resident, steps, warmup)
    res["gpu_launches"] = int((_lib.launch_count() - launches0) * steps // (steps + warmup))
    res["ms_per_step"] = ms / steps
    res["value"] = mesh.num_cells / (ms / steps / 1e3)
    if want_kernels:
        ps = min(steps, 50)
        _lib.profile_enable(True)
        ms_prof = ctx.timed(step_resident, ps, 3)
        res["prof"] = _lib.profile_collect()          # includes the warm-up launches; per-launch means are what we use
        _lib.profile_enable(False)
        res["ms_per_step_instrumented"] = ms_prof / ps
    if want_e2e:
        # e2e goes through the reference-facing API: GNNPartitionScheduler.predict(sample) on a sample whose input /
        # reference fields live in pinned HOST memory, then dataset.reconstruct_from_partition(...)
        # (run_ALDS_3D.py:17-26); both hand back host tensors, so every step pays its H2D and D2H copies.
        from fesr_b200.dataset.GraphDataset import SyntheticDuctDataset
        from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
        ds = SyntheticDuctDataset(mesh_n=mesh_n, num_meshes=1, sub_size=1 << levels, device=dev)
        ds._cache[0] = {"mesh": mesh, "pos": pos, "part": pred.part, "batch": pred.batch, "x": None, "y": None,
                        "occ": pred.occ}          # the decomposition assembled above: same arrays, no second copy
        sched = GNNPartitionScheduler("bench", 1, ds, model, train=True)
        sched.models = [model]
        base = ds.get_one_full_sample(0, materialize=False)
        gid_all = pred.batch.global_ids.cpu().numpy()
        xa_host = torch.from_numpy(mesh.x[gid_all]).pin_memory()
        ya_host = torch.from_numpy(mesh.y[gid_all]).pin_memory()
        sample_h = base.with_host_inputs(xa_host, ya_host)
        host_t = {"predict": [], "reconstruct": [], "wait": []}

        def step_e2e():
            t0 = time.perf_counter()
            p, r, mi, wl = sched.predict(sample_h)
            t1 = time.perf_counter()
            out = ds.reconstruct_from_partition(p, r, 0, mi, wl)
            t2 = time.perf_counter()
            f = out.field_local                  # the stitched prediction on the host (N > 1: this rank's slice of the mesh)
            p.wait()                             # ... and this rank's per-subdomain predictions + weights (packed copy)
            t3 = time.perf_counter()
            host_t["predict"].append(t1 - t0)    # host time to ISSUE the pass (nothing here waits for the GPU)
            host_t["reconstruct"].append(t2 - t1)
            host_t["wait"].append(t3 - t2)
            return f

        ms_e2e = ctx.timed(step_e2e, steps, warmup)
        # whole-job bytes per step: every rank copies ITS rows of x and y in; its rows of the predictions, its
        # subdomain weights and its slice of the stitched field out (the ranks' slices tile the mesh)
        res["e2e"] = {"value": mesh.num_cells / (ms_e2e / steps / 1e3), "unit": "cells/s", "ms_per_step": ms_e2e / steps,
                      "h2d_bytes_per_step": int(xa_host.numel() * 4 + ya_host.numel() * 4),
                      "d2h_bytes_per_step": int(pred.batch.n_tot * 16 + (1 << levels) * 4 + mesh.num_nodes * 16),
                      "api": "GNNPartitionScheduler.predict + dataset.reconstruct_from_partition"
                             + (" (field distributed: every rank keeps its node slice)" if world > 1 else ""),
                      "host_ms_median": {k: 1e3 * statistics.median(v[warmup:] or v) for k, v in host_t.items() if v}}
    return res


def kernel_table(res, pk):
    """Per-kernel-class CUDA-event means of the instrumented pass + algorithmic bytes."""
    pred, d, precision = res["pred"], res["model"].dims, res["precision"]
    sh = pred.shard
    n_s, E_s = sh.csr.n, sh.csr.E
    w = d.w
    zb = 2 if precision in ("f16", "fp16") else 4          # bytes per element of the Z intermediate
    hb = 2 if precision in ("f16", "fp16") else 4          # bytes per element of g and h
    alg = {
        # one fused layer (layer_fused.cu): src index, g row and gathered h row per edge; own h row read, h' row
        # written and two row bounds per node.  Z stays on chip, so it is not in the byte count
        "layer_fused": ("hbm", E_s * (4 + hb * d.kp + hb * d.wp) + n_s * (2 * hb * d.wp + 4) + 4),
        "zbuild": ("hbm", E_s * (4 + hb * d.k1 + hb * w) + n_s * (hb * w + zb * (d.k1 * w + w)) + 4 * (n_s + 1)),
        "node_gemm": ("hbm" if precision != "fp32" else "fp32", n_s * (zb * (d.k1 * w + w) + 4 * w)),
        "edge_hidden": ("hbm", E_s * (4 + hb * d.k1)),
        "stitch": ("hbm", pred.batch.n_tot * 20 + pred.N * 20 // max(pred.world, 1)),
        "node_weight": ("hbm", E_s * (4 + 4 + 32) + n_s * 36),
    }
    kernels = {}
    prof = res["prof"]
    tot = sum(v[0] for v in prof.values()) or 1.0
    for k, (t_ms, cnt) in prof.items():
        if cnt == 0:
            continue
        per = t_ms / cnt
        ent = {"ms_per_launch": per, "launches": int(cnt), "share": t_ms / tot}
        if k in alg:
            ent["bound"] = alg[k][0]
            ent["alg_bytes"] = int(alg[k][1])
            ent["gbs"] = alg[k][1] / (per * 1e-3) / 1e9
            ent["frac_hbm_peak"] = ent["gbs"] / pk["hbm_gbs"]
        if k == "node_gemm":
            ent["tflops"] = 2.0 * n_s * d.zk * d.wp / (per * 1e-3) / 1e12
        kernels[k] = ent
    return kernels


def roofline_of(kernels, pk, precision, mesh_n):
    top = max(kernels, key=lambda k: kernels[k]["share"])
    if top == "node_gemm" and precision == "fp32":
        top = "zbuild"
    k = top if "gbs" in kernels[top] else "zbuild"
    traffic, src = ncu_traffic(k, precision, mesh_n)
    return {"kernel": k, "bound": "hbm", "achieved": kernels[k]["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": kernels[k]["gbs"] / pk["hbm_gbs"], "traffic": traffic, "traffic_source": src,
            "peak_source": pk["src"], "share_of_step": kernels[k]["share"],
            "achieved_is": "algorithmic bytes per launch / CUDA-event mean of the launch, measured live in this run"}


def ncu_traffic(kind, precision, mesh_n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel.  NOT measured by this run: it is
    read from the committed `ncu --set full` capture of the same workload (profiles/*_traffic.json), and labelled so."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
        except (OSError, ValueError):
            continue
        key = f"{kind}:{precision}:n{mesh_n}"
        if key in t:
            return t[key]["dram_bytes_per_launch"], f"committed ncu capture profiles/{name} [{key}] (not re-measured per run)"
    return None, "no committed capture for this kernel / precision / mesh"


def run_cpu_baseline(ctx, args, res, levels):
    """The oracle's per-subdomain loop (reference order, torch CPU, all host cores) on a bounded sample of the SAME
    subdomains the GPU path ran (the device batch is bit-exact with the oracle's own assembly -- tests -- so its arrays
    are what the loop is fed; assembly is outside the timed region in both arms)."""
    pred, mesh = res["pred"], res["mesh"]
    b = pred.batch
    sub = {"node_ptr": b.node_ptr.cpu().numpy(), "edge_ptr": b.edge_ptr.cpu().numpy(), "global_ids": b.global_ids.cpu().numpy(),
           "edge_src": b.edge_src.cpu().numpy().astype(np.int64), "edge_dst": b.edge_dst.cpu().numpy().astype(np.int64),
           "edge_attr": b.edge_attr.cpu().numpy()}
    home_cells = pred.home_cells
    model = cpu_model(args.model)
    S = 1 << levels
    oracle_subdomain_pass(model, sub, mesh, [0], home_cells)                 # untimed: thread pool start-up
    t1, _ = oracle_subdomain_pass(model, sub, mesh, [1 % S], home_cells)
    m = max(1, min(S, int(args.cpu_seconds / max(t1, 1e-3))))
    s_list = [int(v) for v in np.linspace(0, S - 1, m).round()]
    t, c = oracle_subdomain_pass(model, sub, mesh, s_list, home_cells)
    return {"value": c / t, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{m} of {S} subdomains ({c} cells) in {t:.1f} s, torch CPU per-subdomain loop + numpy stitch"}


def train_record(ctx, args):
    """BASELINE config 4: train step = forward + MSELoss + backward + gradient all-reduce (fesr_allreduce_grads) + Adam
    on the 2 044 416-cell duct, the rank's shard of the 512 subdomains as one block-diagonal batch (the MSE is a mean
    over nodes x channels, so the loss definition is unchanged; models/scheduler_gnn.py:398-409)."""
    torch = ctx.torch
    from fesr_b200 import ops
    from fesr_b200.models.training import FlatAdam, train_step
    from fesr_b200.pipeline import make_shard, shard_bounds
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    levels = levels_for(C34_N)
    mesh = make_mesh(C34_N)
    part, batch = ops.assemble(torch.from_numpy(mesh.pos).to(dev), torch.from_numpy(mesh.cells).to(dev), levels)
    bounds = shard_bounds(batch.edge_ptr.cpu().numpy(), world)
    sh = make_shard(batch, bounds[rank], bounds[rank + 1])
    x = torch.from_numpy(mesh.x).to(dev)[sh.global_ids]
    y = torch.from_numpy(mesh.y).to(dev)[sh.global_ids]
    del part
    rec = {"workload": f"BASELINE config 4: run_DS_3D.py --mode=train, --model={args.model}, synthetic duct n={C34_N} "
                       f"({mesh.num_cells} cells), 2^{levels} subdomains sharded x{world}, one block-diagonal batch per rank, "
                       f"Adam lr 5e-4", "cells": mesh.num_cells, "unit": "cells/s", "n_gpus": world, "arms": {}}
    for prec, steps in (("tf32", 5), ("fp32", 2)):
        ctx.free()
        torch.cuda.reset_peak_memory_stats()
        model = make_model(ctx, args.model, prec).train()
        opt = FlatAdam(model, lr=5e-4)
        holder = {}

        def step():
            holder["loss"] = train_step(model, opt, x, sh.csr, sh.edge_attr, y)

        ms = ctx.timed(step, steps, 2)
        rec["arms"][prec] = {"value": mesh.num_cells / (ms / steps / 1e3), "ms_per_step": ms / steps, "steps": steps,
                             "loss": float(holder["loss"]), "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        del model, opt
    return rec


def alds_record(ctx, args):
    """BASELINE config 3: run_ALDS_3D.py -- PCA + k-means routing of the subdomains to 4 per-cluster models, predict,
    overlap stitch -- on the 2 044 416-cell duct through GNNPartitionScheduler.predict + reconstruct_from_partition."""
    torch = ctx.torch
    from fesr_b200.dataset.GraphDataset import SyntheticDuctDataset
    from fesr_b200.models import scheduler_gnn as sg
    from fesr_b200.models.classifier import KMeansClassifier
    from fesr_b200.models.encoder import PCAEncoder
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    k = 4
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix=f"fesr_alds_r{rank}_")
    os.chdir(work)
    try:
        os.makedirs("logs/models/collection_c", exist_ok=True)
        sd = load_weights(args.model)
        out_b = "fc2.bias" if args.model == "neuralop" else "fc_out.bias"
        for i in range(k):
            s = {kk: v.clone() for kk, v in sd.items()}
            s[out_b] = s[out_b] + 0.05 * i          # distinguishable per-cluster models
            torch.save(s, f"logs/models/collection_c/partition_{i}.pth")
        model = make_model(ctx, args.model, args.precision)
        ds = SyntheticDuctDataset(mesh_n=C34_N, num_meshes=1, device=dev)
        x = ds.get_one_full_sample(0, materialize=False)
        enc, clf = PCAEncoder(n_components=2), KMeansClassifier(n_clusters=k)
        b = x.batch
        ptr = b.node_ptr.cpu().numpy()
        xs = x.x_dev.cpu().numpy()
        enc.model.fit(np.stack([xs[ptr[s]:ptr[s] + 280].reshape(-1) for s in range(b.n_sub)]))
        enc._save_model("logs/models/collection_c")
        clf.train(enc.get_latent_space(x), save_model=True, path="logs/models/collection_c")
        sched = sg.GNNPartitionScheduler("c", k, ds, model, train=False, encoder=enc, classifier=clf)
        holder = {}

        def step():
            p, r, mi, wl = sched.predict(x)
            holder["mi"] = mi
            holder["out"] = ds.reconstruct_from_partition(p, r, 0, mi, wl)

        steps = 10
        ms = ctx.timed(step, steps, 3)
        cells = ds._mesh(0)["mesh"].num_cells
        return {"workload": f"BASELINE config 3: run_ALDS_3D.py, PCA(2) + k-means({k}) routing, --model={args.model} "
                            f"{args.precision}, synthetic duct n={C34_N} ({cells} cells), {b.n_sub} subdomains sharded x{world}, "
                            f"scheduler.predict + reconstruct_from_partition (device-resident inputs)",
                "cells": cells, "unit": "cells/s", "n_gpus": world, "value": cells / (ms / steps / 1e3),
                "ms_per_step": ms / steps, "steps": steps,
                "labels_hist": np.bincount(holder["mi"], minlength=k).tolist()}
    finally:
        os.chdir(cwd)


def slim(res, pk):
    """JSON-able summary of a predict_workload result."""
    out = {k: res[k] for k in ("cells", "assembly_ms", "batch_nodes", "batch_edges", "precision", "value", "ms_per_step",
                               "gpu_launches") if k in res}
    if "e2e" in res:
        out["e2e"] = res["e2e"]
    if "prof" in res:
        out["kernels"] = kernel_table(res, pk)
        out["ms_per_step_instrumented"] = res["ms_per_step_instrumented"]
    return out


def run_fesr(args):
    ctx = Ctx()
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    args.gpus = world
    pk = peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    levels = levels_for(args.mesh_n, args.levels)
    ctx.sampler = ClockSampler(ctx.local)
    if rank == 0:
        ctx.sampler.start()

    head = predict_workload(ctx, args.model, args.precision, args.mesh_n, levels, steps, warmup)
    kernels = kernel_table(head, pk)
    roof = roofline_of(kernels, pk, args.precision, args.mesh_n)
    clocks = ctx.sampler.stop() if rank == 0 else None
    ctx.sampler = None
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(ctx, args, head, levels)
    line = {"metric": METRIC, "value": head["value"], "unit": "cells/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32"}.get(args.precision, "f16"), "data": "synthetic",
            "config": workload_config(args, world, levels, head["cells"]), "clocks": clocks, "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "roofline": roof, "kernels": kernels,
            "ms_per_step_instrumented": head["ms_per_step_instrumented"], "cpu_baseline": cpu_baseline,
            "extra": {"assembly_ms": head["assembly_ms"], "batch_nodes": head["batch_nodes"],
                      "batch_edges": head["batch_edges"], "precision": args.precision}}
    del head
    ctx.free()

    if not args.no_extras:
        if world == 1:
            c2_levels = levels_for(C2_N)
            c2 = predict_workload(ctx, args.model, args.precision, C2_N, c2_levels, max(steps, 100), warmup)
            line["extra"]["config2"] = dict(slim(c2, pk), workload=f"BASELINE config 2: run_DS_3D.py --mode=predict "
                                            f"--model={args.model}, synthetic duct n={C2_N}, 2^{c2_levels} subdomains, 1 GPU")
            line["extra"]["config2"]["roofline"] = roofline_of(line["extra"]["config2"]["kernels"], pk, args.precision, C2_N)
            arms = {args.precision: {"value": c2["value"], "ms_per_step": c2["ms_per_step"]}}
            del c2
            ctx.free()
            for prec, st in (("f16", 100), ("tf32", 30), ("fp32", 10)):
                if prec in arms:
                    continue
                r = predict_workload(ctx, args.model, prec, C2_N, c2_levels, st, 3, want_e2e=False, want_kernels=False)
                arms[prec] = {"value": r["value"], "ms_per_step": r["ms_per_step"]}
                del r
                ctx.free()
            line["arms"] = {"workload": "BASELINE config 2 (526 848 cells, 1 GPU), resident predict pass per arithmetic arm; "
                                        "parity gates: fp32 <= 1e-5, tf32 / f16 <= 1e-3 per channel "
                                        "(tests/test_gpu_parity_fullsize.py)", "unit": "cells/s", **arms}
        line["train"] = train_record(ctx, args)
        ctx.free()
        line["alds"] = alds_record(ctx, args)
        ctx.free()
    if rank == 0:
        emit(line)
    if world > 1:
        from fesr_b200 import comm
        ctx.barrier()
        comm.destroy()
        ctx.dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version
    banner, for one) has been routed to stderr."""
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_fesr(a))
