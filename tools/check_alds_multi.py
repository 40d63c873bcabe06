#!/usr/bin/env python
"""Multi-GPU self-check of the ALDS predict path (BASELINE configs 3 / 5):

    torchrun --nproc-per-node N tools/check_alds_multi.py [--mesh-n 44] [--clusters 4] [--model teecnet]

Every rank runs GNNPartitionScheduler.predict + reconstruct_from_partition (a) sharded over the N ranks and
(b) alone on its own GPU (the process group hidden), and the two stitched fields / subdomain weights / labels
must agree.  Prints one JSON line from rank 0 with the timings of both.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", type=int, default=28)
    ap.add_argument("--clusters", type=int, default=4)
    ap.add_argument("--model", default="neuralop")
    ap.add_argument("--precision", default="f16")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch.distributed as dist
    from bench import load_weights
    from fesr_b200.dataset.GraphDataset import SyntheticDuctDataset
    from fesr_b200.models import scheduler_gnn as sg
    from fesr_b200.models.classifier import KMeansClassifier
    from fesr_b200.models.encoder import PCAEncoder
    from fesr_b200.models.model import KernelNN, TEECNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    work = tempfile.mkdtemp(prefix=f"alds_r{rank}_")
    os.chdir(work)
    os.makedirs("logs/models/collection_c", exist_ok=True)
    sd = load_weights(a.model)
    out_b = "fc2.bias" if a.model == "neuralop" else "fc_out.bias"
    for i in range(a.clusters):
        s = {k: v.clone() for k, v in sd.items()}
        s[out_b] = s[out_b] + 0.05 * i
        torch.save(s, f"logs/models/collection_c/partition_{i}.pth")
    model = (KernelNN(43, 43, 5, in_width=4, out_width=4) if a.model == "neuralop" else
             TEECNet(4, 43, 4, num_layers=5, retrieve_weight=False))
    model.precision = a.precision
    ds = SyntheticDuctDataset(mesh_n=a.mesh_n, num_meshes=1, device=dev)
    x = ds.get_one_full_sample(0, materialize=False)
    enc = clf = None
    if a.clusters > 1:
        enc, clf = PCAEncoder(n_components=2), KMeansClassifier(n_clusters=a.clusters)
        # fit on the device batch's first 280 rows of every subdomain (what PCAEncoder.train does with Data lists)
        b = x.batch
        ptr = b.node_ptr.cpu().numpy()
        xs = x.x_dev.cpu().numpy()
        feats = np.stack([xs[ptr[s]:ptr[s] + 280].reshape(-1) for s in range(b.n_sub)])
        enc.model.fit(feats)
        enc._save_model("logs/models/collection_c")
        clf.train(enc.get_latent_space(x), save_model=True, path="logs/models/collection_c")
    sched = sg.GNNPartitionScheduler("c", a.clusters, ds, model, train=False, encoder=enc, classifier=clf)

    def run():
        p, r, mi, wl = sched.predict(x)
        out = ds.reconstruct_from_partition(p, r, 0, mi, wl)
        return p, mi, wl, out

    def timed():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run()
        if world > 1 and sg._dist()[2] > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(a.steps):
            res = run()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / a.steps, res

    ms_multi, (p, mi, wl, out) = timed()
    field = out.field.clone()
    w = torch.stack([t[0] for t in wl]).clone()
    real = sg._dist
    sg._dist = lambda: (None, 0, 1)                 # (b): this rank alone
    x.batch.__dict__.pop("_shards", None)
    ms_single, (p1, mi1, wl1, out1) = timed()
    sg._dist = real
    field1 = out1.field
    w1 = torch.stack([t[0] for t in wl1])
    err = float((field - field1).norm() / field1.norm())
    werr = float((w - w1).abs().max() / w1.abs().max().clamp(min=1e-30))
    ok = err <= 1e-6 and werr <= 1e-5 and np.array_equal(mi, mi1)
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        cells = ds._mesh(0)["mesh"].num_cells
        print(json.dumps({"check": "alds_multi", "ok": bool(flag.item()), "n_gpus": world, "cells": cells,
                          "subdomains": x.batch.n_sub, "clusters": a.clusters, "model": a.model,
                          "precision": a.precision, "labels_hist": np.bincount(mi, minlength=a.clusters).tolist(),
                          "field_rel_l2_vs_single_rank": err, "weight_rel_err": werr,
                          "ms_per_predict_sharded": ms_multi, "ms_per_predict_single_rank": ms_single,
                          "cells_per_s_sharded": cells / (ms_multi / 1e3)}))
    if world > 1:
        dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
