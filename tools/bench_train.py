#!/usr/bin/env python
"""Train-step throughput (BASELINE config 4: forward + MSE + backward + gradient all-reduce + Adam).

    python tools/bench_train.py [--mesh-n 28] [--precision tf32] [--steps 10]
    torchrun --nproc-per-node N tools/bench_train.py ...

All local subdomains of the rank's shard are one block-diagonal batch (the MSE is a mean over
nodes x channels, so the loss definition is unchanged); the mesh is fixed, the ranks share its subdomains (strong).
Prints one JSON line (cells/s over all ranks, max-over-ranks CUDA-event time).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", type=int, default=28)
    ap.add_argument("--levels", type=int, default=7)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--model", default="neuralop")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    from bench import load_weights, make_mesh
    from fesr_b200 import _lib, ops
    from fesr_b200.models.model import KernelNN, TEECNet
    from fesr_b200.models.training import FlatAdam, train_step
    from fesr_b200.pipeline import make_shard, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    levels = a.levels
    mesh = make_mesh(a.mesh_n)
    part, batch = ops.assemble(torch.from_numpy(mesh.pos).to(dev), torch.from_numpy(mesh.cells).to(dev), levels)
    bounds = shard_bounds(batch.edge_ptr.cpu().numpy(), world)
    sh = make_shard(batch, bounds[rank], bounds[rank + 1])
    x = torch.from_numpy(mesh.x).to(dev)[sh.global_ids]
    y = torch.from_numpy(mesh.y).to(dev)[sh.global_ids]
    model = (KernelNN(43, 43, 5, in_width=4, out_width=4) if a.model == "neuralop" else
             TEECNet(4, 43, 4, num_layers=5, retrieve_weight=False))
    model.load_state_dict(load_weights(a.model))
    model = model.to(dev).train()
    model.precision = a.precision
    opt = FlatAdam(model, lr=5e-4)
    for _ in range(a.warmup):
        loss = train_step(model, opt, x, sh.csr, sh.edge_attr, y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if a.profile:
        _lib.profile_enable(True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        loss = train_step(model, opt, x, sh.csr, sh.edge_attr, y)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        line = {"metric": "train-step mesh cells/s (forward + MSE + backward + all-reduce + Adam)",
                "value": mesh.num_cells / (ms / a.steps / 1e3), "unit": "cells/s", "n_gpus": world, "steps": a.steps,
                "ms_per_step": ms / a.steps, "loss": float(loss), "precision": a.precision, "model": a.model,
                "cells": mesh.num_cells, "nodes_batch": batch.n_tot, "edges_batch": batch.e_tot,
                "mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        if a.profile:
            line["kernels_ms_per_step"] = {k: v[0] / a.steps for k, v in _lib.profile_collect().items() if v[1]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
