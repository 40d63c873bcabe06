"""Train step and epoch loop of the scheduler (reference models/scheduler_gnn.py:349-469, the
DistributedDataParallel branch -- the only training branch of the reference that runs, SURVEY.md
3.4e): PyG-style block-diagonal batches, MSELoss, Adam, StepLR stepped on validation epochs,
best-validation checkpointing.  One process per GPU; the gradient all-reduce is ONE NCCL call
on a flat fp32 buffer (DDP's single bucket for a 0.3-1 MB model) issued from here instead of DDP.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class FlatAdam:
    """torch.optim.Adam (default betas/eps, no weight decay -- scheduler_gnn.py:391) over one flat
    buffer: parameters become views of it, gradients are packed, averaged across ranks with one
    all-reduce and applied by one fesr_adam_step launch."""

    def __init__(self, model, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)])
        total = int(self.offsets[-1])
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets[:-1]):
            self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + p.numel()].view_as(p)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.t = 0
        # gradient accumulators shaped like model.param_tensors(): views of the flat gradient buffer, so that
        # fesr_nnconv_backward adds straight into it (no per-parameter pack copies before the all-reduce)
        where = {id(p): (int(o), p) for p, o in zip(self.params, self.offsets[:-1])}

        def view_of(t):
            if t is None:
                return None
            o, p = where[id(t)]
            return self.grad[o:o + p.numel()].view_as(p)

        pt = model.param_tensors() if hasattr(model, "param_tensors") else None
        self.grad_views = None
        if pt is not None and all(id(t) in where for v in pt.values() if v is not None
                                  for t in (v if isinstance(v, (list, tuple)) else [v])):
            self.grad_views = {k: (None if v is None else [view_of(t) for t in v] if isinstance(v, (list, tuple)) else view_of(v))
                               for k, v in pt.items()}
        dist, _, world = _dist()
        if world > 1:
            # DistributedDataParallel broadcasts rank 0's parameters when it wraps the model (scheduler_gnn.py:386);
            # every rank must start from the same point or the averaged gradients are taken at different points
            from .. import comm
            comm.init_from_torch_distributed()
            dist.broadcast(self.flat, src=0)
            ops.invalidate_prepared_weights()

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def pack_grads(self):
        for p, o in zip(self.params, self.offsets[:-1]):
            seg = self.grad[o:o + p.numel()]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))

    def step(self, packed: bool = False):
        """packed=True: the gradients are already in the flat buffer (train_step's direct path)."""
        if not packed:
            self.pack_grads()
        dist, _, world = _dist()
        if world > 1:
            from .. import comm
            comm.allreduce_grads(self.grad)          # DDP averages: sum / world, on the compute stream
        self.t += 1
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.t, self.betas[0],
                      self.betas[1], self.eps)


def train_step(model, opt: FlatAdam, x, graph, edge_attr, y):
    """zero_grad -> forward -> MSELoss -> backward -> (all-reduce) -> Adam.step, :402-409."""
    from .. import _lib
    if opt.grad_views is None or not isinstance(graph, ops.Csr):
        opt.zero_grad()
        out = model(x, graph, edge_attr)
        loss, grad = ops.mse_loss(out.detach(), y)
        out.backward(grad)
        opt.step()
        return loss
    # direct path: forward / backward are one C-ABI call each and the parameter gradients land in the flat buffer
    opt.grad.zero_()
    prec = _lib.PRECISIONS[model.precision]
    tensors = {k: (None if v is None else v.detach() if torch.is_tensor(v) else [t.detach() for t in v])
               for k, v in model.param_tensors().items()}
    out, ws = ops.nnconv_forward(model.dims, tensors, x, graph, edge_attr, prec, keep_for_backward=True)
    loss, grad = ops.mse_loss(out, y)
    ops.nnconv_backward(model.dims, tensors, x, graph, edge_attr, prec, grad, ws, grads=opt.grad_views)
    opt.step(packed=True)
    return loss


def _collate(datas, device):
    """PyG Batch collation of a list of subdomain Data (scheduler_gnn.py:376-381)."""
    from .scheduler_gnn import _as_batch
    csr, ea, _, xs, ys, _ = _as_batch(list(datas), device)
    return xs, csr, ea, ys


def train_subsets(scheduler, subsets, train_config, pretrained=None):
    dist, rank, world = _dist()
    dev = scheduler.device
    trained = []
    for i, subset in enumerate(subsets):
        model = (pretrained[i] if pretrained else scheduler._initialize_model()).to(dev).train()
        n = len(subset)
        gen = torch.Generator().manual_seed(0)
        perm = torch.randperm(n, generator=gen).tolist()            # random_split(subset, [0.8 n, rest]), :102-105
        n_train = int(0.8 * n)
        train_idx, val_idx = perm[:n_train], perm[n_train:]
        if world > 1:                                                # contiguous per-rank slices, remainder dropped, :369-374
            nt, nv = len(train_idx) // world, len(val_idx) // world
            train_idx = train_idx[rank * nt:(rank + 1) * nt]
            val_idx = val_idx[rank * nv:(rank + 1) * nv]
        bs = int(train_config['batch_size'])
        opt = FlatAdam(model, lr=train_config['lr'])
        lr0, step_size, gamma = float(train_config['lr']), int(train_config['step_size']), float(train_config['gamma'])
        sched_steps = 0
        best = np.inf
        path = os.path.join(scheduler._model_dir(), f'partition_{i}.pth')
        cache = {}

        def batch_of(idxs):
            key = tuple(idxs)
            if key not in cache:
                if len(cache) > 64:
                    cache.clear()
                cache[key] = _collate([subset[j] for j in idxs], dev)
            return cache[key]

        for epoch in range(int(train_config['epochs'])):
            model.train()
            order = [train_idx[j] for j in torch.randperm(len(train_idx), generator=gen).tolist()]   # shuffle=True
            losses = []
            for b0 in range(0, len(order), bs):
                xs, csr, ea, ys = batch_of(order[b0:b0 + bs])
                losses.append(train_step(model, opt, xs, csr, ea, ys))
            train_loss = float(torch.stack(losses).mean()) if losses else float('nan')
            if rank == 0:
                print(f'Epoch {epoch}: Train loss: {train_loss}')
            if epoch % int(train_config['val_interval']) == 0:
                model.eval()
                vl = []
                with torch.no_grad():
                    for b0 in range(0, len(val_idx), bs):
                        xs, csr, ea, ys = batch_of(val_idx[b0:b0 + bs])
                        vl.append(ops.mse_loss(model(xs, csr, ea), ys, want_grad=False)[0])
                val_loss = float(torch.stack(vl).mean()) if vl else float('nan')
                if rank == 0:
                    print(f'Epoch {epoch}: Validation loss: {val_loss}')
                if val_loss < best:
                    best = val_loss
                    if rank == 0:
                        os.makedirs(scheduler._model_dir(), exist_ok=True)
                        torch.save({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, path)
                sched_steps += 1                                      # StepLR.step() only here, :459
                opt.lr = lr0 * gamma ** (sched_steps // step_size)
        if rank == 0:
            os.makedirs(scheduler._model_dir(), exist_ok=True)
            torch.save({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, path)   # :467-468
        trained.append(model.eval())
    return trained
