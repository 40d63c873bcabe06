"""Resident predict pipeline: subdomain batch -> model forward -> node weight -> overlap stitch.

This is the device-side core that ``GNNPartitionScheduler.predict`` + ``reconstruct_from_partition``
drive (reference models/scheduler_gnn.py:204-228 and dataset/GraphDataset.py:1308-1409): all
subdomains of a rank's shard run as ONE block-diagonal graph (no per-subdomain launches, no
per-subdomain PCIe copies), and the stitch is one segmented mean keyed by global node id.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, ops


@dataclass
class Shard:
    """A contiguous range of subdomains [s0, s1) re-based to its own block-diagonal batch."""
    s0: int
    s1: int
    node_lo: int
    node_hi: int
    csr: ops.Csr
    edge_attr: torch.Tensor
    node_ptr: torch.Tensor      # [s1-s0+1] int32, re-based
    global_ids: torch.Tensor    # [n_shard] int64


def shard_bounds(edge_ptr_host: np.ndarray, world: int):
    """Contiguous split of the subdomain list into `world` chunks balanced by edge count
    (the reference uses equal-count contiguous chunks, models/scheduler_gnn.py:269-271)."""
    S = edge_ptr_host.size - 1
    total = int(edge_ptr_host[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        s = int(np.searchsorted(edge_ptr_host, target, side="left"))
        s = min(max(s, bounds[-1]), S)
        bounds.append(s)
    bounds.append(S)
    return bounds


def all_gather_rows(local: torch.Tensor, rows, group=None) -> torch.Tensor:
    """All-gather of row blocks of different lengths (rank r contributes rows[r] rows): every rank
    pads its block to max(rows), one equal-size all-gather, then the blocks are concatenated in
    rank order.  Works on NCCL (GPU) and gloo (CPU tests)."""
    import torch.distributed as dist
    world = len(rows)
    if world == 1:
        return local
    mx = max(rows)
    tail = local.shape[1:]
    pad = local.new_zeros((mx, *tail))
    pad[:local.shape[0]] = local
    parts = [local.new_empty((mx, *tail)) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:rows[r]] for r in range(world)], dim=0)


def all_gather_packed(pred_local: torch.Tensor, w_local: torch.Tensor, rows, cnt, group=None):
    """Predictions [rows[r], c] and subdomain weights [cnt[r]] of every rank in ONE collective: each rank packs
    both into one flat fp32 buffer padded to the largest rank, `all_gather_into_tensor`, then the rank blocks
    are unpacked in rank (= subdomain) order.  -> (pred [sum rows, c], weights [sum cnt])."""
    import torch.distributed as dist
    world = len(rows)
    c = int(pred_local.shape[1])
    if world == 1:
        return pred_local, w_local
    mx = max(rows[r] * c + cnt[r] for r in range(world))
    rank = dist.get_rank(group)
    buf = pred_local.new_empty(world, mx)
    mine = buf[rank]
    np_, nw = rows[rank] * c, cnt[rank]
    mine[:np_] = pred_local.reshape(-1)
    mine[np_:np_ + nw] = w_local
    dist.all_gather_into_tensor(buf.view(-1), mine.clone(), group=group)
    pred = torch.cat([buf[r, :rows[r] * c] for r in range(world)]).view(-1, c)
    w = torch.cat([buf[r, rows[r] * c:rows[r] * c + cnt[r]] for r in range(world)])
    return pred, w


def padded_positions(idx: torch.Tensor, rows) -> torch.Tensor:
    """Row index into the rank-order concatenation of the shards -> row index into the padded [world, max(rows)]
    all-gather buffer (int32)."""
    mx = max(rows)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(rows)]), dtype=torch.int64, device=idx.device)
    i = idx.long()
    r = torch.bucketize(i, offs[1:], right=True)
    return (i + r * mx - offs[r]).to(torch.int32).contiguous()


def make_shard(batch: ops.SubdomainBatch, s0: int, s1: int) -> Shard:
    node_ptr_h = batch.node_ptr[[s0, s1]].cpu().numpy()
    edge_ptr_h = batch.edge_ptr[[s0, s1]].cpu().numpy()
    nlo, nhi = int(node_ptr_h[0]), int(node_ptr_h[1])
    elo, ehi = int(edge_ptr_h[0]), int(edge_ptr_h[1])
    if s0 == 0 and s1 == batch.n_sub:
        csr = batch.csr
        return Shard(s0, s1, nlo, nhi, csr, batch.edge_attr, batch.node_ptr, batch.global_ids)
    rowptr = (batch.rowptr[nlo:nhi + 1] - elo).contiguous()
    src = (batch.edge_src[elo:ehi] - nlo).contiguous()
    csr = ops.Csr(rowptr, src, None, nhi - nlo, ehi - elo)
    return Shard(s0, s1, nlo, nhi, csr, batch.edge_attr[elo:ehi].contiguous(),
                 (batch.node_ptr[s0:s1 + 1] - nlo).contiguous(), batch.global_ids[nlo:nhi].contiguous())


class MeshPredictor:
    """Holds one mesh's assembled subdomains on the device and runs predict + stitch."""

    def __init__(self, model, pos: torch.Tensor, cells: torch.Tensor, levels: int,
                 mode: int = _lib.ALL_INTERSECTING, rank: int = 0, world: int = 1, group=None):
        self.model = model
        self.rank, self.world, self.group = rank, world, group
        self.N = int(pos.shape[0])
        self.part, self.batch = ops.assemble(pos, cells, levels, mode)
        edge_ptr_h = self.batch.edge_ptr.cpu().numpy()
        self.bounds = shard_bounds(edge_ptr_h, world)
        self.shard = make_shard(self.batch, self.bounds[rank], self.bounds[rank + 1])
        self.occ = ops.occurrence_build(self.batch.global_ids, self.N)
        node_ptr_h = self.batch.node_ptr.cpu().numpy()
        self.shard_rows = [int(node_ptr_h[self.bounds[r + 1]] - node_ptr_h[self.bounds[r]]) for r in range(world)]
        self.home_cells = np.bincount(self.part.home_leaf.cpu().numpy(), minlength=self.batch.n_sub)

    def gather_inputs(self, field: torch.Tensor) -> torch.Tensor:
        """[N, c] mesh field -> [n_shard, c] per-subdomain copies (what the HDF5 store holds)."""
        return field.index_select(0, self.shard.global_ids)

    @torch.no_grad()
    def forward_shard(self, x_shard: torch.Tensor) -> torch.Tensor:
        return self.model(x_shard, self.shard.csr, self.shard.edge_attr)

    def node_weight(self, pred_shard, y_shard):
        return ops.node_weight(pred_shard, y_shard, self.shard.csr, self.shard.edge_attr, self.shard.node_ptr)

    def all_gather(self, pred_shard: torch.Tensor) -> torch.Tensor:
        """One NCCL all-gather(v) of the per-subdomain predictions (rank order = subdomain order)."""
        return all_gather_rows(pred_shard.contiguous(), self.shard_rows, self.group)

    def stitch(self, pred_all: torch.Tensor, want_merged=False):
        return ops.stitch_mean(pred_all, self.occ, self.batch.global_ids, want_merged=want_merged, want_count=False)

    def _padded_gather(self, pred_shard: torch.Tensor) -> torch.Tensor:
        """The all-gather of step(): every rank's block lands at [r, :rows[r]] of one persistent [world, max rows, c]
        buffer (one copy into the send block + one `all_gather_into_tensor`); the stitch then reads the padded
        buffer directly through an occurrence index remapped once to padded positions, so there is no pad / split /
        concatenate traffic around the collective."""
        import torch.distributed as dist
        c = int(pred_shard.shape[1])
        if getattr(self, "_gbuf", None) is None or self._gbuf.shape[2] != c:
            dev = pred_shard.device
            mx = max(self.shard_rows)
            self._gbuf = torch.zeros(self.world, mx, c, dtype=torch.float32, device=dev)
            self._send = torch.zeros(mx, c, dtype=torch.float32, device=dev)
            self._occ_padded = ops.Occurrence(self.occ.occ_ptr, padded_positions(self.occ.occ_idx, self.shard_rows),
                                              self.occ.N, self.world * mx)
        self._send[:pred_shard.shape[0]].copy_(pred_shard)
        work = dist.all_gather_into_tensor(self._gbuf.view(-1), self._send.view(-1), group=self.group, async_op=True)
        return self._gbuf.view(-1, c), work

    def step(self, x_shard, y_shard=None):
        """forward (+ node weight) + all-gather + stitch; returns (field [N,c], weights [S_shard] | None)."""
        pred = self.forward_shard(x_shard)
        if self.world > 1:
            # the collective runs on NCCL's stream underneath the node-weight kernels (they only need this rank's rows)
            gathered, work = self._padded_gather(pred)
            w = self.node_weight(pred, y_shard) if y_shard is not None else None
            work.wait()
            field, _, _ = ops.stitch_mean(gathered, self._occ_padded, None, want_merged=False, want_count=False)
            return field, w, pred
        w = self.node_weight(pred, y_shard) if y_shard is not None else None
        field, _, _ = self.stitch(pred)
        return field, w, pred
