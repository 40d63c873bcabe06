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


def cluster_major_layout(labels, sizes, edges, world: int, c: int):
    """Host-side layout of the routed (ALDS) predict on several ranks.  The subdomains are sorted by cluster label (stable)
    and that order is cut into `world` chunks balanced by edge count: a rank runs whole runs of one or two clusters instead
    of slivers of all of them (reference fan-out: equal-count contiguous chunks, models/scheduler_gnn.py:269-271).
    Slot r of the [world, slot] all-gather buffer holds rank r's prediction rows in that order ([rows[r], c]) followed by
    its subdomain weights ([cnt[r]]).  Returns a dict:
      sub_rank / sub_off / sub_idx [S]  owner rank, row offset inside the owner's slot, index among the owner's subdomains
      rows / cnt [world], slot          rows and subdomains per rank, floats per slot (a multiple of 4 c)
      row_pos [sum sizes]               row of the [world * slot / c, c] view of the buffer for every batch row
      wpos [S]                          float offset of every subdomain's weight in the flattened buffer"""
    labels = np.asarray(labels)
    sizes = np.asarray(sizes, dtype=np.int64)
    edges = np.asarray(edges, dtype=np.int64)
    S = int(labels.size)
    order = np.argsort(labels, kind="stable")
    bounds = shard_bounds(np.concatenate([[0], np.cumsum(edges[order])]), world)
    sub_rank = np.zeros(S, dtype=np.int64)
    sub_off = np.zeros(S, dtype=np.int64)
    sub_idx = np.zeros(S, dtype=np.int64)
    rows, cnt = [], []
    for r in range(world):
        subs_r = order[bounds[r]:bounds[r + 1]]
        sub_rank[subs_r] = r
        if subs_r.size:
            sub_off[subs_r] = np.concatenate([[0], np.cumsum(sizes[subs_r])[:-1]])
            sub_idx[subs_r] = np.arange(subs_r.size)
        rows.append(int(sizes[subs_r].sum()))
        cnt.append(int(subs_r.size))
    q = 4 * c
    slot = max(q, (max(r_ * c + k_ for r_, k_ in zip(rows, cnt)) + q - 1) // q * q)
    node_ptr = np.concatenate([[0], np.cumsum(sizes)])
    row_pos = np.repeat(sub_rank * (slot // c) + sub_off - node_ptr[:-1], sizes) + np.arange(int(node_ptr[-1]))
    wpos = sub_rank * slot + np.asarray(rows, dtype=np.int64)[sub_rank] * c + sub_idx
    return {"sub_rank": sub_rank, "sub_off": sub_off, "sub_idx": sub_idx, "rows": rows, "cnt": cnt, "slot": int(slot),
            "row_pos": row_pos, "wpos": wpos}


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


class SlotLayout:
    """Padded layout of the one all-gather of a sharded predict pass (fesr_allgatherv_pred): a [world, slot] fp32
    buffer whose slot r holds, back to back, rank r's predictions [rows[r], c], (optionally) its rows of the
    reference field [rows[r], c] and its subdomain weights [cnt[r]].  `slot` is a multiple of 4 c floats, so the
    buffer is also a [world * slot / c, c] row array and a row of any rank has a ROW position in it: the stitch
    reads the gathered buffer in place through an occurrence index remapped once to those positions."""

    def __init__(self, rows, cnt, c: int, with_ref: bool):
        self.rows, self.cnt, self.c, self.with_ref = [int(r) for r in rows], [int(k) for k in cnt], int(c), bool(with_ref)
        self.world = len(self.rows)
        per = [r * self.c * (2 if with_ref else 1) + k for r, k in zip(self.rows, self.cnt)]
        q = 4 * self.c
        self.slot = max(q, (max(per) + q - 1) // q * q)
        self.row_offs = np.concatenate([[0], np.cumsum(self.rows)]).astype(np.int64)      # concatenated row index
        self.sub_offs = np.concatenate([[0], np.cumsum(self.cnt)]).astype(np.int64)

    # float offsets inside slot r
    def pred_off(self, r):
        return 0

    def ref_off(self, r):
        return self.rows[r] * self.c

    def weight_off(self, r):
        return self.rows[r] * self.c * (2 if self.with_ref else 1)

    def row_positions(self, idx: torch.Tensor, ref: bool = False) -> torch.Tensor:
        """Row index into the rank-order concatenation of the shards -> row index into the [world*slot/c, c] view of
        the gathered buffer (int32); ref=True: the position of the same row of the reference block."""
        offs = torch.as_tensor(self.row_offs, device=idx.device)
        i = idx.long()
        r = torch.bucketize(i, offs[1:], right=True)
        pos = i - offs[r] + r * (self.slot // self.c)
        if ref:
            pos = pos + torch.as_tensor(np.asarray(self.rows, dtype=np.int64), device=idx.device)[r]
        return pos.to(torch.int32).contiguous()

    def weight_positions(self, device) -> torch.Tensor:
        """Float offset in the flattened gathered buffer of every subdomain's weight, in subdomain order (int64)."""
        out = np.concatenate([r * self.slot + self.weight_off(r) + np.arange(self.cnt[r], dtype=np.int64)
                              for r in range(self.world)]) if self.world else np.zeros(0, np.int64)
        return torch.as_tensor(out, device=device)


def padded_positions(idx: torch.Tensor, rows) -> torch.Tensor:
    """Row index into the rank-order concatenation of the shards -> row index into the padded [world, max(rows)]
    all-gather buffer (int32)."""
    mx = max(rows)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(rows)]), dtype=torch.int64, device=idx.device)
    i = idx.long()
    r = torch.bucketize(i, offs[1:], right=True)
    return (i + r * mx - offs[r]).to(torch.int32).contiguous()


def node_slice(N: int, rank: int, world: int):
    """The contiguous range of global mesh nodes whose stitched values rank `rank` produces and keeps."""
    return (N * rank) // world, (N * (rank + 1)) // world


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
        if world > 1:
            from . import comm
            comm.init_from_torch_distributed(group)          # libfesr's own NCCL communicator (idempotent)

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

    def _layout(self, c: int):
        """Slot layout + remapped occurrence table of the sharded step (built once per channel count)."""
        if getattr(self, "_lay", None) is None or self._lay.c != c:
            lay = SlotLayout(self.shard_rows, [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)], c, False)
            self._lay = lay
            self._occ_padded = ops.Occurrence(self.occ.occ_ptr, lay.row_positions(self.occ.occ_idx), self.occ.N,
                                              self.world * lay.slot // c)
        return self._lay

    def step(self, x_shard, y_shard=None, full_field: bool = False):
        """forward (+ node weight) + all-gather + stitch -> (field, weights [S_shard] | None, pred [n_shard, c]).
        One rank: field is the whole stitched mesh field [N, c].  Several ranks: the forward writes straight into
        this rank's slot of the gather buffer, ONE in-place fesr_allgatherv_pred (libfesr's communicator, on the
        compute stream) fills the other slots, and the rank stitches ITS slice of the mesh nodes (`node_slice`; full_field=True: all of them --
        every rank then holds the bit-identical whole field)."""
        if self.world == 1:
            pred = self.forward_shard(x_shard)
            w = self.node_weight(pred, y_shard) if y_shard is not None else None
            field, _, _ = self.stitch(pred)
            return field, w, pred
        from . import comm
        c = int(self.model.dims.out_ch)
        lay = self._layout(c)
        dev = x_shard.device
        gbuf = torch.empty(self.world, lay.slot, dtype=torch.float32, device=dev)
        rows = lay.rows[self.rank]
        pred = gbuf[self.rank, :rows * c].view(rows, c)
        with torch.no_grad():
            self.model(x_shard, self.shard.csr, self.shard.edge_attr, out=pred)
        w = self.node_weight(pred, y_shard) if y_shard is not None else None
        comm.allgatherv_pred(gbuf)          # in place, on the compute stream
        rng = None if full_field else node_slice(self.N, self.rank, self.world)
        field, _, _ = ops.stitch_mean(gbuf.view(-1, c), self._occ_padded, None, want_merged=False, want_count=False,
                                      node_range=rng)
        return field, w, pred
