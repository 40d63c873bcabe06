"""GNNPartitionScheduler / GradientbasedLoss with the reference's interface
(models/scheduler_gnn.py:23-514), executing on one B200 per process.

What changes against the reference:
  * predict() runs ALL subdomains of the rank's shard as one block-diagonal batch (the
    reference loops subdomains one at a time with two PCIe copies each, :217-226, :328-340);
  * multi-GPU is one process per GPU (torchrun): contiguous edge-balanced shards and ONE NCCL
    all-gather of the predictions, instead of mp.Process + Manager().dict() pickling (:254-291);
  * every cluster gets its own model copy (the reference appends the same module object for
    every cluster, :42-51, so its clusters alias the last checkpoint);
  * train() is the DDP branch of the reference (:349-469: MSELoss, Adam, StepLR stepped on
    validation epochs) with the gradient all-reduce issued on one flat buffer.
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops
from ..dataset.GraphDataset import SubdomainSample


class TensorList(list):
    """list of per-subdomain CPU tensors that remembers the device-resident concatenation.  The host buffer
    behind the elements may still be filling (asynchronous device -> host copy on a side stream): the first
    element access waits for it.  On a multi-rank run only this rank's own subdomains (`own` = index range) are
    copied eagerly; the first access to any other element fetches the rest (`rest`, a one-shot callable)."""
    padded = None         # (rows view [world*slot/c, c] of the gathered buffer, int64 row position of every batch row)
    ready = None          # torch.cuda.Event recorded after the (own part of the) device -> host copy, or None
    own = None            # (s0, s1): the subdomains that `ready` covers; None = all of them
    rest = None           # callable that copies everything else and blocks until it is there; None = nothing left

    def __init__(self, items=None, make=None, n=0):
        """`make` (with `n` = the length): the per-subdomain views are built on first access -- slicing a host buffer
        into a thousand tensors costs milliseconds of host time that a caller who only hands the list on to
        reconstruct_from_partition (which works from `.dev`) never needs to spend."""
        super().__init__(items if items is not None else ())
        self._make, self._n = make, n
        self._dev = None

    @property
    def dev(self):
        """Device-resident concatenation [sum n_s, c] in subdomain order.  After a sharded predict the rows live in
        the padded all-gather buffer (`padded`); the contiguous copy is only made if somebody asks for it."""
        if self._dev is None and self.padded is not None:
            rows, pos = self.padded
            self._dev = rows.index_select(0, pos)
        return self._dev

    @dev.setter
    def dev(self, t):
        self._dev = t

    def _fill(self):
        if self._make is not None:
            mk, self._make = self._make, None
            list.extend(self, mk())

    def __len__(self):
        return self._n if self._make is not None else list.__len__(self)

    ovf_host = None       # pinned int32 [1]: the forward's fp16 range flag, copied with the results

    def wait(self):
        """Blocks until this rank's own part is on the host.  Raises if the forward that produced it left the fp16
        range (its output is NaN-filled then: fesr.h, fesr_forward_overflow_offset)."""
        if self.ready is not None:
            self.ready.synchronize()
            self.ready = None
            if self.ovf_host is not None and int(self.ovf_host[0]) != 0:
                raise _lib.FesrError("fp16 overflow in the f16 / tf32 arm: an intermediate exceeded 65504 and the output "
                                     "was discarded (NaN).  Normalise the inputs as the reference does "
                                     "(dataset/GraphDataset.py:962-976) or run precision='fp32'.")

    def wait_all(self):
        if self.rest is not None:
            self.rest()
            self.rest = None
        self.wait()

    def __getitem__(self, i):
        self._fill()
        if self.rest is not None:
            j = i + len(self) if isinstance(i, int) and i < 0 else i
            if not (isinstance(j, int) and self.own is not None and self.own[0] <= j < self.own[1]):
                self.wait_all()
        self.wait()
        return super().__getitem__(i)

    def __iter__(self):
        self._fill()
        self.wait_all()
        return super().__iter__()


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class GradientbasedLoss(nn.Module):
    """models/scheduler_gnn.py:472-514, on the device."""

    def __init__(self, max_weight=1.0):
        super().__init__()
        self.max_weight = max_weight

    @staticmethod
    def _csr(edge_index, n):
        return edge_index if isinstance(edge_index, ops.Csr) else ops.csr_build(edge_index, n)

    def forward(self, pred, data, edge_index, edge_attr):
        csr = self._csr(edge_index, pred.shape[0])
        s = ops.node_weight(pred, data, csr, edge_attr, None, clamp_max=float(self.max_weight))
        loss, _ = ops.mse_loss(pred, data, want_grad=False)
        return (loss * s).squeeze(0)

    def compute_node_weight(self, pred, data, edge_index, edge_attr, num_nodes):
        csr = self._csr(edge_index, num_nodes)
        s = ops.node_weight(pred, data, csr, edge_attr, None)
        return s.expand(num_nodes).contiguous()


def _as_batch(x, device, between=None):
    """list[Data] -> (csr, edge_attr, node_ptr, x_dev, y_dev, sizes) as one block-diagonal graph.
    between: called after the copy of x has been ISSUED and before the copy of y (host inputs only) -- the caller
    launches the part of the pass that does not read x there, so that the copy engine and the SMs start together."""
    if isinstance(x, SubdomainSample):
        b = x.batch
        if getattr(b, "_sizes", None) is None:
            b._sizes = np.diff(b.node_ptr.cpu().numpy()).tolist()
        if x.x_host is not None:          # new input / reference fields arriving from the host
            # both copies ride on a side stream: x hides under the edge MLP (which does not read it: the model gets
            # `x_ready`), the reference field -- only needed by the node weight at the very end -- under the layers
            main, side = torch.cuda.current_stream(device), _side_stream(device)
            with torch.cuda.stream(side):
                x_dev = x.x_host.to(device, non_blocking=True)
                x_dev.ready = side.record_event()
            if between is not None:
                between()
            with torch.cuda.stream(side):
                y_dev = x.y_host.to(device, non_blocking=True)
                y_dev.ready = side.record_event()
            x_dev.record_stream(main)
            y_dev.record_stream(main)
            return b.csr, b.edge_attr, b.node_ptr, x_dev, y_dev, b._sizes
        return b.csr, b.edge_attr, b.node_ptr, x.x_dev, x.y_dev, b._sizes
    sizes = [int(d.x.shape[0]) for d in x]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    xs = torch.cat([d.x for d in x]).to(device, dtype=torch.float32)
    ys = torch.cat([d.y for d in x]).to(device, dtype=torch.float32)
    ei = torch.cat([d.edge_index.to(torch.int64) + int(o) for d, o in zip(x, offs[:-1])], dim=1).to(device)
    ea = torch.cat([d.edge_attr.reshape(-1) for d in x]).to(device, dtype=torch.float32)
    csr = ops.csr_build(ei, int(offs[-1]))
    node_ptr = torch.from_numpy(offs.astype(np.int32)).to(device)
    return csr, ea, node_ptr, xs, ys, sizes


_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


def select_subdomains(csr, edge_attr, node_ptr, keep_sub):
    """Block-diagonal sub-batch of the subdomains flagged in keep_sub [S] (device bool)."""
    dev = node_ptr.device
    S = node_ptr.numel() - 1
    sizes = (node_ptr[1:] - node_ptr[:-1]).long()
    sub_of_node = torch.repeat_interleave(torch.arange(S, device=dev), sizes)
    node_keep = keep_sub[sub_of_node]
    new_index = torch.cumsum(node_keep.to(torch.int64), 0) - 1
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    if csr.perm is not None:
        edge_attr = edge_attr.reshape(-1)[csr.perm.long()]
    dst = torch.repeat_interleave(torch.arange(csr.n, device=dev), deg)
    edge_keep = node_keep[dst]
    src_new = new_index[csr.src.long()[edge_keep]].to(torch.int32)
    rowptr_new = torch.zeros(int(node_keep.sum()) + 1, dtype=torch.int32, device=dev)
    rowptr_new[1:] = torch.cumsum(deg[node_keep], 0).to(torch.int32)
    node_ptr_new = torch.zeros(int(keep_sub.sum()) + 1, dtype=torch.int32, device=dev)
    node_ptr_new[1:] = torch.cumsum(sizes[keep_sub], 0).to(torch.int32)
    sub = ops.Csr(rowptr_new, src_new.contiguous(), None, int(rowptr_new.numel() - 1), int(src_new.numel()))
    return sub, edge_attr.reshape(-1)[edge_keep].contiguous(), node_ptr_new, node_keep


class GNNPartitionScheduler():
    def __init__(self, exp_name, num_partitons, dataset, model=None, train=True, encoder=None, classifier=None):
        self.name = exp_name
        self.num_partitions = num_partitons
        if num_partitons != 1:
            self.encoder = encoder
            self.classifier = classifier
        self.model = model
        self.dataset = dataset
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.subsets = self._train_partitions(num_partitons, train)
        if not train:
            self.models = self._load_models()

    def get_sub_dataset(self):
        return self.subsets

    def _model_dir(self):
        return 'logs/models/collection_{}'.format(self.name)

    def _initialize_model(self):
        return copy.deepcopy(self.model)

    def _load_models(self):
        models = []
        for i in range(self.num_partitions):
            model = self._initialize_model()
            path = os.path.join(self._model_dir(), 'partition_{}.pth'.format(i))
            model.load_state_dict(torch.load(path, map_location=torch.device('cpu'), weights_only=True))
            model = model.to(self.device).eval()
            if self.num_partitions > 1:
                model.ws_tag = f"fwd{i}"
            models.append(model)
        return models

    def _train_partitions(self, num_partitions, train):
        if num_partitions == 1:
            return [self.dataset]
        path = self._model_dir()
        if train:
            # every rank fits the same (deterministic) PCA / k-means on the same data; only rank 0 writes the files
            dist, rank, world = _dist()
            save = rank == 0
            if save:
                os.makedirs(path, exist_ok=True)
            self.encoder.train(self.dataset, save_model=save, path=path)
            latent_space = self.encoder.get_latent_space(self.dataset)
            self.classifier.train(latent_space, save_model=save, path=path)
            labels = self.classifier.cluster(latent_space)
            if world > 1:
                dist.barrier()
        else:
            self.encoder.load_model(path)
            self.classifier.load_model(path)
            latent_space = self.encoder.get_latent_space(self.dataset)
            labels = self.classifier.cluster(latent_space)
        subsets = []
        for i in range(num_partitions):
            idx = np.where(np.asarray(labels) == i)[0]
            print(f'Partition {i}: {len(idx)} samples')
            subsets.append([self.dataset[int(j)] for j in idx])
        return subsets

    # ----------------------------------------------------------------------------- predict
    def _route(self, x_dev, node_ptr):
        if self.num_partitions == 1:
            return torch.zeros(node_ptr.numel() - 1, dtype=torch.int32, device=x_dev.device)
        pca = self.encoder.model
        km, sc = self.classifier.model, self.classifier.scaler
        labels, _ = ops.route(x_dev, node_ptr, pca.mean_, pca.components_, sc.mean_, sc.scale_, km.cluster_centers_,
                              rows=self.encoder.min_length)
        return labels

    @torch.no_grad()
    def predict(self, x):
        """-> (pred_y_list, ref_y_list, model_idx, weights_list), as models/scheduler_gnn.py:228,311."""
        if not hasattr(self, 'models'):
            raise ValueError('Models are not trained yet')
        dev = self.device
        dist, rank, world = _dist()
        if world > 1:
            from .. import comm
            comm.init_from_torch_distributed()          # idempotent: libfesr's own NCCL communicator
        if world > 1 and self.num_partitions == 1 and isinstance(x, SubdomainSample):
            return self._predict_sharded(x, rank, world)
        between = None
        if self.num_partitions == 1 and isinstance(x, SubdomainSample) and x.x_host is not None:
            # inputs still on the host: the copy of x is issued first (side stream), then the part of the pass that does
            # not read it (weight preparation + edge MLP) -- the copy engine and the SMs start together; issuing the edge
            # MLP first left x arriving ~70 us after it had finished (tools/dev/e2e_timeline.py)
            between = lambda: self.models[0].edge_phase(x.batch.csr, x.batch.edge_attr)
        csr, edge_attr, node_ptr, x_dev, y_dev, sizes = _as_batch(x, dev, between)
        y_ready, x_ready = getattr(y_dev, "ready", None), getattr(x_dev, "ready", None)
        S = len(sizes)
        if self.num_partitions == 1 and world == 1:          # one model, every subdomain: nothing to select
            labels = self._route(x_dev, node_ptr)
            pred = self.models[0](x_dev, csr, edge_attr, x_ready=x_ready)
            if y_ready is not None:
                torch.cuda.current_stream(dev).wait_event(y_ready)
            weight_s = ops.node_weight(pred, y_dev, csr, edge_attr, node_ptr)
            return self._to_host_lists(x, pred, weight_s, y_dev, sizes, labels)

        if x_ready is not None:
            torch.cuda.current_stream(dev).wait_event(x_ready)
        labels = self._route(x_dev, node_ptr)
        # Routed and / or sharded: the plan -- this rank's contiguous edge-balanced share of the subdomain list and,
        # per cluster, the block-diagonal sub-batch of its subdomains -- depends only on the labels, so it is kept
        # on the device batch and rebuilt only when the routing of this sample changes.  The labels come back from the
        # device asynchronously: the pass is issued with the cached plan while they travel and is checked against them
        # afterwards (a mid-pass `labels.cpu()` left the GPU idle behind the host on every call); only a changed routing
        # -- or the first call -- pays the synchronous path.
        holder = x.batch.__dict__ if isinstance(x, SubdomainSample) else {}
        plan = holder.get("_alds_plan")
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        labels_pin = torch.empty(S, dtype=torch.int32, pin_memory=True)
        routed = main.record_event()
        with torch.cuda.stream(side):
            side.wait_event(routed)
            labels_pin.copy_(labels, non_blocking=True)
            labels_ready = side.record_event()
        labels.record_stream(side)
        usable = plan is not None and plan["world"] == world and plan["rank"] == rank and plan["k"] == self.num_partitions
        if not usable:
            labels_ready.synchronize()
            plan = self._routing_plan(csr, edge_attr, node_ptr, labels_pin.numpy().copy(), rank, world)
            holder["_alds_plan"] = plan
        if y_ready is not None:
            main.wait_event(y_ready)
        out_ch = self.models[0].dims.out_ch

        def run(plan):
            if world > 1:
                return self._run_routed_sharded(plan, x_dev, y_dev, out_ch, rank, world)
            pred = torch.zeros(csr.n, out_ch, dtype=torch.float32, device=dev)
            weight_s = torch.zeros(S, dtype=torch.float32, device=dev)
            for i, c in plan["clusters"]:
                # (torch's index_select / advanced indexing take 84 us per 2.5 MB of 16-byte rows on a B200, whatever the
                # index type -- tools/dev/gather_bench.py; these are 16-byte-row copies at HBM rate)
                xi, yi = ops.gather_rows(x_dev, c["nodes"]), ops.gather_rows(y_dev, c["nodes"])
                pi = self.models[i](xi, c["csr"], c["edge_attr"])
                wi = ops.node_weight(pi, yi, c["csr"], c["edge_attr"], c["node_ptr"])
                ops.scatter_rows(pi, c["nodes"], pred)
                weight_s.index_copy_(0, c["subs"], wi)
            return pred, weight_s

        pred, weight_s = run(plan)
        labels_ready.synchronize()                      # long since there: the copy was issued before the pass
        labels_h = labels_pin.numpy()
        if not np.array_equal(plan["labels"], labels_h):
            # the routing of this sample differs from the cached plan's: what was just issued is discarded.  Every rank
            # sees the same labels (same inputs), so all of them take this branch together (the all-gather inside pairs up)
            plan = self._routing_plan(csr, edge_attr, node_ptr, labels_h.copy(), rank, world)
            holder["_alds_plan"] = plan
            pred, weight_s = run(plan)
        # several ranks: every rank holds the complete arrays on its device after the gather; the host lists are filled on
        # first access (N full device -> host copies through one host were most of the step at 8 ranks)
        own = (0, 0, 0, 0) if world > 1 else None
        return self._to_host_lists(x, pred, weight_s, y_dev, sizes, labels_h.copy(), own=own)

    def _run_routed_sharded(self, plan, x_dev, y_dev, out_ch, rank, world):
        """ALDS on several ranks.  The plan gives every rank a CLUSTER-MAJOR share of the subdomains (sorted by label, cut
        into `world` edge-balanced chunks): a rank then runs one or two per-cluster models instead of all of them on
        slivers of every cluster (20 fused-layer launches of a few tiles each per rank at 8 ranks were launch-bound: the
        step was slower on 8 GPUs than on 4).  Each model writes its predictions and the node-weight kernel its weights
        straight into the rank's slot of the [world, slot] buffer (`out=`), ONE in-place all-gather on libfesr's
        communicator fills the other slots, and the complete arrays in subdomain order are one row gather away."""
        from .. import comm
        dev = self.device
        slot, oc = plan["slot"], out_ch
        gbuf = torch.empty(world, slot, dtype=torch.float32, device=dev)
        mine = gbuf[rank]
        nr, ns = plan["rows"][rank], plan["cnt"][rank]
        pred_v = mine[:nr * oc].view(nr, oc)
        w_v = mine[nr * oc:nr * oc + ns]
        for i, c in plan["clusters"]:
            xi, yi = ops.gather_rows(x_dev, c["nodes"]), ops.gather_rows(y_dev, c["nodes"])
            r0, k0 = c["row0"], c["sub0"]
            pi = pred_v[r0:r0 + int(c["nodes"].numel())]
            self.models[i](xi, c["csr"], c["edge_attr"], out=pi)
            ops.node_weight(pi, yi, c["csr"], c["edge_attr"], c["node_ptr"], out=w_v[k0:k0 + int(c["subs"].numel())])
        comm.allgatherv_pred(gbuf)
        rows = gbuf.view(-1, oc)
        pred_all = ops.gather_rows(rows, plan["pos"]) if oc % 4 == 0 else rows.index_select(0, plan["pos"])
        return pred_all, gbuf.view(-1).index_select(0, plan["wpos"])

    def _routing_plan(self, csr, edge_attr, node_ptr, labels_h, rank, world):
        dev = self.device
        S = node_ptr.numel() - 1
        node_ptr_h = node_ptr.cpu().numpy().astype(np.int64)
        sizes_h = np.diff(node_ptr_h)
        plan = {"world": world, "rank": rank, "k": self.num_partitions, "labels": labels_h.copy()}
        if world > 1:
            # cluster-major order of the subdomains, cut into `world` chunks balanced by edge count (pipeline.py)
            from ..pipeline import cluster_major_layout
            edge_cum = csr.rowptr[node_ptr.long()].cpu().numpy().astype(np.int64)
            oc = int(self.models[0].dims.out_ch)
            lay = cluster_major_layout(labels_h, sizes_h, np.diff(edge_cum), world, oc)
            sub_off, sub_idx = lay["sub_off"], lay["sub_idx"]
            mine = lay["sub_rank"] == rank
            plan.update(rows=lay["rows"], cnt=lay["cnt"], slot=lay["slot"], pos=torch.from_numpy(lay["row_pos"]).to(dev),
                        wpos=torch.from_numpy(lay["wpos"]).to(dev))
        else:
            sub_off = sub_idx = None
            mine = np.ones(S, dtype=bool)
        clusters = []
        for i in range(self.num_partitions):
            keep_h = (labels_h == i) & mine
            if not keep_h.any():
                continue
            keep = torch.from_numpy(keep_h).to(dev)
            sub, ea, nptr, node_keep = select_subdomains(csr, edge_attr, node_ptr, keep)
            c = {"csr": sub, "edge_attr": ea, "node_ptr": nptr, "nodes": node_keep.nonzero().squeeze(1),
                 "subs": keep.nonzero().squeeze(1)}
            if world > 1:      # the cluster's rows / weights are one contiguous run of the rank's slot (cluster-major order)
                first = int(np.flatnonzero(keep_h)[0])
                c["row0"], c["sub0"] = int(sub_off[first]), int(sub_idx[first])
            clusters.append((i, c))
        plan["clusters"] = clusters
        return plan

    def _predict_sharded(self, x, rank, world):
        """One model, several ranks, device-resident decomposition (reference fan-out / fan-in:
        models/scheduler_gnn.py:254-311).  Every rank runs its contiguous edge-balanced share of the subdomains (the
        shard's CSR slice and the slot layout are built once per (mesh, world) and cached on the batch) and copies
        only ITS rows of the input field host -> device.  The forward writes its predictions, and the node-weight
        kernel its subdomain weights, straight into this rank's slot of one [world, slot] buffer; ONE in-place
        fesr_allgatherv_pred on the compute stream (libfesr's NCCL communicator) fills the other slots -- no
        pad / cat / clone, no torch.distributed call.  With host inputs the rank's rows of the reference field ride
        in the same slot (the stitch of ref_y_list needs every rank's rows; they arrive over NVLink, not PCIe)."""
        from .. import comm
        from ..pipeline import SlotLayout, make_shard, shard_bounds
        dev = self.device
        b = x.batch
        model = self.models[0]
        oc = int(model.dims.out_ch)
        host_in = x.x_host is not None
        cache = b.__dict__.setdefault("_shards", {})
        key = (world, rank, oc, host_in)
        c = cache.get(key)
        if c is None:
            node_ptr_h = b.node_ptr.cpu().numpy().astype(np.int64)
            bounds = shard_bounds(b.edge_ptr.cpu().numpy().astype(np.int64), world)
            rows = [int(node_ptr_h[bounds[r + 1]] - node_ptr_h[bounds[r]]) for r in range(world)]
            cnt = [bounds[r + 1] - bounds[r] for r in range(world)]
            lay = SlotLayout(rows, cnt, oc, with_ref=host_in)
            batch_rows = torch.arange(b.n_tot, device=dev)
            c = cache[key] = {"shard": make_shard(b, bounds[rank], bounds[rank + 1]), "lay": lay,
                              "sizes": np.diff(node_ptr_h).tolist(), "b0": bounds[rank],
                              "pos_pred": lay.row_positions(batch_rows).long(),
                              "pos_ref": lay.row_positions(batch_rows, ref=True).long() if host_in else None,
                              "wpos": lay.weight_positions(dev)}
        sh, lay, sizes = c["shard"], c["lay"], c["sizes"]
        lo, hi = sh.node_lo, sh.node_hi
        nr, ns = hi - lo, lay.cnt[rank]
        gbuf = torch.empty(world, lay.slot, dtype=torch.float32, device=dev)
        mine = gbuf[rank]
        pred_v = mine[:nr * oc].view(nr, oc)
        w_v = mine[lay.weight_off(rank):lay.weight_off(rank) + ns]
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        y_ready = x_ready = None
        if host_in:
            with torch.cuda.stream(side):
                xi = x.x_host[lo:hi].to(dev, non_blocking=True)
                x_ready = side.record_event()
            if nr > 0:
                model.edge_phase(sh.csr, sh.edge_attr)       # does not read x: runs while the copy engine brings it
            with torch.cuda.stream(side):
                yi = x.y_host[lo:hi].to(dev, non_blocking=True)
                y_ready = side.record_event()
            xi.record_stream(main)
            yi.record_stream(main)
        else:
            xi, yi = x.x_dev[lo:hi], x.y_dev[lo:hi]
        if nr > 0:
            model(xi, sh.csr, sh.edge_attr, x_ready=x_ready, out=pred_v)
        if y_ready is not None:
            main.wait_event(y_ready)
        if nr > 0:
            ops.node_weight(pred_v, yi, sh.csr, sh.edge_attr, sh.node_ptr, out=w_v)
            if host_in:
                mine[nr * oc:2 * nr * oc].view(nr, oc).copy_(yi)
        comm.allgatherv_pred(gbuf)
        return self._sharded_lists(x, gbuf, c, rank, world, oc, host_in)

    def _sharded_lists(self, x, gbuf, c, rank, world, oc, host_in):
        """The 4-tuple of predict() over the gathered slot buffer.  Only this rank's rows of the predictions and its
        subdomain weights are copied to the host now (one pinned buffer, side stream); the rows the other ranks
        computed -- they hold them on their own hosts -- are fetched on first access."""
        dev = self.device
        lay, sizes, sh = c["lay"], c["sizes"], c["shard"]
        S = len(sizes)
        lo, hi, b0 = sh.node_lo, sh.node_hi, c["b0"]
        nr, ns = hi - lo, lay.cnt[rank]
        n_tot = int(lay.row_offs[-1])
        host = torch.empty(n_tot * oc + S, dtype=torch.float32, pin_memory=True)
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        produced = main.record_event()
        wo = lay.weight_off(rank)
        with torch.cuda.stream(side):
            side.wait_event(produced)
            ovf_host = self._copy_flag(side)
            if nr > 0:
                host[lo * oc:hi * oc].copy_(gbuf[rank, :nr * oc], non_blocking=True)
                host[n_tot * oc + b0:n_tot * oc + b0 + ns].copy_(gbuf[rank, wo:wo + ns], non_blocking=True)
            copied = side.record_event()
        gbuf.record_stream(side)

        def rest(host=host, gbuf=gbuf, side=side):
            with torch.cuda.stream(side):
                for r in range(world):
                    if r == rank or lay.rows[r] == 0:
                        continue
                    r0, r1 = int(lay.row_offs[r]), int(lay.row_offs[r + 1])
                    host[r0 * oc:r1 * oc].copy_(gbuf[r, :(r1 - r0) * oc], non_blocking=True)
                    s0, w0 = int(lay.sub_offs[r]), lay.weight_off(r)
                    host[n_tot * oc + s0:n_tot * oc + s0 + lay.cnt[r]].copy_(gbuf[r, w0:w0 + lay.cnt[r]], non_blocking=True)
            side.synchronize()

        state = {"rest": rest}

        def rest_once(state=state):          # the two lists share one buffer: whoever asks first fetches for both
            if state["rest"] is not None:
                state["rest"]()
                state["rest"] = None

        rows_view = gbuf.view(-1, oc)
        pred_cpu = host[:n_tot * oc].view(n_tot, oc)
        pred_y_list = TensorList(make=lambda: torch.split(pred_cpu, sizes), n=S)
        pred_y_list.padded = (rows_view, c["pos_pred"])
        pred_y_list.layout = lay
        pred_y_list.rank = rank
        pred_y_list.ready = copied
        pred_y_list.ovf_host = ovf_host
        pred_y_list.own = (b0, b0 + ns)
        pred_y_list.rest = rest_once
        if host_in:
            y_host = x.y_host
            ref_y_list = TensorList(make=lambda: torch.split(y_host, sizes), n=S)
            ref_y_list.padded = (rows_view, c["pos_ref"])
            ref_y_list.layout = lay
            ref_y_list.is_ref = True
        else:
            ref_y_list = TensorList(make=lambda: [d.y for d in x], n=S)
            ref_y_list.dev = x.y_dev
        w_cpu = host[n_tot * oc:]
        weights_list = TensorList(make=lambda: [w_cpu[s].expand(sizes[s]) for s in range(S)], n=S)
        weights_list.ready = copied
        weights_list.own = (b0, b0 + ns)
        weights_list.rest = rest_once
        weights_list.padded_flat = (gbuf.view(-1), c["wpos"])
        return pred_y_list, ref_y_list, np.zeros(S, dtype=int), weights_list

    def _copy_flag(self, side):
        """fp16 range flags of the models' last forwards -> one pinned int32 (max over the models), on `side`."""
        flags = [f for f in (ops.overflow_flag(m.dims, self.device, getattr(m, "ws_tag", "fwd")) for m in self.models)
                 if f is not None]
        if not flags:
            return None
        host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        if len(flags) == 1:
            host.copy_(flags[0], non_blocking=True)
        else:
            host.copy_(torch.stack(flags).max(0).values, non_blocking=True)
        return host

    def _to_host_lists(self, x, pred, weight_s, y_dev, sizes, labels, own=None):
        # one packed device -> host copy on the side stream: reconstruct_from_partition works from the device
        # copy (`.dev`), so the host lists only have to be complete when somebody reads them.  `own` =
        # (node_lo, node_hi, sub_lo, sub_hi) on a multi-rank run: only this rank's rows are copied now, the other
        # ranks' rows (they hold them on their own hosts) on first access.
        dev = self.device
        S = len(sizes)
        host = torch.empty(pred.numel() + S, dtype=torch.float32, pin_memory=True)    # (cached host allocator)
        packed = torch.cat([pred.reshape(-1), weight_s])
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        produced = main.record_event()
        rest = None
        with torch.cuda.stream(side):
            side.wait_event(produced)
            ovf_host = self._copy_flag(side)
            if own is None:
                host.copy_(packed, non_blocking=True)
            else:
                oc, np_ = pred.shape[1], pred.numel()
                host[own[0] * oc:own[1] * oc].copy_(packed[own[0] * oc:own[1] * oc], non_blocking=True)
                host[np_ + own[2]:np_ + own[3]].copy_(packed[np_ + own[2]:np_ + own[3]], non_blocking=True)

                def rest(host=host, packed=packed, side=side):
                    with torch.cuda.stream(side):
                        host.copy_(packed, non_blocking=True)
                    side.synchronize()
            copied = side.record_event()
        packed.record_stream(side)
        pred_cpu = host[:pred.numel()].view(pred.shape)
        pred_y_list = TensorList(make=lambda: torch.split(pred_cpu, sizes), n=S)
        pred_y_list.dev = pred
        pred_y_list.ready = copied
        pred_y_list.ovf_host = ovf_host
        if own is not None:
            state = {"rest": rest}

            def rest_once(state=state):          # the two lists share one buffer: whoever asks first fetches for both
                if state["rest"] is not None:
                    state["rest"]()
                    state["rest"] = None
            pred_y_list.own = (own[2], own[3])
            pred_y_list.rest = rest_once
        if isinstance(x, SubdomainSample) and x.y_host is not None:
            y_host = x.y_host
            ref_y_list = TensorList(make=lambda: torch.split(y_host, sizes), n=S)
        else:
            ref_y_list = TensorList(make=lambda: [d.y for d in x], n=S)
        ref_y_list.dev = y_dev
        w_cpu = host[pred.numel():]
        weights_list = TensorList(make=lambda: [w_cpu[s].expand(sizes[s]) for s in range(S)], n=S)
        weights_list.ready = copied
        if own is not None:
            weights_list.own = (own[2], own[3])
            weights_list.rest = rest_once
        if labels is None or self.num_partitions == 1:
            model_idx = np.zeros(S, dtype=int)
        else:      # (the routed path hands over the host copy it already has)
            model_idx = (labels if isinstance(labels, np.ndarray) else labels.cpu().numpy()).astype(int)
        return pred_y_list, ref_y_list, model_idx, weights_list

    # ----------------------------------------------------------------------------- train
    def train(self, train_config, subset_idx=None, start_from_pretrained=False):
        from .training import train_subsets
        subsets = self.subsets if subset_idx is None else [self.subsets[i] for i in subset_idx]
        models = self._load_models() if start_from_pretrained else None
        trained = train_subsets(self, subsets, train_config, models)
        self.models = trained
        return trained
