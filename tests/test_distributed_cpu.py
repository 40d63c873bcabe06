"""CPU, world_size 2, gloo: the host-side sharding / gather / gradient-averaging logic of the
multi-GPU path (the kernels themselves need a B200 and are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fesr_b200.pipeline import SlotLayout, all_gather_packed, all_gather_rows, node_slice, padded_positions, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, edge_ptr, node_ptr, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = shard_bounds(edge_ptr, world)
        rows = [int(node_ptr[bounds[r + 1]] - node_ptr[bounds[r]]) for r in range(world)]
        full = torch.arange(int(node_ptr[-1]) * 4, dtype=torch.float32).reshape(-1, 4)     # "predictions"
        lo = int(node_ptr[bounds[rank]])
        mine = full[lo:lo + rows[rank]].clone()
        got = all_gather_rows(mine, rows)
        ok_gather = bool(torch.equal(got, full))
        # predictions + per-subdomain weights in one packed collective
        cnt = [bounds[r + 1] - bounds[r] for r in range(world)]
        w_full = torch.arange(len(node_ptr) - 1, dtype=torch.float32) + 0.5
        p2, w2 = all_gather_packed(mine, w_full[bounds[rank]:bounds[rank + 1]].clone(), rows, cnt)
        ok_gather = ok_gather and bool(torch.equal(p2, full)) and bool(torch.equal(w2, w_full))
        # gradient averaging as FlatAdam.step does it (sum / world)
        g = torch.full((10,), float(rank + 1))
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        g /= world
        ok_grad = bool(torch.allclose(g, torch.full((10,), (1 + world) / 2.0)))
        ret[rank] = (ok_gather, ok_grad, bounds)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_balance_and_cover():
    rng = np.random.default_rng(0)
    edges = rng.integers(1000, 20000, size=128)
    edge_ptr = np.concatenate([[0], np.cumsum(edges)])
    for world in (1, 2, 4, 8):
        b = shard_bounds(edge_ptr, world)
        assert b[0] == 0 and b[-1] == 128 and all(b[i] <= b[i + 1] for i in range(world))
        per = [edge_ptr[b[r + 1]] - edge_ptr[b[r]] for r in range(world)]
        assert max(per) - min(per) <= 2 * edges.max()
    # more ranks than subdomains: empty shards are allowed, nothing is lost
    b = shard_bounds(np.array([0, 5, 9]), 4)
    assert b[0] == 0 and b[-1] == 2 and len(b) == 5


def test_two_rank_gather_and_grad_average():
    world = 2
    rng = np.random.default_rng(1)
    sizes = rng.integers(300, 900, size=9)
    node_ptr = np.concatenate([[0], np.cumsum(sizes)])
    edge_ptr = np.concatenate([[0], np.cumsum(sizes * 12)])
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, edge_ptr, node_ptr, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok_gather, ok_grad, bounds = ret[r]
        assert ok_gather and ok_grad
    assert ret[0][2] == ret[1][2]


def test_padded_positions_address_the_gather_buffer():
    rows = [5, 0, 9, 3]
    mx, c = max(rows), 4
    full = torch.arange(sum(rows) * c, dtype=torch.float32).reshape(-1, c)
    buf = torch.full((len(rows), mx, c), -1.0)
    o = 0
    for r, n in enumerate(rows):
        buf[r, :n] = full[o:o + n]
        o += n
    idx = torch.randperm(sum(rows)).to(torch.int32)
    pos = padded_positions(idx, rows)
    assert pos.dtype == torch.int32
    assert torch.equal(buf.view(-1, c)[pos.long()], full[idx.long()])


def test_slot_layout_addresses_the_in_place_gather_buffer():
    """Host logic of fesr_allgatherv_pred's padded layout: every rank's predictions / reference rows / subdomain
    weights are found at the positions SlotLayout hands to the stitch and to the host copies."""
    for c, with_ref in ((4, False), (4, True), (3, True)):
        rows, cnt = [5, 0, 9, 3], [2, 0, 3, 1]
        lay = SlotLayout(rows, cnt, c, with_ref)
        assert lay.slot % (4 * c) == 0 and lay.slot >= max(r * c * (2 if with_ref else 1) + k for r, k in zip(rows, cnt))
        full = torch.arange(sum(rows) * c, dtype=torch.float32).reshape(-1, c)
        ref = -full - 1.0
        w = torch.arange(sum(cnt), dtype=torch.float32) + 0.25
        buf = torch.full((len(rows), lay.slot), float("nan"))
        for r in range(len(rows)):
            r0, r1 = int(lay.row_offs[r]), int(lay.row_offs[r + 1])
            buf[r, :rows[r] * c] = full[r0:r1].reshape(-1)
            if with_ref:
                buf[r, lay.ref_off(r):lay.ref_off(r) + rows[r] * c] = ref[r0:r1].reshape(-1)
            s0 = int(lay.sub_offs[r])
            buf[r, lay.weight_off(r):lay.weight_off(r) + cnt[r]] = w[s0:s0 + cnt[r]]
        idx = torch.randperm(sum(rows)).to(torch.int32)
        pos = lay.row_positions(idx)
        assert pos.dtype == torch.int32
        assert torch.equal(buf.view(-1, c)[pos.long()], full[idx.long()])
        if with_ref:
            assert torch.equal(buf.view(-1, c)[lay.row_positions(idx, ref=True).long()], ref[idx.long()])
        assert torch.equal(buf.view(-1)[lay.weight_positions("cpu")], w)
    # the node slices of the ranks tile the mesh
    for N, world in ((10, 4), (896761, 8), (3, 8)):
        sl = [node_slice(N, r, world) for r in range(world)]
        assert sl[0][0] == 0 and sl[-1][1] == N and all(sl[r][1] == sl[r + 1][0] for r in range(world - 1))


def test_tensor_list_is_lazy_and_fetches_remote_rows_on_demand():
    """Host logic of predict()'s return lists: the per-subdomain views are built on first access, a rank's own
    subdomains never trigger the fetch of the other ranks' rows, anything else does exactly once."""
    from fesr_b200.models.scheduler_gnn import TensorList
    built, fetched = [], []
    buf = torch.arange(12.0)
    t = TensorList(make=lambda: (built.append(1), list(torch.split(buf, [3, 4, 5])))[1], n=3)
    t.own = (1, 2)
    t.rest = lambda: fetched.append(1)
    assert isinstance(t, list) and len(t) == 3 and not built
    assert t[1].tolist() == [3.0, 4.0, 5.0, 6.0] and built == [1] and not fetched      # own subdomain
    assert t[-2].shape == (4,) and not fetched                                          # negative index, still own
    assert t[0].shape == (3,) and fetched == [1]                                        # somebody else's rows
    assert t[2].shape == (5,) and fetched == [1] and t.rest is None
    assert [int(v.numel()) for v in t] == [3, 4, 5] and built == [1]
    u = TensorList(make=lambda: [torch.zeros(1), torch.ones(1)], n=2)
    u.rest = lambda: fetched.append(2)
    assert torch.cat(list(u)).tolist() == [0.0, 1.0] and fetched == [1, 2]              # iteration needs everything
    assert TensorList([1, 2])[1] == 2 and len(TensorList()) == 0


# ----------------------------------------------------------------------------- routed (ALDS) predict on several ranks
def _routed_worker(rank, world, port, labels, sizes, edges, ret):
    """What _run_routed_sharded does, with torch on CPU: every rank fills its slot cluster by cluster, one all-gather, one
    row gather into subdomain order."""
    from fesr_b200.pipeline import cluster_major_layout
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = 4
        lay = cluster_major_layout(labels, sizes, edges, world, c)
        node_ptr = np.concatenate([[0], np.cumsum(sizes)])
        n = int(node_ptr[-1])
        full = torch.arange(n * c, dtype=torch.float32).reshape(n, c)             # "predictions" of every batch row
        w_full = torch.arange(len(sizes), dtype=torch.float32) + 0.25              # "weights" of every subdomain
        gbuf = torch.full((world, lay["slot"]), -1.0)
        nr, ns = lay["rows"][rank], lay["cnt"][rank]
        pred_v = gbuf[rank, :nr * c].view(nr, c)
        w_v = gbuf[rank, nr * c:nr * c + ns]
        for s in np.flatnonzero(lay["sub_rank"] == rank):                          # (the GPU path does a whole cluster at once)
            o = int(lay["sub_off"][s])
            pred_v[o:o + sizes[s]] = full[node_ptr[s]:node_ptr[s + 1]]
            w_v[int(lay["sub_idx"][s])] = w_full[s]
        parts = [torch.empty(lay["slot"]) for _ in range(world)]
        dist.all_gather(parts, gbuf[rank].clone())
        g = torch.stack(parts)
        pred_all = g.view(-1, c)[torch.from_numpy(lay["row_pos"])]
        w_all = g.view(-1)[torch.from_numpy(lay["wpos"])]
        ret[rank] = (bool(torch.equal(pred_all, full)), bool(torch.equal(w_all, w_full)))
    finally:
        dist.destroy_process_group()


def test_cluster_major_layout_properties():
    from fesr_b200.pipeline import cluster_major_layout
    rng = np.random.default_rng(3)
    for S, k, world in ((128, 4, 8), (37, 3, 2), (5, 4, 8), (64, 1, 4)):
        labels = rng.integers(0, k, size=S)
        sizes = rng.integers(300, 2000, size=S)
        edges = sizes * rng.integers(10, 15, size=S)
        lay = cluster_major_layout(labels, sizes, edges, world, 4)
        assert sum(lay["rows"]) == sizes.sum() and sum(lay["cnt"]) == S
        assert lay["slot"] % 16 == 0 and all(r * 4 + q <= lay["slot"] for r, q in zip(lay["rows"], lay["cnt"]))
        assert np.unique(lay["row_pos"]).size == sizes.sum()                        # every batch row has its own place
        assert np.unique(lay["wpos"]).size == S
        # weights never collide with prediction rows of the same slot
        rows_as_floats = set((lay["row_pos"][:, None] * 4 + np.arange(4)).ravel().tolist())
        assert not rows_as_floats & set(lay["wpos"].tolist())
        for r in range(world):
            mine = np.flatnonzero(lay["sub_rank"] == r)
            if mine.size == 0:
                continue
            # cluster-major: inside a rank the subdomains of one cluster are one contiguous run of the slot, in index order
            order = mine[np.argsort(lay["sub_idx"][mine])]
            assert np.all(np.diff(labels[order]) >= 0)
            assert len(set(labels[order].tolist())) <= max(1, -(-k // world) + 1)
            off = 0
            for s in order:
                assert lay["sub_off"][s] == off
                off += sizes[s]
        if world > 1 and S >= 4 * world:
            per = [edges[lay["sub_rank"] == r].sum() for r in range(world)]
            assert max(per) - min(per) <= 2 * edges.max()


def test_world2_gloo_routed_slot_gather():
    rng = np.random.default_rng(9)
    S = 48
    labels = rng.integers(0, 4, size=S)
    sizes = rng.integers(50, 400, size=S)
    edges = sizes * 13
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_routed_worker, args=(2, port, labels, sizes, edges, ret), nprocs=2, join=True)
    assert ret[0] == (True, True) and ret[1] == (True, True)
