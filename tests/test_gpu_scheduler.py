"""GPU: scheduler.predict + reconstruct_from_partition (the reference-facing API) vs the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, shipped_state_dict
from oracle import graph as og
from oracle import models as om
from oracle import routing as orr

pytestmark = pytest.mark.gpu


def _setup(tmp_path, n_clusters, shipped, monkeypatch):
    from fesr_b200.dataset.GraphDataset import AnsysDataset
    from fesr_b200.models.model import KernelNN
    monkeypatch.chdir(tmp_path)
    os.makedirs("logs/models/collection_t", exist_ok=True)
    sd = shipped_state_dict(shipped, "neuralop")
    sds = []
    for i in range(n_clusters):
        s = {k: v.clone() for k, v in sd.items()}
        s["fc2.bias"] = s["fc2.bias"] + 0.1 * i          # make the per-cluster models distinguishable
        torch.save(s, f"logs/models/collection_t/partition_{i}.pth")
        sds.append(s)
    ds = AnsysDataset(mesh_n=12, num_meshes=2, sub_size=16)
    model = KernelNN(43, 43, 5, in_width=4, out_width=4)
    return ds, model, sds


def _oracle_predict(ds, idx, sds, labels):
    c = ds._mesh(idx)
    mesh = c["mesh"]
    part = og.kd_partition(mesh.pos, mesh.cells, ds.levels)
    sub = og.build_subdomains(mesh.pos, mesh.cells, part["leaf_ptr"], part["leaf_cells"])
    preds = []
    models = []
    for sd in sds:
        o = om.make_model("neuralop", 43, 5)
        o.load_state_dict(sd)
        models.append(o)
    ws = []
    with torch.no_grad():
        for s in range(sub["node_ptr"].size - 1):
            nl, nh = sub["node_ptr"][s], sub["node_ptr"][s + 1]
            el, eh = sub["edge_ptr"][s], sub["edge_ptr"][s + 1]
            ei = torch.from_numpy(np.stack([sub["edge_src"][el:eh] - nl, sub["edge_dst"][el:eh] - nl]))
            ea = torch.from_numpy(sub["edge_attr"][el:eh]).unsqueeze(1)
            g = sub["global_ids"][nl:nh]
            p = models[labels[s]](torch.from_numpy(mesh.x[g]), ei, ea)
            ws.append(float(om.edge_weight(p.double(), torch.from_numpy(mesh.y[g]).double(), ei, ea.double()).sum()))
            preds.append(p.numpy())
    field, count, merged = og.stitch_mean(np.concatenate(preds), sub["global_ids"], mesh.num_nodes)
    return sub, preds, field, merged, ws


def test_predict_and_stitch_single_cluster(tmp_path, shipped, monkeypatch):
    from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
    ds, model, sds = _setup(tmp_path, 1, shipped, monkeypatch)
    sched = GNNPartitionScheduler("t", 1, ds, model, train=False)
    x = ds.get_one_full_sample(1)
    pred_y_list, ref_y_list, model_idx, weights_list = sched.predict(x)
    assert len(pred_y_list) == len(x) == 16 and model_idx.shape == (16,)
    sub, preds, field, merged, ws = _oracle_predict(ds, 1, sds, np.zeros(16, dtype=int))
    for s in range(16):
        assert pred_y_list[s].shape == preds[s].shape
        assert rel_l2(pred_y_list[s].numpy(), preds[s]) < 1e-5
        assert weights_list[s].shape == (preds[s].shape[0],)
        assert abs(float(weights_list[s][0]) - ws[s]) <= 1e-4 * max(1.0, abs(ws[s]))
    out = ds.reconstruct_from_partition(pred_y_list, ref_y_list, 1, model_idx, weights_list)   # 5-arg call
    assert rel_l2(out.field.numpy(), field) < 1e-5
    assert rel_l2(out.merged.numpy(), merged) < 1e-5
    assert rel_l2(out.ref_field.numpy(), ds._mesh(1)["mesh"].y) < 1e-6
    # generic path: plain list of Data without the attached device batch gives the same numbers
    p2, _, _, _ = sched.predict(list(x))
    assert rel_l2(torch.cat(list(p2)).numpy(), torch.cat(list(pred_y_list)).numpy()) < 1e-6
    # .vtu export (run_ALDS_3D.py:33-38), parsed back: the reference's grid = every partition appended, the averaged
    # arrays written to every copy; each partition's cells in the appended numbering are the mesh cells it holds
    from fesr_b200.dataset.vtu import read_vtu
    out.write_vtu("logs/vtk/t/pred_1.vtu")
    back = read_vtu("logs/vtk/t/pred_1.vtu")
    mesh = ds._mesh(1)["mesh"]
    gids = sub["global_ids"]
    assert back["points"].shape == (gids.size, 3) and np.array_equal(back["points"], mesh.pos[gids])
    pdt = back["point_data"]
    assert set(pdt) == {"velocity", "pressure", "ref_velocity", "ref_pressure", "GlobalPointIds"}
    assert np.array_equal(pdt["GlobalPointIds"], gids)
    assert np.array_equal(pdt["velocity"], out.merged.numpy()[:, :3]) and np.array_equal(pdt["pressure"], out.merged.numpy()[:, 3])
    assert rel_l2(pdt["velocity"], merged[:, :3]) < 1e-5 and rel_l2(pdt["pressure"], merged[:, 3]) < 1e-5
    assert rel_l2(pdt["ref_velocity"], mesh.y[gids][:, :3]) < 1e-6 and rel_l2(pdt["ref_pressure"], mesh.y[gids][:, 3]) < 1e-6
    part = og.kd_partition(mesh.pos, mesh.cells, ds.levels)
    assert back["cells"].shape == (part["leaf_cells"].size, 4) and bool((back["types"] == 10).all())
    assert np.array_equal(gids[back["cells"]], mesh.cells[part["leaf_cells"]])           # same cells, appended numbering
    sub_of_cell = np.repeat(np.arange(16), np.diff(part["leaf_ptr"]))
    lo, hi = sub["node_ptr"][sub_of_cell], sub["node_ptr"][sub_of_cell + 1]
    assert bool(((back["cells"] >= lo[:, None]) & (back["cells"] < hi[:, None])).all())   # inside their own partition
    # the original-mesh variant
    out.write_vtu("logs/vtk/t/pred_1_mesh.vtu", appended=False)
    b2 = read_vtu("logs/vtk/t/pred_1_mesh.vtu")
    assert np.array_equal(b2["points"], mesh.pos) and np.array_equal(b2["cells"], mesh.cells)
    assert np.array_equal(b2["point_data"]["velocity"], out.field.numpy()[:, :3])
    assert np.array_equal(b2["point_data"]["pressure"], out.field.numpy()[:, 3])


def test_alds_routing_and_per_cluster_models(tmp_path, shipped, monkeypatch):
    from fesr_b200.models.classifier import KMeansClassifier
    from fesr_b200.models.encoder import PCAEncoder
    from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
    ds, model, sds = _setup(tmp_path, 3, shipped, monkeypatch)
    enc, clf = PCAEncoder(n_components=2), KMeansClassifier(n_clusters=3)
    x = ds.get_one_full_sample(0)
    enc.train(list(x), save_model=True, path="logs/models/collection_t")
    latent = enc.get_latent_space(x)
    clf.train(latent, save_model=True, path="logs/models/collection_t")
    # oracle routing from the same fitted sklearn objects
    feat = orr.routing_features([d.x.numpy() for d in x])
    labels_ref, latent_ref = orr.route(feat, enc.model.mean_, enc.model.components_, clf.scaler.mean_,
                                       clf.scaler.scale_, clf.model.cluster_centers_)
    assert rel_l2(latent, latent_ref) < 1e-6
    assert np.array_equal(clf.cluster(latent), labels_ref)
    assert len(set(labels_ref.tolist())) == 3
    sched = GNNPartitionScheduler("t", 3, ds, model, train=False, encoder=enc, classifier=clf)
    pred_y_list, ref_y_list, model_idx, weights_list = sched.predict(x)
    assert np.array_equal(model_idx, labels_ref)
    sub, preds, field, merged, ws = _oracle_predict(ds, 0, sds, labels_ref)
    for s in range(16):
        assert rel_l2(pred_y_list[s].numpy(), preds[s]) < 1e-5
    out = ds.reconstruct_from_partition(pred_y_list, ref_y_list, 0, model_idx, weights_list)
    assert rel_l2(out.field.numpy(), field) < 1e-5


def test_routing_vs_reference_sklearn_vectors(golden):
    """fesr_route vs the latent / labels the reference's PCAEncoder + KMeansClassifier produced."""
    from fesr_b200 import ops
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(int(golden["route_mesh_n"]))
    part, batch = ops.assemble(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(),
                               int(golden["route_levels"]))
    x_dev = torch.from_numpy(mesh.x).cuda()[batch.global_ids]
    labels, latent = ops.route(x_dev, batch.node_ptr, golden["route_pca_mean"], golden["route_pca_components"],
                               golden["route_scaler_mean"], golden["route_scaler_scale"], golden["route_centroids"])
    assert rel_l2(latent.cpu().numpy(), golden["route_latent"]) < 1e-5
    assert np.array_equal(labels.cpu().numpy(), golden["route_labels"])


def test_gradient_loss_forward(golden):
    from fesr_b200.models.scheduler_gnn import GradientbasedLoss
    pred = torch.from_numpy(golden["kernelnn_w43_y"]).cuda()
    y = torch.from_numpy(golden["y"]).cuda()
    ei = torch.from_numpy(golden["ref_edge_index"]).cuda()
    ea = torch.from_numpy(golden["ref_edge_attr"]).cuda()
    for mw, key in ((1.0, "gradient_loss"), (4.0, "gradient_loss_mw4")):
        v = float(GradientbasedLoss(max_weight=mw)(pred, y, ei, ea))
        assert abs(v - float(golden[key])) <= 2e-5 * abs(float(golden[key])) + 1e-9
    nw = GradientbasedLoss().compute_node_weight(pred, y, ei, ea, y.shape[0])
    assert nw.shape == (y.shape[0],)


def test_sharded_predict_matches_single_rank(tmp_path, shipped, monkeypatch):
    """The multi-rank predict path (cached shard + slot layout + in-place all-gather), both ranks of a world of 2
    played one after the other on this GPU with the collective emulated: the union of what the ranks contribute
    equals the single-rank result bit for bit, and the two field slices make up the single-rank field."""
    from fesr_b200 import comm
    from fesr_b200.models import scheduler_gnn as sg
    ds, model, sds = _setup(tmp_path, 1, shipped, monkeypatch)
    sched = sg.GNNPartitionScheduler("t", 1, ds, model, train=False)
    base = ds.get_one_full_sample(0)
    p1, r1, mi1, w1 = sched.predict(base)
    full_p, full_w = p1.dev.clone(), torch.stack([w[0] for w in w1]).cuda()
    full_field = ds.reconstruct_from_partition(p1, r1, 0, mi1, w1).field.clone()
    # host-input variant of the same sample (the sharded path copies only its own rows of x)
    c = ds._mesh(0)
    xh, yh = c["x"].cpu().pin_memory(), c["y"].cpu().pin_memory()
    y_full = c["y"]
    contributed = {}
    state = {"rank": 0}

    def fake_allgather(slots, stream=None):
        rank = state["rank"]
        (entry,) = base.batch._shards.values()
        lay = entry["lay"]
        assert slots.shape == (2, lay.slot) and lay.slot % (4 * lay.c) == 0
        nr, ns = lay.rows[rank], lay.cnt[rank]
        wo = lay.weight_off(rank)
        contributed[rank] = (slots[rank, :nr * 4].view(nr, 4).clone(), slots[rank, wo:wo + ns].clone(), lay)
        for r in range(2):
            if r == rank:
                continue
            r0, r1_ = int(lay.row_offs[r]), int(lay.row_offs[r + 1])
            slots[r, :(r1_ - r0) * 4] = full_p[r0:r1_].reshape(-1)
            if lay.with_ref:
                slots[r, (r1_ - r0) * 4:2 * (r1_ - r0) * 4] = y_full[r0:r1_].reshape(-1)
            s0 = int(lay.sub_offs[r])
            slots[r, lay.weight_off(r):lay.weight_off(r) + lay.cnt[r]] = full_w[s0:s0 + lay.cnt[r]]

    monkeypatch.setattr(comm, "allgatherv_pred", fake_allgather)
    N = c["mesh"].num_nodes
    for sample in (base, base.with_host_inputs(xh, yh)):
        contributed.clear()
        slices = {}
        for rank in range(2):
            state["rank"] = rank
            monkeypatch.setattr(sg, "_dist", lambda r=rank: (None, r, 2))
            base.batch.__dict__.pop("_shards", None)          # the cache is per process; here one process plays both
            p, r, mi, w = sched.predict(sample)
            assert len(p) == 16 and mi.shape == (16,)
            # only the rank's own subdomains were copied to the host eagerly; touching one of them does not fetch the rest
            s0, s1 = p.own
            assert 0 <= s0 < s1 <= 16 and p.rest is not None
            own_first = p[s0]
            assert p.rest is not None
            off = int(base.batch.node_ptr[s0])
            assert torch.equal(own_first, full_p[off:off + own_first.shape[0]].cpu())
            out = ds.reconstruct_from_partition(p, r, 0, mi, w)
            lo, hi = out.node_range
            assert (lo, hi) == ((N * rank) // 2, (N * (rank + 1)) // 2)
            slices[rank] = out.field_local.clone()
            assert torch.equal(slices[rank], full_field[lo:hi])                  # the rank's slice, stitched eagerly
            assert rel_l2(out.ref_field_local.numpy(), c["mesh"].y[lo:hi]) < 1e-6
            assert torch.equal(out.field, full_field)                             # the whole field on demand
            assert rel_l2(out.ref_field.numpy(), c["mesh"].y) < 1e-6
            assert torch.equal(torch.cat(list(p)), full_p.cpu())          # iterating fetches everything
            assert p.rest is None
            assert torch.equal(p.dev, full_p)
            assert torch.equal(torch.stack([t[0] for t in w]).cpu(), full_w.cpu())
        assert torch.equal(torch.cat([slices[0], slices[1]]), full_field)
        (pa, wa, lay), (pb, wb, _) = contributed[0], contributed[1]
        assert pa.shape[0] == lay.rows[0] and pb.shape[0] == lay.rows[1] and sum(lay.cnt) == 16 and min(lay.cnt) > 0
        assert torch.equal(torch.cat([pa, pb]), full_p)
        assert torch.equal(torch.cat([wa, wb]).cpu(), full_w.cpu())


def test_stored_subdomains_flow_through_the_gpu_path(tmp_path, shipped, monkeypatch):
    """The reference's on-disk layout (edges in arbitrary order, as its Python set leaves them) -> StoredSubdomainDataset
    -> scheduler.predict + reconstruct_from_partition gives what the GPU-assembled dataset gives."""
    from fesr_b200.dataset.store import StoredSubdomainDataset, save_partitioned
    from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
    from fesr_b200.utils import init_dataset
    ds, model, sds = _setup(tmp_path, 1, shipped, monkeypatch)
    rng = np.random.default_rng(0)
    meshes = []
    for m in range(2):
        subs = []
        for d in ds.get_one_full_sample(m):
            p = torch.from_numpy(rng.permutation(d.edge_index.shape[1]))
            subs.append({"x": d.x, "y": d.y, "pos": d.pos, "edge_index": d.edge_index[:, p], "edge_attr": d.edge_attr[p],
                         "global_node_ids": d.global_node_ids})
        meshes.append(subs)
    path = str(tmp_path / "partitioned.npz")
    save_partitioned(path, meshes)
    stored = init_dataset("stored", root=path)
    assert isinstance(stored, StoredSubdomainDataset) and len(stored) == 32 and stored[17].x.shape == ds[17].x.shape
    sched = GNNPartitionScheduler("t", 1, ds, model, train=False)
    sched2 = GNNPartitionScheduler("t", 1, stored, model, train=False)
    for m in range(2):
        p, r, mi, w = sched.predict(ds.get_one_full_sample(m))
        out = ds.reconstruct_from_partition(p, r, m, mi, w)
        p2, r2, mi2, w2 = sched2.predict(stored.get_one_full_sample(m))
        out2 = stored.reconstruct_from_partition(p2, r2, m, mi2, w2)
        assert rel_l2(torch.cat(list(p2)).numpy(), torch.cat(list(p)).numpy()) < 1e-6
        assert rel_l2(out2.field.numpy(), out.field.numpy()) < 1e-6
        assert rel_l2(out2.ref_field.numpy(), out.ref_field.numpy()) < 1e-6
        assert np.allclose(out2.pos, ds._mesh(m)["mesh"].pos)
        assert torch.allclose(torch.stack([t[0] for t in w2]), torch.stack([t[0] for t in w]), rtol=1e-4, atol=1e-6)


def test_gather_scatter_rows():
    """fesr_gather_rows / fesr_scatter_rows (the sub-batch selection and reorder_predictions of the routed predict,
    models/scheduler_gnn.py:240-251, 302-309): bit-exact row moves, empty index, 4- and 8-float rows."""
    from fesr_b200 import ops
    g = torch.Generator().manual_seed(3)
    for rf in (4, 8):
        src = torch.randn(10007, rf, generator=g).cuda()
        idx = torch.randperm(10007, generator=g)[:6001].cuda()
        got = ops.gather_rows(src, idx)
        assert torch.equal(got, src[idx])
        dst = torch.full((10007, rf), -7.0, device="cuda")
        ops.scatter_rows(got, idx, dst)
        ref = torch.full((10007, rf), -7.0, device="cuda")
        ref[idx] = got
        assert torch.equal(dst, ref)
        assert ops.gather_rows(src, idx[:0]).shape == (0, rf)
