"""CPU: the oracle restatement vs vectors produced by the reference's own code."""
import numpy as np
import torch

from conftest import rel_l2, state_dict_from
from oracle import graph as og
from oracle import models as om
from oracle import routing as orr


def _edges(golden):
    return torch.from_numpy(golden["ref_edge_index"]), torch.from_numpy(golden["ref_edge_attr"])


def test_edge_set_and_attr_match_reference_vtk_to_pyg(golden):
    src, dst, attr = og.build_edges(golden["cells"], golden["pos"])
    n = golden["pos"].shape[0]
    ref = golden["ref_edge_index"]
    ref_key = ref[1] * n + ref[0]
    order = np.argsort(ref_key)
    assert np.array_equal(ref_key[order], dst * n + src)            # same SET, canonical order
    assert np.array_equal(golden["ref_edge_attr"][order, 0].view(np.uint32), attr.view(np.uint32))  # bit-exact fp32


def test_small_width_models_match_reference(golden):
    ei, ea = _edges(golden)
    x = torch.from_numpy(golden["x"])
    for name, kind, w, L in (("kernelnn_w16", "neuralop", 16, 3), ("teecnet_w12", "teecnet", 12, 2),
                             ("kernelnn_w48", "neuralop", 48, 2)):
        m = om.make_model(kind, width=w, num_layers=L)
        m.load_state_dict(state_dict_from(golden, name))
        with torch.no_grad():
            y = m(x, ei, ea).numpy()
        assert rel_l2(y, golden[name + "_y"]) < 1e-6, name


def test_train_step_matches_reference(golden):
    ei, ea = _edges(golden)
    x = torch.from_numpy(golden["x"])
    y = torch.from_numpy(golden["y"])
    for name, kind, w, L in (("kernelnn_w16", "neuralop", 16, 3), ("teecnet_w12", "teecnet", 12, 2)):
        m = om.make_model(kind, width=w, num_layers=L)
        m.load_state_dict(state_dict_from(golden, name))
        opt = torch.optim.Adam(m.parameters(), lr=0.0005)
        loss = om.train_step(m, opt, x, ei, ea, y)
        assert abs(float(loss) - float(golden[name + "_loss"])) <= 1e-6 * abs(float(golden[name + "_loss"]))
        for k, p in m.named_parameters():
            assert rel_l2(p.grad.numpy(), golden[f"{name}_grad::{k}"]) < 1e-4, (name, k)
            assert rel_l2(p.detach().numpy(), golden[f"{name}_sd_after::{k}"]) < 1e-5, (name, k)


def test_node_weight_and_gradient_loss(golden):
    ei, ea = _edges(golden)
    y = torch.from_numpy(golden["y"])
    pred = torch.from_numpy(golden["kernelnn_w43_y"])
    nw = om.compute_node_weight(pred, y, ei, ea, y.shape[0]).numpy()
    assert rel_l2(nw, golden["node_weight"]) < 1e-5
    assert abs(float(om.gradient_loss(pred, y, ei, ea)) - float(golden["gradient_loss"])) < 1e-5 * abs(float(golden["gradient_loss"])) + 1e-9
    assert abs(float(om.gradient_loss(pred, y, ei, ea, 4.0)) - float(golden["gradient_loss_mw4"])) < 1e-5 * abs(float(golden["gradient_loss_mw4"])) + 1e-9


def test_routing_matches_reference_sklearn(golden):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(int(golden["route_mesh_n"]))
    part = og.kd_partition(mesh.pos, mesh.cells, int(golden["route_levels"]))
    sub = og.build_subdomains(mesh.pos, mesh.cells, part["leaf_ptr"], part["leaf_cells"])
    xs = [mesh.x[sub["global_ids"][sub["node_ptr"][s]:sub["node_ptr"][s + 1]]] for s in range(sub["node_ptr"].size - 1)]
    feat = orr.routing_features(xs)
    labels, latent = orr.route(feat, golden["route_pca_mean"], golden["route_pca_components"],
                               golden["route_scaler_mean"], golden["route_scaler_scale"], golden["route_centroids"])
    assert rel_l2(latent, golden["route_latent"]) < 1e-5
    assert np.array_equal(labels, golden["route_labels"])


def test_oracle_interp_gaussian_properties():
    """The restated vtkGaussianKernel interpolation (parity unpinned: no VTK here): partition of unity, null value,
    the rim of the radius is inclusive, weights follow exp(-(sharpness / radius)^2 d^2)."""
    from oracle import graph as og
    src = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]], dtype=np.float32)
    val = np.array([1.0, 3.0, 10.0], dtype=np.float32)
    dst = np.array([[0, 0, 0], [0.5, 0, 0], [5, 5, 5], [2, 0, 0]], dtype=np.float32)
    out, cnt = og.interp_gaussian(src, val, dst, radius=1.0, sharpness=2.0, null_value=-1.0)
    assert cnt.tolist() == [2, 2, 0, 1]                       # (1,0,0) is exactly on the rim of (0,0,0) and of (2,0,0)
    w = np.exp(-4.0)
    assert abs(out[0, 0] - (1.0 + 3.0 * w) / (1.0 + w)) < 1e-6
    assert abs(out[1, 0] - 2.0) < 1e-6 and out[2, 0] == -1.0 and abs(out[3, 0] - 3.0) < 1e-6
    ones, _ = og.interp_gaussian(src, np.ones(3, np.float32), dst[:2], 1.5)
    assert np.allclose(ones, 1.0)


def test_oracle_wall_shear_stress_properties():
    """The restated compute_wss.py chain (parity unpinned: no VTK here): a linear field's gradient is recovered at every
    point, the boundary of the n x n x 4n duct has 2 (2 n^2 + 16 n^2) triangles, tau_wall is tangential."""
    from fesr_b200.dataset.synthetic import make_duct_mesh
    from oracle import graph as og
    n = 3
    m = make_duct_mesh(n)
    A = np.array([[1.0, 2.0, 3.0], [0.5, -1.0, 2.0], [4.0, 0.0, -2.0]])
    r = og.wall_shear_stress(m.pos, m.cells, (m.pos.astype(np.float64) @ A.T).astype(np.float32), 2.0)
    assert np.abs(r["gradient"] - A.reshape(-1)).max() < 1e-4
    assert r["faces"].shape[0] == 2 * (2 * n * n + 16 * n * n)
    assert r["surface_nodes"].size == 2 * (n + 1) ** 2 + 4 * n * (4 * n - 1)
    assert np.abs(np.einsum("ij,ij->i", r["wss"], r["normals"])).max() < 1e-9 * max(1.0, np.abs(r["wss"]).max()) + 1e-9
    tau = 2.0 * (r["normals"] @ (A + A.T).T)
    tw = tau - np.einsum("ij,ij->i", tau, r["normals"])[:, None] * r["normals"]
    assert np.abs(r["wss"] - tw).max() < 1e-3


def test_stitch_matches_the_references_own_averaging_loop():
    """a10 pinned: tests/golden/stitch_vectors.npz holds what dataset/GraphDataset.py:1371-1400 (executed from the
    reference's source by make_golden.py on a duck-typed merged grid) makes of two small ducts' appended partitions --
    up to 8 copies per node, where numpy's summation order of the scalar arrays differs from the vector arrays'."""
    import os
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_vectors.npz")))
    for t in "ab":
        g = z[t + "_global_ids"]
        N = int(g.max()) + 1
        for src, want in ((t + "_pred", t + "_merged"), (t + "_ref", t + "_merged_ref")):
            field, count, merged = og.stitch_mean(z[src], g, N)
            assert np.array_equal(merged.view(np.uint32), z[want].view(np.uint32)), (t, src)     # bit for bit
            assert np.array_equal(field[g].view(np.uint32), z[want].view(np.uint32))
        assert int(count.max()) == int(z[t + "_max_copies"])
    assert int(z["b_max_copies"]) >= 8
