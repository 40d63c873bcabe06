"""GPU parity: overlap stitch (bit-exact vs oracle) and node weight."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import graph as og
from oracle import models as om

pytestmark = pytest.mark.gpu


def _subdomains(n, levels):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(n)
    part = og.kd_partition(mesh.pos, mesh.cells, levels)
    sub = og.build_subdomains(mesh.pos, mesh.cells, part["leaf_ptr"], part["leaf_cells"])
    return mesh, sub


@pytest.mark.parametrize("n,levels", [(3, 2), (6, 4), (6, 5), (13, 4)])
def test_stitch_bit_exact(n, levels):
    from fesr_b200 import ops
    mesh, sub = _subdomains(n, levels)
    rng = np.random.default_rng(0)
    vals = rng.normal(size=(sub["global_ids"].size, 4)).astype(np.float32)
    field, count, merged = og.stitch_mean(vals, sub["global_ids"], mesh.num_nodes)
    gids = torch.from_numpy(sub["global_ids"]).cuda()
    occ = ops.occurrence_build(gids, mesh.num_nodes)
    occ_ptr, occ_idx = og.occurrence_csr(sub["global_ids"], mesh.num_nodes)
    assert np.array_equal(occ.occ_ptr.cpu().numpy(), occ_ptr.astype(np.int32))
    assert np.array_equal(occ.occ_idx.cpu().numpy(), occ_idx.astype(np.int32))
    f, c, mg = ops.stitch_mean(torch.from_numpy(vals).cuda(), occ, gids, want_merged=True)
    assert np.array_equal(f.cpu().numpy().view(np.uint32), field.view(np.uint32))
    assert np.array_equal(c.cpu().numpy(), count)
    assert np.array_equal(mg.cpu().numpy().view(np.uint32), merged.view(np.uint32))
    assert count.max() >= 2                                   # the halo really overlaps


def test_stitch_matches_the_references_own_averaging_loop():
    """a10 pinned on the GPU: fesr_stitch_mean's merged arrays == what the reference's own loop
    (dataset/GraphDataset.py:1371-1400, executed by tests/golden/make_golden.py) produced, bit for bit -- including the
    nodes shared by 8 subdomains, where numpy sums the scalar `pressure` array in its pairwise order."""
    import os
    from fesr_b200 import ops
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_vectors.npz")))
    for t in "ab":
        g = z[t + "_global_ids"]
        N = int(g.max()) + 1
        gids = torch.from_numpy(g).cuda()
        occ = ops.occurrence_build(gids, N)
        for src, want in ((t + "_pred", t + "_merged"), (t + "_ref", t + "_merged_ref")):
            f, c, mg = ops.stitch_mean(torch.from_numpy(z[src]).cuda(), occ, gids, want_merged=True)
            assert np.array_equal(mg.cpu().numpy().view(np.uint32), z[want].view(np.uint32)), (t, src)
            # the scalar (1-channel) entry point follows the scalar order too
            f1, _, _ = ops.stitch_mean(torch.from_numpy(z[src][:, 3:4].copy()).cuda(), occ)
            assert np.array_equal(f1.cpu().numpy()[g, 0].view(np.uint32), z[want][:, 3].view(np.uint32))
        assert int(c.max()) == int(z[t + "_max_copies"])
    # a node range (what a rank of a sharded run stitches) gives the same rows
    f_all, _, _ = ops.stitch_mean(torch.from_numpy(z["b_pred"]).cuda(), occ)
    f_part, c_part, _ = ops.stitch_mean(torch.from_numpy(z["b_pred"]).cuda(), occ, node_range=(100, 900))
    assert torch.equal(f_part, f_all[100:900]) and c_part.shape == (800,)


def test_stitch_of_consistent_copies_is_identity():
    """size-independent property: if every copy holds field[global_id], the stitch returns field."""
    from fesr_b200 import ops
    mesh, sub = _subdomains(13, 4)
    vals = mesh.y[sub["global_ids"]]
    gids = torch.from_numpy(sub["global_ids"]).cuda()
    occ = ops.occurrence_build(gids, mesh.num_nodes)
    f, c, _ = ops.stitch_mean(torch.from_numpy(vals).cuda(), occ)
    assert rel_l2(f.cpu().numpy(), mesh.y) < 1e-6
    assert int(c.min()) >= 1


def test_node_weight_vs_reference_vector(golden):
    from fesr_b200 import ops
    ei = golden["ref_edge_index"]
    n = golden["x"].shape[0]
    csr = ops.csr_build(torch.from_numpy(ei).cuda(), n)
    pred = torch.from_numpy(golden["kernelnn_w43_y"]).cuda()
    y = torch.from_numpy(golden["y"]).cuda()
    s = ops.node_weight(pred, y, csr, torch.from_numpy(golden["ref_edge_attr"]).cuda())
    ref = float(golden["node_weight"][0])
    # the reference sums ~2000 fp32 terms of mixed sign in torch's order; compare against the
    # fp64 value with a tolerance scaled by sum |terms|
    ew = om.edge_weight(torch.from_numpy(golden["kernelnn_w43_y"]).double(), torch.from_numpy(golden["y"]).double(),
                        torch.from_numpy(ei), torch.from_numpy(golden["ref_edge_attr"]).double())
    scale = float(ew.abs().sum())
    assert abs(float(s[0]) - float(ew.sum())) <= 1e-5 * scale
    assert abs(float(s[0]) - ref) <= 1e-5 * scale


def test_node_weight_per_subdomain():
    from fesr_b200 import ops
    mesh, sub = _subdomains(6, 3)
    rng = np.random.default_rng(3)
    n_tot = sub["global_ids"].size
    pred = rng.normal(size=(n_tot, 4)).astype(np.float32)
    y = mesh.y[sub["global_ids"]]
    ei = np.stack([sub["edge_src"], sub["edge_dst"]])
    csr = ops.csr_build(torch.from_numpy(ei).cuda(), n_tot)
    node_ptr = torch.from_numpy(sub["node_ptr"].astype(np.int32)).cuda()
    s = ops.node_weight(torch.from_numpy(pred).cuda(), torch.from_numpy(y).cuda(), csr,
                        torch.from_numpy(sub["edge_attr"]).cuda(), node_ptr).cpu().numpy()
    S = sub["node_ptr"].size - 1
    for k in range(S):
        lo, hi = sub["edge_ptr"][k], sub["edge_ptr"][k + 1]
        nl, nh = sub["node_ptr"][k], sub["node_ptr"][k + 1]
        e = torch.from_numpy(ei[:, lo:hi] - nl)
        ew = om.edge_weight(torch.from_numpy(pred[nl:nh]).double(), torch.from_numpy(y[nl:nh]).double(), e,
                            torch.from_numpy(sub["edge_attr"][lo:hi]).double())
        assert abs(s[k] - float(ew.sum())) <= 1e-5 * float(ew.abs().sum())
