"""GPU parity: subdomain assembly (partition, halo, compaction, edges, CSR) -- bit-exact vs oracle."""
import numpy as np
import pytest
import torch

from oracle import graph as og

pytestmark = pytest.mark.gpu


def _mesh(n):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    return make_duct_mesh(n)


def _check(pos, cells, levels, mode):
    from fesr_b200 import ops
    ref = og.kd_partition(pos, cells, levels, mode)
    sub = og.build_subdomains(pos, cells, ref["leaf_ptr"], ref["leaf_cells"])
    part, batch = ops.assemble(torch.from_numpy(pos).cuda(), torch.from_numpy(cells).cuda(), levels, mode)
    torch.cuda.synchronize()
    assert np.array_equal(part.home_leaf.cpu().numpy(), ref["home_leaf"])
    assert np.array_equal(part.tree_axis.cpu().numpy(), ref["tree_axis"])
    assert np.array_equal(part.tree_split.cpu().numpy(), ref["tree_split"])
    assert np.array_equal(part.leaf_ptr.cpu().numpy(), ref["leaf_ptr"].astype(np.int32))
    assert np.array_equal(part.leaf_cells.cpu().numpy(), ref["leaf_cells"].astype(np.int32))
    assert np.array_equal(batch.node_ptr.cpu().numpy(), sub["node_ptr"].astype(np.int32))
    assert np.array_equal(batch.edge_ptr.cpu().numpy(), sub["edge_ptr"].astype(np.int32))
    assert np.array_equal(batch.global_ids.cpu().numpy(), sub["global_ids"])
    assert np.array_equal(batch.edge_src.cpu().numpy(), sub["edge_src"].astype(np.int32))
    assert np.array_equal(batch.edge_dst.cpu().numpy(), sub["edge_dst"].astype(np.int32))
    assert np.array_equal(batch.edge_attr.cpu().numpy().view(np.uint32), sub["edge_attr"].view(np.uint32))
    assert np.array_equal(batch.rowptr.cpu().numpy(), sub["rowptr"].astype(np.int32))
    return ref, sub


@pytest.mark.parametrize("n,levels", [(2, 0), (2, 1), (3, 3), (6, 4), (13, 4)])
@pytest.mark.parametrize("mode", [og.MODE_ONE_REGION, og.MODE_ALL_INTERSECTING])
def test_assembly_bit_exact(n, levels, mode):
    m = _mesh(n)
    _check(m.pos, m.cells, levels, mode)


def test_more_leaves_than_cells():
    """ragged case: 2^levels > C leaves, so some subdomains are empty."""
    m = _mesh(1)                                    # 24 cells
    ref, sub = _check(m.pos, m.cells, 6, og.MODE_ALL_INTERSECTING)
    assert (np.diff(sub["node_ptr"]) == 0).any()


def test_tied_coordinates():
    """unjittered lattice: many centroids share a coordinate, ties break by cell id."""
    m = _mesh(4)
    pos = np.round(m.pos / 3e-3).astype(np.float32)
    _check(pos, m.cells, 4, og.MODE_ALL_INTERSECTING)
    _check(pos, m.cells, 4, og.MODE_ONE_REGION)


def test_random_soup_of_tets():
    rng = np.random.default_rng(5)
    pos = rng.normal(size=(300, 3)).astype(np.float32)
    cells = np.stack([rng.choice(300, 4, replace=False) for _ in range(900)]).astype(np.int32)
    _check(pos, cells, 3, og.MODE_ALL_INTERSECTING)


def test_500k_properties():
    """full BASELINE size: every cell has a home, every node is covered, halo >= 1 copy."""
    from fesr_b200 import ops
    m = _mesh("500k")
    part, batch = ops.assemble(torch.from_numpy(m.pos).cuda(), torch.from_numpy(m.cells).cuda(), 7)
    S = 128
    assert int(part.leaf_ptr[-1]) >= m.num_cells
    home = part.home_leaf.cpu().numpy()
    counts = np.bincount(home, minlength=S)
    assert counts.max() - counts.min() <= 1                      # exact-median bisection balances the homes
    gids = batch.global_ids.cpu().numpy()
    assert np.array_equal(np.unique(gids), np.arange(m.num_nodes))
    node_ptr = batch.node_ptr.cpu().numpy()
    for s in (0, 17, 127):
        seg = gids[node_ptr[s]:node_ptr[s + 1]]
        assert (np.diff(seg) > 0).all()                          # ascending, unique
    rowptr = batch.rowptr.cpu().numpy()
    assert rowptr[-1] == batch.e_tot and (np.diff(rowptr) >= 0).all()
    src, dst = batch.edge_src.cpu().numpy(), batch.edge_dst.cpu().numpy()
    key = dst.astype(np.int64) * batch.n_tot + src
    assert (np.diff(key) > 0).all()                              # strictly (dst, src) sorted, no duplicates
    # symmetric: every edge has its reverse
    rkey = np.sort(src.astype(np.int64) * batch.n_tot + dst)
    assert np.array_equal(rkey, key)
