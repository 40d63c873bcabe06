"""GPU: size-independent properties of the hot path at BASELINE.json's full sizes (the oracle does not finish
in seconds there): 526 848-cell duct (config 2), 2 044 416 cells (configs 3 / 4), 5 184 000 cells (config 5).

Assembly: partition / halo / CSR invariants on the device arrays.  Stitch: constant and copy fields, linearity,
copy counts.  Forward: determinism, independence of the block-diagonal blocks (a subdomain's rows do not depend on
what else is in the batch), agreement of the arithmetic arms (fp32 <-> f16 fused <-> f16 two-kernel within the
1e-3 gate of north_star), node weight of identical fields.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2, shipped_state_dict

pytestmark = pytest.mark.gpu

SIZES = {"500k": (28, 7), "2M": (44, 9), "5M": (60, 10)}
_CACHE = {}


def _assembled(tag):
    from fesr_b200 import ops
    from fesr_b200.dataset.synthetic import make_duct_mesh
    if tag not in _CACHE:
        _CACHE.clear()                                   # one big mesh resident at a time
        n, levels = SIZES[tag]
        mesh = make_duct_mesh(n)
        pos, cells = torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda()
        part, batch = ops.assemble(pos, cells, levels)
        _CACHE[tag] = (mesh, pos, cells, part, batch, levels)
    return _CACHE[tag]


@pytest.mark.parametrize("tag", ["500k", "2M", "5M"])
def test_assembly_invariants(tag):
    mesh, pos, cells, part, b, levels = _assembled(tag)
    S, C, N = 1 << levels, mesh.num_cells, mesh.num_nodes
    assert b.n_sub == S
    home = part.home_leaf.long()
    assert int(home.min()) >= 0 and int(home.max()) < S
    # kd median split: the home leaves are balanced to within one cell per level
    cnt = torch.bincount(home, minlength=S)
    assert int(cnt.sum()) == C and int(cnt.max() - cnt.min()) <= levels + 1
    # every leaf holds its home cells (+ halo), ascending cell ids inside a leaf, no duplicates
    leaf_ptr, leaf_cells = part.leaf_ptr.long(), part.leaf_cells.long()
    assert int(leaf_ptr[0]) == 0 and int(leaf_ptr[-1]) == leaf_cells.numel()
    leaf_of = torch.repeat_interleave(torch.arange(S, device="cuda"), leaf_ptr[1:] - leaf_ptr[:-1])
    key = leaf_of * C + leaf_cells
    assert bool((key[1:] > key[:-1]).all())
    is_home = home[leaf_cells] == leaf_of
    assert int(is_home.sum()) == C
    # nodes: per-subdomain ascending global ids, all mesh nodes covered, every subdomain node used by one of its cells
    node_ptr, gid = b.node_ptr.long(), b.global_ids
    sub_of = torch.repeat_interleave(torch.arange(S, device="cuda"), node_ptr[1:] - node_ptr[:-1])
    nkey = sub_of * N + gid
    assert bool((nkey[1:] > nkey[:-1]).all())
    assert int(torch.unique(gid).numel()) == N
    ckey = torch.unique((leaf_of.unsqueeze(1) * N + cells.long()[leaf_cells]).reshape(-1))
    assert torch.equal(ckey, nkey)
    # edges: CSR by destination in (subdomain, dst, src) order, no self loops, no duplicates, symmetric, inside the block
    src, dst, rowptr = b.edge_src.long(), b.edge_dst.long(), b.rowptr.long()
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == b.e_tot and bool((rowptr[1:] >= rowptr[:-1]).all())
    deg = rowptr[1:] - rowptr[:-1]
    assert torch.equal(torch.repeat_interleave(torch.arange(b.n_tot, device="cuda"), deg), dst)
    assert int(deg.max()) <= 14 * 2                     # Kuhn-split duct: 14 neighbours, jitter keeps the topology
    ekey = dst * b.n_tot + src
    assert bool((ekey[1:] > ekey[:-1]).all()) and bool((src != dst).all())
    assert torch.equal(sub_of[src], sub_of[dst])
    rkey, _ = torch.sort(src * b.n_tot + dst)
    assert torch.equal(rkey, ekey)
    # edge lengths are the fp32 Euclidean distances of the end points, symmetric bit for bit
    p = pos[gid]
    d = (p[src] - p[dst]).double().pow(2).sum(1).sqrt()
    assert float(((b.edge_attr.double() - d).abs() / d).max()) < 5e-7
    assert float(b.edge_attr.min()) > 0
    # the subdomain edge ranges agree with the node ranges
    assert torch.equal(b.edge_ptr.long(), rowptr[node_ptr])


@pytest.mark.parametrize("tag", ["500k", "5M"])
def test_stitch_properties(tag):
    from fesr_b200 import ops
    mesh, pos, cells, part, b, levels = _assembled(tag)
    N = mesh.num_nodes
    occ = ops.occurrence_build(b.global_ids, N)
    g = torch.Generator(device="cuda").manual_seed(0)
    # constant field -> the same constant, counts = number of subdomains sharing the node (>= 1, sum = n_tot)
    const = torch.full((b.n_tot, 4), 0.375, device="cuda")
    field, count, merged = ops.stitch_mean(const, occ, b.global_ids, want_merged=True)
    assert bool((field == 0.375).all()) and bool((merged == 0.375).all())
    assert int(count.min()) >= 1 and int(count.sum()) == b.n_tot
    assert torch.equal(count.long(), torch.bincount(b.global_ids, minlength=N))
    # copies of a mesh field -> that field (every copy identical: the mean of k equal numbers, <= 1 ulp)
    f = torch.rand(N, 4, device="cuda", generator=g)
    field, _, merged = ops.stitch_mean(f[b.global_ids], occ, b.global_ids, want_merged=True)
    assert float((field - f).abs().max()) <= 1.2e-7
    assert torch.equal(merged, field[b.global_ids])
    # linearity
    u = torch.randn(b.n_tot, 4, device="cuda", generator=g)
    v = torch.randn(b.n_tot, 4, device="cuda", generator=g)
    su, _, _ = ops.stitch_mean(u, occ)
    sv, _, _ = ops.stitch_mean(v, occ)
    suv, _, _ = ops.stitch_mean(2.0 * u - 0.5 * v, occ)
    assert float((suv - (2.0 * su - 0.5 * sv)).abs().max()) < 1e-5
    # interior nodes (one copy) pass through untouched
    single = (count == 1)[b.global_ids]
    assert torch.equal(su[b.global_ids][single], u[single])
    # deterministic
    su2, _, _ = ops.stitch_mean(u, occ)
    assert torch.equal(su, su2)


def _model(shipped, kind="neuralop"):
    from fesr_b200.models.model import KernelNN, TEECNet
    m = KernelNN(43, 43, 5, in_width=4, out_width=4) if kind == "neuralop" else TEECNet(4, 43, 4, num_layers=5,
                                                                                        retrieve_weight=False)
    m.load_state_dict(shipped_state_dict(shipped, kind))
    return m.cuda().eval()


@pytest.mark.parametrize("tag", ["500k", "2M"])
def test_forward_properties_full_size(tag, shipped, monkeypatch):
    from fesr_b200 import ops
    from fesr_b200.models.scheduler_gnn import select_subdomains
    mesh, pos, cells, part, b, levels = _assembled(tag)
    x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
    y = torch.from_numpy(mesh.y).cuda()[b.global_ids]
    m = _model(shipped)
    out = {}
    with torch.no_grad():
        for prec, fuse in (("f16", "3"), ("f16", "0"), ("tf32", "3"), ("fp32", "3")):
            if prec == "fp32" and tag != "500k":
                continue                                     # the CUDA-core arm: once is enough
            monkeypatch.setenv("FESR_FUSE", fuse)
            m.precision = prec
            out[(prec, fuse)] = m(x, b.csr, b.edge_attr)
        monkeypatch.setenv("FESR_FUSE", "3")
        m.precision = "f16"
        again = m(x, b.csr, b.edge_attr)
        assert torch.equal(again, out[("f16", "3")])         # deterministic, bit for bit
        assert bool(torch.isfinite(again).all())
        ref = out.get(("fp32", "3"), out[("tf32", "3")])
        for k, v in out.items():
            err = rel_l2(v.cpu().numpy(), ref.cpu().numpy())
            print(tag, k, f"rel-L2 vs {'fp32' if ('fp32', '3') in out else 'tf32'} arm: {err:.3e}")
            assert err < 1.5e-3, k
        # block-diagonal independence: a few subdomains run alone give the rows they had inside the full batch
        keep = torch.zeros(b.n_sub, dtype=torch.bool, device="cuda")
        keep[[0, b.n_sub // 3, b.n_sub - 1]] = True
        sub, ea, nptr, node_keep = select_subdomains(b.csr, b.edge_attr, b.node_ptr, keep)
        part_out = m(x[node_keep], sub, ea)
        assert rel_l2(part_out.cpu().numpy(), again[node_keep].cpu().numpy()) < 1e-6
        # node weight: identical fields -> 0 for every subdomain; antisymmetry under swapping the fields is not
        # expected (max over channels), but scaling both fields scales the weight
        w0 = ops.node_weight(again, again, b.csr, b.edge_attr, b.node_ptr)
        assert bool((w0 == 0).all())
        w1 = ops.node_weight(again, y, b.csr, b.edge_attr, b.node_ptr)
        w2 = ops.node_weight(2.0 * again, 2.0 * y, b.csr, b.edge_attr, b.node_ptr)
        assert w1.shape == (b.n_sub,) and bool(torch.isfinite(w1).all())
        assert float(((w2 - 2.0 * w1).abs() / w1.abs().clamp(min=1e-6)).max()) < 1e-4


def test_teecnet_arms_agree_full_size(shipped, monkeypatch):
    """TEECNet at 526 848 cells: f16 fused / f16 two-kernel / tf32 against the fp32 arm (north_star gate 1e-3 is
    stated against the fp32 CPU result, which the fp32 arm matches to 1e-5 at the sizes the oracle can run)."""
    mesh, pos, cells, part, b, levels = _assembled("500k")
    x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
    m = _model(shipped, "teecnet")
    out = {}
    with torch.no_grad():
        for prec, fuse in (("fp32", "3"), ("tf32", "3"), ("f16", "0"), ("f16", "3")):
            monkeypatch.setenv("FESR_FUSE", fuse)
            m.precision = prec
            out[(prec, fuse)] = m(x, b.csr, b.edge_attr).cpu().numpy()
    ref = out[("fp32", "3")]
    assert np.isfinite(ref).all()
    for k, v in out.items():
        err = rel_l2(v, ref)
        print("teecnet 500k", k, f"rel-L2 vs fp32 arm: {err:.3e}")
        assert err < 1.5e-3, k


def _gpu_grads(m, x, csr, ea, y, precision):
    m.precision = precision
    for p in m.parameters():
        p.grad = None
    loss = torch.nn.functional.mse_loss(m(x, csr, ea), y)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), {k: p.grad.detach().double().cpu().numpy() for k, p in m.named_parameters()}


@pytest.mark.parametrize("precision,tol,tol_add", [("tf32", 5e-3, 2e-3), ("fp32", 2e-4, 2e-4)])
def test_train_gradients_2M(shipped, precision, tol, tol_add):
    """BASELINE config 4 (2 044 416 cells): (1) on a sample of subdomains the gradients of the MSE loss equal fp64
    autograd of the oracle (reference order, models/model.py + scheduler_gnn.py:402-406) within the arm's gate; (2) at
    full size the loss gradient is additive over a split of the batch into two halves of its subdomains,
    n g = n1 g1 + n2 g2 (MSELoss is a mean over nodes x channels) -- every tiling / split-K range of the backward
    kernels is exercised at the size bench.py's `train` record runs."""
    from fesr_b200.models.model import KernelNN
    from fesr_b200.models.scheduler_gnn import select_subdomains
    from oracle import models as om
    mesh, pos, cells, part, b, levels = _assembled("2M")
    x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
    y = torch.from_numpy(mesh.y).cuda()[b.global_ids]
    sd = shipped_state_dict(shipped, "neuralop")
    m = KernelNN(43, 43, 5, in_width=4, out_width=4)
    m.load_state_dict(sd)
    m = m.cuda().train()
    S = b.n_sub

    # (1) three subdomains against fp64 autograd
    keep = torch.zeros(S, dtype=torch.bool, device="cuda")
    keep[[1, S // 2, S - 2]] = True
    sub, ea, nptr, node_keep = select_subdomains(b.csr, b.edge_attr, b.node_ptr, keep)
    loss_g, g = _gpu_grads(m, x[node_keep], sub, ea, y[node_keep], precision)
    rowptr = sub.rowptr.cpu().numpy().astype(np.int64)
    dst = np.repeat(np.arange(sub.n, dtype=np.int64), np.diff(rowptr))
    ei = torch.from_numpy(np.stack([sub.src.cpu().numpy().astype(np.int64), dst]))
    o = om.make_model("neuralop", 43, 5).double()
    o.load_state_dict({k: v.double() for k, v in sd.items()})
    torch.set_num_threads(16)
    lo = torch.nn.functional.mse_loss(o(x[node_keep].double().cpu(), ei, ea.double().cpu()), y[node_keep].double().cpu())
    lo.backward()
    assert abs(loss_g - float(lo)) <= max(tol, 1e-5) * abs(float(lo))
    worst = 0.0
    for k, p in o.named_parameters():
        err = rel_l2(g[k], p.grad.numpy())
        worst = max(worst, err)
        assert err < tol, (k, err)
    print(f"2M sample ({int(node_keep.sum())} nodes) {precision}: max gradient rel-L2 vs fp64 autograd {worst:.2e}")

    # (2) additivity at full size
    _, g_all = _gpu_grads(m, x, b.csr, b.edge_attr, y, precision)
    halves = []
    for lo_s, hi_s in ((0, S // 2), (S // 2, S)):
        keep = torch.zeros(S, dtype=torch.bool, device="cuda")
        keep[lo_s:hi_s] = True
        sub, ea, nptr, node_keep = select_subdomains(b.csr, b.edge_attr, b.node_ptr, keep)
        _, gh = _gpu_grads(m, x[node_keep], sub, ea, y[node_keep], precision)
        halves.append((int(node_keep.sum()), gh))
    n_all = halves[0][0] + halves[1][0]
    assert n_all == b.n_tot
    worst = 0.0
    for k in g_all:
        comb = (halves[0][0] * halves[0][1][k] + halves[1][0] * halves[1][1][k]) / n_all
        err = rel_l2(g_all[k], comb)
        worst = max(worst, err)
        assert err < tol_add, (k, err)
    print(f"2M full batch ({b.n_tot} nodes) {precision}: additivity over two halves, max rel-L2 {worst:.2e}")
