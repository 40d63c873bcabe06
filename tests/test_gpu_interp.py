"""GPU: Gaussian-kernel point interpolation (low-res -> high-res transfer, the step before the path) vs the oracle's
restatement of vtkPointInterpolator + vtkGaussianKernel (dataset/GraphDataset.py:1041-1105)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import graph as og

pytestmark = pytest.mark.gpu


def _meshes(n_lo, n_hi):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    return make_duct_mesh(n_lo, seed=3), make_duct_mesh(n_hi, seed=4)


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_interp_matches_oracle(channels):
    from fesr_b200 import ops
    lo, hi = _meshes(4, 7)
    # the two ducts have different extents (n x n x 4n cells of the same size): rescale the coarse one onto the fine one
    scale = (hi.pos.max(0) - hi.pos.min(0)) / (lo.pos.max(0) - lo.pos.min(0))
    sp = ((lo.pos - lo.pos.min(0)) * scale + hi.pos.min(0)).astype(np.float32)
    sv = lo.x[:, :channels].copy()
    spacing = float(np.linalg.norm(sp[lo.cells[:, 0]] - sp[lo.cells[:, 1]], axis=1).mean())
    radius = 1.5 * spacing
    ref, cnt_ref = og.interp_gaussian(sp, sv, hi.pos, radius, 2.0)
    out, cnt = ops.interp_gaussian(torch.from_numpy(sp).cuda(), torch.from_numpy(sv).cuda(),
                                   torch.from_numpy(hi.pos).cuda(), radius, 2.0, want_count=True)
    assert out.shape == (hi.num_nodes, channels)
    assert np.array_equal(cnt.cpu().numpy(), cnt_ref)          # the same source points are in range, everywhere
    assert int(cnt_ref.min()) >= 1 and int(cnt_ref.max()) > 8
    assert rel_l2(out.cpu().numpy(), ref) < 1e-5
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 1e-5


def test_interp_edge_cases():
    from fesr_b200 import ops
    rng = np.random.default_rng(0)
    sp = rng.uniform(-1, 1, size=(500, 3)).astype(np.float32)
    sv = rng.normal(size=(500,)).astype(np.float32)
    # targets: the source points themselves, far-away points (no neighbour -> null value), negative coordinates
    dp = np.concatenate([sp[:50], np.full((3, 3), 50.0, np.float32), -sp[50:80]]).astype(np.float32)
    for radius, sharp in ((0.3, 2.0), (0.05, 1.0), (3.0, 4.0)):
        ref, cnt_ref = og.interp_gaussian(sp, sv, dp, radius, sharp, null_value=-7.0)
        out, cnt = ops.interp_gaussian(torch.from_numpy(sp).cuda(), torch.from_numpy(sv).cuda(),
                                       torch.from_numpy(dp).cuda(), radius, sharp, null_value=-7.0, want_count=True)
        assert out.shape == (dp.shape[0],)                      # 1-D values in, 1-D out
        assert np.array_equal(cnt.cpu().numpy(), cnt_ref)
        assert np.all(cnt_ref[50:53] == 0) and np.all(out.cpu().numpy()[50:53] == -7.0)
        assert np.all(cnt_ref[:50] >= 1)
        assert np.allclose(out.cpu().numpy(), ref[:, 0], rtol=2e-5, atol=2e-6)
    # a constant field comes back constant wherever there is a neighbour; deterministic
    one = torch.full((500, 4), 2.5, device="cuda")
    a = ops.interp_gaussian(torch.from_numpy(sp).cuda(), one, torch.from_numpy(dp).cuda(), 0.4)
    b = ops.interp_gaussian(torch.from_numpy(sp).cuda(), one, torch.from_numpy(dp).cuda(), 0.4)
    assert torch.equal(a, b)
    has = ops.interp_gaussian(torch.from_numpy(sp).cuda(), one, torch.from_numpy(dp).cuda(), 0.4, want_count=True)[1] > 0
    assert float((a[has] - 2.5).abs().max()) < 1e-6
    # no source points at all, no targets at all
    e = ops.interp_gaussian(torch.zeros(0, 3, device="cuda"), torch.zeros(0, 1, device="cuda"),
                            torch.from_numpy(dp).cuda(), 0.4, null_value=1.0)
    assert bool((e == 1.0).all())
    z = ops.interp_gaussian(torch.from_numpy(sp).cuda(), torch.from_numpy(sv).cuda(), torch.zeros(0, 3, device="cuda"), 0.4)
    assert z.shape == (0,)
    with pytest.raises(Exception):
        ops.interp_gaussian(torch.from_numpy(sp).cuda(), torch.from_numpy(sv).cuda(), torch.from_numpy(dp).cuda(), -1.0)


def test_interp_full_size_properties():
    """526 848-cell fine mesh <- 65 856-cell coarse mesh: partition of unity, bounds, every fine node covered."""
    from fesr_b200 import ops
    lo, hi = _meshes(14, 28)
    sp = torch.from_numpy(lo.pos).cuda() * 2.0            # same cell size: the coarse duct is half as long per axis
    dp = torch.from_numpy(hi.pos).cuda()
    sv = torch.from_numpy(lo.x).cuda()
    radius = 3.0 * 2.0 * 3e-3                             # 3 x the coarse spacing (reference: 3 * mesh_spacing)
    out, cnt = ops.interp_gaussian(sp, sv, dp, radius, 2.0, want_count=True)
    assert int(cnt.min()) >= 1
    assert bool(torch.isfinite(out).all())
    # a weighted mean with positive weights stays inside the range of the source values, channel by channel
    assert bool((out.max(0).values <= sv.max(0).values + 1e-6).all()) and bool((out.min(0).values >= sv.min(0).values - 1e-6).all())
    ones = ops.interp_gaussian(sp, torch.ones_like(sv[:, :1]), dp, radius, 2.0)
    assert float((ones - 1.0).abs().max()) < 1e-6
    # linear in the values
    a = ops.interp_gaussian(sp, 2.0 * sv - 1.0, dp, radius, 2.0)
    assert float((a - (2.0 * out - 1.0)).abs().max()) < 1e-5


def test_dataset_lagrangian_interpolation_signature():
    """The reference's static method (dataset/GraphDataset.py:1041): (mesh, physics, new_mesh) -> [n_new, 1] float."""
    from fesr_b200.dataset.GraphDataset import AnsysDataset
    lo, hi = _meshes(4, 6)
    sp = lo.pos * (hi.pos.max(0) / lo.pos.max(0))

    class M:
        def __init__(self, pos):
            self.pos = pos
    spacing = 3e-3 * 6 / 4
    out = AnsysDataset._lagrangian_interpolation(M(sp), lo.x[:, 3:4], M(hi.pos), mesh_spacing=spacing)
    assert out.shape == (hi.num_nodes, 1) and out.dtype == torch.float32 and out.is_cuda
    ref, _ = og.interp_gaussian(sp, lo.x[:, 3], hi.pos, 3.0 * spacing, 2.0)
    assert rel_l2(out.cpu().numpy(), ref) < 1e-5
    with pytest.raises(ValueError):
        AnsysDataset._lagrangian_interpolation(M(sp), lo.x[:5, 3], M(hi.pos))
