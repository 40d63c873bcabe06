"""GPU: the two ways the predict path could return a plausible WRONG field, closed.

1. fp16 saturation: the f16 / tf32 arms pack intermediates into fp16 (saturating converts, fp16-accumulated outer
   products).  Un-normalised inputs must not come back as a clipped but plausible field: the pass raises its range
   flag, the output is NaN, and predict()'s host lists raise on first touch.
2. stale prepared weights: fesr_adam_step writes the parameters through a raw pointer; the predict path's cached
   padded / permuted weight copies must be rebuilt after it (validation forward -> train steps -> validation forward).
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2, shipped_state_dict, state_dict_from
from oracle import models as om

pytestmark = pytest.mark.gpu


def _graph(n_side=6):
    from fesr_b200 import ops
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(n_side)
    part, b = ops.assemble(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(), 2)
    x = torch.from_numpy(mesh.x).cuda()[b.global_ids]
    y = torch.from_numpy(mesh.y).cuda()[b.global_ids]
    return mesh, b, x, y


@pytest.mark.parametrize("prec,fuse", [("f16", "3"), ("f16", "0")])
def test_f16_overflow_is_flagged_not_clipped(shipped, monkeypatch, prec, fuse):
    from fesr_b200 import ops
    from fesr_b200.models.model import KernelNN
    mesh, b, x, y = _graph()
    m = KernelNN(43, 43, 5, in_width=4, out_width=4)
    m.load_state_dict(shipped_state_dict(shipped, "neuralop"))
    m = m.cuda().eval()
    m.precision = prec
    monkeypatch.setenv("FESR_FUSE", fuse)
    with torch.no_grad():
        ok = m(x, b.csr, b.edge_attr)
        flag = ops.overflow_flag(m.dims, x.device)
        assert flag is not None and int(flag.item()) == 0 and bool(torch.isfinite(ok).all())
        # the same field un-normalised: at every scale EITHER the pass stays inside the fp16 range and agrees with the
        # fp32 arm, OR it raises the flag and returns NaN -- never a clipped, plausible-looking field
        m32 = KernelNN(43, 43, 5, in_width=4, out_width=4)
        m32.load_state_dict(shipped_state_dict(shipped, "neuralop"))
        m32 = m32.cuda().eval()
        m32.precision = "fp32"
        m32.ws_tag = "fwd_fp32_ref"
        flagged = []
        for scale in (1.0e2, 1.0e3, 1.0e4, 3.0e4, 1.0e5, 1.0e6):
            out = m(x * scale, b.csr, b.edge_attr)
            f = int(ops.overflow_flag(m.dims, x.device).item())
            flagged.append(f)
            if f:
                assert bool(torch.isnan(out).all()), f"scale {scale}: an overflowed pass must not return numbers"
            else:
                ref = m32(x * scale, b.csr, b.edge_attr)
                assert bool(torch.isfinite(ref).all())
                err = rel_l2(out.cpu().numpy(), ref.cpu().numpy())
                assert err < 2e-3, f"scale {scale}: no flag but rel-L2 {err:.2e} vs the fp32 arm"
        assert flagged[0] == 0 and flagged[-1] == 1, flagged
        # the flag belongs to the pass: the next normal pass clears it and reproduces the first result bit for bit
        again = m(x, b.csr, b.edge_attr)
        assert int(ops.overflow_flag(m.dims, x.device).item()) == 0 and torch.equal(again, ok)


def test_predict_lists_raise_on_overflow(tmp_path, shipped, monkeypatch):
    import os
    from fesr_b200._lib import FesrError
    from fesr_b200.dataset.GraphDataset import AnsysDataset
    from fesr_b200.models.model import KernelNN
    from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
    monkeypatch.chdir(tmp_path)
    os.makedirs("logs/models/collection_g", exist_ok=True)
    torch.save(shipped_state_dict(shipped, "neuralop"), "logs/models/collection_g/partition_0.pth")
    ds = AnsysDataset(mesh_n=6, num_meshes=1, sub_size=4)
    sched = GNNPartitionScheduler("g", 1, ds, KernelNN(43, 43, 5, in_width=4, out_width=4), train=False)
    sched.models[0].precision = "f16"
    sample = ds.get_one_full_sample(0)
    p, r, mi, w = sched.predict(sample)
    assert torch.isfinite(torch.cat(list(p))).all()
    c = ds._mesh(0)
    hot = sample.with_host_inputs((c["x"] * 1.0e6).cpu().pin_memory(), c["y"].cpu().pin_memory())
    p, r, mi, w = sched.predict(hot)
    with pytest.raises(FesrError, match="fp16 overflow"):
        p[0]


def test_adam_step_invalidates_prepared_weights(golden):
    """Validation forward, train steps through the raw-pointer Adam kernel, validation forward: the second one must
    use the UPDATED weights (oracle with the updated state_dict), not the copies prepared for the first."""
    from fesr_b200 import ops
    from fesr_b200.models.model import KernelNN
    from fesr_b200.models.training import FlatAdam, train_step
    sd = state_dict_from(golden, "kernelnn_w16")
    m = KernelNN(16, 16, 3, in_width=4, out_width=4)
    m.load_state_dict(sd)
    m = m.cuda()
    x = torch.from_numpy(golden["x"]).cuda()
    y = torch.from_numpy(golden["y"]).cuda()
    ei = torch.from_numpy(golden["ref_edge_index"]).cuda()
    ea = torch.from_numpy(golden["ref_edge_attr"]).cuda()
    csr = ops.csr_build(ei, x.shape[0])

    def oracle_out():
        o = om.make_model("neuralop", 16, 3)
        o.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()})
        with torch.no_grad():
            return o(x.cpu(), ei.cpu(), ea.cpu().reshape(-1, 1)).numpy()

    with torch.no_grad():
        m.eval()
        v0 = m(x, csr, ea)
    assert rel_l2(v0.cpu().numpy(), oracle_out()) < 1e-5
    opt = FlatAdam(m, lr=0.01)
    m.train()
    for _ in range(3):
        train_step(m, opt, x, csr, ea, y)
    with torch.no_grad():
        m.eval()
        v1 = m(x, csr, ea)
    ref1 = oracle_out()
    assert rel_l2(v0.cpu().numpy(), ref1) > 1e-3, "three Adam steps at lr 0.01 must move the output"
    assert rel_l2(v1.cpu().numpy(), ref1) < 1e-5, "predict after Adam ran with stale prepared weights"
    # a second model loaded into recycled storage must not inherit the first one's prepared copies
    del m, opt
    m2 = KernelNN(16, 16, 3, in_width=4, out_width=4)
    m2.load_state_dict(sd)
    m2 = m2.cuda().eval()
    with torch.no_grad():
        v2 = m2(x, csr, ea)
    assert torch.equal(v2, v0)
