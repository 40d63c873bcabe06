"""GPU parity: backward / train step vs gradients produced by the reference's own classes
(tests/golden) and vs torch.autograd of the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, shipped_state_dict, state_dict_from
from oracle import graph as og
from oracle import models as om

pytestmark = pytest.mark.gpu


def _model(kind, w, L):
    from fesr_b200.models.model import KernelNN, TEECNet
    if kind == "neuralop":
        return KernelNN(w, w, L, in_width=4, out_width=4)
    return TEECNet(4, w, 4, num_layers=L, retrieve_weight=False)


def _grads(model, x, ei, ea, y, precision="fp32"):
    model = model.cuda().train()
    model.precision = precision
    for p in model.parameters():
        p.grad = None
    out = model(torch.from_numpy(x).cuda(), torch.from_numpy(ei).cuda(), torch.from_numpy(ea).cuda())
    loss = torch.nn.functional.mse_loss(out, torch.from_numpy(y).cuda())
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}


@pytest.mark.parametrize("name,kind,w,L", [("kernelnn_w16", "neuralop", 16, 3), ("teecnet_w12", "teecnet", 12, 2),
                                           ("kernelnn_w48", "neuralop", 48, 2)])
def test_grads_vs_reference_vectors(golden, name, kind, w, L):
    m = _model(kind, w, L)
    m.load_state_dict(state_dict_from(golden, name))
    loss, g = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    assert abs(loss - float(golden[name + "_loss"])) <= 1e-5 * abs(float(golden[name + "_loss"]))
    for k, v in g.items():
        ref = golden[f"{name}_grad::{k}"]
        assert v.shape == ref.shape, k
        assert rel_l2(v, ref) < 2e-4, (k, rel_l2(v, ref))


@pytest.mark.parametrize("kind", ["neuralop", "teecnet"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tf32", 5e-3)])
def test_grads_w43_vs_oracle_autograd(shipped, kind, precision, tol):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    mesh = make_duct_mesh(5)
    src, dst, ea = og.build_edges(mesh.cells, mesh.pos)
    ei = np.stack([src, dst])
    sd = shipped_state_dict(shipped, kind)
    m = _model(kind, 43, 5)
    m.load_state_dict(sd)
    o = om.make_model(kind, 43, 5).double()
    o.load_state_dict({k: v.double() for k, v in sd.items()})
    out = o(torch.from_numpy(mesh.x).double(), torch.from_numpy(ei), torch.from_numpy(ea).double())
    lo = torch.nn.functional.mse_loss(out, torch.from_numpy(mesh.y).double())
    lo.backward()
    loss, g = _grads(m, mesh.x, ei, ea, mesh.y, precision)
    assert abs(loss - float(lo)) <= max(tol, 1e-5) * abs(float(lo))
    for k, p in o.named_parameters():
        err = rel_l2(g[k], p.grad.numpy())
        assert err < tol, (k, err)


def test_grad_x(golden):
    m = _model("neuralop", 16, 3).cuda()
    sd = state_dict_from(golden, "kernelnn_w16")
    m.load_state_dict(sd)
    o = om.make_model("neuralop", 16, 3)
    o.load_state_dict(sd)
    x = torch.from_numpy(golden["x"]).requires_grad_(True)
    ei, ea = torch.from_numpy(golden["ref_edge_index"]), torch.from_numpy(golden["ref_edge_attr"])
    o(x, ei, ea).square().sum().backward()
    xg = torch.from_numpy(golden["x"]).cuda().requires_grad_(True)
    m(xg, ei.cuda(), ea.cuda()).square().sum().backward()
    assert rel_l2(xg.grad.cpu().numpy(), x.grad.numpy()) < 2e-4


def test_backward_is_deterministic(golden):
    m = _model("neuralop", 16, 3)
    m.load_state_dict(state_dict_from(golden, "kernelnn_w16"))
    _, a = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    _, b = _grads(m, golden["x"], golden["ref_edge_index"], golden["ref_edge_attr"], golden["y"])
    for k in a:
        assert np.array_equal(a[k], b[k]), k


# ----------------------------------------------------------------------------- train step / loop
@pytest.mark.parametrize("name,kind,w,L", [("kernelnn_w16", "neuralop", 16, 3), ("teecnet_w12", "teecnet", 12, 2)])
def test_train_step_vs_reference_vectors(golden, name, kind, w, L):
    """zero_grad -> forward -> MSELoss -> backward -> Adam(lr=5e-4).step (scheduler_gnn.py:402-409)."""
    from fesr_b200 import ops
    from fesr_b200.models.training import FlatAdam, train_step
    m = _model(kind, w, L).cuda().train()
    sd = state_dict_from(golden, name)
    m.load_state_dict(sd)
    opt = FlatAdam(m, lr=0.0005)
    ei = torch.from_numpy(golden["ref_edge_index"]).cuda()
    csr = ops.csr_build(ei, golden["x"].shape[0])
    loss = train_step(m, opt, torch.from_numpy(golden["x"]).cuda(), csr, torch.from_numpy(golden["ref_edge_attr"]).cuda(),
                      torch.from_numpy(golden["y"]).cuda())
    assert abs(float(loss) - float(golden[name + "_loss"])) <= 1e-5 * abs(float(golden[name + "_loss"]))
    after = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for k, v in after.items():
        ref_after, before = golden[f"{name}_sd_after::{k}"], sd[k].numpy()
        assert rel_l2(v, ref_after) < 1e-6, k
        # the Adam update itself (first step: -lr * g / (|g| + eps)); entries with |g| ~ eps are ill-conditioned
        d, dref = v - before, ref_after - before
        assert np.abs(d - dref).max() <= 0.02 * 0.0005 + 1e-9, (k, np.abs(d - dref).max())


def test_scheduler_train_loop_writes_checkpoint_and_learns(tmp_path, monkeypatch, shipped):
    from fesr_b200.dataset.GraphDataset import AnsysDataset
    from fesr_b200.models.model import KernelNN
    from fesr_b200.models.scheduler_gnn import GNNPartitionScheduler
    monkeypatch.chdir(tmp_path)
    ds = AnsysDataset(mesh_n=6, num_meshes=1, sub_size=8)
    torch.manual_seed(0)
    model = KernelNN(16, 16, 3, in_width=4, out_width=4)
    sched = GNNPartitionScheduler("tr", 1, ds, model, train=True)
    cfg = {"epochs": 6, "batch_size": 2, "lr": 0.005, "step_size": 30, "gamma": 0.1, "log_interval": 1, "val_interval": 2}
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        sched.train(cfg)
    losses = [float(l.split("Train loss:")[1]) for l in buf.getvalue().splitlines() if "Train loss" in l]
    assert len(losses) == 6 and losses[-1] < losses[0]
    sd = torch.load("logs/models/collection_tr/partition_0.pth", weights_only=True)
    assert set(sd) == set(model.state_dict())
    # the checkpoint loads back through the predict path
    sched2 = GNNPartitionScheduler("tr", 1, ds, KernelNN(16, 16, 3, in_width=4, out_width=4), train=False)
    p, r, mi, w = sched2.predict(ds.get_one_full_sample(0))
    assert len(p) == 8 and all(torch.isfinite(t).all() for t in p)


def test_run_script_end_to_end(tmp_path, shipped):
    """python run_DS_3D.py --mode predict ... : flags, timers and the .vtu output of the entry script."""
    import subprocess, sys, os
    from conftest import ROOT
    os.makedirs(tmp_path / "logs/models/collection_duct_neuralop", exist_ok=True)
    torch.save(shipped_state_dict(shipped, "neuralop"), tmp_path / "logs/models/collection_duct_neuralop/partition_0.pth")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_DS_3D.py"), "--mode", "predict", "--model", "neuralop",
                        "--dataset", "duct", "--exp_name", "duct_neuralop", "--exp_config",
                        os.path.join(ROOT, "configs/exp_config/teecnet_duct.yaml"), "--train_config",
                        os.path.join(ROOT, "configs/train_config/teecnet.yaml")], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Prediction time:" in r.stdout and "Reconstruction time:" in r.stdout and "Prediction done!" in r.stdout
    assert os.path.getsize(tmp_path / "logs/vtk/duct_neuralop/pred_1.vtu") > 10000


def test_backward_kernel_switches_agree_in_their_own_process(shipped):
    """The tcgen05 weight gradient and the all-layer edge gradient (default on) against the per-layer mma.sync kernels they
    replaced (FESR_WGRAD_TC=0 FESR_EDGE_GRAD_LAYERS=0; the switches are read once per process, hence the subprocess), on a
    ragged graph: isolated nodes, a hub of 150 in-edges (ten 16-edge chunks), a node count that is no multiple of 64."""
    import os
    import subprocess
    import sys
    import tempfile
    from conftest import ROOT
    rng = np.random.default_rng(11)
    n = 64 * 9 + 21
    deg = rng.integers(0, 24, size=n)
    deg[7] = 150
    dst = np.repeat(np.arange(n), deg)
    src = rng.integers(0, n, size=dst.size)
    ei = np.stack([src, dst]).astype(np.int64)
    ea = rng.uniform(2e-3, 6e-3, size=dst.size).astype(np.float32)
    x = rng.uniform(0.0, 1.0, size=(n, 4)).astype(np.float32)
    y = rng.uniform(0.0, 1.0, size=(n, 4)).astype(np.float32)
    sd = shipped_state_dict(shipped, "neuralop")
    m = _model("neuralop", 43, 5)
    m.load_state_dict(sd)
    loss, g = _grads(m, x, ei, ea, y, "tf32")
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "in.npz"), x=x, y=y, ei=ei, ea=ea, **{"sd::" + k: v.numpy() for k, v in sd.items()})
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "from fesr_b200.models.model import KernelNN\n"
            f"z = np.load({os.path.join(td, 'in.npz')!r})\n"
            "m = KernelNN(43, 43, 5, in_width=4, out_width=4)\n"
            "m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd::')})\n"
            "m = m.cuda().train(); m.precision = 'tf32'\n"
            "out = m(torch.from_numpy(z['x']).cuda(), torch.from_numpy(z['ei']).cuda(), torch.from_numpy(z['ea']).cuda())\n"
            "torch.nn.functional.mse_loss(out, torch.from_numpy(z['y']).cuda()).backward()\n"
            f"np.savez({os.path.join(td, 'out.npz')!r}, **{{k: p.grad.cpu().numpy() for k, p in m.named_parameters()}})\n")
        env = dict(os.environ, FESR_WGRAD_TC="0", FESR_EDGE_GRAD_LAYERS="0")
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        old = dict(np.load(os.path.join(td, "out.npz")))
    assert set(old) == set(g)
    for k, v in g.items():
        assert np.isfinite(v).all(), k
        # both are tf32-class evaluations of the same gradient (5e-3 against fp64 autograd above)
        assert rel_l2(v, old[k]) < 2e-3, (k, rel_l2(v, old[k]))
