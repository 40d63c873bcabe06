"""Generate golden vectors by executing the REFERENCE's own Python classes.

Run once in the build container (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed from /root/reference, unmodified, and how:
  * models/model.py (KernelNN, NNConv_old, DenseNet, TEECNet, KernelConv) -- imported as a
    module.  Its only missing dependency, torch_geometric, is replaced by the ~40-line stub
    below that restates PyG 2.6.1 ``MessagePassing(aggr='mean')`` semantics (flow
    source_to_target: x_j = x[edge_index[0]], x_i = x[edge_index[1]], reduce over
    edge_index[1] with dim_size = x.size(0), sum / clamp(count, 1)) and
    ``inits.reset / inits.uniform``.
  * models/scheduler_gnn.py:472-514 GradientbasedLoss -- the class node is pulled out of the
    file with ``ast`` and compiled as-is (the module itself imports vtk/matplotlib/pyg).
  * dataset/GraphDataset.py:838-869 AnsysDataset.vtk_to_pyg -- same ``ast`` extraction, fed a
    duck-typed stand-in for vtkUnstructuredGrid (GetNumberOfPoints/GetPoint/GetCell/...).
  * models/encoder.py PCAEncoder and models/classifier.py KMeansClassifier -- imported for
    real (sklearn is installed).
Shipped checkpoints logs/models/collection_duct_{neuralop,teecnet}/partition_0.pth provide
the w=43 weights; they are re-saved as ``shipped_w43_weights.npz`` (arrays keyed
``neuralop::<state_dict key>`` / ``teecnet::<key>``) so that the GPU box, which has no
/root/reference, can run the w=43 parity tests and the benchmark on realistic weights.
"""
from __future__ import annotations

import ast
import importlib.util
import inspect
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# ---------------------------------------------------------------- torch_geometric stub
def _install_pyg_stub():
    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add", flow="source_to_target", **kwargs):
            super().__init__()
            self.aggr = aggr
            assert flow == "source_to_target"

        def propagate(self, edge_index, **kwargs):
            n = kwargs["x"].size(0)
            src, dst = edge_index[0], edge_index[1]
            margs = {}
            for name in inspect.signature(self.message).parameters:
                if name.endswith("_j"):
                    margs[name] = kwargs[name[:-2]].index_select(0, src)
                elif name.endswith("_i"):
                    margs[name] = kwargs[name[:-2]].index_select(0, dst)
                else:
                    margs[name] = kwargs[name]
            msg = self.message(**margs)
            out = torch.zeros(n, msg.size(1), dtype=msg.dtype).index_add_(0, dst, msg)
            if self.aggr == "mean":
                cnt = torch.zeros(n, dtype=msg.dtype).index_add_(0, dst, torch.ones_like(dst, dtype=msg.dtype))
                out = out / cnt.clamp(min=1).unsqueeze(-1)
            elif self.aggr != "add":
                raise NotImplementedError(self.aggr)
            uargs = {k: kwargs[k] for k in list(inspect.signature(self.update).parameters)[1:]}
            return self.update(out, **uargs)

    def reset(value):
        if hasattr(value, "reset_parameters"):
            value.reset_parameters()
        else:
            for child in value.children() if hasattr(value, "children") else []:
                reset(child)

    def uniform(size, value):
        if isinstance(value, torch.Tensor):
            bound = 1.0 / np.sqrt(size)
            value.data.uniform_(-bound, bound)

    class Data:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    pyg = types.ModuleType("torch_geometric")
    pyg_nn = types.ModuleType("torch_geometric.nn")
    pyg_inits = types.ModuleType("torch_geometric.nn.inits")
    pyg_data = types.ModuleType("torch_geometric.data")
    pyg_nn.MessagePassing = MessagePassing
    pyg_inits.reset, pyg_inits.uniform = reset, uniform
    pyg_data.Data = Data
    pyg.nn, pyg.data = pyg_nn, pyg_data
    pyg_nn.inits = pyg_inits
    sys.modules.update({"torch_geometric": pyg, "torch_geometric.nn": pyg_nn,
                        "torch_geometric.nn.inits": pyg_inits, "torch_geometric.data": pyg_data})
    return Data


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _extract(path, class_name, func_name=None, ns=None):
    """Compile one class (or one of its functions) of a reference file without importing it."""
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            if func_name is None:
                mod = ast.Module(body=[node], type_ignores=[])
            else:
                fn = [f for f in node.body if isinstance(f, ast.FunctionDef) and f.name == func_name][0]
                fn.decorator_list = []
                mod = ast.Module(body=[fn], type_ignores=[])
            ns = dict(ns or {})
            exec(compile(mod, path, "exec"), ns)
            return ns[func_name or class_name]
    raise KeyError(class_name)


# ---------------------------------------------------------------- fake vtk grid
class _FakeCell:
    def __init__(self, ids):
        self.ids = ids

    def GetNumberOfPoints(self):
        return len(self.ids)

    def GetPointId(self, j):
        return int(self.ids[j])


class _FakeGrid:
    def __init__(self, pos, cells):
        self.pos, self.cells = pos, cells

    def GetNumberOfPoints(self):
        return self.pos.shape[0]

    def GetPoint(self, i):
        return tuple(float(v) for v in self.pos[i])

    def GetNumberOfCells(self):
        return self.cells.shape[0]

    def GetCell(self, i):
        return _FakeCell(self.cells[i])


# ---------------------------------------------------------------- a10: the reference's averaging loop
class _FakeIdList:
    def __init__(self):
        self.ids = []

    def GetNumberOfIds(self):
        return len(self.ids)

    def GetId(self, j):
        return int(self.ids[j])


class _FakePointLocator:
    """vtkStaticPointLocator stand-in: FindPointsWithinRadius through scipy's cKDTree; ids ascending (VTK returns
    them in bucket order -- within one bucket ascending point id; the coincident copies of a node share a bucket)."""

    def SetDataSet(self, grid):
        from scipy.spatial import cKDTree
        self.tree = cKDTree(grid.points.astype(np.float64))

    def AutomaticOn(self):
        pass

    def BuildLocator(self):
        pass

    def FindPointsWithinRadius(self, r, x, id_list):
        id_list.ids = sorted(self.tree.query_ball_point(np.asarray(x, dtype=np.float64), r))


class _FakePointData:
    def __init__(self, arrays):
        self.arrays = arrays                     # name -> numpy (what vtk_to_numpy hands back: [n, 3] or [n])
        self.names = list(arrays)

    def GetNumberOfArrays(self):
        return len(self.names)

    def GetArrayName(self, i):
        return self.names[i]

    def GetArray(self, name):
        return self.arrays[name]


class _FakeMergedGrid:
    """The vtkAppendDataSets output: all partitions' points appended, with their point arrays."""

    def __init__(self, points, arrays):
        self.points, self.pd = points, _FakePointData(arrays)

    def GetNumberOfPoints(self):
        return self.points.shape[0]

    def GetPoint(self, i):
        return tuple(float(v) for v in self.points[i])

    def GetPointData(self):
        return self.pd


def _reference_averaging(points, arrays):
    """Executes the statements of AnsysDataset.reconstruct_from_partition between `locator = ...` and the end of the
    per-point loop (dataset/GraphDataset.py:1371-1400) -- pulled out of the file with `ast`, unmodified -- on a
    duck-typed merged grid.  Returns the averaged arrays (`array_np` of the reference)."""
    path = os.path.join(REF, "dataset/GraphDataset.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "AnsysDataset"][0]
    fn = [f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "reconstruct_from_partition"][0]
    first = [i for i, st in enumerate(fn.body) if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "locator"][0]
    last = [i for i, st in enumerate(fn.body) if isinstance(st, ast.For) and getattr(st.target, "id", "") == "i"
            and i > first][0]
    body = fn.body[first:last + 1]
    assert body[0].lineno == 1371 and body[-1].end_lineno == 1400, (body[0].lineno, body[-1].end_lineno)
    vtk = types.SimpleNamespace(vtkStaticPointLocator=_FakePointLocator, vtkIdList=_FakeIdList)
    ns = {"vtk": vtk, "np": np, "vtk_to_numpy": lambda a: a,
          "merged_grid": _FakeMergedGrid(points, {k: v.copy() for k, v in arrays.items()})}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns["array_np"]


def make_stitch_golden():
    """a10 fixture: per-subdomain predictions / reference rows of two small ducts appended the way the reference
    appends its partitions (subdomain order), averaged by the reference's own loop."""
    from fesr_b200.dataset.synthetic import make_duct_mesh
    from oracle import graph as og
    out = {}
    for tag, n, levels in (("a", 3, 2), ("b", 6, 5)):
        mesh = make_duct_mesh(n)
        part = og.kd_partition(mesh.pos, mesh.cells, levels)
        sub = og.build_subdomains(mesh.pos, mesh.cells, part["leaf_ptr"], part["leaf_cells"])
        gids = sub["global_ids"]
        rng = np.random.default_rng(7 + n)
        # per-copy predictions differ between the subdomains that share a node (as real predictions do)
        pred = (mesh.y[gids] + rng.normal(0.0, 0.05, size=(gids.size, 4))).astype(np.float32)
        ref = mesh.y[gids].astype(np.float32)
        arrays = {"velocity": pred[:, :3].copy(), "pressure": pred[:, 3].copy(),
                  "ref_velocity": ref[:, :3].copy(), "ref_pressure": ref[:, 3].copy()}
        avg = _reference_averaging(mesh.pos[gids], arrays)
        cnt = np.bincount(gids, minlength=mesh.num_nodes)
        out.update({f"{tag}_mesh_n": np.int64(n), f"{tag}_levels": np.int64(levels), f"{tag}_global_ids": gids.astype(np.int64),
                    f"{tag}_pred": pred, f"{tag}_ref": ref, f"{tag}_max_copies": np.int64(cnt.max()),
                    f"{tag}_merged": np.concatenate([avg["velocity"], avg["pressure"][:, None]], axis=1).astype(np.float32),
                    f"{tag}_merged_ref": np.concatenate([avg["ref_velocity"], avg["ref_pressure"][:, None]], axis=1).astype(np.float32)})
        print("stitch golden", tag, "points", gids.size, "copies histogram", np.bincount(cnt).tolist())
    path = os.path.join(HERE, "stitch_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    from fesr_b200.dataset.synthetic import make_duct_mesh
    from oracle import graph as og

    Data = _install_pyg_stub()
    ref_model = _load_module("ref_model", os.path.join(REF, "models/model.py"))
    GLoss = _extract(os.path.join(REF, "models/scheduler_gnn.py"), "GradientbasedLoss",
                     ns={"torch": torch, "nn": torch.nn})
    vtk_to_pyg = _extract(os.path.join(REF, "dataset/GraphDataset.py"), "AnsysDataset", "vtk_to_pyg",
                          ns={"torch": torch, "np": np, "Data": Data})

    out = {}
    # ---- graph build (a1) on a 3x3x12 duct: 648 cells, 208 nodes
    mesh = make_duct_mesh(3)
    d = vtk_to_pyg(_FakeGrid(mesh.pos, mesh.cells))
    ei = d.edge_index.numpy().astype(np.int64)            # reference set order (arbitrary)
    ea = d.edge_attr.numpy().astype(np.float32)           # [E,1]
    out.update(mesh_n=np.int64(3), pos=mesh.pos, cells=mesh.cells, x=mesh.x, y=mesh.y,
               ref_edge_index=ei, ref_edge_attr=ea)

    x = torch.from_numpy(mesh.x)
    y = torch.from_numpy(mesh.y)
    eit = torch.from_numpy(ei)
    eat = torch.from_numpy(ea)

    # ---- models with the shipped w=43 checkpoints (a6, a7)
    torch.manual_seed(0)
    knn = ref_model.KernelNN(width=43, ker_width=43, depth=5, in_width=4, out_width=4)
    knn.load_state_dict(torch.load(os.path.join(REF, "logs/models/collection_duct_neuralop/partition_0.pth"),
                                   map_location="cpu", weights_only=True))
    tee = ref_model.TEECNet(4, out_channels=4, width=43, num_layers=5, retrieve_weight=False)
    tee.load_state_dict(torch.load(os.path.join(REF, "logs/models/collection_duct_teecnet/partition_0.pth"),
                                   map_location="cpu", weights_only=True))
    with torch.no_grad():
        out["kernelnn_w43_y"] = knn.eval()(x, eit, eat).numpy()
        out["teecnet_w43_y"] = tee.eval()(x, eit, eat.squeeze(1)).numpy()   # [E] edge_attr as Duct stores it

    # ---- seeded small-width models (state_dicts travel in the fixture)
    for name, ctor in (("kernelnn_w16", lambda: ref_model.KernelNN(width=16, ker_width=16, depth=3, in_width=4, out_width=4)),
                       ("teecnet_w12", lambda: ref_model.TEECNet(4, out_channels=4, width=12, num_layers=2, retrieve_weight=False)),
                       ("kernelnn_w48", lambda: ref_model.KernelNN(width=48, ker_width=48, depth=2, in_width=4, out_width=4))):
        torch.manual_seed(1234)
        m = ctor()
        for k, v in m.state_dict().items():
            out[f"{name}_sd::{k}"] = v.numpy().copy()
        with torch.no_grad():
            out[f"{name}_y"] = m.eval()(x, eit, eat).numpy()
        # one training step, reference order (scheduler_gnn.py:398-408), Adam lr from teecnet.yaml:3
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=0.0005)
        opt.zero_grad()
        o = m(x, eit, eat)
        loss = torch.nn.MSELoss()(o, y)
        loss.backward()
        out[f"{name}_loss"] = loss.detach().numpy()
        for k, p in m.named_parameters():
            out[f"{name}_grad::{k}"] = p.grad.numpy().copy()
        opt.step()
        for k, v in m.state_dict().items():
            out[f"{name}_sd_after::{k}"] = v.numpy().copy()

    # ---- GradientbasedLoss (a9)
    crit = GLoss()
    pred = torch.from_numpy(out["kernelnn_w43_y"])
    out["node_weight"] = crit.compute_node_weight(pred, y, eit, eat, x.shape[0]).numpy()
    out["gradient_loss"] = crit(pred, y, eit, eat).numpy()
    out["gradient_loss_mw4"] = GLoss(max_weight=4)(pred, y, eit, eat).numpy()

    # ---- ALDS routing (a4): the reference's PCAEncoder + KMeansClassifier, for real
    sys.path.insert(0, REF)
    enc_mod = _load_module("ref_encoder", os.path.join(REF, "models/encoder.py"))
    cls_mod = _load_module("ref_classifier", os.path.join(REF, "models/classifier.py"))
    big = make_duct_mesh(12)
    part = og.kd_partition(big.pos, big.cells, 4)
    sub = og.build_subdomains(big.pos, big.cells, part["leaf_ptr"], part["leaf_cells"])
    xs = [torch.from_numpy(big.x[sub["global_ids"][sub["node_ptr"][s]:sub["node_ptr"][s + 1]]])
          for s in range(16)]
    assert min(t.shape[0] for t in xs) >= 280
    dataset = [Data(x=t) for t in xs]
    enc = enc_mod.PCAEncoder(n_components=2)
    enc._train_graph([Data(x=t[:280]) for t in xs])           # train() cuts to the min length; predict cuts to 280
    latent = enc.get_latent_space(dataset)
    clf = cls_mod.KMeansClassifier(n_clusters=4)
    clf.train(latent)
    labels = clf.cluster(latent)
    out.update(route_mesh_n=np.int64(12), route_levels=np.int64(4),
               route_pca_mean=enc.model.mean_, route_pca_components=enc.model.components_,
               route_scaler_mean=clf.scaler.mean_, route_scaler_scale=clf.scaler.scale_,
               route_centroids=clf.model.cluster_centers_, route_latent=latent,
               route_labels=np.asarray(labels, dtype=np.int64))

    weights = {}
    for tag, m in (("neuralop", knn), ("teecnet", tee)):
        for k, v in m.state_dict().items():
            weights[f"{tag}::{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "shipped_w43_weights.npz"), **weights)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    print("labels", labels)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stitch":
        make_stitch_golden()
    else:
        main()
        make_stitch_golden()
