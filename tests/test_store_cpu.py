"""CPU: the on-disk subdomain store (reference layout mesh_{m}/subdomain_{i}/{x,y,pos,edge_index,edge_attr,global_node_ids},
dataset/GraphDataset.py:1128-1133, 1245-1284) -- container round trip and validation (no GPU, no h5py needed)."""
import numpy as np
import pytest
import torch

from fesr_b200.data import Data
from fesr_b200.dataset.store import FIELDS, load_partitioned, save_partitioned


def _meshes(rng, n_mesh=2, n_sub=3):
    out = []
    for m in range(n_mesh):
        subs = []
        for i in range(n_sub):
            n = 4 + i + m
            subs.append(Data(x=torch.from_numpy(rng.normal(size=(n, 4)).astype(np.float32)),
                             y=torch.from_numpy(rng.normal(size=(n, 4)).astype(np.float32)),
                             pos=torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)),
                             edge_index=torch.tensor([[0, 1, 2, 1], [1, 2, 3, 0]]),
                             edge_attr=torch.ones(4, 1), global_node_ids=torch.arange(n) + 10 * i))
        out.append(subs)
    return out


def test_npz_round_trip(tmp_path):
    meshes = _meshes(np.random.default_rng(0))
    path = tmp_path / "store.npz"
    save_partitioned(str(path), meshes)
    back = load_partitioned(str(path))
    assert sorted(back) == [0, 1] and all(len(back[m]) == 3 for m in back)
    for m, subs in enumerate(meshes):
        for i, d in enumerate(subs):
            for k in FIELDS:
                assert np.array_equal(back[m][i][k], getattr(d, k).numpy()), (m, i, k)
    only = load_partitioned(str(path), mesh_indices=[1])
    assert sorted(only) == [1]


def test_store_validation(tmp_path):
    meshes = _meshes(np.random.default_rng(1), n_mesh=1)
    path = tmp_path / "bad.npz"
    save_partitioned(str(path), meshes)
    z = dict(np.load(str(path)))
    del z["mesh_0/subdomain_1/edge_attr"]
    np.savez(str(path), **z)
    with pytest.raises(ValueError, match="missing"):
        load_partitioned(str(path))
    z = {k: v for k, v in dict(np.load(str(tmp_path / "bad.npz"))).items() if "subdomain_1" not in k}
    np.savez(str(path), **z)
    with pytest.raises(ValueError, match="indices"):
        load_partitioned(str(path))


def test_hdf5_store_round_trip_without_h5py(tmp_path, monkeypatch):
    """The reference's container (.h5) through the built-in reader / writer: same arrays, dtypes and shapes."""
    monkeypatch.setenv("FESR_HDF5", "min")
    meshes = _meshes(np.random.default_rng(2), n_mesh=3, n_sub=5)
    path = tmp_path / "data.h5"
    save_partitioned(str(path), meshes)
    assert open(path, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    back = load_partitioned(str(path))
    assert sorted(back) == [0, 1, 2] and all(len(back[m]) == 5 for m in back)
    for m, subs in enumerate(meshes):
        for i, d in enumerate(subs):
            for k in FIELDS:
                want = getattr(d, k).numpy()
                assert back[m][i][k].dtype == want.dtype and np.array_equal(back[m][i][k], want), (m, i, k)
    assert sorted(load_partitioned(str(path), mesh_indices=[2])) == [2]
