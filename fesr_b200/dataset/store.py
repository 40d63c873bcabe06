"""On-disk subdomain store with the reference's layout (dataset/GraphDataset.py:1128-1133, 1245-1284):

    mesh_{m}/subdomain_{i}/{x, y, pos, edge_index, edge_attr, global_node_ids}

The reference writes it as HDF5 groups through h5py.  Two containers behind one reader / writer:

  * ``.h5`` / ``.hdf5`` -- the reference's own files.  Read and written through ``hdf5_min`` (this package's
    dependency-free implementation of the HDF5 subset h5py's defaults produce: version-0 superblock, symbol-table
    groups, contiguous / compact / unfiltered-chunked datasets of little-endian numbers); when h5py happens to be
    importable it is used instead (``FESR_HDF5=min`` forces the built-in reader);
  * ``.npz``           -- the same hierarchy with the group path as the key (``mesh_0/subdomain_3/x``).

``StoredSubdomainDataset`` serves ``get_one_full_sample`` / ``reconstruct_from_partition`` from such a store: the
stored subdomains (any edge order, as the reference's Python ``set`` leaves it) become one block-diagonal device batch
through fesr_csr_build, and the stitch works from the stored ``global_node_ids`` -- the reference's decomposition is
used as stored, no GPU re-assembly.
"""
from __future__ import annotations

import re

import numpy as np
import torch

from .. import ops
from ..data import Data
from .GraphDataset import StitchedMesh, SubdomainSample

FIELDS = ("x", "y", "pos", "edge_index", "edge_attr", "global_node_ids")
_KEY = re.compile(r"^mesh_(\d+)/subdomain_(\d+)/(\w+)$")


def _is_h5(path):
    return str(path).endswith((".h5", ".hdf5"))


def _h5py():
    """h5py if it is importable and not switched off, else None (the built-in hdf5_min is used)."""
    import os
    if os.environ.get("FESR_HDF5", "") == "min":
        return None
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def save_partitioned(path, meshes):
    """meshes: list (one entry per mesh) of lists of Data / dicts with the FIELDS above."""
    def arr(d, k):
        v = d[k] if isinstance(d, dict) else getattr(d, k)
        return v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)

    if _is_h5(path):
        h5py = _h5py()
        if h5py is None:
            from .hdf5_min import write_hdf5
            write_hdf5(path, {f"mesh_{m}": {f"subdomain_{i}": {k: arr(d, k) for k in FIELDS} for i, d in enumerate(subs)}
                              for m, subs in enumerate(meshes)})
            return
        with h5py.File(path, "w") as f:
            for m, subs in enumerate(meshes):
                gm = f.create_group(f"mesh_{m}")
                for i, d in enumerate(subs):
                    g = gm.create_group(f"subdomain_{i}")
                    for k in FIELDS:
                        g.create_dataset(k, data=arr(d, k))
        return
    out = {}
    for m, subs in enumerate(meshes):
        for i, d in enumerate(subs):
            for k in FIELDS:
                out[f"mesh_{m}/subdomain_{i}/{k}"] = arr(d, k)
    np.savez(path, **out)


def load_partitioned(path, mesh_indices=None):
    """-> {mesh index: [dict(FIELDS -> numpy array) per subdomain, in subdomain order]}"""
    meshes = {}
    if _is_h5(path):
        h5py = _h5py()
        if h5py is None:
            from .hdf5_min import Hdf5File
            with Hdf5File(path) as f:
                for name in f.keys():
                    if not name.startswith("mesh_"):
                        continue
                    m = int(name.split("_")[1])
                    if mesh_indices is not None and m not in mesh_indices:
                        continue
                    subs = sorted((s for s in f.keys(name) if s.startswith("subdomain_")), key=lambda s: int(s.split("_")[1]))
                    meshes[m] = []
                    for s_ in subs:
                        have = f.keys(f"{name}/{s_}")
                        missing = [k for k in FIELDS if k not in have]
                        if missing:
                            raise ValueError(f"{name}/{s_}: missing {missing}")
                        meshes[m].append({k: f[f"{name}/{s_}/{k}"] for k in FIELDS})
            return meshes
        with h5py.File(path, "r") as f:
            for name, gm in f.items():
                m = int(name.split("_")[1])
                if mesh_indices is not None and m not in mesh_indices:
                    continue
                subs = sorted(gm.keys(), key=lambda s: int(s.split("_")[1]))
                meshes[m] = [{k: np.asarray(gm[s][k]) for k in FIELDS} for s in subs]
        return meshes
    with np.load(path) as z:
        tmp = {}
        for key in z.files:
            mt = _KEY.match(key)
            if not mt:
                continue
            m, i, k = int(mt.group(1)), int(mt.group(2)), mt.group(3)
            if mesh_indices is not None and m not in mesh_indices:
                continue
            tmp.setdefault(m, {}).setdefault(i, {})[k] = z[key]
    for m, subs in tmp.items():
        order = sorted(subs)
        if order != list(range(len(order))):
            raise ValueError(f"mesh_{m}: subdomain indices are not 0..{len(order) - 1}")
        for i in order:
            missing = [k for k in FIELDS if k not in subs[i]]
            if missing:
                raise ValueError(f"mesh_{m}/subdomain_{i}: missing {missing}")
        meshes[m] = [subs[i] for i in order]
    return meshes


def batch_from_subdomains(subs, device):
    """Stored subdomains of one mesh -> (SubdomainBatch in the canonical (subdomain, dst, src) order, x, y on the device)."""
    sizes = np.array([int(np.asarray(d["x"]).shape[0]) for d in subs], dtype=np.int64)
    node_ptr = np.concatenate([[0], np.cumsum(sizes)])
    x = torch.from_numpy(np.concatenate([np.asarray(d["x"], dtype=np.float32) for d in subs])).to(device)
    y = torch.from_numpy(np.concatenate([np.asarray(d["y"], dtype=np.float32) for d in subs])).to(device)
    gids = torch.from_numpy(np.concatenate([np.asarray(d["global_node_ids"]).astype(np.int64).reshape(-1) for d in subs])).to(device)
    ei = np.concatenate([np.asarray(d["edge_index"]).astype(np.int64).reshape(2, -1) + o for d, o in zip(subs, node_ptr[:-1])], axis=1)
    for d, n in zip(subs, sizes):
        e = np.asarray(d["edge_index"])
        if e.size and (e.min() < 0 or e.max() >= n):
            raise ValueError("edge_index of a stored subdomain points outside its node range")
    ea = torch.from_numpy(np.concatenate([np.asarray(d["edge_attr"], dtype=np.float32).reshape(-1) for d in subs])).to(device)
    n_tot = int(node_ptr[-1])
    csr = ops.csr_build(torch.from_numpy(ei).to(device), n_tot)
    ea_sorted = ea[csr.perm.long()].contiguous()
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(n_tot, device=device, dtype=torch.int32), deg)
    nptr = torch.from_numpy(node_ptr.astype(np.int32)).to(device)
    batch = ops.SubdomainBatch(node_ptr=nptr, edge_ptr=csr.rowptr[nptr.long()].contiguous(), global_ids=gids, edge_src=csr.src,
                               edge_dst=dst, edge_attr=ea_sorted, rowptr=csr.rowptr, n_sub=len(subs), n_tot=n_tot, e_tot=csr.E)
    return batch, x, y


class StoredSubdomainDataset:
    """The reference's partitioned dataset served from a store file (``root`` = path of the .npz / .h5)."""

    def __init__(self, root, device=None, **kwargs):
        self.root = root
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._meshes = load_partitioned(root)
        if not self._meshes:
            raise ValueError(f"{root}: no mesh_*/subdomain_* entries")
        self._order = sorted(self._meshes)
        self.sub_size = max(len(v) for v in self._meshes.values())
        self._cache = {}

    def _mesh(self, idx):
        if idx < 0 or idx >= len(self._order):
            raise IndexError(f"Mesh index {idx} out of range. Maximum index is {len(self._order) - 1}.")
        if idx not in self._cache:
            subs = self._meshes[self._order[idx]]
            batch, x, y = batch_from_subdomains(subs, self.device)
            N = int(batch.global_ids.max().item()) + 1
            datas = [Data(x=torch.from_numpy(np.asarray(d["x"], dtype=np.float32)), y=torch.from_numpy(np.asarray(d["y"], dtype=np.float32)),
                          pos=torch.from_numpy(np.asarray(d["pos"], dtype=np.float32)),
                          edge_index=torch.from_numpy(np.asarray(d["edge_index"]).astype(np.int64)),
                          edge_attr=torch.from_numpy(np.asarray(d["edge_attr"], dtype=np.float32)),
                          global_node_ids=torch.from_numpy(np.asarray(d["global_node_ids"]).astype(np.int64))) for d in subs]
            # positions of the global nodes: any copy (they coincide)
            pos = np.zeros((N, 3), dtype=np.float32)
            for d in subs:
                pos[np.asarray(d["global_node_ids"]).astype(np.int64).reshape(-1)] = np.asarray(d["pos"], dtype=np.float32)
            self._cache[idx] = {"batch": batch, "x": x, "y": y, "datas": datas, "N": N, "pos": pos,
                                "occ": ops.occurrence_build(batch.global_ids, N)}
        return self._cache[idx]

    def len(self):
        return sum(len(v) for v in self._meshes.values())

    __len__ = len

    def get(self, idx):
        idx = int(idx)
        for m in range(len(self._order)):
            n = len(self._meshes[self._order[m]])
            if idx < n:
                return self._mesh(m)["datas"][idx]
            idx -= n
        raise IndexError("index out of range")

    __getitem__ = get

    def get_one_full_sample(self, idx, materialize=True):
        c = self._mesh(idx)
        return SubdomainSample(c["datas"], c["batch"], c["x"], c["y"], idx, c["N"])

    def reconstruct_from_partition(self, subdomain_data_list, subdomain_ref_list, subdomain_idx, model_idx=None,
                                   weights_list=None):
        from .GraphDataset import stitch_lists
        c = self._mesh(subdomain_idx)
        if "gids_cpu" not in c:
            c["gids_cpu"] = c["batch"].global_ids.cpu()
        cells = np.zeros((0, 4), dtype=np.int32)            # the store keeps graphs, not cells
        return stitch_lists(c["batch"], c["occ"], c["pos"], cells, c["gids_cpu"], c, subdomain_data_list,
                            subdomain_ref_list, self.device)
