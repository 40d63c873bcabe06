"""Dataset side of the hot path: subdomain assembly and overlap stitching on the GPU.

Mirrors the interface of the reference's ``AnsysDataset`` (dataset/GraphDataset.py:751-1484) for
the methods on the path -- ``len``/``get`` (:760-797), ``get_one_full_sample`` (:1464-1484),
``reconstruct_from_partition`` (:1308-1409) -- on synthetic duct meshes, because the reference's
ANSYS data is not shipped (README.md:26) and its ingest needs VTK.  The decomposition that the
reference runs once through vtkRedistributeDataSetFilter and stores in HDF5 (:1183-1306) is a
GPU pass here (fesr_partition_cells / fesr_build_subdomains), kept resident on the device.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from .. import _lib, ops
from ..data import Data
from . import synthetic


class SubdomainSample(list):
    """``list[Data]`` (CPU tensors, what the reference's loader returns) that also carries the
    device-resident block-diagonal batch, so predict/stitch never go through per-subdomain copies."""

    def __init__(self, datas, batch, x_dev, y_dev, mesh_idx, num_nodes):
        super().__init__(datas)
        self.batch = batch
        self.x_dev = x_dev
        self.y_dev = y_dev
        self.x_host = None        # set by with_host_inputs(): predict() then copies x / y host -> device
        self.y_host = None
        self.mesh_idx = mesh_idx
        self.num_nodes = num_nodes

    def with_host_inputs(self, x_host, y_host):
        """Same subdomains, but the per-subdomain input / reference fields come from (pinned) host
        memory on every predict() call -- a new time step on a mesh whose decomposition is resident."""
        out = SubdomainSample(list(self), self.batch, None, None, self.mesh_idx, self.num_nodes)
        out.x_host, out.y_host = x_host, y_host
        return out


class StitchedMesh:
    """Result of reconstruct_from_partition: the appended partitions with averaged point data
    (what the reference returns as a vtkUnstructuredGrid) plus the field on the original mesh.

    Every array is copied to the host on first access.  After a sharded (multi-rank) predict the stitched field is
    DISTRIBUTED: this rank has stitched -- and `field_local` / `ref_field_local` return -- only its own contiguous
    slice `node_range` of the mesh nodes; the whole-mesh arrays (`field`, `ref_field`, `merged`, ...) are computed
    from the gathered predictions when somebody asks for them (`lazy`: name -> callable returning the device array)."""

    def __init__(self, pos, cells, dev_arrays, global_ids, lazy=None, node_range=None, appended_cells=None):
        self.pos, self.cells = pos, cells            # original mesh (numpy)
        self._appended_cells = appended_cells        # callable -> [sum cells, 4] int64 in appended numbering, or None
        self._dev = dict(dev_arrays)                 # name -> device tensor; copied to the host on first access
        self._lazy = dict(lazy or {})
        self._host = {}
        self.global_ids = global_ids
        self.node_range = node_range if node_range is not None else (0, int(np.asarray(pos).shape[0]))

    def _get(self, name):
        if name not in self._host:
            if name not in self._dev:
                if name.endswith("_local") and name[:-6] in self._dev:          # one rank: the slice is everything
                    self._dev[name] = self._dev[name[:-6]]
                else:
                    self._dev[name] = self._lazy[name]()
            t = self._dev[name]
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            self._host[name] = host
        return self._host[name]

    field_local = property(lambda self: self._get("field_local"))          # [node_range, 4] this rank's slice
    ref_field_local = property(lambda self: self._get("ref_field_local"))
    field = property(lambda self: self._get("field"))            # [N, 4] prediction on the original mesh
    ref_field = property(lambda self: self._get("ref_field"))
    merged = property(lambda self: self._get("merged"))          # [sum n_s, 4] appended partitions, averaged
    merged_ref = property(lambda self: self._get("merged_ref"))
    count = property(lambda self: self._get("count"))

    @property
    def point_data(self):
        m, r = self.merged, self.merged_ref
        return {"velocity": m[:, :3], "pressure": m[:, 3], "ref_velocity": r[:, :3], "ref_pressure": r[:, 3]}

    def GetNumberOfPoints(self):
        return int(self.global_ids.shape[0])

    def write_vtu(self, path, appended: bool | None = None):
        """Writes the result as VTK XML UnstructuredGrid (run_ALDS_3D.py:33-38).
        appended=True: the grid the reference writes -- every partition appended (sum n_s points at their copies'
        positions, each partition's cells in its own point numbering), point arrays `velocity`, `pressure`,
        `ref_velocity`, `ref_pressure` = the averages written back to every copy, plus `GlobalPointIds`.
        appended=False: the ORIGINAL mesh (N points) with the stitched fields -- smaller, same information.
        Default: the reference's appended grid when this dataset knows the partitions' cells, else the original mesh."""
        from .vtu import write_vtu
        if appended is None:
            appended = self._appended_cells is not None
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if appended:
            if self._appended_cells is None:
                raise ValueError("this dataset does not keep the partitions' cells: write_vtu(path, appended=False)")
            m, r = self.merged.numpy(), self.merged_ref.numpy()
            gids = self.global_ids.numpy() if torch.is_tensor(self.global_ids) else np.asarray(self.global_ids)
            write_vtu(path, np.asarray(self.pos)[gids], self._appended_cells(),
                      {"velocity": m[:, :3], "pressure": m[:, 3], "ref_velocity": r[:, :3], "ref_pressure": r[:, 3],
                       "GlobalPointIds": gids.astype(np.int64)})
            return
        f, r = self.field.numpy(), self.ref_field.numpy()
        write_vtu(path, np.asarray(self.pos), np.asarray(self.cells),
                  {"velocity": f[:, :3], "pressure": f[:, 3], "ref_velocity": r[:, :3], "ref_pressure": r[:, 3]})


def appended_cells_of(part, batch, cells, N):
    """Cells of every partition in the appended grid's point numbering (vtkAppendDataSets over the partitions,
    dataset/GraphDataset.py:1324-1367): cell j of subdomain s keeps its four corners, renumbered to the rows of
    subdomain s's copies in the batch.  One searchsorted over the (subdomain, global id) keys, on the device."""
    S = batch.n_sub
    leaf_ptr = part.leaf_ptr.long()
    sub_of_cell = torch.repeat_interleave(torch.arange(S, device=leaf_ptr.device), leaf_ptr[1:] - leaf_ptr[:-1])
    corners = cells.long()[part.leaf_cells.long()]                                     # [pairs, 4] global ids
    node_ptr = batch.node_ptr.long()
    sub_of_node = torch.repeat_interleave(torch.arange(S, device=leaf_ptr.device), node_ptr[1:] - node_ptr[:-1])
    node_key = sub_of_node * N + batch.global_ids                                      # ascending
    want = (sub_of_cell.unsqueeze(1) * N + corners).reshape(-1)
    pos = torch.searchsorted(node_key, want)
    assert bool((node_key[pos] == want).all())
    return pos.reshape(-1, 4).cpu().numpy()


class SyntheticDuctDataset:
    """``num_meshes`` synthetic ducts, each decomposed into 2^levels overlapping subdomains on the GPU.

    kwargs follow the reference's flat exp_config (configs/exp_config/*.yaml): ``partition``,
    ``sub_size`` (requested number of subdomains; rounded up to a power of two exactly as
    vtkRedistributeDataSetFilter does), plus ``mesh_n`` (duct cross-section in hexes) and
    ``num_meshes``; unknown keys are ignored, as in the reference's constructors.
    """

    def __init__(self, root=None, transform=None, pre_transform=None, partition=True, sub_size=None, mesh_n=13,
                 num_meshes=4, boundary_mode="all", device=None, length_factor=1, **kwargs):
        self.root = root
        self.partition = partition
        self.mesh_n = int(mesh_n)
        self.num_meshes = int(num_meshes)
        self.length_factor = int(length_factor)
        self.mode = _lib.ALL_INTERSECTING if boundary_mode == "all" else _lib.ONE_REGION
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        probe = synthetic.make_positions(self.mesh_n).shape[0]
        if sub_size:
            self.levels = max(0, math.ceil(math.log2(int(sub_size))))
        else:
            self.levels = synthetic.default_kd_levels(probe)
        self.sub_size = 1 << self.levels
        self._cache = {}

    # -- assembly ---------------------------------------------------------------------------
    def _mesh(self, idx):
        if idx >= self.num_meshes or idx < 0:
            raise IndexError(f"Mesh index {idx} out of range. Maximum index is {self.num_meshes - 1}.")
        if idx not in self._cache:
            mesh = (synthetic.make_duct_mesh(self.mesh_n, seed=idx) if self.length_factor == 1 else
                    synthetic.make_duct_mesh_long(self.mesh_n, self.length_factor, seed=idx))
            dev = self.device
            pos = torch.from_numpy(mesh.pos).to(dev)
            cells = torch.from_numpy(mesh.cells).to(dev)
            part, batch = ops.assemble(pos, cells, self.levels, self.mode)
            x_dev = torch.from_numpy(mesh.x).to(dev).index_select(0, batch.global_ids)
            y_dev = torch.from_numpy(mesh.y).to(dev).index_select(0, batch.global_ids)
            self._cache[idx] = {"mesh": mesh, "pos": pos, "part": part, "batch": batch, "x": x_dev, "y": y_dev,
                                "occ": ops.occurrence_build(batch.global_ids, mesh.num_nodes)}
        return self._cache[idx]

    def len(self):
        return self.num_meshes * self.sub_size

    __len__ = len

    def _datas(self, c):
        b = c["batch"]
        node_ptr = b.node_ptr.cpu().numpy()
        edge_ptr = b.edge_ptr.cpu().numpy()
        x, y = c["x"].cpu(), c["y"].cpu()
        gids = b.global_ids.cpu()
        pos = torch.from_numpy(c["mesh"].pos)[gids]
        src, dst = b.edge_src.cpu().long(), b.edge_dst.cpu().long()
        ea = b.edge_attr.cpu()
        out = []
        for s in range(b.n_sub):
            nl, nh, el, eh = node_ptr[s], node_ptr[s + 1], edge_ptr[s], edge_ptr[s + 1]
            out.append(Data(x=x[nl:nh], y=y[nl:nh], pos=pos[nl:nh],
                            edge_index=torch.stack([src[el:eh] - nl, dst[el:eh] - nl]),
                            edge_attr=ea[el:eh].unsqueeze(1), global_node_ids=gids[nl:nh]))
        return out

    def get(self, idx):
        mesh_idx, sub_idx = divmod(int(idx), self.sub_size)
        c = self._mesh(mesh_idx)
        if not c.get("datas"):              # also after get_one_full_sample(materialize=False) left it empty
            c["datas"] = self._datas(c)
        return c["datas"][sub_idx]

    __getitem__ = get

    def get_one_full_sample(self, idx, materialize=True):
        """All subdomains of mesh ``idx`` (reference :1464-1484), with the device batch attached.
        materialize=False skips building the per-subdomain CPU ``Data`` list (large meshes)."""
        c = self._mesh(idx)
        if not c.get("datas"):
            c["datas"] = self._datas(c) if materialize else []
        return SubdomainSample(c["datas"], c["batch"], c["x"], c["y"], idx, c["mesh"].num_nodes)

    # -- low-res -> high-res transfer ---------------------------------------------------------
    @staticmethod
    def _lagrangian_interpolation(mesh, physics, new_mesh, mesh_spacing=0.012, sharpness=2.0):
        """Field of `mesh` interpolated at the points of `new_mesh` (reference :1041-1105: vtkPointInterpolator with
        a Gaussian kernel of radius 3 * mesh_spacing and sharpness 2; mesh_spacing is hard-coded to 0.012 there and a
        keyword here).  `mesh` / `new_mesh`: anything with a `.pos` [n, 3] array (or the array itself); `physics`:
        [n] or [n, 1] values at the points of `mesh`.  -> float tensor [n_new, 1] on the current CUDA device."""
        dev = torch.device("cuda", torch.cuda.current_device())

        def pts(m):
            return torch.as_tensor(np.asarray(getattr(m, "pos", m)), dtype=torch.float32).to(dev)

        src, dst = pts(mesh), pts(new_mesh)
        vals = torch.as_tensor(np.asarray(physics), dtype=torch.float32).reshape(-1).to(dev)
        if vals.shape[0] != src.shape[0]:
            raise ValueError("Mismatch: physics array length must match the number of points in the original mesh.")
        if dst.shape[0] == 0:
            raise ValueError("New mesh has no points to interpolate.")
        return ops.interp_gaussian(src, vals, dst, 3.0 * float(mesh_spacing), sharpness).reshape(-1, 1)

    # -- stitch -----------------------------------------------------------------------------
    def reconstruct_from_partition(self, subdomain_data_list, subdomain_ref_list, subdomain_idx, model_idx=None,
                                   weights_list=None):
        """Mean over all subdomain copies of every mesh node (reference :1308-1409).  Accepts the
        5-argument call of run_ALDS_3D.py:26 (model_idx / weights are carried but, as in the
        reference, not used by the averaging)."""
        c = self._mesh(subdomain_idx)
        if "gids_cpu" not in c:
            c["gids_cpu"] = c["batch"].global_ids.cpu()
        def appended(c=c):
            if "appended_cells" not in c:
                c["appended_cells"] = appended_cells_of(c["part"], c["batch"], torch.from_numpy(c["mesh"].cells).to(self.device),
                                                        c["mesh"].num_nodes)
            return c["appended_cells"]

        return stitch_lists(c["batch"], c["occ"], c["mesh"].pos, c["mesh"].cells, c["gids_cpu"], c,
                            subdomain_data_list, subdomain_ref_list, self.device, appended_cells=appended)


def stitch_lists(b, occ, pos, cells, gids_cpu, cache, pred_list, ref_list, dev, appended_cells=None):
    """The averaging of reconstruct_from_partition over predict()'s return lists (shared by the dataset classes).
    Lists that came out of a sharded predict carry the padded all-gather buffer (`padded`): the stitch reads it in
    place through an occurrence index remapped once per layout, and only this rank's slice of the mesh nodes is
    stitched eagerly (StitchedMesh.field_local); everything else is computed on demand."""

    def to_dev(lst):
        dev_t = getattr(lst, "dev", None)
        if dev_t is not None:
            return dev_t
        t = torch.cat([torch.as_tensor(v, dtype=torch.float32) for v in lst], dim=0)
        if t.shape[0] != b.n_tot:
            raise ValueError(f"expected {b.n_tot} rows over all subdomains, got {t.shape[0]}")
        return t.to(dev)

    lay = getattr(pred_list, "layout", None)
    if lay is not None and getattr(pred_list, "padded", None) is not None and lay.world > 1:
        from .. import comm
        from ..pipeline import node_slice
        rank = getattr(pred_list, "rank", None)
        if rank is None:
            rank = comm.rank() if comm.ready() else 0
        rng = node_slice(occ.N, rank, lay.world)
        rows_view = pred_list.padded[0]
        key = ("occ_padded", lay.world, lay.c, lay.slot, lay.with_ref, tuple(lay.rows))
        if key not in cache:
            cache[key] = (ops.Occurrence(occ.occ_ptr, lay.row_positions(occ.occ_idx), occ.N, int(rows_view.shape[0])),
                          ops.Occurrence(occ.occ_ptr, lay.row_positions(occ.occ_idx, ref=True), occ.N,
                                         int(rows_view.shape[0])) if lay.with_ref else None)
        occ_p, occ_r = cache[key]
        ref_padded = getattr(ref_list, "padded", None) is not None and occ_r is not None

        def stitched(which, full):
            if which == "pred":
                vals, oc_ = rows_view, occ_p
            elif ref_padded:
                vals, oc_ = ref_list.padded[0], occ_r
            else:
                vals, oc_ = to_dev(ref_list), occ
            return ops.stitch_mean(vals, oc_, None, want_merged=False, want_count=True, node_range=None if full else rng)

        f_loc, cnt_loc, _ = stitched("pred", False)
        r_loc, _, _ = stitched("ref", False)
        full = {}

        def whole(which):
            if which not in full:
                full[which] = stitched(which, True)
            return full[which]

        def merged_of(which):
            return whole(which)[0].index_select(0, b.global_ids)

        lazy = {"field": lambda: whole("pred")[0], "ref_field": lambda: whole("ref")[0], "count": lambda: whole("pred")[1],
                "merged": lambda: merged_of("pred"), "merged_ref": lambda: merged_of("ref")}
        return StitchedMesh(pos, cells, {"field_local": f_loc, "ref_field_local": r_loc, "count_local": cnt_loc},
                            gids_cpu, lazy=lazy, node_range=rng, appended_cells=appended_cells)

    pred, ref = to_dev(pred_list), to_dev(ref_list)
    field, count, merged = ops.stitch_mean(pred, occ, b.global_ids, want_merged=True)
    rfield, _, rmerged = ops.stitch_mean(ref, occ, b.global_ids, want_merged=True)
    return StitchedMesh(pos, cells, {"field": field, "ref_field": rfield, "merged": merged, "merged_ref": rmerged,
                                     "count": count}, gids_cpu, appended_cells=appended_cells)


class AnsysDataset(SyntheticDuctDataset):
    """The reference's AnsysDataset reads proprietary Fluent data through VTK and HDF5; neither
    the data nor those libraries exist here, so this name resolves to the synthetic duct with the
    Ansys boundary mode (AssignToAllIntersectingRegions, reference :1219)."""

    def __init__(self, root=None, **kwargs):
        kwargs.setdefault("boundary_mode", "all")
        super().__init__(root, **kwargs)


class DuctAnalysisDataset(SyntheticDuctDataset):
    """Synthetic duct with the Duct boundary mode (AssignToOneRegion, reference :565)."""

    def __init__(self, root=None, **kwargs):
        kwargs.setdefault("boundary_mode", "one")
        super().__init__(root, **kwargs)
