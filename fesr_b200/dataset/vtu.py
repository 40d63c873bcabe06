"""VTK XML UnstructuredGrid (.vtu) writer / reader without VTK.

The reference ends `--mode pred` by handing the grid `reconstruct_from_partition` returns -- ALL partitions appended
(sum n_s points, every partition's cells with its own point numbering), point arrays `velocity`, `pressure`,
`ref_velocity`, `ref_pressure` averaged over coincident points -- to vtkXMLUnstructuredGridWriter
(run_ALDS_3D.py:33-38, dataset/GraphDataset.py:1324-1409).  `write_vtu` emits that grid as one Piece in the XML
format's "appended raw" mode (binary blocks behind one UInt64 byte count each: what ParaView / VTK read fastest, and
a 5 M-cell grid stays a few hundred MB instead of gigabytes of ASCII).  `read_vtu` parses such a file back (the
round-trip check of the tests, and a VTK-free way for downstream scripts such as compute_wss.py to load predictions).
"""
from __future__ import annotations

import re
import struct

import numpy as np

_VTK_TYPES = {"float32": "Float32", "float64": "Float64", "int64": "Int64", "int32": "Int32", "uint8": "UInt8"}
_NP_TYPES = {v: k for k, v in _VTK_TYPES.items()}
VTK_TETRA = 10


def write_vtu(path, points, cells, point_data: dict, cell_type: int = VTK_TETRA):
    """points [n, 3] float32; cells [c, k] integer (k points per cell, one cell type); point_data name -> [n] / [n, m]."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    n, c = points.shape[0], cells.shape[0]
    k = cells.shape[1] if cells.ndim == 2 and c else 4
    blocks, header = [], []
    offset = 0

    def add(name, arr, ncomp=None, extra=""):
        nonlocal offset
        arr = np.ascontiguousarray(arr)
        tname = _VTK_TYPES[arr.dtype.name]
        comp = f' NumberOfComponents="{ncomp}"' if ncomp else ""
        nm = f' Name="{name}"' if name else ""
        header.append(f'<DataArray type="{tname}"{nm}{comp} format="appended" offset="{offset}"{extra}/>')
        raw = arr.tobytes()
        blocks.append(struct.pack("<Q", len(raw)) + raw)
        offset += 8 + len(raw)

    out = ['<?xml version="1.0"?>',
           '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
           '<UnstructuredGrid>', f'<Piece NumberOfPoints="{n}" NumberOfCells="{c}">']
    vectors = [nm for nm, a in point_data.items() if np.ndim(a) == 2 and np.shape(a)[1] == 3]
    scalars = [nm for nm, a in point_data.items() if np.ndim(a) == 1]
    attr = (f' Vectors="{vectors[0]}"' if vectors else "") + (f' Scalars="{scalars[0]}"' if scalars else "")
    out.append(f"<PointData{attr}>")
    for nm, a in point_data.items():
        a = np.asarray(a)
        if a.shape[0] != n:
            raise ValueError(f"point array {nm!r} has {a.shape[0]} rows for {n} points")
        add(nm, a, a.shape[1] if a.ndim == 2 else None)
        out.append(header.pop())
    out.append("</PointData>")
    out.append("<Points>")
    add("", points, 3)
    out.append(header.pop())
    out.append("</Points>")
    out.append("<Cells>")
    add("connectivity", cells.reshape(-1))
    out.append(header.pop())
    add("offsets", np.arange(1, c + 1, dtype=np.int64) * k)
    out.append(header.pop())
    add("types", np.full(c, cell_type, dtype=np.uint8))
    out.append(header.pop())
    out += ["</Cells>", "</Piece>", "</UnstructuredGrid>", '<AppendedData encoding="raw">']
    with open(path, "wb") as fh:
        fh.write(("\n".join(out) + "\n_").encode("ascii"))
        for b in blocks:
            fh.write(b)
        fh.write(b"\n</AppendedData>\n</VTKFile>\n")


_ARRAY = re.compile(rb'<DataArray\s+([^>]*?)/?>')
_ATTR = re.compile(rb'(\w+)="([^"]*)"')


def read_vtu(path):
    """-> dict(points [n, 3], cells [c, k] (uniform cell size), types [c], point_data {name: array}).  Handles the
    appended-raw files `write_vtu` makes and inline-ASCII files (the previous writer / hand-written fixtures)."""
    data = open(path, "rb").read()
    cut = data.find(b"<AppendedData")
    head = data if cut < 0 else data[:cut]
    appended = None
    if cut >= 0:
        if b'encoding="raw"' not in data[cut:cut + 80]:
            raise ValueError("only raw appended data is supported")
        appended = data.index(b"_", cut) + 1
    m = re.search(rb'<Piece\s+NumberOfPoints="(\d+)"\s+NumberOfCells="(\d+)"', head)
    if not m:
        raise ValueError(f"{path}: no <Piece> with point / cell counts")
    n, c = int(m.group(1)), int(m.group(2))
    wide = b'header_type="UInt64"' in head[:400]

    def section(tag):
        a, b = head.find(b"<" + tag), head.find(b"</" + tag + b">")
        return head[a:b] if a >= 0 and b >= 0 else b""

    def arrays(sec):
        out = []
        for mm in _ARRAY.finditer(sec):
            at = {k.decode(): v.decode() for k, v in _ATTR.findall(mm.group(1))}
            dt = np.dtype(_NP_TYPES[at["type"]])
            if at.get("format") == "appended":
                p = appended + int(at["offset"])
                nbytes = struct.unpack_from("<Q" if wide else "<I", data, p)[0]
                arr = np.frombuffer(data, dtype=dt, count=nbytes // dt.itemsize, offset=p + (8 if wide else 4)).copy()
            else:
                end = sec.find(b"</DataArray>", mm.end())
                arr = np.array(sec[mm.end():end].split(), dtype=dt)
            nc = int(at.get("NumberOfComponents", "1"))
            out.append((at.get("Name", ""), arr.reshape(-1, nc) if nc > 1 else arr))
        return out

    pd = dict(arrays(section(b"PointData")))
    pts = arrays(section(b"Points"))[0][1].reshape(n, 3)
    cl = dict(arrays(section(b"Cells")))
    offs = cl["offsets"].astype(np.int64)
    k = int(offs[0]) if c else 0
    if c and not np.array_equal(offs, np.arange(1, c + 1) * k):
        raise ValueError("mixed cell sizes")
    return {"points": pts, "cells": cl["connectivity"].astype(np.int64).reshape(c, k) if c else np.zeros((0, 4), np.int64),
            "types": cl["types"], "point_data": pd}
