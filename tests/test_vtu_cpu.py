"""CPU: the VTK-free .vtu writer / reader (fesr_b200/dataset/vtu.py; reference output stage run_ALDS_3D.py:33-38)."""
import numpy as np

from fesr_b200.dataset.vtu import read_vtu, write_vtu


def test_appended_raw_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(11, 3)).astype(np.float32)
    cells = rng.integers(0, 11, size=(7, 4))
    pd = {"velocity": rng.normal(size=(11, 3)).astype(np.float32), "pressure": rng.normal(size=11).astype(np.float32),
          "GlobalPointIds": np.arange(11, dtype=np.int64) * 3}
    p = tmp_path / "a.vtu"
    write_vtu(str(p), pts, cells, pd)
    raw = p.read_bytes()
    assert raw.startswith(b'<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid"') and b'header_type="UInt64"' in raw
    assert b'<AppendedData encoding="raw">\n_' in raw and raw.rstrip().endswith(b"</VTKFile>")
    # first appended block: UInt64 byte count, then the velocity array
    start = raw.index(b"_", raw.index(b"<AppendedData")) + 1
    assert int.from_bytes(raw[start:start + 8], "little") == 11 * 3 * 4
    assert np.array_equal(np.frombuffer(raw, np.float32, 33, start + 8).reshape(11, 3), pd["velocity"])
    back = read_vtu(str(p))
    assert np.array_equal(back["points"], pts) and np.array_equal(back["cells"], cells) and bool((back["types"] == 10).all())
    for k, v in pd.items():
        assert back["point_data"][k].dtype == v.dtype and np.array_equal(back["point_data"][k], v)


def test_reads_inline_ascii(tmp_path):
    p = tmp_path / "ascii.vtu"
    p.write_text('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n'
                 '<UnstructuredGrid><Piece NumberOfPoints="4" NumberOfCells="1">\n<PointData Scalars="pressure">\n'
                 '<DataArray type="Float32" Name="pressure" format="ascii">\n1 2 3 4.5\n</DataArray>\n</PointData>\n'
                 '<Points><DataArray type="Float32" NumberOfComponents="3" format="ascii">\n0 0 0 1 0 0 0 1 0 0 0 1\n'
                 '</DataArray></Points>\n<Cells>\n<DataArray type="Int32" Name="connectivity" format="ascii">\n0 1 2 3\n'
                 '</DataArray>\n<DataArray type="Int32" Name="offsets" format="ascii">\n4\n</DataArray>\n'
                 '<DataArray type="UInt8" Name="types" format="ascii">\n10\n</DataArray>\n</Cells>\n</Piece>'
                 '</UnstructuredGrid>\n</VTKFile>\n')
    back = read_vtu(str(p))
    assert back["points"].shape == (4, 3) and back["cells"].tolist() == [[0, 1, 2, 3]]
    assert back["point_data"]["pressure"].tolist() == [1.0, 2.0, 3.0, 4.5]
