"""CPU: fesr_b200/dataset/hdf5_min.py -- the HDF5 subset of the reference's store (dataset/GraphDataset.py:1245-1284,
1470-1484), checked against structures laid out BY HAND from the HDF5 File Format Specification (there is no libhdf5 /
h5py in this image to produce a foreign file), and through round trips that exercise multi-level group B-trees,
compact / contiguous / chunked layouts and every supported number type."""
import struct

import numpy as np
import pytest

from fesr_b200.dataset.hdf5_min import SIGNATURE, UNDEF, Hdf5File, Hdf5FormatError, read_hdf5, write_hdf5


def _hand_built_file():
    """One group-less file written byte by byte from the specification: superblock v0, root group (symbol table with
    one leaf B-tree node and one SNOD), one contiguous 2 x 3 float32 dataset named "x" and one compact int64 dataset
    "ids" whose object header continues in a continuation block."""
    buf = bytearray(96)
    def put(data, align=8):
        while len(buf) % align:
            buf.append(0)
        a = len(buf)
        buf.extend(data)
        return a
    def msg(t, body, flags=0):
        body = body + b"\x00" * (-len(body) % 8)
        return struct.pack("<HHB3x", t, len(body), flags) + body
    x = np.arange(6, dtype="<f4").reshape(2, 3) * 0.5
    x_addr = put(x.tobytes())
    f32 = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    space = struct.pack("<BBB5xQQ", 1, 2, 0, 2, 3)
    layout = struct.pack("<BBQQ", 3, 1, x_addr, 24)
    body = msg(1, space) + msg(3, f32, 1) + msg(8, layout)
    x_hdr = put(struct.pack("<BxHII4x", 1, 3, 1, len(body)) + body)
    # "ids": first chunk holds dataspace + a continuation message; the rest lives in the continuation block
    ids = np.array([7, -1, 2 ** 40], dtype="<i8")
    i64 = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)
    cont_body = msg(3, i64, 1) + msg(0, b"\x00" * 8) + msg(8, struct.pack("<BBH", 3, 0, 24) + ids.tobytes())
    cont_addr = put(cont_body)
    first = msg(1, struct.pack("<BBB5xQ", 1, 1, 0, 3)) + msg(0x10, struct.pack("<QQ", cont_addr, len(cont_body)))
    ids_hdr = put(struct.pack("<BxHII4x", 1, 5, 1, len(first)) + first)
    # root group: heap ("" at 0, "ids" at 8, "x" at 16), SNOD with two entries sorted by name, one leaf TREE node
    heap_data = put(b"\x00" * 8 + b"ids\x00" + b"\x00" * 4 + b"x\x00" + b"\x00" * 6)
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, 24, 1, heap_data))
    snod = b"SNOD" + struct.pack("<BxH", 1, 2) + struct.pack("<QQII16x", 8, ids_hdr, 0, 0) + struct.pack("<QQII16x", 16, x_hdr, 0, 0)
    snod += b"\x00" * (8 + 40 * 8 - len(snod))
    snod_addr = put(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, 16)
    tree += b"\x00" * (24 + 8 + 16 * 32 - len(tree))
    tree_addr = put(tree)
    root_body = msg(0x11, struct.pack("<QQ", tree_addr, heap))
    root_hdr = put(struct.pack("<BxHII4x", 1, 1, 1, len(root_body)) + root_body)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF) + struct.pack("<QQIIQQ", 0, root_hdr, 1, 0, tree_addr, heap)
    buf[:96] = sb
    return bytes(buf), x, ids


def test_reads_a_file_laid_out_by_hand_from_the_specification(tmp_path):
    data, x, ids = _hand_built_file()
    p = tmp_path / "hand.h5"
    p.write_bytes(data)
    with Hdf5File(str(p)) as f:
        assert f.keys() == ["ids", "x"] and f.visit() == ["ids", "x"] and not f.is_group("x") and f.is_group("")
        got = f["x"]
        assert got.dtype == np.float32 and got.shape == (2, 3) and np.array_equal(got, x)
        got = f["ids"]                                   # compact layout behind an object-header continuation
        assert got.dtype == np.int64 and np.array_equal(got, ids)
        with pytest.raises(KeyError):
            f["nope"]


def test_writer_emits_the_same_structures_as_the_hand_built_file(tmp_path):
    """The writer's superblock / object headers parse with the reader AND match the hand-built bytes field by field
    where the layout is forced (superblock constants, message encodings)."""
    p = tmp_path / "w.h5"
    x = np.arange(6, dtype=np.float32).reshape(2, 3) * 0.5
    write_hdf5(str(p), {"x": x, "ids": np.array([7, -1, 2 ** 40])}, compact=("ids",))
    raw = p.read_bytes()
    hand, _, _ = _hand_built_file()
    assert raw[:24] == hand[:24]                          # signature, versions, offset / length sizes, K values, flags
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)      # end-of-file address
    assert raw.count(b"SNOD") == 1 and raw.count(b"TREE") == 1 and raw.count(b"HEAP") == 1
    f32 = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    assert f32 in raw                                     # IEEE float32 datatype message body, as the spec lays it out
    got = read_hdf5(str(p))
    assert np.array_equal(got["x"], x) and got["ids"].dtype == np.int64 and got["ids"].tolist() == [7, -1, 2 ** 40]


def test_reference_layout_with_many_subdomains_uses_multi_level_btrees(tmp_path):
    """mesh_{m}/subdomain_{i}/... with 300 subdomains: 38 symbol-table nodes -> a two-level group B-tree."""
    rng = np.random.default_rng(0)
    tree = {}
    for m in range(2):
        tree[f"mesh_{m}"] = {f"subdomain_{i}": {"x": rng.normal(size=(3 + i % 5, 4)).astype(np.float32),
                                               "edge_index": rng.integers(0, 3, size=(2, 6)),
                                               "global_node_ids": np.arange(3 + i % 5) + i} for i in range(300)}
    p = tmp_path / "big.h5"
    write_hdf5(str(p), tree)
    raw = p.read_bytes()
    levels = [raw[i + 5] for i in range(len(raw) - 8) if raw[i:i + 4] == b"TREE" and raw[i + 4] == 0]
    assert max(levels) == 1 and levels.count(0) >= 4
    with Hdf5File(str(p)) as f:
        assert f.keys() == ["mesh_0", "mesh_1"]
        names = f.keys("mesh_1")
        assert len(names) == 300 and names == sorted(names) and "subdomain_299" in names
        for m in (0, 1):
            for i in (0, 7, 8, 9, 150, 299):
                for k, v in tree[f"mesh_{m}"][f"subdomain_{i}"].items():
                    got = f[f"mesh_{m}/subdomain_{i}/{k}"]
                    assert got.dtype == v.dtype and np.array_equal(got, v)
        assert len(f.visit("mesh_0")) == 900


@pytest.mark.parametrize("dtype", ["<f4", "<f8", "<f2", "<i8", "<i4", "<i2", "<u1", "<u8"])
def test_number_types_and_edge_shapes(tmp_path, dtype):
    rng = np.random.default_rng(1)
    a = (rng.normal(size=(5, 3)) * 100).astype(dtype)
    p = tmp_path / "t.h5"
    write_hdf5(str(p), {"g": {"a": a, "empty": np.zeros((0, 4), dtype=dtype), "scalar": np.asarray(a[0, 0]),
                              "one_d": a[:, 0].copy()}})
    with Hdf5File(str(p)) as f:
        for k, want in (("a", a), ("empty", np.zeros((0, 4), dtype=dtype)), ("scalar", np.asarray(a[0, 0])), ("one_d", a[:, 0])):
            got = f["g/" + k]
            assert got.dtype == np.dtype(dtype) and got.shape == want.shape and np.array_equal(got, want)


def test_unfiltered_chunked_dataset(tmp_path):
    """h5py writes chunked storage when maxshape / chunks is given; read it when no filter is applied.  The chunk
    B-tree (node type 1) is appended by hand to a file from the writer."""
    a = np.arange(7 * 5, dtype="<f4").reshape(7, 5)
    p = tmp_path / "c.h5"
    write_hdf5(str(p), {"a": a})
    buf = bytearray(p.read_bytes())
    chunk = (4, 3)
    def put(data):
        while len(buf) % 8:
            buf.append(0)
        o = len(buf)
        buf.extend(data)
        return o
    entries = []
    for r in range(0, 7, 4):
        for c in range(0, 5, 3):
            blk = np.zeros(chunk, dtype="<f4")
            part = a[r:r + 4, c:c + 3]
            blk[:part.shape[0], :part.shape[1]] = part
            entries.append((r, c, put(blk.tobytes())))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
    for r, c, addr in entries:
        node += struct.pack("<IIQQQ", 48, 0, r, c, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 8, 6, 0)
    bt = put(node)
    # replace the contiguous layout message of "a" by a chunked one (same 24-byte body size)
    old = struct.pack("<BB", 3, 1)
    i = buf.find(struct.pack("<HHB3x", 8, 24, 0) + old)
    assert i > 0
    new_body = struct.pack("<BBBQIII", 3, 2, 3, bt, 4, 3, 4)
    buf[i + 8:i + 8 + len(new_body)] = new_body
    struct.pack_into("<Q", buf, 40, len(buf))
    p.write_bytes(bytes(buf))
    with Hdf5File(str(p)) as f:
        assert np.array_equal(f["a"], a)


def test_unsupported_features_raise_by_name(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(Hdf5FormatError, match="signature"):
        Hdf5File(str(p))
    write_hdf5(str(p), {"a": np.zeros(3, dtype=np.float32)})
    raw = bytearray(p.read_bytes())
    raw[8] = 2                                            # superblock version 2 (libver='latest')
    p.write_bytes(bytes(raw))
    with pytest.raises(Hdf5FormatError, match="superblock version 2"):
        Hdf5File(str(p))
    with pytest.raises(Hdf5FormatError, match="cannot store"):
        write_hdf5(str(p), {"s": np.array(["a", "b"])})
