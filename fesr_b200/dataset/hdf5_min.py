"""A dependency-free reader (and matching writer) for the HDF5 subset the reference's subdomain store uses.

The reference writes `partition/data.h5` with h5py defaults (dataset/GraphDataset.py:1245-1284 `create_group` /
`create_dataset(name, data=array)`, :1128-1133 `copy` into `mesh_{m}` groups) and reads it back with
`f[f'mesh_{m}/subdomain_{i}'][name][:]` (:1470-1484).  h5py is not installed here and cannot be, so the `.h5` branch of
`store.py` goes through this module.  With h5py's default `libver='earliest'` such a file is

    superblock version 0  ->  root group symbol-table entry
    groups                :  version-1 object header with a Symbol Table message -> B-tree v1 (node type 0, any depth)
                             -> symbol-table nodes ("SNOD") -> link names in the group's local heap ("HEAP")
    datasets              :  version-1 object header (with continuation blocks) carrying Dataspace (v1 / v2),
                             Datatype (fixed / floating point, v1..v3), Data Layout v3 -- contiguous (what
                             `create_dataset(data=...)` makes), compact, or chunked WITHOUT filters (B-tree v1, type 1)

which is exactly what is implemented, from the HDF5 File Format Specification (version 1.1 / 2.0 structures named
above).  Anything else -- superblock >= 2, version-2 ("OHDR") object headers, link-info groups, filtered chunks,
compound / string / variable-length types -- raises `Hdf5FormatError` naming the feature, never a wrong array.

`write_hdf5` emits the same structures (multi-level group B-trees, sorted names, heap offsets as keys) so that stores
written here open in h5py / libhdf5, and is what builds the test fixtures: there is no libhdf5 in this image to
cross-check against, so the evidence for interoperability is the specification, not a round trip through h5py --
stated as such in DESIGN.md.
"""
from __future__ import annotations

import mmap
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4          # symbol-table node holds up to 2 K = 8 entries
INTERNAL_K = 16     # group B-tree node holds up to 2 K = 32 children


class Hdf5FormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------------------ reader
class Hdf5File:
    """Read-only view of an HDF5 file of the subset above.  `f["mesh_0/subdomain_3/x"]` -> numpy array (a copy);
    `f.keys("mesh_0")` -> sorted child names; `f.is_group(path)`; `f.visit()` -> every dataset path."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        try:
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError as e:
            self._fh.close()
            raise Hdf5FormatError(f"{path}: empty file") from e
        self._groups = {}
        b = self._buf
        if b[:8] != SIGNATURE:
            raise Hdf5FormatError(f"{path}: not an HDF5 file (no signature at offset 0; user blocks are not supported)")
        ver = b[8]
        if ver > 1:
            raise Hdf5FormatError(f"{path}: superblock version {ver} (libver='latest' file); this reader handles the "
                                  "version 0 / 1 superblock h5py writes by default")
        self._so, self._sl = b[13], b[14]
        if (self._so, self._sl) != (8, 8):
            raise Hdf5FormatError(f"offset / length sizes {self._so} / {self._sl}: only 8 / 8 is supported")
        pos = 24 + (4 if ver == 1 else 0)           # version 1 adds indexed-storage K + reserved
        self._base = self._u64(pos)
        if self._base != 0:
            raise Hdf5FormatError("non-zero base address")
        root_entry = pos + 32
        self._root = self._entry(root_entry)

    def close(self):
        self._buf.close()
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- primitives
    def _u16(self, p):
        return struct.unpack_from("<H", self._buf, p)[0]

    def _u32(self, p):
        return struct.unpack_from("<I", self._buf, p)[0]

    def _u64(self, p):
        return struct.unpack_from("<Q", self._buf, p)[0]

    def _entry(self, p):
        """Symbol-table entry -> (name offset, object header address, cache type, btree, heap)."""
        name_off, ohdr, cache = self._u64(p), self._u64(p + 8), self._u32(p + 16)
        btree = heap = None
        if cache == 1:
            btree, heap = self._u64(p + 24), self._u64(p + 32)
        return name_off, ohdr, cache, btree, heap

    def _messages(self, addr):
        """Messages of a version-1 object header (following continuation blocks): [(type, flags, offset, size)]."""
        b = self._buf
        if b[addr:addr + 4] == b"OHDR":
            raise Hdf5FormatError("version-2 object header (libver='latest'); only version-1 headers are supported")
        if b[addr] != 1:
            raise Hdf5FormatError(f"object header version {b[addr]} at {addr}")
        nmsg, size = self._u16(addr + 2), self._u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u16(p), self._u16(p + 2), b[p + 4]
                body = p + 8
                out.append((mtype, flags, body, msize))
                if mtype == 0x0010:                                  # continuation
                    blocks.append((self._u64(body), self._u64(body + 8)))
                p = body + msize
        return out

    def _group_tables(self, entry):
        _, ohdr, cache, btree, heap = entry
        if cache == 1:
            return btree, heap
        for mtype, _, body, _ in self._messages(ohdr):
            if mtype == 0x0011:
                return self._u64(body), self._u64(body + 8)
            if mtype in (0x0002, 0x0006):
                raise Hdf5FormatError("link-info / link-message group (libver='latest'); only symbol-table groups are supported")
        return None

    def _children(self, entry):
        """name -> symbol-table entry of every link of a group."""
        key = entry[1]
        if key in self._groups:
            return self._groups[key]
        tables = self._group_tables(entry)
        if tables is None:
            raise Hdf5FormatError("not a group")
        btree, heap = tables
        b = self._buf
        if b[heap:heap + 4] != b"HEAP":
            raise Hdf5FormatError("local heap signature missing")
        heap_data = self._u64(heap + 24)
        out = {}

        def walk(node):
            if b[node:node + 4] == b"SNOD":
                n = self._u16(node + 6)
                for i in range(n):
                    e = self._entry(node + 8 + 40 * i)
                    s = heap_data + e[0]
                    name = b[s:b.find(b"\x00", s)].decode("utf-8")
                    out[name] = e
                return
            if b[node:node + 4] != b"TREE":
                raise Hdf5FormatError(f"B-tree node signature missing at {node}")
            if b[node + 4] != 0:
                raise Hdf5FormatError("group B-tree of the wrong node type")
            used = self._u16(node + 6)
            p = node + 24
            for i in range(used):
                walk(self._u64(p + 8 + 16 * i))          # key_i (8) child_i (8) ... key_used

        walk(btree)
        self._groups[key] = out
        return out

    def _resolve(self, path):
        entry = self._root
        for part in [p for p in path.split("/") if p]:
            kids = self._children(entry)
            if part not in kids:
                raise KeyError(path)
            entry = kids[part]
        return entry

    def is_group(self, path=""):
        try:
            return self._group_tables(self._resolve(path)) is not None
        except Hdf5FormatError:
            return False

    def keys(self, path=""):
        return sorted(self._children(self._resolve(path)))

    def visit(self, path=""):
        """Every dataset path below `path`, depth first, names sorted."""
        out = []
        for k in self.keys(path):
            p = f"{path}/{k}" if path else k
            if self.is_group(p):
                out.extend(self.visit(p))
            else:
                out.append(p)
        return out

    # -- datasets
    def _dtype(self, body):
        b = self._buf
        cls, ver = b[body] & 0x0F, b[body] >> 4
        bits0 = b[body + 1]
        size = self._u32(body + 4)
        if ver not in (1, 2, 3):
            raise Hdf5FormatError(f"datatype message version {ver}")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            signed = bool(bits0 & 0x08)
            prec = self._u16(body + 10)
            if prec != 8 * size or size not in (1, 2, 4, 8):
                raise Hdf5FormatError(f"fixed-point type with {prec} bits in {size} bytes")
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:
            prec, esize, msize = self._u16(body + 10), b[body + 13], b[body + 15]
            if (size, prec, esize, msize) == (4, 32, 8, 23):
                return np.dtype(order + "f4")
            if (size, prec, esize, msize) == (8, 64, 11, 52):
                return np.dtype(order + "f8")
            if (size, prec, esize, msize) == (2, 16, 5, 10):
                return np.dtype(order + "f2")
            raise Hdf5FormatError(f"floating-point type size {size} exponent {esize} mantissa {msize}")
        raise Hdf5FormatError(f"datatype class {cls} (only fixed- and floating-point numbers are supported)")

    def __getitem__(self, path):
        entry = self._resolve(path)
        b = self._buf
        shape = dtype = layout = None
        for mtype, _, body, msize in self._messages(entry[1]):
            if mtype == 0x0001:
                ver, rank = b[body], b[body + 1]
                if ver == 1:
                    p = body + 8
                elif ver == 2:
                    if b[body + 3] == 2:
                        raise Hdf5FormatError("null dataspace")
                    p = body + 4
                else:
                    raise Hdf5FormatError(f"dataspace message version {ver}")
                shape = tuple(self._u64(p + 8 * i) for i in range(rank))
            elif mtype == 0x0003:
                dtype = self._dtype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                raise Hdf5FormatError("filtered (compressed) dataset: filter pipelines are not supported")
            elif mtype == 0x0011:
                raise KeyError(f"{path} is a group")
        if shape is None or dtype is None or layout is None:
            raise Hdf5FormatError(f"{path}: dataset header without dataspace / datatype / layout")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * dtype.itemsize
        ver, cls = b[layout], b[layout + 1]
        if ver != 3:
            raise Hdf5FormatError(f"data layout message version {ver}")
        if cls == 1:                                                   # contiguous
            addr, size = self._u64(layout + 2), self._u64(layout + 10)
            if addr == UNDEF:
                if count:
                    return np.zeros(shape, dtype=dtype.newbyteorder("="))     # never written: fill value 0
                return np.empty(shape, dtype=dtype.newbyteorder("="))
            if size < nbytes:
                raise Hdf5FormatError(f"{path}: storage of {size} bytes for {nbytes} bytes of data")
            arr = np.frombuffer(b, dtype=dtype, count=count, offset=addr)
        elif cls == 0:                                                 # compact
            size = self._u16(layout + 2)
            if size < nbytes:
                raise Hdf5FormatError(f"{path}: compact storage of {size} bytes for {nbytes} bytes of data")
            arr = np.frombuffer(b, dtype=dtype, count=count, offset=layout + 4)
        elif cls == 2:
            return self._read_chunked(path, layout, shape, dtype)
        else:
            raise Hdf5FormatError(f"layout class {cls}")
        return arr.reshape(shape).astype(dtype.newbyteorder("="), copy=True)

    def _read_chunked(self, path, layout, shape, dtype):
        b = self._buf
        nd = b[layout + 2]                       # rank + 1
        btree = self._u64(layout + 3)
        cdims = tuple(self._u32(layout + 11 + 4 * i) for i in range(nd))
        if nd - 1 != len(shape) or cdims[-1] != dtype.itemsize:
            raise Hdf5FormatError(f"{path}: chunk dimensionality does not match the dataspace")
        out = np.zeros(shape, dtype=dtype.newbyteorder("="))
        if btree == UNDEF:
            return out
        chunk = cdims[:-1]
        key_size = 8 + 8 * nd

        def walk(node):
            if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
                raise Hdf5FormatError("chunk B-tree node malformed")
            level, used = b[node + 5], self._u16(node + 6)
            p = node + 24
            for i in range(used):
                k = p + i * (key_size + 8)
                child = self._u64(k + key_size)
                if level > 0:
                    walk(child)
                    continue
                csize, mask = self._u32(k), self._u32(k + 4)
                if mask != 0 or csize != int(np.prod(chunk)) * dtype.itemsize:
                    raise Hdf5FormatError(f"{path}: filtered chunk")
                off = tuple(self._u64(k + 8 + 8 * j) for j in range(nd - 1))
                data = np.frombuffer(b, dtype=dtype, count=int(np.prod(chunk)), offset=child).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, chunk, shape))
                out[sl] = data[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(btree)
        return out


def read_hdf5(path, prefix=""):
    """{dataset path: array} of every dataset below `prefix`."""
    with Hdf5File(path) as f:
        return {p: f[p] for p in f.visit(prefix)}


# ------------------------------------------------------------------------------------------------------------ writer
def _pad8(n):
    return (n + 7) & ~7


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                 # superblock, filled at the end

    def alloc(self, data: bytes, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def message(mtype, body: bytes, flags=0):
        body = body + b"\x00" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, messages):
        body = b"".join(messages)
        hdr = struct.pack("<BxHII4x", 1, len(messages), 1, len(body))
        return self.alloc(hdr + body)

    # -- dataset
    def dataset(self, arr: np.ndarray, compact=False):
        shape = arr.shape                                              # (ascontiguousarray turns a 0-d array into 1-d)
        arr = np.ascontiguousarray(arr).reshape(shape)
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        arr = arr.astype(dt, copy=False)
        size = dt.itemsize
        if dt.kind == "f":
            esz, msz = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
            bias = (1 << (esz - 1)) - 1
            dtype_msg = struct.pack("<BBBBI", 0x11, 0x20, 8 * size - 1, 0, size) + \
                struct.pack("<HHBBBBI", 0, 8 * size, msz, esz, 0, msz, bias)
        elif dt.kind in "iu":
            dtype_msg = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, size) + \
                struct.pack("<HH", 0, 8 * size)
        else:
            raise Hdf5FormatError(f"cannot store dtype {dt}")
        space_msg = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill_msg = struct.pack("<BBBB", 2, 2, 2, 0)                      # v2: late allocation, fill if set, undefined
        raw = arr.tobytes()
        if compact:
            if len(raw) > 0xFFF0:
                raise Hdf5FormatError("compact datasets hold less than 64 KiB")
            layout_msg = struct.pack("<BBH", 3, 0, len(raw)) + raw
        else:
            addr = self.alloc(raw) if raw else UNDEF
            layout_msg = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        return self.object_header([self.message(0x0001, space_msg), self.message(0x0003, dtype_msg, flags=1),
                                   self.message(0x0005, fill_msg), self.message(0x0008, layout_msg)])

    # -- group
    def group(self, children: dict):
        """children: name -> (object header address, btree, heap) [btree / heap None for datasets] -> same triple."""
        names = sorted(children, key=lambda s: s.encode("utf-8"))
        heap_data = bytearray(8)                                          # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            raw = n.encode("utf-8") + b"\x00"
            heap_data += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))     # free list: none (1)

        def entry(n):
            ohdr, bt, hp = children[n]
            if bt is None:
                return struct.pack("<QQII16x", offs[n], ohdr, 0, 0)
            return struct.pack("<QQIIQQ", offs[n], ohdr, 1, 0, bt, hp)

        # leaves: symbol-table nodes of up to 2 * LEAF_K entries
        level = []
        for i in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            body = b"".join(entry(n) for n in part)
            body += b"\x00" * (40 * 2 * LEAF_K - len(body))
            addr = self.alloc(b"SNOD" + struct.pack("<BxH", 1, len(part)) + body)
            level.append((addr, offs[part[-1]] if part else 0))           # (child address, heap offset of its largest name)
        depth = 0
        while True:
            nodes = []
            for i in range(0, len(level), 2 * INTERNAL_K):
                part = level[i:i + 2 * INTERNAL_K]
                body = struct.pack("<Q", 0 if i == 0 else level[i - 1][1])                  # key 0: less than every name below
                for child, last in part:
                    body += struct.pack("<QQ", child, last)
                body += b"\x00" * (8 + 16 * 2 * INTERNAL_K - len(body))
                nodes.append((b"TREE" + struct.pack("<BBH", 0, depth, len(part)), body, part[-1][1]))
            placed = []
            for j, (head, body, last) in enumerate(nodes):
                placed.append([self.alloc(head + struct.pack("<QQ", UNDEF, UNDEF) + body), last])
            for j in range(len(placed)):                                   # sibling links
                left = placed[j - 1][0] if j > 0 else UNDEF
                right = placed[j + 1][0] if j + 1 < len(placed) else UNDEF
                struct.pack_into("<QQ", self.buf, placed[j][0] + 8, left, right)
            level = [tuple(p) for p in placed]
            depth += 1
            if len(level) == 1:
                break
        btree = level[0][0]
        ohdr = self.object_header([self.message(0x0011, struct.pack("<QQ", btree, heap))])
        return ohdr, btree, heap

    def finish(self, root):
        ohdr, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, ohdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(path, tree: dict, compact=()):
    """tree: nested dict, leaves are arrays (e.g. {"mesh_0": {"subdomain_0": {"x": ..., ...}}}); `compact`: dataset
    names stored with the compact layout (in the object header) instead of contiguous."""
    w = _Writer()

    def build(node):
        kids = {}
        for name, v in node.items():
            if "/" in name or not name:
                raise Hdf5FormatError(f"bad link name {name!r}")
            if isinstance(v, dict):
                kids[name] = build(v)
            else:
                kids[name] = (w.dataset(np.asarray(v), compact=name in compact), None, None)
        return w.group(kids)

    data = w.finish(build(tree))
    with open(path, "wb") as fh:
        fh.write(data)
