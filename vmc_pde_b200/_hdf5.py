"""Minimal HDF5 reader / writer for the `infos.hdf5` files of the reference (util.py:29-32: one flat group, one
dataset per key, `f.create_dataset(key, data=np.array(value))`; readers visualization.py:141-280, paper_plot/*.py).

h5py is not available in this image, so the subset of the format those files use is implemented directly from the
HDF5 file-format specification (version 0 superblock, version 1 object headers, symbol-table groups with a version 1
B-tree and a local heap, contiguous little-endian fixed-point / IEEE datasets):

  * `read(path)`  -> {name: ndarray}; understands what h5py 2.x/3.x writes with default settings for such files
    (also chunked layouts without filters are refused loudly rather than misread);
  * `write(path, {name: array})` writes the same structures (one root group, contiguous datasets) to the specification, the
    way libhdf5 laid them out in the reference's own files (which the reader is tested against); meant to be opened by
    h5py / the reference's plotting scripts, but with no libhdf5 in this image that read-back itself is UNTESTED here.

Pure NumPy; no device code (this is the on-disk side of SURVEY 8f rank 2, not the hot path).
"""
import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise ValueError(f"superblock version {buf[8]} is not supported (the reference's files use version 0)")
        self.so, self.sl = buf[13], buf[14]           # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise ValueError("only 8-byte offsets and lengths are supported")
        # base address, free-space address, end of file, driver info; then the root symbol-table entry
        self.base = self.u64(24)
        self.root = self.sym_entry(24 + 32)

    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]

    def sym_entry(self, o):
        name_off, header, cache = self.u64(o), self.u64(o + 8), self.u32(o + 16)
        e = {"name_off": name_off, "header": header, "cache": cache}
        if cache == 1:
            e["btree"], e["heap"] = self.u64(o + 24), self.u64(o + 32)
        return e

    def heap_data(self, addr):
        if self.b[addr:addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return self.u64(addr + 24)                      # address of the data segment

    def heap_str(self, seg, off):
        end = self.b.index(b"\0", seg + off)
        return self.b[seg + off:end].decode()

    def group_entries(self, btree, heap):
        seg = self.heap_data(heap)
        out = []

        def walk(addr):
            if self.b[addr:addr + 4] != b"TREE":
                raise ValueError("bad B-tree signature")
            ntype, level, used = self.b[addr + 4], self.b[addr + 5], self.u16(addr + 6)
            if ntype != 0:
                raise ValueError("expected a group B-tree")
            p = addr + 8 + 16                            # skip sibling addresses
            for i in range(used):
                child = self.u64(p + 8 + i * 16)         # key_i (8) child_i (8) ...
                if level > 0:
                    walk(child)
                else:
                    if self.b[child:child + 4] != b"SNOD":
                        raise ValueError("bad symbol-table node signature")
                    n = self.u16(child + 6)
                    for k in range(n):
                        e = self.sym_entry(child + 8 + k * 40)
                        out.append((self.heap_str(seg, e["name_off"]), e))

        walk(btree)
        return out

    def messages(self, addr):
        """(type, offset, size) of the messages of a version-1 object header, following continuation blocks."""
        if self.b[addr] != 1:
            raise ValueError(f"object header version {self.b[addr]} is not supported")
        nmsg, hsize = self.u16(addr + 2), self.u32(addr + 8)
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize = self.u16(p), self.u16(p + 2)
                body = p + 8
                msgs.append((mtype, body, msize))
                if mtype == 0x10:
                    blocks.append((self.u64(body), self.u64(body + 8)))
                p = body + msize
        return msgs

    def dataset(self, header):
        shape, dtype, data = None, None, None
        for mtype, o, size in self.messages(header):
            if mtype == 0x1:                             # dataspace
                ver, rank, flags = self.b[o], self.b[o + 1], self.b[o + 2]
                p = o + (8 if ver == 1 else 4)
                shape = tuple(self.u64(p + 8 * i) for i in range(rank))
            elif mtype == 0x3:                           # datatype
                cls, bits0 = self.b[o] & 0x0F, self.b[o + 1]
                nbytes = self.u32(o + 4)
                if bits0 & 1:
                    raise ValueError("big-endian data is not supported")
                if cls == 1:
                    dtype = np.dtype(f"<f{nbytes}")
                elif cls == 0:
                    dtype = np.dtype(("<i" if bits0 & 0x08 else "<u") + str(nbytes))
                else:
                    raise ValueError(f"datatype class {cls} is not supported")
            elif mtype == 0x8:                           # data layout
                ver = self.b[o]
                if ver == 3:
                    lclass = self.b[o + 1]
                    if lclass == 1:
                        data = (self.u64(o + 2), self.u64(o + 10))
                    elif lclass == 0:
                        n = self.u16(o + 2)
                        data = ("compact", o + 4, n)
                    else:
                        raise ValueError("chunked datasets are not supported")
                else:
                    raise ValueError(f"data layout version {ver} is not supported")
        if shape is None or dtype is None or data is None:
            raise ValueError("incomplete dataset header")
        count = int(np.prod(shape)) if shape else 1
        if data[0] == "compact":
            raw = self.b[data[1]:data[1] + data[2]]
        elif data[0] == _UNDEF or count == 0:
            raw = b""
        else:
            raw = self.b[self.base + data[0]:self.base + data[0] + count * dtype.itemsize]
        return np.frombuffer(raw, dtype=dtype, count=count if raw else 0).reshape(shape).copy()

    def read_all(self):
        r = self.root
        if r["cache"] != 1:
            # fetch the symbol-table message of the root object header
            for mtype, o, size in self.messages(r["header"]):
                if mtype == 0x11:
                    r = {"btree": self.u64(o), "heap": self.u64(o + 8)}
        return {name: self.dataset(e["header"]) for name, e in self.group_entries(r["btree"], r["heap"])}


def read(path):
    """All datasets of the root group of `path` as {name: ndarray}."""
    with open(path, "rb") as fh:
        return _Reader(fh.read()).read_all()


# ------------------------------------------------------------------------------------------------ writer
def _pad8(n):
    return (n + 7) // 8 * 8


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        elif dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            raise TypeError("float16/longdouble datasets are not supported")
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], dt.itemsize) + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError(f"unsupported dtype {dt}")


def _msg(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


def write(path, datasets):
    """Write {name: array-like} as contiguous datasets of one root group (the layout util.store_infos produces)."""
    items = []
    for name, value in datasets.items():
        a = np.asarray(value)
        if a.dtype == object:
            raise TypeError(f"dataset {name!r} is ragged / object-typed")
        if a.dtype.kind == "b":
            a = a.astype(np.int8)
        if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
            a = a.astype(np.float64)
        items.append((str(name), np.array(a, order="C").astype(a.dtype.newbyteorder("<"), copy=False)))
    items.sort(key=lambda kv: kv[0].encode())            # symbol-table nodes are ordered by name
    n = len(items)
    LEAF_K, INT_K = 4, 16                                 # group leaf node K, internal node K (library defaults)
    per_node = 2 * LEAF_K
    nodes = [items[i:i + per_node] for i in range(0, n, per_node)] or [[]]
    if len(nodes) > 2 * INT_K:
        raise ValueError(f"more than {2 * INT_K * per_node} datasets need a two-level group B-tree (not written here)")

    # local heap data: offset 0 holds the empty string
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for name, _ in items:
        name_off[name] = len(heap)
        raw = name.encode() + b"\0"
        heap += raw + b"\0" * (_pad8(len(raw)) - len(raw))
    heap_size = _pad8(len(heap)) + 16
    heap += b"\0" * (heap_size - len(heap))
    free_off = heap_size - 16
    struct.pack_into("<QQ", heap, free_off, 1, 16)        # one free block at the end (next = 1: last), its size

    # lay the file out: superblock, root header, B-tree, heap header, heap data, symbol nodes, headers, raw data
    pos = 8 + 8 + 4 + 4 + 32 + 40                         # superblock v0 with 8-byte offsets: 96 bytes
    sb_size = pos
    root_hdr_addr = pos
    root_hdr = _object_header([_msg(0x11, struct.pack("<QQ", 0, 0))])   # patched below
    pos += _pad8(len(root_hdr))
    btree_addr = pos
    btree_size = 8 + 16 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8
    pos += btree_size
    heap_addr = pos
    pos += 32
    heap_data_addr = pos
    pos += len(heap)
    snod_size = 8 + per_node * 40
    snod_addrs = []
    for _ in nodes:
        snod_addrs.append(pos)
        pos += snod_size
    hdr_addr, hdrs = {}, {}
    for name, a in items:
        rank = a.ndim
        space = struct.pack("<BBBB", 1, rank, 0, 0) + b"\0" * 4 + b"".join(struct.pack("<Q", s) for s in a.shape)
        hdrs[name] = [_msg(0x1, space), _msg(0x3, _dtype_msg(a.dtype), flags=1), None]
        hdr_addr[name] = pos
        pos += 16 + len(hdrs[name][0]) + len(hdrs[name][1]) + 8 + _pad8(18)
    data_addr = {}
    for name, a in items:
        pos = _pad8(pos)
        data_addr[name] = pos if a.size else _UNDEF
        pos += a.nbytes
    eof = pos

    out = bytearray(eof)
    out[0:8] = _SIG
    struct.pack_into("<BBBBBBBBHHI", out, 8, 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INT_K, 0)
    struct.pack_into("<QQQQ", out, 24, 0, _UNDEF, eof, _UNDEF)
    struct.pack_into("<QQII", out, 56, 0, root_hdr_addr, 1, 0)
    struct.pack_into("<QQ", out, 56 + 24, btree_addr, heap_addr)
    assert sb_size == 96
    root_hdr = _object_header([_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))])
    out[root_hdr_addr:root_hdr_addr + len(root_hdr)] = root_hdr
    # B-tree node (group, leaf level 0): keys are heap offsets of the largest name of each child (key 0 = 0)
    struct.pack_into("<4sBBHQQ", out, btree_addr, b"TREE", 0, 0, len(nodes) if n else 0, _UNDEF, _UNDEF)
    p = btree_addr + 24
    struct.pack_into("<Q", out, p, 0)
    p += 8
    for node, addr in zip(nodes, snod_addrs):
        if not node:
            break
        struct.pack_into("<QQ", out, p, addr, name_off[node[-1][0]])
        p += 16
    struct.pack_into("<4sBBBBQQQ", out, heap_addr, b"HEAP", 0, 0, 0, 0, len(heap), free_off, heap_data_addr)
    out[heap_data_addr:heap_data_addr + len(heap)] = heap
    for node, addr in zip(nodes, snod_addrs):
        struct.pack_into("<4sBBH", out, addr, b"SNOD", 1, 0, len(node))
        for k, (name, _) in enumerate(node):
            struct.pack_into("<QQII", out, addr + 8 + 40 * k, name_off[name], hdr_addr[name], 0, 0)
    for name, a in items:
        layout = struct.pack("<BBQQ", 3, 1, data_addr[name], a.nbytes)
        hdrs[name][2] = _msg(0x8, layout)
        h = _object_header(hdrs[name])
        out[hdr_addr[name]:hdr_addr[name] + len(h)] = h
        if a.size:
            out[data_addr[name]:data_addr[name] + a.nbytes] = a.tobytes()
    with open(path, "wb") as fh:
        fh.write(out)
