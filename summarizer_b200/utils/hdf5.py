"""Minimal HDF5 reader / writer in pure Python + numpy — no h5py, no libhdf5 (SURVEY.md §8f NEXT-3).

Covers what the reference touches through h5py: the dataset files of datasets/README.md:5-42 (groups ``/video_k`` with
float / integer arrays, integer scalars and a name string) which it opens with ``h5py.File(path, "r")``
(models/__init__.py:15), and the ``<split>_preds.h5`` file ``predict_dataset`` writes with ``create_group`` /
``create_dataset(name, data=...)`` (models/__init__.py:149-177) and ``summary.py:40-43`` reads back.

File format (HDF5 File Format Specification, version 1.1 structures — what libhdf5 writes by default and h5py 2.10
therefore produces): superblock version 0; "old style" groups = object header (version 1) with a Symbol Table message
-> v1 B-tree of symbol-table nodes + local heap with the link names; datasets = object header with Dataspace (v1),
Datatype (v1), Fill Value (v2) and Data Layout (v3) messages.

* writer: contiguous layout only, numeric little-endian types (int8..64, uint8..64, float32/64), fixed-length strings;
  every structure is laid out exactly as the specification words it so that libhdf5 / h5py open the result.
* reader: additionally follows object-header continuation blocks, compact and chunked layouts (v1 chunk B-tree, gzip and
  shuffle filters), big-endian numeric types and variable-length strings (global heap) — i.e. what default h5py files
  of this schema can contain.  Anything else (new-style groups, v2 object headers, other filters) raises
  ``NotImplementedError`` naming the feature.

API: the h5py subset the reference uses — ``File(path, mode)`` as context manager, ``keys() / __contains__ / __getitem__``
with ``/``-separated paths, ``create_group``, ``create_dataset(name, data=...)``, ``Dataset[...]``, ``Dataset[()]``,
``.shape / .dtype``.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 32, 16            # symbol-table node holds 2*LEAF_K entries, a group B-tree node 2*INTERNAL_K children
HEAP_FREE_NULL = 1                     # libhdf5's H5HL_FREE_NULL: end of the local heap's free list

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_FILTERS = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8, 0xB
MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0x10, 0x11


def _pad8(n):
    return (n + 7) & ~7


# =====================================================================================================================
# datatype message <-> numpy dtype
# =====================================================================================================================
def encode_datatype(dt):
    """Datatype message (version 1) of a numpy dtype."""
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00            # bit 0: little endian (0), bit 3: signed
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        sign, eloc, esize, msize, bias = (31, 23, 8, 23, 127) if dt.itemsize == 4 else (63, 52, 11, 52, 1023)
        # bit field: byte 0 = little endian, mantissa normalisation 2 (implied msb) in bits 4-5; byte 1 = sign bit position
        return (struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize)
                + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias))
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))   # null-padded, ASCII
    raise TypeError(f"hdf5 writer: unsupported dtype {dt}")


def decode_datatype(buf):
    """-> (numpy dtype or ("vlen_str",), element size in the file)."""
    cls, ver = buf[0] & 0x0F, buf[0] >> 4
    b0, b1 = buf[1], buf[2]
    size = struct.unpack_from("<I", buf, 4)[0]
    if ver not in (1, 2, 3):
        raise NotImplementedError(f"hdf5: datatype message version {ver}")
    order = ">" if (b0 & 1) else "<"
    if cls == 0:
        return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}"), size
    if cls == 1:
        if size not in (2, 4, 8):
            raise NotImplementedError(f"hdf5: {size}-byte floating point")
        return np.dtype(f"{order}f{size}"), size
    if cls == 3:
        return np.dtype(f"S{size}"), size
    if cls == 9:
        if (b0 & 0x0F) != 1:
            raise NotImplementedError("hdf5: variable-length sequences (only variable-length strings are read)")
        return ("vlen_str", "utf-8" if (b1 & 0x0F) == 1 else "ascii"), size
    raise NotImplementedError(f"hdf5: datatype class {cls}")


# =====================================================================================================================
# writer
# =====================================================================================================================
class _WGroup:
    def __init__(self, file, name):
        self._file, self.name, self._children = file, name, {}

    def _split(self, path):
        return [p for p in path.split("/") if p]

    def create_group(self, path):
        g = self
        for part in self._split(path):
            nxt = g._children.get(part)
            if nxt is None:
                nxt = g._children[part] = _WGroup(self._file, part)
            elif not isinstance(nxt, _WGroup):
                raise ValueError(f"{part} already exists and is not a group")
            g = nxt
        return g

    require_group = create_group

    def create_dataset(self, path, data=None, shape=None, dtype=None):
        parts = self._split(path)
        g = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
        if parts[-1] in g._children:
            raise ValueError(f"Unable to create dataset (name already exists): {parts[-1]}")
        if data is None:
            data = np.zeros(shape, dtype=dtype or np.float32)
        if isinstance(data, str):
            data = data.encode("utf-8")
        arr = np.asarray(data, dtype=dtype) if dtype is not None else np.asarray(data)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind == "b":
            arr = arr.astype(np.uint8)
        if arr.dtype.kind == "f" and arr.dtype.itemsize == 2:
            arr = arr.astype(np.float32)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        encode_datatype(arr.dtype)                           # raises early for unsupported types
        d = g._children[parts[-1]] = _WDataset(parts[-1], np.ascontiguousarray(arr) if arr.ndim else arr)
        return d

    def __getitem__(self, path):
        node = self
        for part in self._split(path):
            node = node._children[part]
        return node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except (KeyError, AttributeError):
            return False

    def keys(self):
        return list(self._children.keys())


class _WDataset:
    def __init__(self, name, arr):
        self.name, self._arr = name, arr
        self.shape, self.dtype = arr.shape, arr.dtype

    def __getitem__(self, idx):
        return self._arr[idx] if idx is not Ellipsis and idx != () else (self._arr[()] if self._arr.ndim == 0 else self._arr.copy())


class _Writer:
    """Lays the tree out bottom-up (children before their group) and streams it to the file."""

    def __init__(self, fh):
        self.fh, self.eof = fh, 96                            # superblock v0 with 8-byte offsets / lengths is 96 bytes
        self.leaf_k = LEAF_K

    def put(self, blob, align=8):
        addr = _pad8(self.eof) if align == 8 else self.eof
        self.fh.seek(addr)
        self.fh.write(blob)
        self.eof = addr + len(blob)
        return addr

    @staticmethod
    def message(mtype, body):
        body = body + b"\0" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def object_header(self, messages):
        chunk = b"".join(messages)
        # version 1 prefix: version, reserved, #messages, reference count, chunk-0 size, then 4 bytes so that the
        # messages start 8-byte aligned
        return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(chunk)) + chunk

    def dataset(self, d):
        arr = d._arr
        raw = arr.tobytes()
        data_addr = self.put(raw) if raw else UNDEF
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", n) for n in arr.shape)
        fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)           # v2: late allocation, write fill if set, default (size 0) fill
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        hdr = self.object_header([self.message(MSG_DATASPACE, space), self.message(MSG_DATATYPE, encode_datatype(arr.dtype)),
                                  self.message(MSG_FILL, fill), self.message(MSG_LAYOUT, layout)])
        return self.put(hdr), None

    def group(self, g):
        names = sorted(g._children, key=lambda s: s.encode("utf-8"))
        entries = []                                          # (name bytes, header address, (btree, heap) or None)
        for n in names:
            child = g._children[n]
            addr, scratch = self.group(child) if isinstance(child, _WGroup) else self.dataset(child)
            entries.append((n.encode("utf-8"), addr, scratch))
        # local heap: offset 0 holds the empty string (the B-tree's first key), then the names, then one free block
        seg, offsets = bytearray(8), []
        for nb, _, _ in entries:
            offsets.append(len(seg))
            seg += nb + b"\0" * (_pad8(len(nb) + 1) - len(nb))
        free_at = len(seg)
        seg += struct.pack("<QQ", HEAP_FREE_NULL, 16)         # free block: next = end of list, size of this block
        heap_addr = _pad8(self.eof)
        self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_at, heap_addr + 32) + bytes(seg))
        # symbol-table nodes of up to 2*leaf_k entries, in name order
        per = 2 * self.leaf_k
        snods, keys = [], [0]
        for i0 in range(0, len(entries), per):
            part = entries[i0:i0 + per]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for j, (nb, addr, scratch) in enumerate(part):
                if scratch is not None:
                    body += struct.pack("<QQII", offsets[i0 + j], addr, 1, 0) + struct.pack("<QQ", *scratch)
                else:
                    body += struct.pack("<QQII", offsets[i0 + j], addr, 0, 0) + b"\0" * 16
            body += b"\0" * (8 + per * 40 - len(body))
            snods.append(self.put(body))
            keys.append(offsets[i0 + len(part) - 1])          # key right of child i: its largest name
        if len(snods) > 2 * INTERNAL_K:
            raise ValueError(f"hdf5 writer: more than {2 * INTERNAL_K * per} links in one group")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, child in enumerate(snods):
            node += struct.pack("<QQ", keys[i], child)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\0" * (24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8 - len(node))
        btree_addr = self.put(node)
        hdr = self.object_header([self.message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree_addr, heap_addr))])
        return self.put(hdr), (btree_addr, heap_addr)

    def finish(self, root):
        biggest = [0]

        def walk(g):
            biggest[0] = max(biggest[0], len(g._children))
            for c in g._children.values():
                if isinstance(c, _WGroup):
                    walk(c)
        walk(root)
        while 2 * self.leaf_k * 2 * INTERNAL_K < biggest[0]:
            self.leaf_k *= 2
        root_addr, (btree, heap) = self.group(root)
        eof = _pad8(self.eof)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.fh.seek(0)
        self.fh.write(sb)
        if eof > self.eof:
            self.fh.seek(eof - 1)
            self.fh.write(b"\0")


# =====================================================================================================================
# reader
# =====================================================================================================================
class _Reader:
    def __init__(self, fh):
        self.fh = fh
        fh.seek(0, 2)
        self.size = fh.tell()
        base = 0
        while True:                                           # the superblock may sit at 0, 512, 1024, ...
            if base + 8 > self.size:
                raise OSError("not an HDF5 file (signature not found)")
            if self.read(base, 8) == SIGNATURE:
                break
            base = 512 if base == 0 else base * 2
        ver = self.read(base + 8, 1)[0]
        if ver not in (0, 1):
            raise NotImplementedError(f"hdf5: superblock version {ver} (files written with libver='latest' are not supported)")
        so, sl = self.read(base + 13, 2)
        if (so, sl) != (8, 8):
            raise NotImplementedError(f"hdf5: {so}-byte offsets / {sl}-byte lengths")
        self.leaf_k, self.internal_k = struct.unpack("<HH", self.read(base + 16, 4))
        p = base + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack("<QQQQ", self.read(p, 32))
        self.root_entry = self.sym_entry(self.read(p + 32, 40))

    def read(self, addr, n):
        self.fh.seek(addr)
        b = self.fh.read(n)
        if len(b) != n:
            raise OSError(f"hdf5: truncated file (wanted {n} bytes at {addr})")
        return b

    @staticmethod
    def sym_entry(b):
        name_off, hdr, cache = struct.unpack_from("<QQI", b, 0)
        return dict(name_off=name_off, header=hdr, cache=cache, scratch=struct.unpack_from("<QQ", b, 24))

    # ---- object headers ----------------------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, body bytes)] of a version-1 object header, continuation blocks followed."""
        addr += self.base
        pre = self.read(addr, 16)
        if pre[:4] == b"OHDR":
            raise NotImplementedError("hdf5: version-2 object headers (file written with libver='latest')")
        if pre[0] != 1:
            raise NotImplementedError(f"hdf5: object header version {pre[0]}")
        n_msgs, _, size = struct.unpack_from("<HII", pre, 2)
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < n_msgs:
            at, left = blocks.pop(0)
            buf = self.read(at, left)
            p = 0
            while p + 8 <= len(buf) and len(out) < n_msgs:
                mtype, msize, flags = struct.unpack_from("<HHB", buf, p)
                body = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == MSG_CONTINUATION:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off + self.base, ln))
                if flags & 0x02:
                    raise NotImplementedError("hdf5: shared header messages")
                out.append((mtype, body))
        return out

    # ---- groups ------------------------------------------------------------------------------------------------------
    def heap_name(self, heap_addr, off):
        h = self.read(heap_addr + self.base, 32)
        if h[:4] != b"HEAP":
            raise OSError("hdf5: bad local heap signature")
        seg_size, _, seg_addr = struct.unpack_from("<QQQ", h, 8)
        seg = self.read(seg_addr + self.base, seg_size)
        return seg[off:seg.index(b"\0", off)].decode("utf-8")

    def group_links(self, btree_addr, heap_addr):
        """{name: symbol-table entry} by walking the group's v1 B-tree."""
        h = self.read(heap_addr + self.base, 32)
        if h[:4] != b"HEAP":
            raise OSError("hdf5: bad local heap signature")
        seg_size, _, seg_addr = struct.unpack_from("<QQQ", h, 8)
        seg = self.read(seg_addr + self.base, seg_size)
        out = {}

        def node(addr):
            hd = self.read(addr + self.base, 24)
            if hd[:4] != b"TREE":
                raise OSError("hdf5: bad B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", hd, 4)
            if ntype != 0:
                raise OSError("hdf5: group B-tree expected")
            body = self.read(addr + self.base + 24, used * 16 + 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, 16 * i + 8)[0]
                if level > 0:
                    node(child)
                    continue
                sn = self.read(child + self.base, 8)
                if sn[:4] != b"SNOD":
                    raise OSError("hdf5: bad symbol table node signature")
                n = struct.unpack_from("<H", sn, 6)[0]
                ents = self.read(child + self.base + 8, 40 * n)
                for j in range(n):
                    e = self.sym_entry(ents[40 * j:40 * j + 40])
                    if e["cache"] == 2:
                        continue                              # symbolic link: not followed
                    name = seg[e["name_off"]:seg.index(b"\0", e["name_off"])].decode("utf-8")
                    out[name] = e
        node(btree_addr)
        return out

    def open(self, header_addr, name):
        msgs = self.messages(header_addr)
        kinds = {t for t, _ in msgs}
        if MSG_SYMBOL_TABLE in kinds:
            body = next(b for t, b in msgs if t == MSG_SYMBOL_TABLE)
            return Group(self, name, *struct.unpack_from("<QQ", body, 0))
        if MSG_LAYOUT in kinds:
            return Dataset(self, name, msgs)
        if 0x2 in kinds or 0x6 in kinds:
            raise NotImplementedError("hdf5: new-style (link message / dense) groups — file written with libver='latest'")
        raise OSError(f"hdf5: object {name!r} is neither a group nor a dataset")

    # ---- raw data ----------------------------------------------------------------------------------------------------
    def chunks(self, btree_addr, rank):
        """[(offsets tuple, address, stored size, filter mask)] from a v1 chunk B-tree (node type 1)."""
        out = []

        def node(addr):
            hd = self.read(addr + self.base, 24)
            if hd[:4] != b"TREE":
                raise OSError("hdf5: bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", hd, 4)
            ksz = 8 + 8 * (rank + 1)
            body = self.read(addr + self.base + 24, used * (ksz + 8) + ksz)
            for i in range(used):
                p = i * (ksz + 8)
                csize, fmask = struct.unpack_from("<II", body, p)
                offs = struct.unpack_from(f"<{rank + 1}Q", body, p + 8)[:rank]
                child = struct.unpack_from("<Q", body, p + ksz)[0]
                if level > 0:
                    node(child)
                else:
                    out.append((offs, child, csize, fmask))
        if btree_addr != UNDEF:
            node(btree_addr)
        return out

    def global_heap_object(self, coll_addr, index):
        hd = self.read(coll_addr + self.base, 16)
        if hd[:4] != b"GCOL":
            raise OSError("hdf5: bad global heap signature")
        total = struct.unpack_from("<Q", hd, 8)[0]
        buf = self.read(coll_addr + self.base, total)
        p = 16
        while p + 16 <= total:
            idx, _, _, size = struct.unpack_from("<HHIQ", buf, p)
            if idx == 0:
                break
            if idx == index:
                return buf[p + 16:p + 16 + size]
            p += 16 + _pad8(size)
        raise OSError("hdf5: global heap object not found")


class Group:
    def __init__(self, reader, name, btree, heap):
        self._r, self.name, self._btree, self._heap = reader, name, btree, heap
        self._links = None

    def _entries(self):
        if self._links is None:
            self._links = self._r.group_links(self._btree, self._heap)
        return self._links

    def keys(self):
        return list(self._entries().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._entries())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            ents = node._entries()
            if part not in ents:
                raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
            node = node._r.open(ents[part]["header"], (node.name.rstrip("/") + "/" + part))
        return node

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class Dataset:
    def __init__(self, reader, name, msgs):
        self._r, self.name = reader, name
        self._filters = []
        for t, b in msgs:
            if t == MSG_DATASPACE:
                ver, rank, flags = b[0], b[1], b[2]
                if ver == 1:
                    self.shape = struct.unpack_from(f"<{rank}Q", b, 8) if rank else ()
                elif ver == 2:
                    if b[3] == 2:
                        raise NotImplementedError("hdf5: null dataspace")
                    self.shape = struct.unpack_from(f"<{rank}Q", b, 4) if rank else ()
                else:
                    raise NotImplementedError(f"hdf5: dataspace version {ver}")
            elif t == MSG_DATATYPE:
                self._type, self._esize = decode_datatype(b)
            elif t == MSG_LAYOUT:
                self._layout = b
            elif t == MSG_FILTERS:
                self._filters = self._parse_filters(b)
        self.dtype = self._type if isinstance(self._type, np.dtype) else np.dtype(object)
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @staticmethod
    def _parse_filters(b):
        ver, n = b[0], b[1]
        if ver != 1:
            raise NotImplementedError(f"hdf5: filter pipeline version {ver}")
        p, out = 8, []
        for _ in range(n):
            fid, nlen, _, ncd = struct.unpack_from("<HHHH", b, p)
            p += 8 + _pad8(nlen)
            cd = struct.unpack_from(f"<{ncd}I", b, p)
            p += 4 * ncd + (4 if ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _raw(self):
        """All element bytes of the dataset, in C order."""
        lay, r = self._layout, self._r
        total = self.size * self._esize
        if lay[0] != 3:
            raise NotImplementedError(f"hdf5: data layout message version {lay[0]}")
        if lay[1] == 0:                                       # compact
            n = struct.unpack_from("<H", lay, 2)[0]
            return bytes(lay[4:4 + n])
        if lay[1] == 1:                                       # contiguous
            addr, n = struct.unpack_from("<QQ", lay, 2)
            return b"\0" * total if addr == UNDEF else r.read(addr + r.base, min(n, total))
        if lay[1] != 2:
            raise NotImplementedError(f"hdf5: layout class {lay[1]}")
        rank = lay[2] - 1                                     # chunked: the last dimension is the element size
        btree = struct.unpack_from("<Q", lay, 3)[0]
        cdims = struct.unpack_from(f"<{rank}I", lay, 11)
        out = np.zeros(self.shape, dtype=np.uint8 if False else f"V{self._esize}")
        for offs, addr, csize, fmask in r.chunks(btree, rank):
            buf = r.read(addr + r.base, csize)
            for k, (fid, cd) in reversed(list(enumerate(self._filters))):
                if fmask & (1 << k):
                    continue
                if fid == 1:
                    buf = zlib.decompress(buf)
                elif fid == 2:                                # shuffle: bytes of the elements were transposed
                    es = cd[0] if cd else self._esize
                    n = len(buf) // es
                    buf = np.frombuffer(buf[:n * es], np.uint8).reshape(es, n).T.tobytes() + buf[n * es:]
                elif fid == 3:                                # fletcher32 checksum trails the chunk
                    buf = buf[:-4]
                else:
                    raise NotImplementedError(f"hdf5: filter id {fid}")
            chunk = np.frombuffer(buf, dtype=f"V{self._esize}", count=int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
            out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out.tobytes()

    def _array(self):
        raw = self._raw()
        if isinstance(self._type, np.dtype):
            a = np.frombuffer(raw, dtype=self._type, count=self.size).reshape(self.shape)
            return a.astype(a.dtype.newbyteorder("=")) if a.dtype.byteorder == ">" else a.copy()
        enc = self._type[1]                                   # variable-length strings: (length, heap address, index)
        vals = []
        for i in range(self.size):
            _, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
            vals.append(self._r.global_heap_object(addr, idx).decode(enc) if addr not in (0, UNDEF) else "")
        a = np.empty(self.size, dtype=object)
        a[:] = vals
        return a.reshape(self.shape)

    def __getitem__(self, idx):
        a = self._array()
        if idx is Ellipsis:
            return a if a.ndim else a[()]
        if isinstance(idx, tuple) and idx == ():
            return a[()] if a.ndim == 0 else a
        return a[idx]

    def __array__(self, dtype=None, copy=None):
        a = self._array()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        if not self.shape:
            raise TypeError("Attempt to take len() of scalar dataset")
        return self.shape[0]

    def tolist(self):
        return self._array().tolist()


class File:
    """``File(path, "r")`` -> read-only root group; ``File(path, "w")`` -> in-memory tree written on ``close()``."""

    def __init__(self, path, mode="r"):
        self.filename, self.mode = path, mode
        if mode == "r":
            self._fh = open(path, "rb")
            self._reader = _Reader(self._fh)
            e = self._reader.root_entry
            if e["cache"] == 1:
                self._root = Group(self._reader, "/", *e["scratch"])
            else:
                self._root = self._reader.open(e["header"], "/")
        elif mode == "w":
            self._fh = open(path, "wb")
            self._root = _WGroup(self, "/")
        else:
            raise ValueError("hdf5.File: mode must be 'r' or 'w'")

    def __getattr__(self, name):
        if name in ("keys", "items", "create_group", "require_group", "create_dataset"):
            return getattr(self._root, name)
        raise AttributeError(name)

    def __getitem__(self, path):
        return self._root if path in ("/", "") else self._root[path]

    def __contains__(self, path):
        return path in self._root

    def __iter__(self):
        return iter(self._root.keys())

    def __len__(self):
        return len(self._root.keys())

    def close(self):
        if self._fh is None:
            return
        if self.mode == "w":
            _Writer(self._fh).finish(self._root)
        self._fh.close()
        self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# =====================================================================================================================
# specification-level checker (used by the tests; independent walk of the structures the writer emits)
# =====================================================================================================================
def check_file(path):
    """Walks a file the way libhdf5's H5F / H5G / H5O loaders do and asserts the invariants they enforce: signature,
    versions, sizes of offsets, superblock EOF, 8-byte alignment of every object header / message, B-tree and symbol-table
    node sizes implied by the K values, ascending link names, key[i] <= names of child i <= key[i+1], local-heap free
    list inside the data segment, contiguous data inside the file.  Returns the number of objects visited."""
    with open(path, "rb") as fh:
        buf = fh.read()
    assert buf[:8] == SIGNATURE, "signature"
    ver_sb, ver_fs, ver_root, _, ver_shm, so, sl, _ = struct.unpack_from("<8B", buf, 8)
    assert (ver_sb, ver_fs, ver_root, ver_shm, so, sl) == (0, 0, 0, 0, 8, 8), "superblock versions / sizes"
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", buf, 16)
    assert leaf_k > 0 and internal_k > 0 and flags == 0
    base, free, eof, driver = struct.unpack_from("<QQQQ", buf, 24)
    assert base == 0 and free == UNDEF and driver == UNDEF and eof == len(buf), "superblock addresses"
    visited = [0]

    def header(addr):
        assert addr % 8 == 0 and addr + 16 <= eof, "object header alignment"
        ver, _, n_msgs, refcnt, size = struct.unpack_from("<BBHII", buf, addr)
        assert ver == 1 and refcnt >= 1 and size % 8 == 0 and addr + 16 + size <= eof
        p, out = addr + 16, []
        for _ in range(n_msgs):
            mtype, msize, mflags = struct.unpack_from("<HHB", buf, p)
            assert msize % 8 == 0 and p + 8 + msize <= addr + 16 + size, "message size"
            out.append((mtype, buf[p + 8:p + 8 + msize]))
            p += 8 + msize
        assert p == addr + 16 + size, "messages must fill chunk 0 exactly"
        return out

    def group(hdr_addr, scratch):
        visited[0] += 1
        msgs = header(hdr_addr)
        assert [t for t, _ in msgs] == [MSG_SYMBOL_TABLE]
        btree, heap = struct.unpack_from("<QQ", msgs[0][1], 0)
        if scratch is not None:
            assert scratch == (btree, heap), "cached symbol-table info must match the header message"
        assert buf[heap:heap + 4] == b"HEAP" and buf[heap + 4] == 0
        seg_size, free_head, seg_addr = struct.unpack_from("<QQQ", buf, heap + 8)
        assert seg_size % 8 == 0 and seg_addr + seg_size <= eof and buf[seg_addr] == 0, "heap segment / empty first name"
        while free_head != HEAP_FREE_NULL:
            assert free_head % 8 == 0 and free_head + 16 <= seg_size, "free block inside the segment"
            nxt, fsize = struct.unpack_from("<QQ", buf, seg_addr + free_head)
            assert fsize >= 16 and free_head + fsize <= seg_size
            free_head = nxt
        name_at = lambda off: buf[seg_addr + off:buf.index(b"\0", seg_addr + off)]
        assert buf[btree:btree + 4] == b"TREE"
        ntype, level, used, left, right = struct.unpack_from("<BBHQQ", buf, btree + 4)
        assert (ntype, level, left, right) == (0, 0, UNDEF, UNDEF) and used <= 2 * internal_k
        assert btree + 24 + 2 * internal_k * 8 + (2 * internal_k + 1) * 8 <= eof, "full-size B-tree node allocated"
        prev = b""
        for i in range(used):
            k_left, child, k_right = struct.unpack_from("<QQQ", buf, btree + 24 + 16 * i)
            assert buf[child:child + 4] == b"SNOD" and buf[child + 4] == 1
            n = struct.unpack_from("<H", buf, child + 6)[0]
            assert 0 < n <= 2 * leaf_k and child + 8 + 2 * leaf_k * 40 <= eof, "full-size symbol-table node allocated"
            for j in range(n):
                off, ohdr, cache, _ = struct.unpack_from("<QQII", buf, child + 8 + 40 * j)
                nm = name_at(off)
                assert nm > prev, "link names strictly ascending"
                assert name_at(k_left) < nm <= name_at(k_right) or (i == 0 and k_left == 0 and nm <= name_at(k_right)), "B-tree keys bracket the names"
                prev = nm
                if cache == 1:
                    group(ohdr, struct.unpack_from("<QQ", buf, child + 8 + 40 * j + 24))
                else:
                    assert cache == 0
                    dataset(ohdr)

    def dataset(hdr_addr):
        visited[0] += 1
        msgs = dict(header(hdr_addr))
        assert set(msgs) == {MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT}
        sp = msgs[MSG_DATASPACE]
        assert sp[0] == 1 and sp[2] == 0
        dims = struct.unpack_from(f"<{sp[1]}Q", sp, 8)
        dt, esize = decode_datatype(msgs[MSG_DATATYPE])
        assert isinstance(dt, np.dtype) and dt.itemsize == esize
        lay = msgs[MSG_LAYOUT]
        assert lay[0] == 3 and lay[1] == 1
        addr, size = struct.unpack_from("<QQ", lay, 2)
        assert size == int(np.prod(dims, dtype=np.int64)) * esize if dims else size == esize
        assert (addr == UNDEF and size == 0) or (addr % 8 == 0 and addr + size <= eof), "raw data inside the file"
        assert msgs[MSG_FILL][0] == 2

    _, root_hdr, cache, _ = struct.unpack_from("<QQII", buf, 56)
    assert cache == 1
    group(root_hdr, struct.unpack_from("<QQ", buf, 80))
    return visited[0]
