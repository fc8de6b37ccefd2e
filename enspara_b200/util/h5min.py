"""A small, dependency-free HDF5 reader/writer for the files on either side of the clustering
path: the reference reads features and mdtraj ``.h5`` trajectories and writes assignments /
distances through PyTables (/root/reference/enspara/ra/ra.py:45-220,
/root/reference/enspara/mpi/io.py:16-65); neither PyTables nor h5py exists in this image.

Reader: superblock v0/v1, v1 object headers, old-style groups (symbol table + B-tree v1 + local
heap), contiguous / compact / chunked (B-tree v1) layouts, deflate + shuffle (+ fletcher32
skipped) filters, fixed-point and floating-point datatypes.  That is what PyTables / h5py write
by default (``libver='earliest'``).  Anything else raises ``H5FormatError`` loudly.

Writer: superblock v0, one root group, contiguous numeric datasets (no filters), laid out after
the HDF5 file-format specification (version 0 structures) for ``ra.save`` style outputs.  Verified
in this repository only by reading the files back with the reader above -- libhdf5 / h5py /
PyTables are not available in the build image to cross-check.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters, attrs):
        self._f = f
        self.name = name
        self.shape = tuple(shape)
        self.dtype = dtype
        self._layout = layout
        self._filters = filters
        self.attrs = attrs

    def read(self):
        return self._f._read_dataset(self)

    def __getitem__(self, key):
        return self.read()[key]

    def __len__(self):
        return self.shape[0] if self.shape else 0


class _Group:
    def __init__(self, f, name, links, attrs):
        self._f = f
        self.name = name
        self._links = links          # name -> object header address
        self.attrs = attrs

    def keys(self):
        return list(self._links)

    def __contains__(self, k):
        return k in self._links

    def __iter__(self):
        return iter(self._links)

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, _Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._object(node._links[part], node.name.rstrip("/") + "/" + part)
        return node


class File(_Group):
    """``File(path)['/coordinates'].read()`` -> numpy array."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self._b = fh.read()
        b = self._b
        if b[:8] != _SIG:
            raise H5FormatError("%s is not an HDF5 file" % path)
        ver = b[8]
        if ver not in (0, 1):
            raise H5FormatError("superblock version %d is not supported (only 0/1)" % ver)
        self._so, self._sl = b[13], b[14]
        if (self._so, self._sl) != (8, 8):
            raise H5FormatError("only 8-byte offsets/lengths are supported")
        p = 24 if ver == 0 else 28
        self._base = struct.unpack_from("<Q", b, p)[0]
        p += 32                                  # base, free-space, eof, driver
        # root symbol table entry
        _, root_ohdr, cache, _ = struct.unpack_from("<QQII", b, p)
        self._cache = {}
        root = self._object(root_ohdr, "/")
        if not isinstance(root, _Group):
            raise H5FormatError("root object is not a group")
        super().__init__(self, "/", root._links, root.attrs)

    # -- low level -----------------------------------------------------------------------
    def _messages(self, addr):
        b = self._b
        addr += self._base
        ver = b[addr]
        if ver != 1:
            raise H5FormatError("object header version %d is not supported (only v1)" % ver)
        nmsg, _, hsize = struct.unpack_from("<HII", b, addr + 2)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:       # continuation
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off + self._base, ln))
                out.append((mtype, body, mflags))
        return out

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        shape = dtype = layout = None
        filters = []
        attrs = {}
        stab = None
        for mtype, body, mflags in msgs:
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                if mflags & 2:
                    raise H5FormatError("shared datatype messages are not supported")
                dtype = self._datatype(body)[0]
            elif mtype == 0x08:
                layout = self._layout_msg(body)
            elif mtype == 0x0B:
                filters = self._filters_msg(body)
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(body)
                    attrs[k] = v
                except H5FormatError:
                    pass                 # attributes are informational here
            elif mtype == 0x11:
                stab = struct.unpack_from("<QQ", body, 0)
            elif mtype in (0x02, 0x06):
                raise H5FormatError("new-style (link message) groups are not supported")
        if stab is not None:
            obj = _Group(self, name, self._group_links(*stab), attrs)
        elif layout is not None and dtype is not None and shape is not None:
            obj = _Dataset(self, name, shape, dtype, layout, filters, attrs)
        else:
            obj = _Group(self, name, {}, attrs)
        self._cache[addr] = obj
        return obj

    def _heap_data(self, heap_addr):
        b = self._b
        p = heap_addr + self._base
        if b[p:p + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        size, _, data_addr = struct.unpack_from("<QQQ", b, p + 8)
        return b[data_addr + self._base:data_addr + self._base + size]

    def _group_links(self, btree_addr, heap_addr):
        heap = self._heap_data(heap_addr)
        links = {}

        def walk(addr):
            b = self._b
            p = addr + self._base
            sig = b[p:p + 4]
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
                if ntype != 0:
                    raise H5FormatError("expected a group B-tree node")
                q = p + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", b, q + 8)[0]   # key_i, child_i
                    q += 16
                    walk(child)
            elif sig == b"SNOD":
                n = struct.unpack_from("<H", b, p + 6)[0]
                q = p + 8
                for i in range(n):
                    name_off, ohdr = struct.unpack_from("<QQ", b, q)
                    q += 40
                    end = heap.index(b"\0", name_off)
                    links[heap[name_off:end].decode("utf8")] = ohdr
            else:
                raise H5FormatError("unexpected node signature %r in a group" % sig)
        walk(btree_addr)
        return links

    # -- messages ------------------------------------------------------------------------
    @staticmethod
    def _dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
        else:
            raise H5FormatError("dataspace version %d" % ver)
        return struct.unpack_from("<%dQ" % rank, body, p) if rank else ()

    @staticmethod
    def _datatype(body):
        cls = body[0] & 0x0F
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if (bits & 1) else "<"
        if cls == 0:      # fixed point
            signed = bool(bits & 0x08)
            return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), 8 + 4
        if cls == 1:      # floating point
            if size not in (2, 4, 8):
                raise H5FormatError("float size %d" % size)
            return np.dtype("%sf%d" % (order, size)), 8 + 12
        if cls == 3:      # fixed-length string
            return np.dtype("S%d" % size), 8
        raise H5FormatError("datatype class %d is not supported" % cls)

    @staticmethod
    def _layout_msg(body):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                size = struct.unpack_from("<H", body, 2)[0]
                return ("compact", body[4:4 + size])
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", body, 2)
                return ("contiguous", addr, size)
            if cls == 2:
                nd = body[2]
                addr = struct.unpack_from("<Q", body, 3)[0]
                dims = struct.unpack_from("<%dI" % nd, body, 11)
                return ("chunked", addr, dims)
            raise H5FormatError("layout class %d" % cls)
        if ver in (1, 2):
            nd, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", body, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % nd, body, p)
            p += 4 * nd
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                return ("chunked", addr, dims)
            size = struct.unpack_from("<I", body, p)[0]
            return ("compact", body[p + 4:p + 4 + size])
        raise H5FormatError("data layout version %d is not supported" % ver)

    @staticmethod
    def _filters_msg(body):
        ver, n = body[0], body[1]
        out = []
        if ver == 1:
            p = 8
            for _ in range(n):
                fid, nlen, flags, ncd = struct.unpack_from("<HHHH", body, p)
                p += 8 + ((nlen + 7) // 8) * 8
                cd = struct.unpack_from("<%dI" % ncd, body, p)
                p += 4 * ncd + (4 if ncd % 2 else 0)
                out.append((fid, cd))
        elif ver == 2:
            p = 2
            for _ in range(n):
                fid = struct.unpack_from("<H", body, p)[0]
                p += 2
                nlen = 0
                if fid >= 256:
                    nlen = struct.unpack_from("<H", body, p)[0]
                    p += 2
                flags, ncd = struct.unpack_from("<HH", body, p)
                p += 4 + nlen
                cd = struct.unpack_from("<%dI" % ncd, body, p)
                p += 4 * ncd
                out.append((fid, cd))
        else:
            raise H5FormatError("filter pipeline version %d" % ver)
        return out

    def _attribute(self, body):
        ver = body[0]
        if ver != 1:
            raise H5FormatError("attribute version %d" % ver)
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        pad = lambda x: (x + 7) // 8 * 8   # noqa: E731
        p = 8
        name = body[p:p + nsz].split(b"\0")[0].decode("utf8")
        p += pad(nsz)
        dt = self._datatype(body[p:p + tsz])[0]
        p += pad(tsz)
        shape = self._dataspace(body[p:p + ssz]) if ssz >= 8 else ()
        p += pad(ssz)
        cnt = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(body, dtype=dt, count=cnt, offset=p)
        if dt.kind == "S":
            val = [v.split(b"\0")[0].decode("utf8", "replace") for v in val]
        return name, (val[0] if not shape else val)

    # -- data ----------------------------------------------------------------------------
    def _read_dataset(self, ds):
        kind = ds._layout[0]
        dt = ds.dtype
        count = int(np.prod(ds.shape)) if ds.shape else 1
        if kind == "compact":
            return np.frombuffer(ds._layout[1], dtype=dt, count=count).reshape(ds.shape).copy()
        if kind == "contiguous":
            addr = ds._layout[1]
            if addr == _UNDEF:
                return np.zeros(ds.shape, dtype=dt)
            return np.frombuffer(self._b, dtype=dt, count=count,
                                 offset=addr + self._base).reshape(ds.shape).copy()
        _, addr, cdims = ds._layout
        rank = len(ds.shape)
        if len(cdims) != rank + 1:
            raise H5FormatError("chunk rank mismatch")
        chunk = tuple(int(c) for c in cdims[:rank])
        out = np.zeros(ds.shape, dtype=dt)
        if addr == _UNDEF:
            return out
        self._walk_chunks(addr, rank, chunk, ds, out)
        return out

    def _walk_chunks(self, addr, rank, chunk, ds, out):
        b = self._b
        p = addr + self._base
        if b[p:p + 4] != b"TREE":
            raise H5FormatError("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        if ntype != 1:
            raise H5FormatError("expected a chunk B-tree node")
        keysz = 8 + 8 * (rank + 1)
        q = p + 24
        for _ in range(used):
            csize, fmask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from("<%dQ" % (rank + 1), b, q + 8)
            child = struct.unpack_from("<Q", b, q + keysz)[0]
            q += keysz + 8
            if level > 0:
                self._walk_chunks(child, rank, chunk, ds, out)
                continue
            raw = b[child + self._base:child + self._base + csize]
            raw = self._defilter(raw, ds._filters, fmask, ds.dtype.itemsize)
            arr = np.frombuffer(raw, dtype=ds.dtype, count=int(np.prod(chunk))).reshape(chunk)
            sel_out, sel_in = [], []
            for d in range(rank):
                lo = int(offs[d])
                hi = min(lo + chunk[d], ds.shape[d])
                sel_out.append(slice(lo, hi))
                sel_in.append(slice(0, hi - lo))
            out[tuple(sel_out)] = arr[tuple(sel_in)]

    @staticmethod
    def _defilter(raw, filters, mask, itemsize):
        for i in reversed(range(len(filters))):
            if mask & (1 << i):
                continue
            fid, cd = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // itemsize
                a = np.frombuffer(raw, dtype=np.uint8, count=n * itemsize)
                raw = a.reshape(itemsize, n).T.tobytes() + raw[n * itemsize:]
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise H5FormatError("HDF5 filter id %d is not supported" % fid)
        return raw


def read(path, dataset):
    """One dataset of an HDF5 file as a numpy array (native byte order)."""
    a = File(path)[dataset].read()
    return a.astype(a.dtype.newbyteorder("="), copy=False)


def list_datasets(path):
    f = File(path)
    out = {}

    def walk(g, prefix):
        for k in g.keys():
            o = g[k]
            name = prefix + k
            if isinstance(o, _Group):
                walk(o, name + "/")
            else:
                out[name] = (o.shape, o.dtype)
    walk(f, "/")
    return out


# ---------------------------------------------------------------------------------------------
# writer: flat root group of contiguous numeric arrays
# ---------------------------------------------------------------------------------------------
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        elif dt.itemsize == 8:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        else:
            raise H5FormatError("cannot write float%d" % (dt.itemsize * 8))
        return struct.pack("<BBBBI", 0x10 | 1, bits[0], bits[1], bits[2], dt.itemsize) + prop
    raise H5FormatError("cannot write dtype %s" % dt)


def _msg(mtype, body, flags=0):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def write(path, arrays):
    """Write ``{name: ndarray}`` as contiguous datasets of the root group (superblock v0, v1
    object headers, one symbol-table node)."""
    names = sorted(arrays)
    if len(names) > 2 * 16:
        # one leaf node holds 2K entries; grow K instead of building a deeper tree
        leaf_k = (len(names) + 1) // 2
    else:
        leaf_k = 16
    arrs = {k: np.ascontiguousarray(arrays[k]) for k in names}
    for k in names:
        if arrs[k].dtype.byteorder == ">":
            arrs[k] = arrs[k].astype(arrs[k].dtype.newbyteorder("<"))

    buf = bytearray()

    def align():
        buf.extend(b"\0" * (-len(buf) % 8))

    def ohdr(msgs):
        body = b"".join(msgs)
        return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body

    SB = 96   # superblock v0 (56) + root symbol table entry (40)
    buf.extend(b"\0" * SB)
    # local heap data: names
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for k in names:
        name_off[k] = len(heap_data)
        heap_data.extend(k.encode("utf8") + b"\0")
        heap_data.extend(b"\0" * (-len(heap_data) % 8))
    heap_data.extend(b"\0" * 16)   # room for a free block
    free_off = len(heap_data) - 16
    struct.pack_into("<QQ", heap_data, free_off, 1, 16)    # next-free = 1 (none), size

    # datasets: object header then raw data
    ohdr_addr = {}
    for k in names:
        a = arrs[k]
        align()
        space = struct.pack("<BBBBI", 1, a.ndim, 0, 0, 0) + struct.pack("<%dQ" % a.ndim,
                                                                       *a.shape)
        msgs = [_msg(0x01, space), _msg(0x03, _dtype_msg(a.dtype), 1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0))]
        hdr_len = 16 + sum(len(m) for m in msgs) + 8 + 24
        data_addr = len(buf) + hdr_len
        data_addr += -data_addr % 8
        layout = struct.pack("<BBQQ", 3, 1, data_addr if a.nbytes else _UNDEF, a.nbytes)
        msgs.append(_msg(0x08, layout))
        ohdr_addr[k] = len(buf)
        buf.extend(ohdr(msgs))
        align()
        assert len(buf) <= data_addr
        buf.extend(b"\0" * (data_addr - len(buf)))
        buf.extend(a.tobytes())

    # heap
    align()
    heap_addr = len(buf)
    heap_data_addr = heap_addr + 32
    buf.extend(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), free_off,
                                     heap_data_addr))
    buf.extend(heap_data)
    # symbol table node
    align()
    snod_addr = len(buf)
    buf.extend(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
    for k in names:          # names are sorted, as the B-tree requires
        buf.extend(struct.pack("<QQII", name_off[k], ohdr_addr[k], 0, 0) + b"\0" * 16)
    buf.extend(b"\0" * (40 * (2 * leaf_k - len(names))))
    # B-tree root (one child)
    align()
    btree_addr = len(buf)
    buf.extend(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF))
    internal_k = 16
    keys = struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]] if names else 0)
    buf.extend(keys + b"\0" * ((2 * internal_k + 1) * 8 + 2 * internal_k * 8 - len(keys)))
    # root group object header
    align()
    root_addr = len(buf)
    buf.extend(ohdr([_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))]))
    align()
    eof = len(buf)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == SB, len(sb)
    buf[:SB] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(buf))
