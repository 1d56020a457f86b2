"""A small HDF5 reader / writer in NumPy, for the two HDF5 uses of the reference (h5py is not installable here):

* Keras checkpoints: ``ModelCheckpoint('...{epoch:02d}-{val_loss:.4f}.h5')`` / ``model.save('fov_s2s_tanh.h5')`` /
  ``model.load_weights('....h5')`` (mycode/FoV_seq2seq.py:108, mycode/FoV_seq2seq_mu_var.py:239,256,
  mycode/all_share_lstm.py:48) - ``layer_names`` / ``weight_names`` attributes, one group per layer, float32 datasets;
* cached sample tensors: ``save2hdf5`` / ``load_h5`` (mycode/utility.py:868-880) - one dataset per key.

It covers the subset of the HDF5 File Format Specification (v1.1 structures, the ones libhdf5's default
"earliest" format writes and h5py 2.x / Keras 2.2 therefore produce): superblock v0 / v1 (v2 / v3 read only),
version-1 object headers with continuation blocks (version-2 headers with compact link messages read only),
symbol-table groups (v1 B-tree + SNOD + local heap), contiguous / compact / chunked datasets (v1 chunk B-tree,
deflate / shuffle / fletcher32 filters), fixed-point / IEEE float / fixed- and variable-length string datatypes,
attributes v1-v3.  Not covered: compound / enum / reference / array types, dense link storage (fractal heaps),
layout v4 chunk indices, external storage, virtual datasets.

The reader is pinned to a file libhdf5 wrote (scipy's ``testhdf5_7.4_GLNX86.mat``: 512-byte user block, superblock
v0, v1 headers, layout v2 - tests/test_h5lite.py); the writer emits the same structures and is checked by reading
its files back.  Host-side I/O only: nothing here touches the GPU.
"""
import mmap
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


# ------------------------------------------------------------------------------------------------------------------ #
# reading
# ------------------------------------------------------------------------------------------------------------------ #
class _Type:
    """Decoded datatype message: a NumPy dtype, or a variable-length string marker."""
    __slots__ = ("dtype", "vlen_str", "size", "utf8")

    def __init__(self, dtype, size, vlen_str=False, utf8=False):
        self.dtype, self.size, self.vlen_str, self.utf8 = dtype, size, vlen_str, utf8


def _parse_datatype(b, off=0):
    cls_ver, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, off)
    cls = cls_ver & 0x0F
    if cls == 0:                                   # fixed point
        order = ">" if b0 & 1 else "<"
        kind = "i" if b0 & 8 else "u"
        return _Type(np.dtype("%s%s%d" % (order, kind, size)), size)
    if cls == 1:                                   # floating point; IEEE layouts only
        order = ">" if b0 & 1 else "<"
        prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<xxHBBBBI", b, off + 8)
        ieee = {2: (10, 5, 0, 10, 15), 4: (23, 8, 0, 23, 127), 8: (52, 11, 0, 52, 1023)}
        if size not in ieee or (eloc, esize, mloc, msize, bias) != ieee[size] or prec != 8 * size:
            raise H5Error("non-IEEE floating-point type")
        return _Type(np.dtype("%sf%d" % (order, size)), size)
    if cls == 3:                                   # fixed-length string
        return _Type(np.dtype("S%d" % size), size, utf8=bool((b0 >> 4) & 1))
    if cls == 9:                                   # variable length: strings only
        if (b0 & 0x0F) != 1:
            raise H5Error("variable-length sequences are not supported")
        return _Type(None, size, vlen_str=True, utf8=bool(b1 & 1))
    raise H5Error("datatype class %d is not supported" % cls)


def _parse_dataspace(b, off=0):
    ver, rank, flags = struct.unpack_from("<BBB", b, off)
    if ver == 1:
        p = off + 8
    elif ver == 2:
        if b[off + 3] == 2:
            return None                            # null dataspace
        p = off + 4
    else:
        raise H5Error("dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, b, p)) if rank else ()


class _Node:
    """An object header's messages: [(type, flags, bytes)]."""

    def __init__(self, f, addr):
        self.f, self.addr, self.msgs = f, addr, []
        buf = f.buf
        a = f.base + addr
        if buf[a:a + 4] == b"OHDR":
            self._read_v2(a)
        else:
            self._read_v1(a)

    def _read_v1(self, a):
        buf = self.f.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", buf, a)
        if ver != 1:
            raise H5Error("object header version %d at %d" % (ver, a))
        blocks = [(a + 16, hsize)]
        while blocks and len(self.msgs) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(self.msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, p)
                body = bytes(buf[p + 8:p + 8 + msize])
                p += 8 + msize
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", body)
                    blocks.append((self.f.base + off, ln))
                self.msgs.append((mtype, mflags, body))

    def _read_v2(self, a):
        buf = self.f.buf
        ver, flags = struct.unpack_from("<BB", buf, a + 4)
        if ver != 2:
            raise H5Error("object header version %d" % ver)
        p = a + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        w = 1 << (flags & 3)
        size0 = int.from_bytes(buf[p:p + w], "little")
        p += w
        blocks = [(p, size0)]
        track = bool(flags & 4)
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = struct.unpack_from("<BHB", buf, p)
                p += 4 + (2 if track else 0)
                body = bytes(buf[p:p + msize])
                p += msize
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", body)
                    blocks.append((self.f.base + off + 4, ln - 8))    # skip "OCHK", drop the checksum
                self.msgs.append((mtype, mflags, body))

    def find(self, mtype):
        return [m[2] for m in self.msgs if m[0] == mtype]


class _Attrs(dict):
    pass


class _Object:
    def __init__(self, f, addr, name):
        self.file, self.name = f, name
        self._node = _Node(f, addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = _Attrs()
            for body in self._node.find(0x0C):
                k, v = self.file._parse_attribute(body)
                self._attrs[k] = v
        return self._attrs


class Group(_Object):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        links = {}
        f = self.file
        st = self._node.find(0x11)
        if st:
            btree, heap = struct.unpack_from("<QQ", st[0])
            hbuf = f._local_heap(heap)
            f._walk_group_btree(btree, hbuf, links)
        for body in self._node.find(0x06):                       # new-style compact links
            nm, addr = f._parse_link(body)
            if addr is not None:
                links[nm] = addr
        for body in self._node.find(0x02):
            ver, fl = struct.unpack_from("<BB", body)
            p = 2 + (8 if fl & 1 else 0)
            fheap = struct.unpack_from("<Q", body, p)[0]
            if fheap != UNDEF:
                raise H5Error("dense link storage (fractal heap) is not supported: %s" % self.name)
        self._links = links

    def keys(self):
        self._load()
        return list(self._links.keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, path):
        obj = self
        if path.startswith("/"):
            obj = self.file.root
        for part in [p for p in path.split("/") if p]:
            if not isinstance(obj, Group):
                raise KeyError(path)
            obj._load()
            if part not in obj._links:
                raise KeyError(path)
            obj = self.file._open(obj._links[part], (obj.name.rstrip("/") + "/" + part))
        return obj

    def get(self, path, default=None):
        try:
            return self[path]
        except KeyError:
            return default

    def visit_datasets(self, prefix=""):
        """[(path, Dataset)] below this group, depth first in stored (name) order."""
        out = []
        for k in self.keys():
            o = self[k]
            if isinstance(o, Group):
                out += o.visit_datasets(prefix + k + "/")
            else:
                out.append((prefix + k, o))
        return out


class Dataset(_Object):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        n = self._node
        self._type = _parse_datatype(n.find(0x03)[0])
        self.shape = _parse_dataspace(n.find(0x01)[0])
        self.dtype = self._type.dtype if not self._type.vlen_str else np.dtype(object)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._read())                                 # a scalar dataset decodes to a NumPy scalar
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        a = self._read()
        return a if key == () or key is Ellipsis else a[key]

    @property
    def value(self):
        return self._read()

    def _read(self):
        f, t = self.file, self._type
        if self.shape is None:
            return np.zeros((0,), t.dtype or object)
        count = int(np.prod(self.shape, dtype=np.int64))
        lay = self._node.find(0x08)[0]
        ver = lay[0]
        esz = t.size
        raw = None
        if ver in (1, 2):
            ndim, cls = lay[1], lay[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", lay, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % ndim, lay, p)
            p += 4 * ndim
            if cls == 0:
                n = struct.unpack_from("<I", lay, p)[0]
                raw = lay[p + 4:p + 4 + n]
            elif cls == 1:
                raw = None if addr == UNDEF else f.buf[f.base + addr:f.base + addr + count * esz]
            else:
                raw = self._read_chunked(addr, dims[:-1], esz)
        elif ver == 3:
            cls = lay[1]
            if cls == 0:
                n = struct.unpack_from("<H", lay, 2)[0]
                raw = lay[4:4 + n]
            elif cls == 1:
                addr, _ = struct.unpack_from("<QQ", lay, 2)
                raw = None if addr == UNDEF else f.buf[f.base + addr:f.base + addr + count * esz]
            elif cls == 2:
                ndim = lay[2]
                addr = struct.unpack_from("<Q", lay, 3)[0]
                dims = struct.unpack_from("<%dI" % ndim, lay, 11)
                raw = self._read_chunked(addr, dims[:-1], esz)
            else:
                raise H5Error("layout class %d" % cls)
        else:
            raise H5Error("data layout version %d is not supported (file written with libver='latest'?)" % ver)
        if raw is None:                                              # never written: fill value (zeros)
            raw = bytes(count * esz)
        return f._decode(raw, t, self.shape)

    def _read_chunked(self, btree, chunk, esz):
        f = self.file
        shape = self.shape
        rank = len(shape)
        filters = []
        for body in self._node.find(0x0B):
            filters = _parse_filters(body)
        out = np.zeros(shape, np.dtype("V%d" % esz))
        if btree == UNDEF:
            return out.tobytes()
        csize = int(np.prod(chunk, dtype=np.int64)) * esz
        for nbytes, mask, offs, addr in f._walk_chunk_btree(btree, rank):
            data = bytes(f.buf[f.base + addr:f.base + addr + nbytes])
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    data = zlib.decompress(data)
                elif fid == 2:
                    w = cd[0] if cd else esz
                    n = len(data) // w
                    data = np.frombuffer(data[:n * w], np.uint8).reshape(w, n).T.tobytes() + data[n * w:]
                elif fid == 3:
                    data = data[:-4]
                else:
                    raise H5Error("filter %d is not supported" % fid)
            if len(data) != csize:
                raise H5Error("chunk of %d bytes, expected %d" % (len(data), csize))
            c = np.frombuffer(data, np.dtype("V%d" % esz)).reshape(chunk)
            sl_out, sl_in = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + chunk[d], shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = c[tuple(sl_in)]
        return out.tobytes()


def _parse_filters(body):
    ver, n = body[0], body[1]
    p = 8 if ver == 1 else 2
    out = []
    for _ in range(n):
        fid = struct.unpack_from("<H", body, p)[0]
        p += 2
        nlen = 0
        if ver == 1 or fid >= 256:
            nlen = struct.unpack_from("<H", body, p)[0]
            p += 2
        _, ncd = struct.unpack_from("<HH", body, p)
        p += 4
        if ver == 1:
            nlen = (nlen + 7) & ~7
        p += nlen
        cd = struct.unpack_from("<%dI" % ncd, body, p)
        p += 4 * ncd
        if ver == 1 and ncd % 2:
            p += 4
        out.append((fid, cd))
    return out


class File(Group):
    """Read-only view of an HDF5 file (``h5lite.File(path)`` ~ ``h5py.File(path, 'r')``)."""

    def __init__(self, path_or_bytes):
        self._mm = self._fh = None
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = memoryview(bytes(path_or_bytes))
        else:
            # memory-mapped: the sample caches of the reference run to gigabytes; only the bytes a dataset read
            # touches are paged in
            self._fh = open(path_or_bytes, "rb")
            try:
                self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
            except ValueError:
                self._fh.close()
                raise H5Error("not an HDF5 file (empty): %s" % path_or_bytes)
            self.buf = memoryview(self._mm)
        buf = self.buf
        a = 0
        while True:                                                  # user block: 0, 512, 1024, ...
            if buf[a:a + 8] == SIG:
                break
            a = 512 if a == 0 else a * 2
            if a + 8 > len(buf):
                raise H5Error("not an HDF5 file (no signature)")
        ver = buf[a + 8]
        self._heaps, self._gcols = {}, {}
        if ver in (0, 1):
            so, sl = buf[a + 13], buf[a + 14]
            if (so, sl) != (8, 8):
                raise H5Error("only 8-byte offsets / lengths are supported")
            p = a + 24 + (4 if ver == 1 else 0)
            self.base = struct.unpack_from("<Q", buf, p)[0]
            root = struct.unpack_from("<Q", buf, p + 32 + 8)[0]
        elif ver in (2, 3):
            if (buf[a + 9], buf[a + 10]) != (8, 8):
                raise H5Error("only 8-byte offsets / lengths are supported")
            self.base = struct.unpack_from("<Q", buf, a + 12)[0]
            root = struct.unpack_from("<Q", buf, a + 12 + 24)[0]
        else:
            raise H5Error("superblock version %d" % ver)
        self.file = self
        self.root = self
        super().__init__(self, root, "/")

    def close(self):
        if self._mm is not None:
            try:
                self.buf.release()
                self._mm.close()
            except BufferError:                                      # a view of the mapping is still alive somewhere:
                pass                                                 # the mapping goes with its last reference
            self._fh.close()
            self._mm = self._fh = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    # -- helpers ---------------------------------------------------------------------------------------------------- #
    def _open(self, addr, name):
        node = _Node(self, addr)
        if node.find(0x08) or node.find(0x03):
            return Dataset(self, addr, name)
        return Group(self, addr, name)

    def _local_heap(self, addr):
        if addr not in self._heaps:
            a = self.base + addr
            if self.buf[a:a + 4] != b"HEAP":
                raise H5Error("bad local heap at %d" % a)
            size, _, daddr = struct.unpack_from("<QQQ", self.buf, a + 8)
            self._heaps[addr] = bytes(self.buf[self.base + daddr:self.base + daddr + size])
        return self._heaps[addr]

    def _walk_group_btree(self, addr, heap, links):
        a = self.base + addr
        buf = self.buf
        if buf[a:a + 4] == b"SNOD":
            n = struct.unpack_from("<H", buf, a + 6)[0]
            for i in range(n):
                noff, oaddr = struct.unpack_from("<QQ", buf, a + 8 + 40 * i)
                end = heap.index(b"\0", noff)
                links[heap[noff:end].decode("utf8")] = oaddr
            return
        if buf[a:a + 4] != b"TREE":
            raise H5Error("bad group B-tree node at %d" % a)
        ntype, level, used = struct.unpack_from("<BBH", buf, a + 4)
        if ntype != 0:
            raise H5Error("group B-tree node of type %d" % ntype)
        p = a + 24
        for i in range(used):
            child = struct.unpack_from("<Q", buf, p + 8 + 16 * i)[0]
            self._walk_group_btree(child, heap, links)

    def _walk_chunk_btree(self, addr, rank):
        a = self.base + addr
        buf = self.buf
        if buf[a:a + 4] != b"TREE":
            raise H5Error("bad chunk B-tree node at %d" % a)
        ntype, level, used = struct.unpack_from("<BBH", buf, a + 4)
        if ntype != 1:
            raise H5Error("chunk B-tree node of type %d" % ntype)
        ksz = 8 + 8 * (rank + 1)
        p = a + 24
        for i in range(used):
            nbytes, mask = struct.unpack_from("<II", buf, p)
            offs = struct.unpack_from("<%dQ" % rank, buf, p + 8)
            child = struct.unpack_from("<Q", buf, p + ksz)[0]
            p += ksz + 8
            if level == 0:
                yield nbytes, mask, offs, child
            else:
                yield from self._walk_chunk_btree(child, rank)

    def _global_heap_object(self, addr, index):
        if addr not in self._gcols:
            a = self.base + addr
            buf = self.buf
            if buf[a:a + 4] != b"GCOL":
                raise H5Error("bad global heap collection at %d" % a)
            total = struct.unpack_from("<Q", buf, a + 8)[0]
            objs = {}
            p = a + 16
            while p + 16 <= a + total:
                idx, _, _, size = struct.unpack_from("<HHIQ", buf, p)
                if idx == 0:
                    break
                objs[idx] = bytes(buf[p + 16:p + 16 + size])
                p += 16 + ((size + 7) & ~7)
            self._gcols[addr] = objs
        return self._gcols[addr][index]

    def _decode(self, raw, t, shape):
        count = int(np.prod(shape, dtype=np.int64)) if shape is not None else 0
        if t.vlen_str:
            vals = []
            for i in range(count):
                ln, gaddr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                s = self._global_heap_object(gaddr, idx)[:ln] if ln else b""
                vals.append(s.decode("utf8") if t.utf8 else s)
            if shape == ():
                return vals[0]
            a = np.empty(count, object)
            a[:] = vals
            return a.reshape(shape)
        a = np.frombuffer(bytes(raw[:count * t.size]), t.dtype, count).reshape(shape)
        if shape == ():
            return a[()]
        return a.copy()

    def _parse_attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        p = 8
        if ver == 3:
            p += 1
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = bytes(body[p:p + nsz]).split(b"\0")[0].decode("utf8")
        p += pad(nsz)
        t = _parse_datatype(body, p)
        p += pad(tsz)
        shape = _parse_dataspace(body, p)
        p += pad(ssz)
        if shape is None:
            return name, None
        return name, self._decode(body[p:], t, shape)

    def _parse_link(self, body):
        ver, fl = struct.unpack_from("<BB", body)
        p = 2
        ltype = 0
        if fl & 8:
            ltype = body[p]
            p += 1
        if fl & 4:
            p += 8
        if fl & 16:
            p += 1
        w = 1 << (fl & 3)
        n = int.from_bytes(body[p:p + w], "little")
        p += w
        name = bytes(body[p:p + n]).decode("utf8")
        p += n
        if ltype != 0:
            return name, None                                        # soft / external links are skipped
        return name, struct.unpack_from("<Q", body, p)[0]


# ------------------------------------------------------------------------------------------------------------------ #
# writing
# ------------------------------------------------------------------------------------------------------------------ #
def _pad8(b):
    return b + bytes(-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (2, 4, 8):
        sz = dt.itemsize
        eloc, esize, msize, bias = {2: (10, 5, 10, 15), 4: (23, 8, 23, 127), 8: (52, 11, 52, 1023)}[sz]
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * sz - 1, 0, sz, 0, 8 * sz, eloc, esize, 0, msize, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 8 if dt.kind == "i" else 0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "b":
        return struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 1, 0, 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))     # null-padded ASCII, as h5py's 'S'
    raise H5Error("cannot store dtype %s" % dt)


def _space_msg(shape):
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _as_storable(value):
    """NumPy array in a storable little-endian dtype (str / bytes / lists of them -> fixed-length 'S')."""
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, (list, tuple)) and value and all(isinstance(v, (str, bytes)) for v in value):
        value = [v.encode("utf8") if isinstance(v, str) else v for v in value]
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    if a.dtype.kind == "S" and a.dtype.itemsize == 0:
        a = a.astype("S1")
    if a.dtype.kind == "b":
        a = a.astype(np.uint8)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return np.ascontiguousarray(a).reshape(a.shape)                 # ascontiguousarray makes 0-d arrays 1-d


class _WNode:
    def __init__(self):
        self.attrs = {}


class _WGroup(_WNode):
    def __init__(self):
        super().__init__()
        self.children = {}


class _WDataset(_WNode):
    def __init__(self, data, chunks=None, compression=None, shuffle=False):
        super().__init__()
        self.data = _as_storable(data)
        self.chunks, self.compression, self.shuffle = chunks, compression, shuffle


class _AttrProxy:
    def __init__(self, node):
        self._n = node

    def __setitem__(self, k, v):
        self._n.attrs[k] = v

    def __getitem__(self, k):
        return self._n.attrs[k]


class _GroupWriter:
    def __init__(self, node):
        self._node = node

    @property
    def attrs(self):
        return _AttrProxy(self._node)

    def _walk(self, path, create=True):
        parts = [p for p in path.split("/") if p]
        node = self._node
        for part in parts[:-1]:
            if part not in node.children:
                node.children[part] = _WGroup()
            node = node.children[part]
            if not isinstance(node, _WGroup):
                raise H5Error("%s is a dataset" % part)
        return node, parts[-1]

    def create_group(self, path):
        node, last = self._walk(path)
        if last in node.children:
            raise H5Error("name already exists: %s" % path)
        node.children[last] = _WGroup()
        return _GroupWriter(node.children[last])

    def require_group(self, path):
        node, last = self._walk(path)
        if last not in node.children:
            node.children[last] = _WGroup()
        return _GroupWriter(node.children[last])

    def create_dataset(self, path, data=None, shape=None, dtype=None, chunks=None, compression=None, shuffle=False):
        if data is None:
            data = np.zeros(shape, dtype or np.float32)
        elif dtype is not None:
            data = np.asarray(data, dtype)
        node, last = self._walk(path)
        if last in node.children:
            raise H5Error("name already exists: %s" % path)
        d = _WDataset(data, chunks, compression, shuffle)
        node.children[last] = d
        return _GroupWriter(d)


class Writer(_GroupWriter):
    """Build a tree (``create_group`` / ``create_dataset`` / ``.attrs[...] = ...``), then ``close()`` writes the file.

    Layout: superblock v0 at 0, one version-1 object header per object, one symbol-table B-tree node + ONE symbol
    node + local heap per group (the superblock's group-leaf K is raised to hold the largest group in a single
    symbol node, which libhdf5 honours), contiguous (or chunked + deflate) raw data.
    """

    def __init__(self, path):
        super().__init__(_WGroup())
        self.path = path
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close()
        return False

    def close(self):
        if self._closed:
            return
        self._closed = True
        self._leaf_k = max(4, (self._max_group(self._node) + 1) // 2)
        self._int_k = 16
        with open(self.path, "wb") as fh:                            # streamed: one dataset in memory at a time
            self._fh, self._pos = fh, 96
            fh.write(bytes(96))
            root_hdr, btree, heap = self._emit_group(self._node)
            sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self._leaf_k, self._int_k, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, self._pos, UNDEF)
            sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
            assert len(sb) == 96
            fh.seek(0)
            fh.write(sb)
        self._fh = None

    # -- emitters ---------------------------------------------------------------------------------------------------- #
    def _max_group(self, g):
        m = len(g.children)
        for c in g.children.values():
            if isinstance(c, _WGroup):
                m = max(m, self._max_group(c))
        return m

    def _alloc(self, blob):
        pad = -self._pos % 8
        if pad:
            self._fh.write(bytes(pad))
        addr = self._pos + pad
        self._fh.write(blob)
        self._pos = addr + len(blob)
        return addr

    def _attr_msgs(self, node):
        msgs = []
        for k, v in node.attrs.items():
            a = _as_storable(v)
            name = k.encode("utf8") + b"\0"
            dt, sp = _dtype_msg(a.dtype), _space_msg(a.shape)
            body = struct.pack("<BBHHH", 1, 0, len(name), len(dt), len(sp)) + _pad8(name) + _pad8(dt) + _pad8(sp)
            body += a.tobytes()
            if len(body) > 64000:
                raise H5Error("attribute %r too large for an object header (%d bytes); Keras chunks such "
                              "attributes as name0, name1, ..." % (k, len(body)))
            msgs.append((0x0C, 0, body))
        return msgs

    def _emit_header(self, msgs):
        blob = b""
        for mtype, flags, body in msgs:
            body = _pad8(body)
            blob += struct.pack("<HHB3x", mtype, len(body), flags) + body
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(blob))
        return self._alloc(hdr + blob)

    def _emit_group(self, g):
        entries = []
        for name in sorted(g.children, key=lambda s: s.encode("utf8")):
            c = g.children[name]
            if isinstance(c, _WGroup):
                addr, bt, hp = self._emit_group(c)
                entries.append((name, addr, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self._emit_dataset(c), 0, bytes(16)))
        # local heap: "" at offset 0, then the names, one free block at the end
        data = bytearray(8)
        offs = []
        for name, *_ in entries:
            offs.append(len(data))
            data += _pad8(name.encode("utf8") + b"\0")
        free_off = len(data)
        data += struct.pack("<QQ", 1, 16)
        heap_data = self._alloc(bytes(data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(data), free_off, heap_data))
        # one symbol node
        snod = bytearray(8 + 2 * self._leaf_k * 40)
        snod[:8] = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
        for i, ((name, addr, cache, scratch), off) in enumerate(zip(entries, offs)):
            struct.pack_into("<QQII", snod, 8 + 40 * i, off, addr, cache, 0)
            snod[8 + 40 * i + 24:8 + 40 * i + 40] = scratch
        snod_addr = self._alloc(bytes(snod))
        # one B-tree node (level 0) with that single child; an empty group has no entries
        node = bytearray(24 + (2 * self._int_k + 1) * 8 + 2 * self._int_k * 8)
        node[:24] = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF)
        if entries:
            struct.pack_into("<QQQ", node, 24, 0, snod_addr, offs[-1])
        btree = self._alloc(bytes(node))
        msgs = [(0x11, 0, struct.pack("<QQ", btree, heap))] + self._attr_msgs(g)
        return self._emit_header(msgs), btree, heap

    def _emit_dataset(self, d):
        a = d.data
        msgs = [(0x01, 0, _space_msg(a.shape)), (0x03, 1, _dtype_msg(a.dtype)),
                (0x05, 0, struct.pack("<BBBB", 2, 1 if not d.chunks else 3, 0, 0))]
        esz = a.dtype.itemsize
        if d.chunks and a.ndim and a.size:
            chunk = tuple(int(min(c, s)) for c, s in zip(d.chunks, a.shape))
            filt = []
            if d.shuffle:
                filt.append((2, b"shuffle\0", (esz,)))
            if d.compression:
                level = d.compression if isinstance(d.compression, int) else 4
                filt.append((1, b"deflate\0", (level,)))
            if filt:
                body = struct.pack("<BB6x", 1, len(filt))
                for fid, nm, cd in filt:
                    body += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + _pad8(nm)
                    body += b"".join(struct.pack("<I", c) for c in cd) + (bytes(4) if len(cd) % 2 else b"")
                msgs.append((0x0B, 0, body))
            recs = []
            grid = [range(0, s, c) for s, c in zip(a.shape, chunk)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                lo = [grid[i][o] for i, o in enumerate(offs)]
                block = np.zeros(chunk, a.dtype)
                src = a[tuple(slice(l, l + c) for l, c in zip(lo, chunk))]
                block[tuple(slice(0, n) for n in src.shape)] = src
                raw = block.tobytes()
                if d.shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, esz).T.tobytes()
                if d.compression:
                    raw = zlib.compress(raw, level)
                recs.append((len(raw), tuple(lo), self._alloc(raw)))
            root = self._emit_chunk_tree(recs, a.shape, chunk)
            lay = struct.pack("<BBBQ", 3, 2, a.ndim + 1, root) + b"".join(struct.pack("<I", c) for c in chunk + (esz,))
        else:
            addr = self._alloc(a.reshape(-1).view(np.uint8).data if a.dtype.kind != "S" else a.tobytes()) if a.size else UNDEF
            lay = struct.pack("<BBQQ", 3, 1, addr, a.size * esz)
        msgs.append((0x08, 0, lay))
        return self._emit_header(msgs + self._attr_msgs(d))

    def _emit_chunk_tree(self, recs, shape, chunk, k=32):
        """v1 B-tree over chunk records [(nbytes, offsets, addr)] in row-major order; returns the root address."""
        rank = len(shape)
        ksz = 8 + 8 * (rank + 1)
        past = tuple(((s + c - 1) // c) * c for s, c in zip(shape, chunk))       # key after the last chunk

        def key(nbytes, offs):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + bytes(8)

        level = 0
        items = [(n, o, ad) for n, o, ad in recs]                                 # (first nbytes, first offs, child)
        while True:
            nodes = []
            groups = [items[i:i + 2 * k] for i in range(0, len(items), 2 * k)]
            addrs = []
            base = self._pos + (-self._pos % 8)
            nsize = 24 + (2 * k + 1) * ksz + 2 * k * 8
            for gi in range(len(groups)):
                addrs.append(base + gi * nsize)
            for gi, grp in enumerate(groups):
                node = bytearray(nsize)
                left = addrs[gi - 1] if gi else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                node[:24] = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
                p = 24
                for n, o, ad in grp:
                    node[p:p + ksz] = key(n, o)
                    struct.pack_into("<Q", node, p + ksz, ad)
                    p += ksz + 8
                if gi + 1 < len(groups):
                    nxt = groups[gi + 1][0]
                    node[p:p + ksz] = key(nxt[0], nxt[1])
                else:
                    node[p:p + ksz] = key(0, past)
                got = self._alloc(bytes(node))
                assert got == addrs[gi]
                nodes.append((grp[0][0], grp[0][1], got))
            if len(nodes) == 1:
                return nodes[0][2]
            items = nodes
            level += 1


# ------------------------------------------------------------------------------------------------------------------ #
# the reference's two helpers (mycode/utility.py:868-880)
# ------------------------------------------------------------------------------------------------------------------ #
def load_h5(path_name, key):
    """``np.array(h5py.File(path_name, 'r').get(key))`` (mycode/utility.py:874-880)."""
    with File(path_name) as f:
        d = f.get(key)
        if d is None:
            return np.array(None)
        return np.array(d)


def save2hdf5(path_name, key, data_to_store):
    """``h5py.File(path_name, 'a').create_dataset(key, data=...)`` (mycode/utility.py:868-871): the existing datasets
    are read back and the file is rewritten with the new key added (append mode without in-place allocation)."""
    import os
    existing = []
    if os.path.exists(path_name):
        with File(path_name) as f:
            existing = [(p, np.array(d)) for p, d in f.visit_datasets()]
        if any(p == key.strip("/") for p, _ in existing):
            raise H5Error("unable to create dataset (name already exists): %s" % key)
    with Writer(path_name) as w:
        for p, a in existing:
            w.create_dataset(p, data=a)
        w.create_dataset(key, data=np.asarray(data_to_store))


# ------------------------------------------------------------------------------------------------------------------ #
# Keras weight files
# ------------------------------------------------------------------------------------------------------------------ #
def _attr_list(attrs, name):
    """Keras' ``load_attributes_from_hdf5_group``: ``name`` or its chunks ``name0, name1, ...``."""
    if name in attrs:
        vals = list(np.atleast_1d(attrs[name]))
    else:
        vals, i = [], 0
        while "%s%d" % (name, i) in attrs:
            vals += list(np.atleast_1d(attrs["%s%d" % (name, i)]))
            i += 1
    return [v.decode("utf8") if isinstance(v, bytes) else str(v) for v in vals]


def read_keras_weights(path):
    """[(layer_name, [(weight_name, array), ...]), ...] in the file's ``layer_names`` / ``weight_names`` order, from a
    ``save_weights`` file or the ``model_weights`` group of a ``model.save`` file (Keras 2.2 ``saving.py`` layout)."""
    with File(path) as f:
        g = f["model_weights"] if "layer_names" not in f.attrs and "model_weights" in f else f
        if "layer_names" not in g.attrs and "layer_names0" not in g.attrs:
            raise H5Error("%s holds no Keras layer_names attribute" % path)
        out = []
        for lname in _attr_list(g.attrs, "layer_names"):
            lg = g[lname]
            ws = [(wn, np.array(lg[wn])) for wn in _attr_list(lg.attrs, "weight_names")]
            out.append((lname, ws))
        return out


def _put_keras_layers(g, layers):
    g.attrs["layer_names"] = np.array([l.encode("utf8") for l, _ in layers]) if layers else np.zeros((0,), "S1")
    for lname, ws in layers:
        lg = g.require_group(lname)
        lg.attrs["weight_names"] = (np.array([n.encode("utf8") for n, _ in ws]) if ws else np.zeros((0,), "S1"))
        for wn, arr in ws:
            lg.create_dataset(wn, data=np.asarray(arr))


def write_keras_weights(path, layers, keras_version="2.2.4", backend="tensorflow"):
    """``layers``: [(layer_name, [(weight_name, array), ...]), ...] -> a file Keras' ``load_weights`` layout describes:
    root attributes ``layer_names`` / ``backend`` / ``keras_version``, one group per layer with ``weight_names`` and
    one dataset per weight at ``/<layer>/<weight_name>`` (weight names carry their own ``layer/`` prefix)."""
    with Writer(path) as w:
        _put_keras_layers(w, layers)
        w.attrs["backend"] = backend
        w.attrs["keras_version"] = keras_version


def write_keras_model(path, layers, optimizer_weights=None, training_config=None, model_config=None,
                      keras_version="2.2.4", backend="tensorflow"):
    """Keras 2.2 ``model.save`` layout (``saving._serialize_model``): the layers under ``/model_weights``, the
    optimizer's ``[(name, array), ...]`` under ``/optimizer_weights`` (attribute ``weight_names``), ``model_config`` /
    ``training_config`` JSON strings as root attributes."""
    import json
    with Writer(path) as w:
        w.attrs["keras_version"] = keras_version
        w.attrs["backend"] = backend
        w.attrs["model_config"] = json.dumps(model_config or {})
        if training_config is not None:
            w.attrs["training_config"] = json.dumps(training_config)
        mg = w.create_group("model_weights")
        _put_keras_layers(mg, layers)
        mg.attrs["backend"] = backend
        mg.attrs["keras_version"] = keras_version
        if optimizer_weights:
            og = w.create_group("optimizer_weights")
            og.attrs["weight_names"] = np.array([n.encode("utf8") for n, _ in optimizer_weights])
            for n, a in optimizer_weights:
                og.create_dataset(n, data=np.asarray(a))


def read_keras_optimizer(path):
    """(training_config dict or None, [(name, array), ...]) of a ``model.save`` file; ([], None) parts when absent."""
    import json
    with File(path) as f:
        cfg = f.attrs.get("training_config")
        if cfg is not None:
            cfg = json.loads(cfg.decode("utf8") if isinstance(cfg, bytes) else cfg)
        og = f.get("optimizer_weights")
        ws = []
        if og is not None:
            ws = [(n, np.array(og[n])) for n in _attr_list(og.attrs, "weight_names")]
        return cfg, ws


# ------------------------------------------------------------------------------------------------------------------ #
# inspection (``python -m longterm360fov_b200.h5lite checkpoint.h5`` ~ ``h5ls -r`` with attributes)
# ------------------------------------------------------------------------------------------------------------------ #
def describe(path):
    """One line per group / dataset / attribute of the file, depth first: what a Keras checkpoint or a sample cache
    holds, without h5py."""
    lines = []

    def attrs_of(obj, indent):
        for k, v in obj.attrs.items():
            a = np.asarray(v)
            txt = repr(v.decode("utf8", "replace") if isinstance(v, bytes) else v)
            if a.ndim:
                txt = "%s%s" % (a.dtype, list(a.shape))
            elif len(txt) > 60:
                txt = txt[:57] + "..."
            lines.append("%s@%s = %s" % (indent, k, txt))

    def walk(g, indent):
        attrs_of(g, indent)
        for k in g.keys():
            o = g[k]
            if isinstance(o, Group):
                lines.append("%s%s/" % (indent, k))
                walk(o, indent + "  ")
            else:
                lines.append("%s%s  %s%s" % (indent, k, o.dtype, list(o.shape) if o.shape is not None else "null"))
                attrs_of(o, indent + "  ")

    with File(path) as f:
        walk(f, "")
    return lines


if __name__ == "__main__":
    import sys
    for _p in sys.argv[1:]:
        print(_p)
        print("\n".join("  " + ln for ln in describe(_p)))
