"""Minimal read-only HDF5 parser, enough for tombo-style k-mer model files (no h5py in this image).

Replaces the ``h5py.File(...)`` use in the reference's ``nadavca/kmer_model.py:13-29``: it reads the root
attributes (``central_pos``) and one chunked + deflate-compressed 1-D compound dataset (``model`` with fields
``kmer``/``mean``/``sd``).  Supported subset of the format (what h5py's default ``libver='earliest'`` writes):
superblock v0/v1, v1 object headers (with continuation blocks), v1 group B-trees + local heaps, dataspace v1/v2,
fixed-point / floating-point / string / compound datatypes, contiguous, compact and chunked (v1 B-tree) layouts,
deflate and shuffle filters, v1-v3 attribute messages.  Anything else raises ``NotImplementedError``.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        self.O = 8  # size of offsets
        self.L = 8  # size of lengths

    def u(self, pos, size):
        return int.from_bytes(self.buf[pos:pos + size], "little")

    def off(self, pos):
        return self.u(pos, self.O)

    def length(self, pos):
        return self.u(pos, self.L)


def _pad8(x):
    return (x + 7) & ~7


def _parse_datatype(r, pos):
    """Return (numpy dtype, bytes consumed)."""
    b0 = r.buf[pos]
    cls, version = b0 & 0x0F, b0 >> 4
    bits = r.u(pos + 1, 3)
    size = r.u(pos + 4, 4)
    p = pos + 8
    if cls == 0:  # fixed point
        signed = (bits >> 3) & 1
        big = bits & 1
        dt = np.dtype(("%s%s%d" % (">" if big else "<", "i" if signed else "u", size)))
        return dt, p + 4 - pos
    if cls == 1:  # floating point
        big = bits & 1
        dt = np.dtype("%sf%d" % (">" if big else "<", size))
        return dt, p + 12 - pos
    if cls == 3:  # string
        return np.dtype("S%d" % size), p - pos
    if cls == 6:  # compound
        nmembers = bits & 0xFFFF
        names, formats, offsets = [], [], []
        for _ in range(nmembers):
            end = r.buf.index(b"\0", p)
            name = r.buf[p:end].decode("ascii")
            if version < 3:
                p += _pad8(end - p + 1)
            else:
                p = end + 1
            if version < 3:
                moff = r.u(p, 4)
                p += 4
            else:
                nbytes = max(1, (size.bit_length() + 7) // 8)
                moff = r.u(p, nbytes)
                p += nbytes
            if version == 1:
                p += 1 + 3 + 4 + 4 + 16  # dimensionality, reserved, permutation, reserved, 4 dim sizes
            mdt, used = _parse_datatype(r, p)
            p += used
            names.append(name)
            formats.append(mdt)
            offsets.append(moff)
        dt = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
        return dt, p - pos
    if cls == 9:  # variable length (only strings are materialised, see _vlen_string)
        base, used = _parse_datatype(r, p)
        dt = np.dtype([("len", "<u4"), ("heap", "<u%d" % r.O), ("index", "<u4")])
        dt = np.dtype(dt, metadata={"vlen_string": (bits & 0xF) == 1})
        return dt, p + used - pos
    raise NotImplementedError("HDF5 datatype class %d not supported" % cls)


def _vlen_string(r, rec):
    """Fetch one variable-length string from a global heap collection (GCOL)."""
    addr, index = int(rec["heap"]), int(rec["index"])
    if r.buf[addr:addr + 4] != b"GCOL":
        raise ValueError("bad global heap signature")
    end = addr + r.length(addr + 8)
    p = addr + 8 + r.L
    while p + 8 + r.L <= end:
        idx = r.u(p, 2)
        size = r.length(p + 8)
        if idx == index:
            return bytes(r.buf[p + 8 + r.L:p + 8 + r.L + size]).decode("utf-8")
        if idx == 0:
            break
        p += 8 + r.L + _pad8(size)
    raise KeyError("global heap object %d not found" % index)


def _parse_dataspace(r, pos):
    version = r.buf[pos]
    rank = r.buf[pos + 1]
    flags = r.buf[pos + 2]
    if version == 1:
        p = pos + 8
    elif version == 2:
        p = pos + 4
    else:
        raise NotImplementedError("dataspace version %d" % version)
    dims = [r.length(p + i * r.L) for i in range(rank)]
    p += rank * r.L
    if flags & 1:
        p += rank * r.L
    return tuple(dims), p - pos


class _Object:
    """Parsed v1 object header: list of (type, data_pos, size)."""

    def __init__(self, r, addr):
        self.r = r
        self.messages = []
        if r.buf[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("HDF5 v2 object headers are not supported (file written with libver='latest')")
        version = r.buf[addr]
        if version != 1:
            raise NotImplementedError("object header version %d" % version)
        nmsg = r.u(addr + 2, 2)
        hsize = r.u(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        while blocks and len(self.messages) < nmsg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(self.messages) < nmsg:
                mtype = r.u(p, 2)
                msize = r.u(p + 2, 2)
                data = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((r.off(data), r.length(data + r.O)))
                self.messages.append((mtype, data, msize))
                p = data + msize

    def find(self, mtype):
        return [(d, s) for t, d, s in self.messages if t == mtype]

    def attributes(self):
        out = {}
        r = self.r
        for d, _ in self.find(0x0C):
            version = r.buf[d]
            name_size = r.u(d + 2, 2)
            dt_size = r.u(d + 4, 2)
            ds_size = r.u(d + 6, 2)
            p = d + 8
            if version == 3:
                p += 1
            name = r.buf[p:p + name_size].split(b"\0")[0].decode("ascii")
            step = _pad8 if version == 1 else (lambda x: x)
            p += step(name_size)
            dt, _ = _parse_datatype(r, p)
            p += step(dt_size)
            dims, _ = _parse_dataspace(r, p)
            p += step(ds_size)
            count = int(np.prod(dims)) if dims else 1
            val = np.frombuffer(r.buf, dtype=dt, count=count, offset=p)
            if dt.metadata and "vlen_string" in dt.metadata:
                strs = [_vlen_string(r, rec) for rec in val]
                out[name] = strs if dims else strs[0]
            else:
                out[name] = val.reshape(dims) if dims else val[0]
        return out


def _walk_group(r, btree_addr, heap_addr):
    """Yield (name, object header address) of a v1 group."""
    assert r.buf[heap_addr:heap_addr + 4] == b"HEAP"
    heap_data = r.off(heap_addr + 8 + 2 * r.L)

    def name_at(o):
        s = heap_data + o
        return r.buf[s:r.buf.index(b"\0", s)].decode("ascii")

    def node(addr):
        sig = r.buf[addr:addr + 4]
        if sig == b"TREE":
            level = r.buf[addr + 5]
            used = r.u(addr + 6, 2)
            p = addr + 8 + 2 * r.O
            for i in range(used):
                child = r.off(p + r.L + i * (r.L + r.O))
                yield from node(child)
            del level
        elif sig == b"SNOD":
            count = r.u(addr + 6, 2)
            p = addr + 8
            for i in range(count):
                e = p + i * (2 * r.O + 24)
                yield name_at(r.off(e)), r.off(e + r.O)
        else:
            raise ValueError("bad group node signature %r" % sig)

    yield from node(btree_addr)


def _chunks(r, addr, rank):
    """Yield (offsets, size, filter_mask, address) from a v1 chunk B-tree."""
    assert r.buf[addr:addr + 4] == b"TREE", "bad chunk b-tree"
    level = r.buf[addr + 5]
    used = r.u(addr + 6, 2)
    p = addr + 8 + 2 * r.O
    key = 8 + 8 * (rank + 1)
    for i in range(used):
        k = p + i * (key + r.O)
        size = r.u(k, 4)
        mask = r.u(k + 4, 4)
        offs = tuple(r.u(k + 8 + 8 * j, 8) for j in range(rank))
        child = r.off(k + key)
        if level > 0:
            yield from _chunks(r, child, rank)
        else:
            yield offs, size, mask, child


def _read_dataset(r, obj):
    (dt_pos, _), = obj.find(0x03)
    (ds_pos, _), = obj.find(0x01)
    (lay_pos, _), = obj.find(0x08)
    dtype, _ = _parse_datatype(r, dt_pos)
    dims, _ = _parse_dataspace(r, ds_pos)
    filters = []
    for fpos, _ in obj.find(0x0B):
        version = r.buf[fpos]
        nfilters = r.buf[fpos + 1]
        p = fpos + (8 if version == 1 else 2)
        for _ in range(nfilters):
            fid = r.u(p, 2)
            if version == 1 or fid >= 256:
                name_len = r.u(p + 2, 2)
                p += 2
            else:
                name_len = 0
            ncd = r.u(p + 4, 2)
            p += 6
            if version == 1:
                name_len = _pad8(name_len)
            p += name_len
            cd = [r.u(p + 4 * j, 4) for j in range(ncd)]
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            filters.append((fid, cd))
    lver = r.buf[lay_pos]
    if lver != 3:
        raise NotImplementedError("data layout message version %d" % lver)
    lclass = r.buf[lay_pos + 1]
    count = int(np.prod(dims)) if dims else 1
    if lclass == 0:  # compact
        size = r.u(lay_pos + 2, 2)
        raw = r.buf[lay_pos + 4:lay_pos + 4 + size]
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(dims).copy()
    if lclass == 1:  # contiguous
        addr = r.off(lay_pos + 2)
        return np.frombuffer(r.buf, dtype=dtype, count=count, offset=addr).reshape(dims).copy()
    if lclass == 2:  # chunked
        rank = r.buf[lay_pos + 2] - 1
        btree = r.off(lay_pos + 3)
        cdims = tuple(r.u(lay_pos + 3 + r.O + 4 * j, 4) for j in range(rank))
        if rank != 1:
            raise NotImplementedError("only 1-D chunked datasets are supported")
        out = np.zeros(dims, dtype=dtype)
        for offs, size, mask, addr in _chunks(r, btree, rank):
            raw = bytes(r.buf[addr:addr + size])
            for idx in range(len(filters) - 1, -1, -1):
                if mask & (1 << idx):
                    continue
                fid, cd = filters[idx]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle
                    es = cd[0] if cd else dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    nel = len(a) // es
                    raw = a[:nel * es].reshape(es, nel).T.tobytes() + a[nel * es:].tobytes()
                elif fid == 3:  # fletcher32 checksum trailer
                    raw = raw[:-4]
                else:
                    raise NotImplementedError("HDF5 filter id %d" % fid)
            chunk = np.frombuffer(raw, dtype=dtype, count=cdims[0])
            lo = offs[0]
            hi = min(lo + cdims[0], dims[0])
            out[lo:hi] = chunk[:hi - lo]
        return out
    raise NotImplementedError("layout class %d" % lclass)


class File:
    """``File(path).attrs`` -> dict of root attributes; ``File(path)[name]`` -> numpy array of a root dataset."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            buf = fh.read()
        if buf[:8] != _SIG:
            raise ValueError("%s is not an HDF5 file" % path)
        r = _Reader(buf)
        version = buf[8]
        if version > 1:
            raise NotImplementedError("HDF5 superblock version %d is not supported" % version)
        r.O, r.L = buf[13], buf[14]
        p = 24 + (4 if version == 1 else 0)
        p += 4 * r.O  # base, free-space, eof, driver info addresses
        root_header = r.off(p + r.O)
        self._r = r
        self._root = _Object(r, root_header)
        self.attrs = self._root.attributes()
        (st, _), = self._root.find(0x11)
        self._members = dict(_walk_group(r, r.off(st), r.off(st + r.O)))

    def keys(self):
        return list(self._members)

    def __getitem__(self, name):
        return _read_dataset(self._r, _Object(self._r, self._members[name]))
