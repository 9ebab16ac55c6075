"""A small read-only HDF5 reader (NumPy only) for the input files of the path.

The reference reads two kinds of HDF5 input through h5py: the GBM background spectral templates
(``cosmogrb/data/gbm_backgrounds/<det>.h5``, sampler/background.py:46-60) and popsynth population files
(universe/universe.py:45-47).  h5py is a hard dependency of the reference but may be absent where only the engine is
installed (it is in the build image), so this module decodes the subset of the format those files use:

  * superblock version 0 / 1, "old style" groups (symbol table message -> v1 B-tree -> symbol nodes -> local heap),
    "new style" compact groups (link messages);
  * version-1 and version-2 object headers with continuation blocks;
  * datasets: fixed-point, floating-point, fixed-length string, variable-length string (global heap) and enum (the
    h5py ``bool``) types, scalar / simple dataspaces, compact, contiguous and chunked layouts (v1 B-tree chunk index),
    filters deflate (zlib), shuffle, fletcher32 and LZF (h5py's filter 32000, decoded here);
  * attributes (message versions 1-3) of groups and datasets.

Tested on real files (tests/test_hdf5_lite.py): everything h5py's default ``libver="earliest"`` writes -- version-0
superblock, version-1 object headers, old-style groups -- which is what the reference's inputs are.  The version-2
object header / link-message branch follows the format specification but no file of that kind was at hand.
Anything else raises ``HDF5Error`` naming the unsupported feature -- nothing is guessed.  ``File`` mimics the few h5py
idioms the callers use: ``f["a/b"]``, ``.keys()``, ``.attrs``, ``ds[()]``, ``ds.shape``, ``ds.dtype``, ``visititems``.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(ValueError):
    pass


def lzf_decompress(src, out_len):
    """liblzf's format: control byte < 32 -> literal run of ctrl + 1 bytes; else a back reference of length
    (ctrl >> 5) + 2 (7 -> one more length byte) at distance ((ctrl & 31) << 8 | next byte) + 1."""
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            run = ctrl + 1
            if op + run > out_len or ip + run > n:
                raise HDF5Error("corrupt LZF stream (literal run)")
            out[op:op + run] = src[ip:ip + run]
            ip += run
            op += run
        else:
            length = ctrl >> 5
            if length == 7:
                length += src[ip]
                ip += 1
            ref = op - ((ctrl & 0x1F) << 8) - src[ip] - 1
            ip += 1
            length += 2
            if ref < 0 or op + length > out_len:
                raise HDF5Error("corrupt LZF stream (back reference)")
            if ref + length <= op:
                out[op:op + length] = out[ref:ref + length]
            else:                                   # overlapping copy: byte by byte, as memmove would not
                for k in range(length):
                    out[op + k] = out[ref + k]
            op += length
    if op != out_len:
        raise HDF5Error(f"LZF stream decoded to {op} bytes, expected {out_len}")
    return bytes(out)


class _Reader(object):
    def __init__(self, raw):
        self.raw = raw
        if raw[:8] != SIGNATURE:
            raise HDF5Error("not an HDF5 file (signature at offset 0 only)")
        ver = raw[8]
        if ver not in (0, 1):
            raise HDF5Error(f"superblock version {ver} is not supported")
        self.O, self.L = raw[13], raw[14]
        if self.O != 8 or self.L != 8:
            raise HDF5Error(f"offset / length sizes {self.O} / {self.L} are not supported")
        pos = 24 + (4 if ver == 1 else 0)
        self.base = self.u(pos, 8)
        if self.base:
            raise HDF5Error("a non-zero base address (user block) is not supported")
        pos += 4 * 8                                 # base, free-space info, end of file, driver info
        # root group symbol table entry
        self.root_header = self.u(pos + 8, 8)
        self._gheaps = {}

    def u(self, pos, n):
        return int.from_bytes(self.raw[pos:pos + n], "little")

    # -- object headers ------------------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, body bytes)] of the object header at ``addr`` (v1 or v2), continuation blocks followed."""
        raw = self.raw
        out = []
        if raw[addr:addr + 4] == b"OHDR":
            ver, flags = raw[addr + 4], raw[addr + 5]
            if ver != 2:
                raise HDF5Error(f"object header version {ver}")
            pos = addr + 6
            if flags & 0x20:
                pos += 16
            if flags & 0x10:
                pos += 4
            szl = 1 << (flags & 3)
            chunk0 = self.u(pos, szl)
            pos += szl
            blocks = [(pos, pos + chunk0)]
            tracked = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = raw[p], self.u(p + 1, 2), raw[p + 3]
                    p += 4 + (2 if tracked else 0)
                    body = raw[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        off, ln = self.u(p - msize, 8), self.u(p - msize + 8, 8)
                        blocks.append((off + 4, off + ln - 4))         # "OCHK" ... checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, body))
            return out
        ver = raw[addr]
        if ver != 1:
            raise HDF5Error(f"object header version {ver} at {addr}")
        n_msgs, size = self.u(addr + 2, 2), self.u(addr + 8, 4)
        blocks = [(addr + 16, addr + 16 + size)]
        while blocks and len(out) < n_msgs + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize, mflags = self.u(p, 2), self.u(p + 2, 2), raw[p + 4]
                body = raw[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    blocks.append((self.u(p - msize, 8), self.u(p - msize, 8) + self.u(p - msize + 8, 8)))
                elif mtype != 0:
                    out.append((mtype, mflags, body))
        return out

    # -- groups --------------------------------------------------------------------------------------------------
    def heap_string(self, heap_addr, offset):
        raw = self.raw
        if raw[heap_addr:heap_addr + 4] != b"HEAP":
            raise HDF5Error("local heap signature")
        data = self.u(heap_addr + 24, 8)
        end = raw.index(b"\0", data + offset)
        return raw[data + offset:end].decode("utf-8")

    def group_btree(self, addr, heap_addr, out):
        raw = self.raw
        if raw[addr:addr + 4] == b"SNOD":
            n = self.u(addr + 6, 2)
            p = addr + 8
            for _ in range(n):
                out[self.heap_string(heap_addr, self.u(p, 8))] = self.u(p + 8, 8)
                p += 40
            return
        if raw[addr:addr + 4] != b"TREE" or raw[addr + 4] != 0:
            raise HDF5Error("group B-tree node signature")
        used = self.u(addr + 6, 2)
        p = addr + 24 + 8                            # past the siblings and key 0
        for _ in range(used):
            self.group_btree(self.u(p, 8), heap_addr, out)
            p += 16

    def links(self, header_addr):
        """{name: object header address} of a group, in name order."""
        out = {}
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x11:
                self.group_btree(int.from_bytes(body[0:8], "little"), int.from_bytes(body[8:16], "little"), out)
            elif mtype == 0x06:
                ver, flags = body[0], body[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                szl = 1 << (flags & 3)
                nlen = int.from_bytes(body[p:p + szl], "little")
                p += szl
                name = body[p:p + nlen].decode("utf-8")
                p += nlen
                if ver != 1 or ltype != 0:
                    raise HDF5Error(f"link {name!r}: only hard links are supported")
                out[name] = int.from_bytes(body[p:p + 8], "little")
            elif mtype == 0x02:
                if int.from_bytes(body[-16:-8] if len(body) >= 18 else b"\xff" * 8, "little") != UNDEF:
                    raise HDF5Error("dense (fractal-heap) groups are not supported")
        return dict(sorted(out.items()))

    def is_group(self, header_addr):
        return not any(t == 0x08 for t, _, _ in self.messages(header_addr))

    # -- datatypes / dataspaces ----------------------------------------------------------------------------------
    def datatype(self, body, p=0):
        """-> (numpy dtype or ('vlen_str',) / ('enum_bool', base dtype), bytes consumed)"""
        cls, ver = body[p] & 0x0F, body[p] >> 4
        bits = body[p + 1] | (body[p + 2] << 8) | (body[p + 3] << 16)
        size = int.from_bytes(body[p + 4:p + 8], "little")
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits & 0x08 else 'u'}{size}"), 8 + 4
        if cls == 1:
            if size not in (2, 4, 8):
                raise HDF5Error(f"{size}-byte floating point")
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:
            return np.dtype(f"S{size}"), 8
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise HDF5Error("variable-length sequences (only variable-length strings are supported)")
            _, used = self.datatype(body, p + 8)
            return ("vlen_str",), 8 + used
        if cls == 8:
            base, used = self.datatype(body, p + 8)
            n = bits & 0xFFFF
            q = p + 8 + used
            names = []
            for _ in range(n):
                if ver >= 3:
                    end = body.index(b"\0", q)
                    names.append(body[q:end])
                    q = end + 1
                else:
                    end = body.index(b"\0", q)
                    names.append(body[q:end])
                    q += (end - q) // 8 * 8 + 8
            q += n * base.itemsize
            if sorted(names) == [b"FALSE", b"TRUE"] and base.itemsize == 1:
                return ("enum_bool", base), q - p
            raise HDF5Error("enumerations other than the h5py bool")
        raise HDF5Error(f"datatype class {cls} is not supported")

    @staticmethod
    def dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:
                return None                          # null dataspace
            p = 4
        else:
            raise HDF5Error(f"dataspace version {ver}")
        return tuple(int.from_bytes(body[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    # -- data ----------------------------------------------------------------------------------------------------
    def gheap_object(self, addr, index):
        if addr not in self._gheaps:
            raw = self.raw
            if raw[addr:addr + 4] != b"GCOL":
                raise HDF5Error("global heap signature")
            size = self.u(addr + 8, 8)
            objs, p = {}, addr + 16
            while p + 16 <= addr + size:
                idx, osz = self.u(p, 2), self.u(p + 8, 8)
                if idx == 0:
                    break
                objs[idx] = raw[p + 16:p + 16 + osz]
                p += 16 + (osz + 7) // 8 * 8
            self._gheaps[addr] = objs
        return self._gheaps[addr][index]

    def decode(self, dt, shape, data):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, tuple) and dt[0] == "vlen_str":
            out = np.empty(n, dtype=object)
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", data, 16 * i)
                out[i] = self.gheap_object(addr, idx)[:ln] if addr else b""
            return out.reshape(shape) if shape else out[0]
        if isinstance(dt, tuple) and dt[0] == "enum_bool":
            a = np.frombuffer(data, dtype=dt[1], count=n).astype(bool)
            return a.reshape(shape) if shape else a[0]
        a = np.frombuffer(data, dtype=dt, count=n)
        if dt.kind != "S":
            a = a.astype(dt.newbyteorder("="))
        return a.reshape(shape).copy() if shape else a[0]

    def chunks(self, addr, rank, out):
        """[(offsets, filter mask, address, size)] of a v1 chunk B-tree"""
        raw = self.raw
        if addr == UNDEF:
            return
        if raw[addr:addr + 4] != b"TREE" or raw[addr + 4] != 1:
            raise HDF5Error("chunk B-tree node signature")
        level, used = raw[addr + 5], self.u(addr + 6, 2)
        p = addr + 24
        ksz = 8 + 8 * (rank + 1)
        for _ in range(used):
            size, mask = self.u(p, 4), self.u(p + 4, 4)
            offs = tuple(self.u(p + 8 + 8 * i, 8) for i in range(rank))
            child = self.u(p + ksz, 8)
            if level:
                self.chunks(child, rank, out)
            else:
                out.append((offs, mask, child, size))
            p += ksz + 8

    def read_dataset(self, header_addr):
        dt = shape = layout = None
        filters = []
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x03:
                dt, _ = self.datatype(body)
            elif mtype == 0x01:
                shape = self.dataspace(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                ver, nf = body[0], body[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = int.from_bytes(body[p:p + 2], "little")
                    if ver == 1 or fid >= 256:
                        nlen = int.from_bytes(body[p + 2:p + 4], "little")
                        p += 4
                    else:
                        nlen = 0
                        p += 2
                    ncd = int.from_bytes(body[p + 2:p + 4], "little")
                    p += 4 + (((nlen + 7) // 8 * 8) if ver == 1 else nlen)
                    cd = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd)]
                    p += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
                    filters.append((fid, cd))
        if dt is None or layout is None:
            raise HDF5Error("dataset without datatype / layout message")
        if shape is None:
            return None
        esize = 16 if isinstance(dt, tuple) and dt[0] == "vlen_str" else (dt[1].itemsize if isinstance(dt, tuple) else dt.itemsize)
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver != 3:
            raise HDF5Error(f"data layout message version {ver}")
        cls = layout[1]
        if cls == 0:
            size = int.from_bytes(layout[2:4], "little")
            return self.decode(dt, shape, layout[4:4 + size])
        if cls == 1:
            addr = int.from_bytes(layout[2:10], "little")
            if addr == UNDEF:
                return self.decode(dt, shape, bytes(n * esize))
            return self.decode(dt, shape, self.raw[addr:addr + n * esize])
        if cls != 2:
            raise HDF5Error(f"layout class {cls}")
        rank = layout[2] - 1
        btree = int.from_bytes(layout[3:11], "little")
        cdims = tuple(int.from_bytes(layout[11 + 4 * i:15 + 4 * i], "little") for i in range(rank))
        as_bool = isinstance(dt, tuple) and dt[0] == "enum_bool"
        if as_bool:
            dt = dt[1]
        elif isinstance(dt, tuple):
            raise HDF5Error("chunked variable-length datasets")
        full = np.zeros(shape, dtype=dt.newbyteorder("=") if dt.kind != "S" else dt)
        found = []
        self.chunks(btree, rank, found)
        cbytes = int(np.prod(cdims)) * esize
        for offs, mask, addr, size in found:
            buf = self.raw[addr:addr + size]
            for k in range(len(filters) - 1, -1, -1):
                if mask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    buf = zlib.decompress(buf)
                elif fid == 2:
                    w = cd[0] if cd else esize
                    a = np.frombuffer(buf, dtype=np.uint8)
                    m = len(a) // w
                    buf = a[:m * w].reshape(w, m).T.tobytes() + a[m * w:].tobytes()
                elif fid == 3:
                    buf = buf[:-4]
                elif fid == 32000:
                    buf = lzf_decompress(buf, cd[2] if len(cd) > 2 else cbytes)
                else:
                    raise HDF5Error(f"filter {fid} is not supported")
            chunk = np.frombuffer(buf, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            full[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return full.astype(bool) if as_bool else full

    def attributes(self, header_addr):
        out = {}
        for mtype, _, body in self.messages(header_addr):
            if mtype != 0x0C:
                continue
            ver = body[0]
            nsz, tsz, ssz = (int.from_bytes(body[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
            if ver == 1:
                pad = lambda v: (v + 7) // 8 * 8
                p = 8
            elif ver in (2, 3):
                if body[1] & 3:
                    raise HDF5Error("shared attribute datatypes / dataspaces")
                pad = lambda v: v
                p = 8 + (1 if ver == 3 else 0)
            else:
                raise HDF5Error(f"attribute message version {ver}")
            name = body[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += pad(nsz)
            dt, _ = self.datatype(body[p:p + tsz])
            p += pad(tsz)
            shape = self.dataspace(body[p:p + ssz])
            p += pad(ssz)
            out[name] = None if shape is None else self.decode(dt, shape, body[p:])
        return out


def _plain(v):
    """h5py-like Python values: bytes of variable-length strings as ``str``, 0-d results as scalars"""
    if isinstance(v, bytes):
        return v.decode("utf-8")
    if isinstance(v, np.ndarray) and v.dtype == object:
        return np.array([x.decode("utf-8") if isinstance(x, bytes) else x for x in v.ravel()], dtype=object).reshape(v.shape)
    return v


class Dataset(object):
    def __init__(self, reader, addr, name):
        self._r, self._addr, self.name = reader, addr, name
        self._value = None
        self._loaded = False

    def _load(self):
        if not self._loaded:
            self._value = self._r.read_dataset(self._addr)
            self._loaded = True
        return self._value

    def __getitem__(self, idx):
        v = self._load()
        if isinstance(v, np.ndarray) and v.dtype == object:
            return v[idx]                            # variable-length strings come back as bytes, like h5py 3
        return v if idx == () else v[idx]

    shape = property(lambda self: np.shape(self._load()))
    dtype = property(lambda self: np.asarray(self._load()).dtype)

    @property
    def attrs(self):
        return {k: _plain(v) for k, v in self._r.attributes(self._addr).items()}

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._load(), dtype=dtype)

    def __len__(self):
        return len(self._load())


class Group(object):
    def __init__(self, reader, addr, name="/"):
        self._r, self._addr, self.name = reader, addr, name
        self._links = None

    def _l(self):
        if self._links is None:
            self._links = self._r.links(self._addr)
        return self._links

    def keys(self):
        return list(self._l())

    def __iter__(self):
        return iter(self._l())

    def __len__(self):
        return len(self._l())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def _child(self, name):
        links = self._l()
        if name not in links:
            raise KeyError(f"{name!r} not in {self.name}")
        addr = links[name]
        full = self.name.rstrip("/") + "/" + name
        return Group(self._r, addr, full) if self._r.is_group(addr) else Dataset(self._r, addr, full)

    def __getitem__(self, path):
        obj = self
        for part in [p for p in str(path).split("/") if p]:
            if not isinstance(obj, Group):
                raise KeyError(path)
            obj = obj._child(part)
        return obj

    def items(self):
        return [(k, self._child(k)) for k in self._l()]

    def values(self):
        return [self._child(k) for k in self._l()]

    @property
    def attrs(self):
        return {k: _plain(v) for k, v in self._r.attributes(self._addr).items()}

    def visititems(self, func):
        def rec(g, prefix):
            for k, v in g.items():
                r = func(prefix + k, v)
                if r is not None:
                    return r
                if isinstance(v, Group):
                    r = rec(v, prefix + k + "/")
                    if r is not None:
                        return r
            return None

        return rec(self, "")


class File(Group):
    """``with File(path) as f: f["group/dataset"][()]`` -- read-only."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise HDF5Error("hdf5_lite is read-only")
        with open(str(path), "rb") as fh:
            raw = fh.read()
        r = _Reader(raw)
        Group.__init__(self, r, r.root_header, "/")
        self.filename = str(path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass
