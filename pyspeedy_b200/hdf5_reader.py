"""Minimal pure-Python reader for the NetCDF-4 / HDF5 files pySPEEDY uses as boundary conditions.

The reference reads ``example_bc.nc`` and ``sst_anomaly.nc`` with xarray + netCDF4 (pyspeedy/speedy.py:277,329); neither
library exists in this image, so ``Speedy.set_bc(bc_file="*.nc")`` goes through this module.  It implements the subset of
the HDF5 file format (HDF5 File Format Specification, version 3.0) that netCDF-4 writes for such files:

  * superblock version 0-3, version-1 and version-2 object headers with continuation blocks;
  * groups with compact links (link messages) or old-style symbol tables (B-tree v1 + local heap); dense link storage
    (fractal heap) is read for the common single-direct-block case;
  * datasets: fixed-point and IEEE floating-point types, simple dataspaces, contiguous / compact / chunked layout (version-1
    B-tree chunk index, layout message versions 3 and 4 with B-tree v1 or single-chunk index), filters deflate (1),
    shuffle (2), fletcher32 (3); fill value and the ``_FillValue`` attribute;
  * attributes with numeric or fixed-length string values stored compactly in the object header.

``load(path)`` returns {name: ndarray} with the file's dimension order and ``_FillValue`` replaced by NaN for floating
types -- what ``xr.load_dataset(path)[name].values`` gives.  Anything outside the subset raises ``NotImplementedError``.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, path):
        with open(path, "rb") as fp:
            self.b = fp.read()
        if self.b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError(f"{path}: not an HDF5 file")
        ver = self.b[8]
        if ver in (0, 1):
            self.so, self.sl = self.b[13], self.b[14]
            p = 24 + (4 if ver == 1 else 0)
            base = self.u(p, self.so)
            p += 4 * self.so  # base, free-space info, end of file, driver info
            p += self.so  # root symbol table entry: link name offset
            self.root = self.u(p, self.so) + base
        elif ver in (2, 3):
            self.so, self.sl = self.b[9], self.b[10]
            base = self.u(12, self.so)
            self.root = self.u(12 + 3 * self.so, self.so) + base
        else:
            raise NotImplementedError(f"superblock version {ver}")
        if self.so != 8 or self.sl != 8:
            raise NotImplementedError("only 8-byte offsets and lengths")

    def u(self, p, n):
        return int.from_bytes(self.b[p:p + n], "little")

    # ---- object headers -------------------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, payload bytes)] of the object header at addr (versions 1 and 2, continuation blocks followed)."""
        b, out = self.b, []
        if b[addr:addr + 4] == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            n = 1 << (flags & 3)
            size0 = self.u(p, n)
            p += n
            blocks = [(p, p + size0)]
            track = bool(flags & 4)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], self.u(p + 1, 2), b[p + 3]
                    p += 4 + (2 if track else 0)
                    data = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:  # continuation: OCHK block = signature, messages, checksum
                        off, ln = struct.unpack("<QQ", data[:16])
                        assert b[off:off + 4] == b"OCHK"
                        blocks.append((off + 4, off + ln - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        if b[addr] != 1:
            raise NotImplementedError(f"object header version {b[addr]} at {addr}")
        nmsg, size0 = self.u(addr + 2, 2), self.u(addr + 8, 4)
        blocks = [(addr + 16, addr + 16 + size0)]
        while blocks and len(out) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize, mflags = self.u(p, 2), self.u(p + 2, 2), b[p + 4]
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    off, ln = struct.unpack("<QQ", data[:16])
                    blocks.append((off, off + ln))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------
    def _parse_link(self, d):
        flags = d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        n = 1 << (flags & 3)
        ln = int.from_bytes(d[p:p + n], "little")
        p += n
        name = d[p:p + ln].decode()
        p += ln
        if ltype != 0:
            return name, None  # soft / external links: not needed
        return name, int.from_bytes(d[p:p + 8], "little")

    def links(self, addr):
        """{name: object header address} of the group at addr."""
        out = {}
        for mtype, _, d in self.messages(addr):
            if mtype == 0x06:
                name, a = self._parse_link(d)
                if a is not None:
                    out[name] = a
            elif mtype == 0x02:  # link info: dense storage in a fractal heap
                flags = d[1]
                p = 2 + (8 if flags & 1 else 0)
                heap = int.from_bytes(d[p:p + 8], "little")
                if heap != UNDEF:
                    out.update(self._dense_links(heap))
            elif mtype == 0x11:  # symbol table message: B-tree v1 + local heap
                btree, heap = struct.unpack("<QQ", d[:16])
                out.update(self._symtab_links(btree, heap))
        return out

    def _dense_links(self, heap):
        b = self.b
        assert b[heap:heap + 4] == b"FRHP"
        # header: sig, ver, heap id len (2), filter len (2), flags (1), max managed size (4), then 8-byte fields ...
        hflags = b[heap + 9]
        p = heap + 5
        p += 2 + 2 + 1 + 4
        p += 8 * 12  # next huge id, huge btree, free space, fs manager, managed space, alloc, iter offset, nobjs, huge size/n, tiny size/n
        width = self.u(p, 2)
        start_size, max_direct = self.u(p + 2, 8), self.u(p + 10, 8)
        max_heap_bits, start_rows = self.u(p + 18, 2), self.u(p + 20, 2)
        root, cur_rows = self.u(p + 22, 8), self.u(p + 30, 2)
        if cur_rows != 0:
            raise NotImplementedError("fractal heap with an indirect root block (very large groups)")
        assert b[root:root + 4] == b"FHDB"
        off_bytes = (max_heap_bits + 7) // 8
        p = root + 5 + 8 + off_bytes + (4 if hflags & 2 else 0)  # signature, version, heap header address, block offset[, checksum]
        end = root + start_size
        out = {}
        while p < end and b[p] == 1:  # link messages, back to back
            name, a = self._parse_link(b[p:end])
            flags = b[p + 1]
            q = p + 2 + (1 if flags & 8 else 0) + (8 if flags & 4 else 0) + (1 if flags & 0x10 else 0)
            n = 1 << (flags & 3)
            ln = self.u(q, n)
            p = q + n + ln + 8
            if a is not None:
                out[name] = a
        return out

    def _symtab_links(self, btree, heap):
        b = self.b
        assert b[heap:heap + 4] == b"HEAP"
        data = self.u(heap + 24, 8)
        out = {}

        def node(a):
            assert b[a:a + 4] == b"TREE"
            level, n = b[a + 5], self.u(a + 6, 2)
            p = a + 8 + 16
            for i in range(n):
                child = self.u(p + 8 + i * 16, 8)
                if level > 0:
                    node(child)
                else:
                    assert b[child:child + 4] == b"SNOD"
                    for e in range(self.u(child + 6, 2)):
                        q = child + 8 + e * 40
                        off, oh = self.u(q, 8), self.u(q + 8, 8)
                        end = b.index(b"\0", data + off)
                        out[b[data + off:end].decode()] = oh
        node(btree)
        return out

    # ---- datasets ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(d):
        cls, bits0, size = d[0] & 0x0F, d[1], int.from_bytes(d[4:8], "little")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise NotImplementedError(f"datatype class {cls}")

    @staticmethod
    def _shape(d):
        ver, rank = d[0], d[1]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(d[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    def _chunks_btree_v1(self, addr, rank):
        """[(offsets, address, size, filter mask)] of a version-1 chunk B-tree."""
        b, out = self.b, []

        def node(a):
            assert b[a:a + 4] == b"TREE" and b[a + 4] == 1
            level, n = b[a + 5], self.u(a + 6, 2)
            p = a + 8 + 16
            ksz = 8 + 8 * (rank + 1)
            for i in range(n):
                k = p + i * (ksz + 8)
                size, mask = self.u(k, 4), self.u(k + 4, 4)
                offs = tuple(self.u(k + 8 + 8 * j, 8) for j in range(rank))
                child = self.u(k + ksz, 8)
                if level > 0:
                    node(child)
                else:
                    out.append((offs, child, size, mask))
        node(addr)
        return out

    def _unfilter(self, raw, filters, mask, esize):
        for i, (fid, cd) in reversed(list(enumerate(filters))):
            if mask & (1 << i):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // esize
                raw = np.frombuffer(raw[:n * esize], dtype=np.uint8).reshape(esize, n).T.tobytes() + raw[n * esize:]
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise NotImplementedError(f"HDF5 filter {fid}")
        return raw

    def attributes(self, addr):
        out = {}
        for mtype, _, d in self.messages(addr):
            if mtype != 0x0C:
                continue
            ver = d[0]
            nlen, tlen, slen = self.u_(d, 2, 2), self.u_(d, 4, 2), self.u_(d, 6, 2)
            p = 8 + (1 if ver == 3 else 0)
            pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
            name = d[p:p + nlen].split(b"\0")[0].decode()
            p += pad(nlen)
            try:
                dt = self._dtype(d[p:p + tlen])
            except NotImplementedError:
                continue
            p += pad(tlen)
            shape = self._shape(d[p:p + slen]) if d[p + 1] else ()
            p += pad(slen)
            n = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(d[p:p + n * dt.itemsize], dtype=dt)
            out[name] = val.reshape(shape) if shape else val[0]
        return out

    @staticmethod
    def u_(d, p, n):
        return int.from_bytes(d[p:p + n], "little")

    def dataset(self, addr):
        dt = shape = layout = None
        filters = []
        for mtype, _, d in self.messages(addr):
            if mtype == 0x01:
                shape = self._shape(d)
            elif mtype == 0x03:
                dt = self._dtype(d)
            elif mtype == 0x08:
                layout = d
            elif mtype == 0x0B:
                ver, nf = d[0], d[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = self.u_(d, p, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self.u_(d, p + 2, 2)
                        p += 2
                    else:
                        nlen = 0
                    ncd = self.u_(d, p + 4, 2)
                    p += 6 + (((nlen + 7) // 8 * 8) if ver == 1 else nlen)
                    cd = [self.u_(d, p + 4 * i, 4) for i in range(ncd)]
                    p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                    filters.append((fid, cd))
        if dt is None or shape is None or layout is None:
            return None
        ver, cls = layout[0], layout[1]
        if ver not in (3, 4):
            raise NotImplementedError(f"data layout version {ver}")
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:  # compact
            size = self.u_(layout, 2, 2)
            return np.frombuffer(layout[4:4 + size], dtype=dt).reshape(shape).copy()
        if cls == 1:  # contiguous
            a = self.u_(layout, 2, 8)
            if a == UNDEF:
                return np.zeros(shape, dtype=dt)
            return np.frombuffer(self.b[a:a + n * dt.itemsize], dtype=dt).reshape(shape).copy()
        if cls != 2:
            raise NotImplementedError(f"data layout class {cls}")
        out = np.zeros(shape, dtype=dt)
        if ver == 3:
            rank = layout[2] - 1
            a = self.u_(layout, 3, 8)
            cdims = tuple(self.u_(layout, 11 + 4 * i, 4) for i in range(rank))
            chunks = self._chunks_btree_v1(a, rank) if a != UNDEF else []
        else:
            flags, rank, enc = layout[2], layout[3] - 1, layout[4]
            cdims = tuple(self.u_(layout, 5 + enc * i, enc) for i in range(rank))
            p = 5 + enc * (rank + 1)
            itype = layout[p]
            p += 1
            if itype == 1:  # single chunk
                if flags & 2:
                    size, mask = self.u_(layout, p, 8), self.u_(layout, p + 8, 4)
                    p += 12
                else:
                    size, mask = int(np.prod(cdims)) * dt.itemsize, 0
                chunks = [((0,) * rank, self.u_(layout, p, 8), size, mask)]
            else:
                raise NotImplementedError(f"chunk index type {itype}")
        for offs, a, size, mask in chunks:
            raw = self._unfilter(self.b[a:a + size], filters, mask, dt.itemsize)
            c = np.frombuffer(raw[:int(np.prod(cdims)) * dt.itemsize], dtype=dt).reshape(cdims)
            sl = tuple(slice(o, min(o + cd, s)) for o, cd, s in zip(offs, cdims, shape))
            out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out


def load(path, decode_fill=True):
    """{variable name: array} of the root group's datasets (dimension order of the file).  With ``decode_fill`` the
    ``_FillValue`` / ``missing_value`` of floating-point variables becomes NaN, as xarray's CF decoding does."""
    f = _File(path)
    out = {}
    for name, addr in f.links(f.root).items():
        a = f.dataset(addr)
        if a is None:
            continue
        a = a.astype(a.dtype.newbyteorder("="))
        if decode_fill and a.dtype.kind == "f":
            att = f.attributes(addr)
            for key in ("_FillValue", "missing_value"):
                if key in att:
                    a = a.copy()
                    a[a == np.asarray(att[key]).astype(a.dtype)] = np.nan
        out[name] = a
    return out
