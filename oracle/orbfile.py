"""Minimal reader for VOTCA .orb checkpoint files (HDF5) - test infrastructure.

The reference's integration tests (xtp/src/tests/CMakeLists.txt:336-416) compare the .orb written by
`xtp_tools -e dftgwbse` with checked-in files.  Neither h5py nor the HDF5 C library exist in this
environment, so this module restates just enough of the published HDF5 file format (superblock v2/v3,
version-2 object headers with continuation chunks, compact and dense (fractal heap) link storage,
dataspace / datatype / layout messages, contiguous and compact little-endian numeric datasets, numeric and
fixed-length string attributes) to read what Orbitals::WriteToCpt stores (orbitals.cc:990-1063, layout of
matrices: checkpointwriter.h:204-253 - file dataset (rows, cols) in C order equals the Eigen matrix).

    f = OrbFile(path); f.keys("/QMdata"); f.read("/QMdata/mos/eigenvalues"); f.attrs("/QMdata")
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def lookup3(data, initval=0):
    """Bob Jenkins' lookup3 hashlittle (public domain), the metadata checksum of the HDF5 file format."""
    M = 0xFFFFFFFF

    def rot(x, k):
        return ((x << k) | (x >> (32 - k))) & M

    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & M
    p = 0
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & M
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & M
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & M
        a = (a - c) & M; a ^= rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 4); b = (b + a) & M
        p += 12
        n -= 12
    if n == 0:
        return c
    t = bytes(data[p:p + n]) + b"\0" * (12 - n)
    a = (a + int.from_bytes(t[0:4], "little")) & M
    b = (b + int.from_bytes(t[4:8], "little")) & M
    c = (c + int.from_bytes(t[8:12], "little")) & M
    c ^= b; c = (c - rot(b, 14)) & M
    a ^= c; a = (a - rot(c, 11)) & M
    b ^= a; b = (b - rot(a, 25)) & M
    c ^= b; c = (c - rot(b, 16)) & M
    a ^= c; a = (a - rot(c, 4)) & M
    b ^= a; b = (b - rot(a, 14)) & M
    c ^= b; c = (c - rot(b, 24)) & M
    return c


class OrbFile:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver = b[8]
        if ver not in (2, 3) or b[9] != 8 or b[10] != 8:
            raise ValueError("only superblock v2/v3 with 8-byte offsets is supported")
        self.base = struct.unpack_from("<Q", b, 12)[0]
        self.root = struct.unpack_from("<Q", b, 36)[0]
        self._objs = {}
        self._heap_links = None

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """[(type, flags, payload bytes)] of the version-2 object header at addr, continuation chunks followed."""
        if addr in self._objs:
            return self._objs[addr]
        b = self.b
        if b[addr:addr + 4] != b"OHDR" or b[addr + 4] != 2:
            raise ValueError(f"no version-2 object header at {addr}")
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nb = 1 << (flags & 3)
        size0 = int.from_bytes(b[p:p + nb], "little")
        p += nb
        track_order = bool(flags & 0x04)
        chunks = [(p, p + size0)]
        msgs = []
        while chunks:
            p, end = chunks.pop(0)
            while p + 4 <= end:
                mtype = b[p]
                msize = struct.unpack_from("<H", b, p + 1)[0]
                mflags = b[p + 3]
                p += 4
                if track_order:
                    p += 2
                body = b[p:p + msize]
                p += msize
                if mtype == 0x10:  # continuation: offset, length of an OCHK block (signature first, checksum last)
                    off, length = struct.unpack_from("<QQ", body, 0)
                    if b[off:off + 4] != b"OCHK":
                        raise ValueError("bad continuation chunk")
                    chunks.append((off + 4, off + length - 4))
                elif mtype != 0:
                    msgs.append((mtype, mflags, body))
        self._objs[addr] = msgs
        return msgs

    @staticmethod
    def _parse_link(body, p=0):
        """One link message (also the record format of dense link storage).  Returns (name, address, next offset)."""
        if body[p] != 1:
            raise ValueError("link version")
        flags = body[p + 1]
        p += 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nb = 1 << (flags & 3)
        nlen = int.from_bytes(body[p:p + nb], "little")
        p += nb
        name = body[p:p + nlen].decode()
        p += nlen
        if ltype != 0:
            raise ValueError("only hard links are supported")
        addr = struct.unpack_from("<Q", body, p)[0]
        return name, addr, p + 8

    def _scan_heaps(self):
        """Dense link storage: link records packed in the direct blocks of the group's fractal heap.  Rather than
        walking the heap's block tree, every direct block in the file is attributed to its heap header."""
        if self._heap_links is not None:
            return
        b = self.b
        self._heap_links = {}
        self._heap_attrs = {}
        pos = b.find(b"FHDB")
        while pos >= 0:
            if b[pos + 4] == 0:
                heap = struct.unpack_from("<Q", b, pos + 5)[0]
                if b[heap:heap + 4] == b"FRHP":
                    hflags = b[heap + 9]
                    max_heap_bits = struct.unpack_from("<H", b, heap + 5 + 2 + 2 + 1 + 4 + 12 * 8 + 2 + 8 + 8)[0]
                    p = pos + 13 + (max_heap_bits + 7) // 8
                    if hflags & 0x02:
                        p += 4
                    links = self._heap_links.setdefault(heap, {})
                    attrs = self._heap_attrs.setdefault(heap, {})
                    while p < len(b) and b[p] == 1:
                        try:
                            name, addr, nxt = self._parse_link(b, p)
                        except (ValueError, UnicodeDecodeError, struct.error, IndexError):
                            break
                        if not name or addr >= len(b) or b[addr:addr + 4] != b"OHDR":
                            break
                        links[name] = addr
                        p = nxt
                    while p < len(b) and b[p] == 3 and not links:
                        try:
                            name, val, nxt = self._parse_attribute(b, p)
                        except (ValueError, NotImplementedError, UnicodeDecodeError, struct.error, IndexError):
                            break
                        if not name or not name.isprintable():
                            break
                        attrs[name] = val
                        p = nxt
            pos = b.find(b"FHDB", pos + 4)

    def _links(self, addr):
        out = {}
        for mtype, _, body in self._messages(addr):
            if mtype == 0x06:
                name, a, _ = self._parse_link(body)
                out[name] = a
            elif mtype == 0x02:  # link info: fractal heap of a densely stored group
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                heap = struct.unpack_from("<Q", body, p)[0]
                if heap != UNDEF:
                    self._scan_heaps()
                    out.update(self._heap_links.get(heap, {}))
        return out

    def _resolve(self, path):
        addr = self.root
        for part in [s for s in path.split("/") if s]:
            links = self._links(addr)
            if part not in links:
                raise KeyError(f"{part!r} not found below {path!r}; have {sorted(links)}")
            addr = links[part]
        return addr

    def keys(self, path="/"):
        return sorted(self._links(self._resolve(path)))

    def is_group(self, path):
        return not any(t == 0x08 for t, _, _ in self._messages(self._resolve(path)))

    # ------------------------------------------------------------------ datatypes / dataspaces
    @staticmethod
    def _dtype(body):
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits0 = body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        if cls == 0:  # fixed point
            if bits0 & 1:
                raise ValueError("big-endian data is not supported")
            return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
        if cls == 1:  # IEEE float
            if bits0 & 1:
                raise ValueError("big-endian data is not supported")
            return np.dtype("<f" + str(size))
        if cls == 3:  # fixed-length string
            return np.dtype("S" + str(size))
        raise NotImplementedError(f"HDF5 datatype class {cls} (version {ver})")

    @staticmethod
    def _dataspace(body):
        ver, rank = body[0], body[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:  # null dataspace
                return None
            p = 4
        else:
            raise NotImplementedError("dataspace version")
        return tuple(struct.unpack_from("<Q", body, p + 8 * i)[0] for i in range(rank))

    # ------------------------------------------------------------------ datasets
    def read(self, path):
        """Dataset as a NumPy array with the shape stored in the file."""
        msgs = self._messages(self._resolve(path))
        dims = dt = layout = None
        for mtype, _, body in msgs:
            if mtype == 0x01:
                dims = self._dataspace(body)
            elif mtype == 0x03:
                dt = self._dtype(body)
            elif mtype == 0x08:
                layout = body
        if dt is None or layout is None:
            raise ValueError(f"{path} is not a dataset")
        if dims is None:
            return np.zeros(0, dtype=dt)
        n = int(np.prod(dims)) if dims else 1
        ver, cls = layout[0], layout[1]
        if ver not in (3, 4):
            raise NotImplementedError("data layout version")
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            raw = layout[4:4 + size]
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return np.zeros(dims, dtype=dt)
            raw = self.b[self.base + addr:self.base + addr + size]
        else:
            raise NotImplementedError("chunked datasets are not supported")
        return np.frombuffer(raw, dtype=dt, count=n).reshape(dims).copy()

    def _parse_attribute(self, body, p=0):
        """One attribute message (also the record format of dense attribute storage).
        Returns (name, value or None, next offset)."""
        ver = body[p]
        if ver not in (1, 2, 3):
            raise ValueError("attribute version")
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, p + 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
        name = bytes(body[q:q + nsz]).split(b"\0")[0].decode()
        q += pad(nsz)
        tbody = bytes(body[q:q + tsz])
        q += pad(tsz)
        dims = self._dataspace(bytes(body[q:q + ssz]))
        q += pad(ssz)
        n = int(np.prod(dims)) if dims else 1
        cls = tbody[0] & 0x0F
        size = struct.unpack_from("<I", tbody, 4)[0]
        if cls == 9:  # variable length (strings): 4-byte length + global heap id (8 + 4) per element
            val = None
            if tbody[1] & 0x0F == 1 and n == 1:
                val = self._global_heap_object(*struct.unpack_from("<IQI", body, q))
            return name, val, q + 16 * n
        dt = self._dtype(tbody)
        val = np.frombuffer(bytes(body[q:q + n * size]), dtype=dt, count=n)
        return name, (val[0] if n == 1 else val.copy()), q + n * size

    def _global_heap_object(self, length, coll, index):
        b = self.b
        if b[coll:coll + 4] != b"GCOL":
            return None
        size = struct.unpack_from("<Q", b, coll + 8)[0]
        p, end = coll + 16, coll + size
        while p + 16 <= end:
            idx, _, _, osz = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + length].decode(errors="replace")
            p += 16 + (osz + 7) // 8 * 8
        return None

    def attrs(self, path):
        """Attributes of an object as {name: value} (numeric, fixed and variable-length strings), stored compactly
        in the object header or densely in a fractal heap."""
        out = {}
        for mtype, _, body in self._messages(self._resolve(path)):
            if mtype == 0x0C:
                try:
                    name, val, _ = self._parse_attribute(body)
                except NotImplementedError:
                    continue
                out[name] = val
            elif mtype == 0x15:  # attribute info: fractal heap of densely stored attributes
                flags = body[1]
                q = 2 + (2 if flags & 1 else 0)
                heap = struct.unpack_from("<Q", body, q)[0]
                if heap != UNDEF:
                    self._scan_heaps()
                    out.update(self._heap_attrs.get(heap, {}))
        return out

    def verify_checksums(self):
        """Recomputes the lookup3 checksum of the superblock and of every object-header chunk reachable from the
        root group.  Returns the number of blocks checked; raises on the first mismatch."""
        b = self.b
        count = 0

        def check(start, end, what):
            nonlocal count
            stored = struct.unpack_from("<I", b, end)[0]
            if lookup3(b[start:end]) != stored:
                raise ValueError(f"checksum mismatch in {what} at {start}")
            count += 1

        check(0, 44, "superblock")
        seen, todo = set(), [self.root]
        while todo:
            addr = todo.pop()
            if addr in seen:
                continue
            seen.add(addr)
            flags = b[addr + 5]
            p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            nb = 1 << (flags & 3)
            size0 = int.from_bytes(b[p:p + nb], "little")
            check(addr, p + nb + size0, "object header")
            for mtype, _, body in self._messages(addr):
                if mtype == 0x10:
                    off, length = struct.unpack_from("<QQ", body, 0)
                    check(off, off + length - 4, "continuation chunk")
            todo.extend(self._links(addr).values())
        return count

    def walk(self, path="/"):
        """All dataset paths below path."""
        found = []
        for k in self.keys(path):
            child = path.rstrip("/") + "/" + k
            if self.is_group(child):
                found += self.walk(child)
            else:
                found.append(child)
        return found
