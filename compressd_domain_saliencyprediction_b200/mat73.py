"""Minimal reader for MATLAB v7.3 .mat files (HDF5), enough for the reference's video_database.mat (main.m:3, 9-18, 325-346).

No h5py / libhdf5 in this image, so this walks the file format directly — the subset MATLAB R2012-R2015 writes:
superblock version 0 behind a 512-byte user block, version-1 object headers (with continuation blocks), old-style groups
(symbol-table message -> version-1 B-tree of SNOD nodes + local heap), dataspace / datatype / layout (compact, contiguous,
chunked through a version-1 chunk B-tree) / filter-pipeline messages, the deflate filter (and shuffle), fixed-point, floating
point and object-reference datatypes.  MATLAB conventions on top: arrays are stored with reversed dimensions (column-major),
char arrays as uint16 code units, structs as groups, cell arrays as datasets of object references into /#refs#.

    db = mat73.load("video_database.mat")["video_database"]
    db["video_names_list"]            -> list of str
    db["videos_info"]["size"]         -> float array [33][2]
    db["fixdata"]                     -> float array [N][6]
"""
import struct
import zlib

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class Mat73Error(ValueError):
    pass


class _File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        # the superblock sits at 0, 512, 1024, ... (MATLAB: 512-byte text header)
        off = 0
        while off < len(self.b) and self.b[off:off + 8] != b"\x89HDF\r\n\x1a\n":
            off = 512 if off == 0 else off * 2
        if off >= len(self.b):
            raise Mat73Error("no HDF5 superblock (not a v7.3 .mat file)")
        sb = self.b[off:]
        ver = sb[8]
        if ver not in (0, 1):
            raise Mat73Error(f"superblock version {ver} not supported (MATLAB writes 0)")
        self.so, self.sl = sb[13], sb[14]                      # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise Mat73Error("only 8-byte offsets and lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", sb, p)[0]          # base address: every address below is relative to it
        if self.base == 0:
            self.base = off
        p += 4 * 8                                              # base, free-space, end-of-file, driver-info addresses
        self.root = self._ste(off + p)                          # root group symbol table entry

    # ---- primitives
    def u(self, fmt, a):
        return struct.unpack_from("<" + fmt, self.b, a)

    def addr(self, a):
        return a + self.base

    def _ste(self, a):
        name_off, hdr, cache = self.u("QQI", a)
        btree = heap = None
        if cache == 1:
            btree, heap = self.u("QQ", a + 24)
        return {"name_off": name_off, "hdr": hdr, "btree": btree, "heap": heap}

    # ---- object header (version 1)
    def messages(self, hdr_addr):
        a = self.addr(hdr_addr)
        ver, _, nmsg, _, size = self.u("BBHII", a)
        if ver != 1:
            raise Mat73Error(f"object header version {ver} at {hdr_addr:#x} not supported")
        out = []
        blocks = [(a + 16, size)]
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = self.u("HHB", p)
                body = p + 8
                if mtype == 0x10:                               # continuation
                    off, ln = self.u("QQ", body)
                    blocks.append((self.addr(off), ln))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # ---- groups
    def _heap_name(self, heap_addr, off):
        a = self.addr(heap_addr)
        if self.b[a:a + 4] != b"HEAP":
            raise Mat73Error("bad local heap")
        data = self.u("Q", a + 24)[0]
        s = self.addr(data) + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode("ascii")

    def _group_nodes(self, btree_addr):
        a = self.addr(btree_addr)
        if self.b[a:a + 4] == b"SNOD":
            yield a
            return
        if self.b[a:a + 4] != b"TREE":
            raise Mat73Error("bad group B-tree node")
        _ntype, level, used = self.u("BBH", a + 4)
        p = a + 8 + 16
        for i in range(used):
            child = self.u("Q", p + 8)[0]                       # key (8), child (8), ...
            if level == 0:
                yield self.addr(child)
            else:
                yield from self._group_nodes(child)
            p += 16

    def group_members(self, hdr_addr):
        """name -> object header address for an old-style group."""
        btree = heap = None
        for mtype, body, _ in self.messages(hdr_addr):
            if mtype == 0x11:
                btree, heap = self.u("QQ", body)
        if btree is None:
            return None
        out = {}
        for node in self._group_nodes(btree):
            n = self.u("H", node + 6)[0]
            for i in range(n):
                e = self._ste(node + 8 + 40 * i)
                out[self._heap_name(heap, e["name_off"])] = e["hdr"]
        return out

    # ---- datasets
    def _dtype(self, body):
        cv, b0, _b1, _b2, size = self.u("BBBBI", body)
        cls = cv & 15
        if cls == 0:                                            # fixed point
            signed = bool(b0 & 8)
            return np.dtype(("<i" if signed else "<u") + str(size)), "num"
        if cls == 1:
            return np.dtype("<f" + str(size)), "num"
        if cls == 7:
            return np.dtype("<u8"), "ref"
        raise Mat73Error(f"datatype class {cls} not supported")

    def _chunks(self, btree_addr, rank):
        a = self.addr(btree_addr)
        if self.b[a:a + 4] != b"TREE":
            raise Mat73Error("bad chunk B-tree node")
        ntype, level, used = self.u("BBH", a + 4)
        if ntype != 1:
            raise Mat73Error("not a chunk B-tree")
        keysz = 8 + 8 * (rank + 1)
        p = a + 8 + 16
        for i in range(used):
            csize, fmask = self.u("II", p)
            offs = self.u(f"{rank + 1}Q", p + 8)
            child = self.u("Q", p + keysz)[0]
            if level == 0:
                yield csize, fmask, offs[:rank], self.addr(child)
            else:
                yield from self._chunks(child, rank)
            p += keysz + 8

    def read_dataset(self, hdr_addr):
        """-> (numpy array in the file's (C) dimension order, kind, attributes)"""
        shape = dt = kind = None
        layout = None
        filters = []
        attrs = {}
        for mtype, body, msize in self.messages(hdr_addr):
            if mtype == 0x1:
                ver, rank, flags = self.u("BBB", body)
                p = body + (8 if ver == 1 else 4)
                shape = self.u(f"{rank}Q", p) if rank else ()
            elif mtype == 0x3:
                dt, kind = self._dtype(body)
            elif mtype == 0x8:
                ver, cls = self.u("BB", body)
                if ver != 3:
                    raise Mat73Error(f"layout message version {ver} not supported")
                if cls == 0:
                    n = self.u("H", body + 2)[0]
                    layout = ("compact", body + 4, n)
                elif cls == 1:
                    a, n = self.u("QQ", body + 2)
                    layout = ("contiguous", a, n)
                else:
                    nd = self.u("B", body + 2)[0]
                    bt = self.u("Q", body + 3)[0]
                    dims = self.u(f"{nd}I", body + 11)
                    layout = ("chunked", bt, dims)
            elif mtype == 0xB:
                ver, nf = self.u("BB", body)
                p = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid, nlen, _fl, ncv = self.u("HHHH", p)
                    p += 8
                    if ver == 1 or fid >= 256:
                        p += (nlen + 7) // 8 * 8
                    p += 4 * ncv
                    if ver == 1 and ncv % 2:
                        p += 4
                    filters.append(fid)
            elif mtype == 0xC:
                try:
                    k, v = self._attribute(body)
                    attrs[k] = v
                except Mat73Error:
                    pass
        if shape is None or dt is None or layout is None:
            raise Mat73Error(f"object at {hdr_addr:#x} is not a dataset")
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            arr = np.frombuffer(self.b, dt, n, layout[1])
        elif layout[0] == "contiguous":
            arr = np.zeros(n, dt) if layout[1] == _UNDEF else np.frombuffer(self.b, dt, n, self.addr(layout[1]))
        else:
            bt, dims = layout[1], layout[2]
            rank = len(dims) - 1
            cshape = tuple(int(d) for d in dims[:rank])
            arr = np.zeros(shape, dt)
            if bt != _UNDEF:
                for csize, fmask, offs, ca in self._chunks(bt, rank):
                    raw = self.b[ca:ca + csize]
                    for idx, fid in reversed(list(enumerate(filters))):
                        if fmask & (1 << idx):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:                          # shuffle
                            a8 = np.frombuffer(raw, np.uint8)
                            raw = a8.reshape(dt.itemsize, -1).T.tobytes()
                        elif fid == 3:                          # fletcher32: trailing checksum
                            raw = raw[:-4]
                        else:
                            raise Mat73Error(f"filter {fid} not supported")
                    ch = np.frombuffer(raw, dt, int(np.prod(cshape))).reshape(cshape)
                    sl = tuple(slice(int(o), min(int(o) + c, s)) for o, c, s in zip(offs, cshape, shape))
                    arr[sl] = ch[tuple(slice(0, s.stop - s.start) for s in sl)]
            arr = arr.reshape(-1)
        return np.array(arr).reshape(shape), kind, attrs

    def _attribute(self, body):
        ver, _, nsz, dsz, ssz = self.u("BBHHH", body)
        if ver != 1:
            raise Mat73Error("attribute version")
        p = body + 8
        name = self.b[p:p + nsz].split(b"\x00")[0].decode("ascii")
        p += (nsz + 7) // 8 * 8
        cv, _b0, _b1, _b2, size = self.u("BBBBI", p)
        dtp = p
        p += (dsz + 7) // 8 * 8
        p += (ssz + 7) // 8 * 8
        if cv & 15 == 3:                                        # string
            return name, self.b[p:p + size].split(b"\x00")[0].decode("ascii")
        dt, _ = self._dtype(dtp)
        return name, np.frombuffer(self.b, dt, 1, p)[0]


def _to_matlab(f, hdr_addr, depth=0):
    members = f.group_members(hdr_addr)
    if members is not None:                                     # struct (or the root)
        return {k: _to_matlab(f, a, depth + 1) for k, a in members.items() if not k.startswith("#")}
    arr, kind, attrs = f.read_dataset(hdr_addr)
    cls = attrs.get("MATLAB_class", "")
    if kind == "ref":                                           # cell array: dereference every element
        vals = [_to_matlab(f, int(r), depth + 1) for r in arr.reshape(-1)]
        return vals
    if attrs.get("MATLAB_empty", 0):
        return "" if cls == "char" else np.zeros((0, 0))
    if cls == "char":
        return arr.astype(np.uint16).tobytes().decode("utf-16le")
    a = arr.T                                                   # stored with reversed dimensions
    return a.astype(np.float64) if cls == "double" else a


def load(path):
    """Every variable of a MATLAB v7.3 file as nested dict / list / numpy array / str (MATLAB's orientation)."""
    f = _File(path)
    return _to_matlab(f, f.root["hdr"])
