"""TEST INFRASTRUCTURE: a minimal writer of the HDF5 file format (no h5py / libhdf5 in this image), independent of the reader
in kiwi_b200/csrc/gfdb_hdf_host.cpp, producing the structures HDF5 1.6/1.8 write by default ("HDF5 File Format Specification"
of The HDF Group): superblock version 0, version-1 object headers, groups as symbol tables (version-1 B-tree of type 0, local
heap, symbol table nodes with at most 2*K_leaf entries, B-tree nodes with at most 2*K_internal children), contiguous dataset
layout (layout message version 3), version-1 attribute messages, object references.  On top of it: Kiwi's database layout
(gfdb_io_hdf.f90): <base>.index with scalar datasets and <base>.<i>.chunk with the "index" dataset of references and one
dataset /gf/<ixc>/<iz>/<ig> per trace carrying the attributes "pofs" and "ofs" (trace_to_storable, sparse_trace.f90:814-847)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
K_LEAF, K_INTERNAL = 4, 16


def pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace / layout message bodies ------------------------------------------------------------
def dt_float32():
    return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def dt_int32():
    return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 4) + struct.pack("<HH", 0, 32)


def dt_objref():
    return struct.pack("<BBBBI", 0x17, 0x00, 0, 0, 8)


def dataspace(dims):
    return struct.pack("<BBBB4x", 1, len(dims), 0, 0) + b"".join(struct.pack("<Q", d) for d in dims)


class Dataset:
    def __init__(self, dtype, dims, raw, attrs=()):
        self.dtype, self.dims, self.raw, self.attrs = dtype, list(dims), raw, list(attrs)   # attrs: (name, dtype, dims, raw)
        self.addr = None


class Group:
    def __init__(self):
        self.children = {}
        self.addr = None


class File:
    def __init__(self):
        self.buf = bytearray(b"\0" * 96)      # room for the superblock (O = L = 8)
        self.root = Group()

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        a = len(self.buf)
        self.buf += data
        return a

    # -- object headers --------------------------------------------------------------------------------------
    def _ohdr(self, messages):
        body = b""
        for mtype, mbody in messages:
            mb = pad8(mbody)
            body += struct.pack("<HHB3x", mtype, len(mb), 0) + mb
        hdr = struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4
        return self.alloc(hdr + body)

    def write_dataset(self, ds):
        nbytes = len(ds.raw)
        daddr = self.alloc(ds.raw) if nbytes else UNDEF
        msgs = [(0x0001, dataspace(ds.dims)), (0x0003, ds.dtype), (0x0008, struct.pack("<BBQQ", 3, 1, daddr, nbytes))]
        for name, adt, adims, araw in ds.attrs:
            nm = name.encode() + b"\0"
            ads = dataspace(adims)
            msgs.append((0x000C, struct.pack("<BBHHH", 1, 0, len(nm), len(adt), len(ads)) + pad8(nm) + pad8(adt) + pad8(ads) + araw))
        ds.addr = self._ohdr(msgs)
        return ds.addr

    def write_group(self, grp):
        # children first (their object header addresses go into the symbol table)
        for name in sorted(grp.children):
            c = grp.children[name]
            if c.addr is None:
                self.write_group(c) if isinstance(c, Group) else self.write_dataset(c)
        names = sorted(grp.children)      # symbol table order = strcmp order of the names
        # local heap: offset 0 holds the empty string, every name padded to a multiple of 8
        seg = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(seg)
            seg += pad8(n.encode() + b"\0")
        free = len(seg)
        seg += struct.pack("<QQ", 1, 16)    # one free block at the end: next = 1 (last), size = 16
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free, seg_addr))
        # symbol table nodes
        leaves = []                            # (address, heap offset of the largest name)
        for i in range(0, max(len(names), 1), 2 * K_LEAF):
            part = names[i:i + 2 * K_LEAF]
            ent = b""
            for n in part:
                c = grp.children[n]
                if isinstance(c, Group):
                    ent += struct.pack("<QQII", offs[n], c.addr, 1, 0) + struct.pack("<QQ", c.btree, c.heap)
                else:
                    ent += struct.pack("<QQII16x", offs[n], c.addr, 0, 0)
            ent += b"\0" * (40 * (2 * K_LEAF - len(part)))
            leaves.append((self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + ent), offs[part[-1]] if part else 0))
        # B-tree levels
        level, nodes = 0, leaves
        while True:
            parents = []
            for i in range(0, len(nodes), 2 * K_INTERNAL):
                part = nodes[i:i + 2 * K_INTERNAL]
                body = struct.pack("<Q", 0)
                for a, key in part:
                    body += struct.pack("<QQ", a, key)
                body += b"\0" * (16 * (2 * K_INTERNAL - len(part)))
                node = b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF) + body
                parents.append((self.alloc(node), part[-1][1]))
            if len(parents) == 1:
                grp.btree = parents[0][0]
                break
            nodes, level = parents, level + 1
        grp.heap = heap_addr
        grp.addr = self._ohdr([(0x0011, struct.pack("<QQ", grp.btree, grp.heap))])
        return grp.addr

    def tobytes(self):
        self.write_group(self.root)
        eof = len(self.buf)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, K_LEAF, K_INTERNAL, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, self.root.addr, 1, 0) + struct.pack("<QQ", self.root.btree, self.root.heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


# ---- Kiwi's database on top of it ---------------------------------------------------------------------------------
def scalar(dtype, fmt, value):
    return Dataset(dtype, [], struct.pack(fmt, value))


def write_kiwi_gfdb(base, nx, nz, ng, dt, dx, dz, firstx, firstz, nxc, traces, with_first=True):
    """traces: dict (ix, iz, ig) -> list of (first sample index, float32 array) strips, 1-based indices as in the Fortran."""
    nchunks = -(-nx // nxc)
    f = File()
    for name, val in (("dt", dt), ("dx", dx), ("dz", dz)) + ((("firstx", firstx), ("firstz", firstz)) if with_first else ()):
        f.root.children[name] = scalar(dt_float32(), "<f", val)
    for name, val in (("nchunks", nchunks), ("nx", nx), ("nxc", nxc), ("nz", nz), ("ng", ng)):
        f.root.children[name] = scalar(dt_int32(), "<i", val)
    with open(base + ".index", "wb") as fh:
        fh.write(f.tobytes())
    for ichunk in range(1, nchunks + 1):
        nxcthis = nx - (ichunk - 1) * nxc if ichunk == nchunks else nxc
        f = File()
        gf = Group()
        f.root.children["gf"] = gf
        placed = {}
        for ixc in range(1, nxcthis + 1):
            ix = (ichunk - 1) * nxc + ixc
            for iz in range(1, nz + 1):
                for ig in range(1, ng + 1):
                    strips = traces.get((ix, iz, ig))
                    if not strips:
                        continue
                    packed = np.concatenate([np.asarray(d, np.float32) for (_, d) in strips])
                    pofs = np.cumsum([1] + [len(d) for (_, d) in strips[:-1]]).astype(np.int32)
                    ofs = np.array([o for (o, _) in strips], np.int32)
                    ds = Dataset(dt_float32(), [packed.size], packed.tobytes(),
                                 [("pofs", dt_int32(), [len(strips)], pofs.tobytes()), ("ofs", dt_int32(), [len(strips)], ofs.tobytes())])
                    gx = gf.children.setdefault(str(ixc), Group())
                    gz = gx.children.setdefault(str(iz), Group())
                    gz.children[str(ig)] = ds
                    placed[(ixc, iz, ig)] = ds
        # datasets first, so that their addresses are known to the index of references
        f.write_group(gf)
        refs = np.zeros((nxcthis, nz, ng), np.uint64)
        for (ixc, iz, ig), ds in placed.items():
            refs[ixc - 1, iz - 1, ig - 1] = ds.addr
        f.root.children["index"] = Dataset(dt_objref(), [nxcthis, nz, ng], refs.tobytes())
        with open("%s.%d.chunk" % (base, ichunk), "wb") as fh:
            fh.write(f.tobytes())
