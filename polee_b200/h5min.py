"""h5min -- the few HDF5 features Polee's two file formats use, in pure Python.

No HDF5 library exists in the build image, and the device library deliberately stops at the C ABI (writing `.prep.h5`
stays the reference's `write_approximation`).  The harness still has to move data across that boundary:

  * read_likelihood_matrix(path)   `--likelihood-matrix` dumps      (src/rnaseq_sample.jl:505-519 writes them,
                                    :34-47 reads them)               -> RNASeqSample fields
  * read_prep(path)                 a prepared sample `.prep.h5`     (src/likelihood-approximation.jl:61-87)
  * write_prep(path, ...)           the same file, written here      (datasets n, m, effective_lengths, every
                                    parameter vector; group `metadata` with the version / provenance attributes)

Reader: superblock v0, v1 object headers (+ continuation blocks), groups stored as link messages, as a single
fractal-heap direct block (dense storage) or as old-style symbol tables, contiguous and chunked + deflate layouts,
fixed-point / floating-point / fixed-length string types, attributes.  That covers the reference's own test files
(test/dataset/*.h5, written by HDF5.jl) and the files write_prep produces.

Writer: the most conservative encoding of the format specification (readable by every libhdf5 since 1.0): superblock
v0, old-style groups (v1 B-tree + local heap + one symbol-table node), v1 object headers, contiguous little-endian
datasets, fixed-length NUL-terminated string attributes.  It is checked by round-tripping through the reader and by
structural assertions on the bytes (tests/test_host_logic.py); libhdf5 itself is not available here to read it back,
which INTEGRATION.md states.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


# ------------------------------------------------------------------ reading
def read_object_header(buf, addr):
    """v1 object header -> list of (type, body); follows continuation messages."""
    ver, _, nmsgs, _refc, hsize = struct.unpack_from("<BBHII", buf, addr)
    assert ver == 1, ver
    msgs, blocks = [], [(addr + 16, hsize)]
    while blocks and len(msgs) < nmsgs:
        off, size = blocks.pop(0)
        end = off + size
        while off + 8 <= end and len(msgs) < nmsgs:
            t, s, _fl = struct.unpack_from("<HHB", buf, off)
            body = buf[off + 8: off + 8 + s]
            msgs.append((t, body))
            if t == 0x10:
                o, ln = struct.unpack_from("<QQ", body, 0)
                blocks.append((o, ln))
            off += 8 + s
    return msgs


def parse_link(body):
    """link message v1 -> (name, object header address) for hard links."""
    ver, flags = body[0], body[1]
    assert ver == 1
    p = 2
    ltype = 0
    if flags & 0x08:
        ltype = body[p]; p += 1
    if flags & 0x04:
        p += 8
    if flags & 0x10:
        p += 1
    w = 1 << (flags & 3)
    nlen = int.from_bytes(body[p:p + w], "little"); p += w
    name = body[p:p + nlen].decode(); p += nlen
    assert ltype == 0
    return name, struct.unpack_from("<Q", body, p)[0]


def parse_dtype(body):
    cls = body[0] & 0x0F
    size = struct.unpack_from("<I", body, 4)[0]
    if cls == 0:
        signed = bool(body[1] & 0x08)
        return np.dtype(("<i" if signed else "<u") + str(size))
    if cls == 1:
        return np.dtype("<f" + str(size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise ValueError("unsupported datatype class %d" % cls)


def parse_dataspace(body):
    ver, rank = body[0], body[1]
    off = 8 if ver == 1 else 4
    return [struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank)]


def read_chunks(buf, addr, ndims, out, filtered=True):
    """v1 B-tree (node type 1) over (deflate-compressed) chunks."""
    assert buf[addr:addr + 4] == b"TREE", buf[addr:addr + 4]
    ntype, level, nent = struct.unpack_from("<BBH", buf, addr + 4)
    assert ntype == 1
    p = addr + 24
    keysz = 8 + 8 * ndims
    for _ in range(nent):
        nbytes, _mask = struct.unpack_from("<II", buf, p)
        offs = struct.unpack_from("<%dQ" % ndims, buf, p + 8)
        child = struct.unpack_from("<Q", buf, p + keysz)[0]
        if level == 0:
            raw = buf[child:child + nbytes]
            out.append((offs[0], zlib.decompress(raw) if filtered else raw))
        else:
            read_chunks(buf, child, ndims, out, filtered)
        p += keysz + 8


def read_dataset(buf, addr):
    msgs = read_object_header(buf, addr)
    dims = dtype = layout = None
    filtered = False
    for t, b in msgs:
        if t == 0x01:
            dims = parse_dataspace(b)
        elif t == 0x03:
            dtype = parse_dtype(b)
        elif t == 0x08:
            layout = b
        elif t == 0x0B:
            filtered = True
    count = int(np.prod(dims)) if dims else 1
    assert layout[0] == 3
    cls = layout[1]
    if cls == 1:
        a, _sz = struct.unpack_from("<QQ", layout, 2)
        assert not filtered
        return np.frombuffer(buf, dtype, count, a).copy()
    if cls == 2:
        nd = layout[2]
        bt = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, layout, 11)
        chunks = []
        read_chunks(buf, bt, nd, chunks, filtered)
        arr = np.empty(count, dtype)
        for off, raw in chunks:
            a = np.frombuffer(raw, dtype)
            nn = min(len(a), count - off, cdims[0])
            arr[off:off + nn] = a[:nn]
        return arr
    raise ValueError("layout class %d" % cls)


def _symbol_table_links(buf, btree, heap):
    """old-style group: v1 B-tree (node type 0) -> SNODs; names in the local heap"""
    assert buf[heap:heap + 4] == b"HEAP"
    data = struct.unpack_from("<Q", buf, heap + 24)[0]
    links = {}

    def name_at(off):
        end = buf.index(b"\0", data + off)
        return buf[data + off:end].decode()

    def walk(addr):
        if buf[addr:addr + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", buf, addr + 6)[0]
            for i in range(nsym):
                noff, ohdr = struct.unpack_from("<QQ", buf, addr + 8 + 40 * i)
                links[name_at(noff)] = ohdr
            return
        assert buf[addr:addr + 4] == b"TREE" and buf[addr + 4] == 0
        nent = struct.unpack_from("<H", buf, addr + 6)[0]
        for i in range(nent):
            walk(struct.unpack_from("<Q", buf, addr + 24 + 8 + 16 * i)[0])

    walk(btree)
    return links


def group_links(buf, addr):
    """name -> object header address for the group whose object header is at addr"""
    links = {}
    for t, b in read_object_header(buf, addr):
        if t == 0x06:
            name, a = parse_link(b)
            links[name] = a
        elif t == 0x11:
            bt, hp = struct.unpack_from("<QQ", b, 0)
            links.update(_symbol_table_links(buf, bt, hp))
    return links


def root_links(buf):
    assert buf[:8] == SIGNATURE and buf[8] == 0, "superblock v0 expected"
    root = struct.unpack_from("<Q", buf, 56 + 8)[0]
    links = group_links(buf, root)
    if links:
        return links
    # dense link storage: scan the single fractal-heap direct block
    p = buf.find(b"FHDB")
    assert p >= 0
    q = p
    while True:
        q = buf.find(b"\x01\x10\x01", q + 1)
        if q < 0:
            break
        nlen = buf[q + 3]
        name = buf[q + 4:q + 4 + nlen]
        if 0 < nlen < 64 and name.isascii() and name.replace(b"_", b"a").isalnum():
            addr = struct.unpack_from("<Q", buf, q + 4 + nlen)[0]
            if addr < len(buf):
                links[name.decode()] = addr
    return links


def read_attributes(buf, addr):
    """attribute messages (v1) of the object header at addr -> {name: numpy scalar / array / str}"""
    out = {}
    for t, b in read_object_header(buf, addr):
        if t != 0x0C or b[0] != 1:
            continue
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
        pad = lambda x: (x + 7) & ~7  # noqa: E731
        p = 8
        name = b[p:p + nsz].split(b"\0")[0].decode(); p += pad(nsz)
        dtype = parse_dtype(b[p:p + tsz]); p += pad(tsz)
        dims = parse_dataspace(b[p:p + ssz]); p += pad(ssz)
        count = int(np.prod(dims)) if dims else 1
        val = np.frombuffer(b, dtype, count, p)
        if dtype.kind == "S":
            out[name] = val[0].split(b"\0")[0].decode()
        else:
            out[name] = val[0] if not dims else val.copy()
    return out


def read_likelihood_matrix(path):
    """-> dict(m, n, colptr, rowval, nzval, effective_lengths): the arrays of a SparseMatrixCSC{Float32,UInt32} exactly
    as Julia stores them (1-based), ready for RNASeqSample(...)  (src/rnaseq_sample.jl:34-47)"""
    buf = open(path, "rb").read()
    links = root_links(buf)
    out = {"m": int(read_dataset(buf, links["m"])[0]), "n": int(read_dataset(buf, links["n"])[0])}
    for k in ("colptr", "rowval", "nzval", "effective_lengths"):
        out[k] = read_dataset(buf, links[k])
    assert len(out["colptr"]) == out["n"] + 1 and out["colptr"][-1] == len(out["rowval"]) + 1
    return out


def read_prep(path):
    """-> dict of every root dataset of a .prep.h5 (scalars unwrapped) + "metadata": {attribute: value} when readable"""
    buf = open(path, "rb").read()
    links = root_links(buf)
    out = {}
    for k, a in links.items():
        if k == "metadata":
            try:
                out["metadata"] = read_attributes(buf, a)
            except Exception:      # HDF5.jl stores strings as variable-length (global heap): not needed by the harness
                out["metadata"] = {}
            continue
        v = read_dataset(buf, a)
        msgs = dict((t, b) for t, b in read_object_header(buf, a))
        out[k] = v[0] if not parse_dataspace(msgs[0x01]) else v
    return out


# ------------------------------------------------------------------ writing
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10 | 0, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        exp_bits, man_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        sign = 8 * dt.itemsize - 1
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, dt.itemsize)   # NUL-terminated, ASCII
    raise ValueError(dt)


def _dataspace_msg(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _File:
    def __init__(self):
        self.buf = bytearray(96)          # superblock goes in at the end

    def append(self, b):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += b
        return addr

    def dataset(self, arr):
        arr = np.asarray(arr)
        scalar = arr.ndim == 0
        data = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        daddr = self.append(data)
        msgs = [_msg(0x01, _dataspace_msg(() if scalar else arr.shape)), _msg(0x03, _dtype_msg(arr.dtype)),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                      # fill value v2: late alloc, undefined
                _msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, len(data)))]          # layout v3, contiguous
        return self.append(_object_header(msgs))

    def group(self, links, attributes=()):
        """links: {name: object header address}; old-style group = local heap + one SNOD + a one-entry B-tree"""
        names = sorted(links)
        heap_data = bytearray(b"\0" * 8)                                           # offset 0: the empty string
        offs = {}
        for nm in names:
            offs[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\0")
        data_addr = self.append(bytes(heap_data))
        heap_addr = self.append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))  # free list: none
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for nm in names:
            snod += struct.pack("<QQII16x", offs[nm], links[nm], 0, 0)
        snod += b"\0" * (40 * (2 * LEAF_K - len(names)))
        snod_addr = self.append(bytes(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            tree += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])             # key0, child0, key1
        tree += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree))
        tree_addr = self.append(bytes(tree))
        msgs = [_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))]
        for name, val in attributes:
            if isinstance(val, str):
                raw = val.encode() + b"\0"
                arr = np.frombuffer(raw, np.dtype("S%d" % len(raw)))[0]
                dt, data = np.dtype("S%d" % len(raw)), raw
            else:
                arr = np.asarray(val)
                dt, data = arr.dtype, arr.tobytes()
            nm = name.encode() + b"\0"
            tmsg, smsg = _dtype_msg(dt), _dataspace_msg(())
            msgs.append(_msg(0x0C, struct.pack("<BBHHH", 1, 0, len(nm), len(tmsg), len(smsg)) + _pad8(nm) + _pad8(tmsg)
                             + _pad8(smsg) + data))
        return self.append(_object_header(msgs)), tree_addr, heap_addr

    def finish(self, root):
        ohdr, tree, heap = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", tree, heap)   # root symbol table entry (cached)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


LEAF_K = 16        # symbol table nodes hold up to 2 * LEAF_K entries
INTERNAL_K = 16

PREPARED_SAMPLE_FORMAT_VERSION = 2   # src/constants.jl


def write_prep(path, params, n, m, effective_lengths, approximation="Polee.LogitSkewNormalPTTApprox", gfffilename="",
               gffhash="", fafilename="", fahash="", date="", args=""):
    """write_approximation (src/likelihood-approximation.jl:61-87): datasets n, m (Int64 scalars), effective_lengths
    (Float32[n]) and every key of `params` (mu / omega / alpha Float32[n-1], node_parent_idxs / node_js Int32[2n-1]);
    group `metadata` with attributes version, approximation, gfffilename, gffhash, fafilename, fahash, date, args."""
    f = _File()
    links = {"n": f.dataset(np.int64(n)), "m": f.dataset(np.int64(m)),
             "effective_lengths": f.dataset(np.asarray(effective_lengths, np.float32))}
    for key, val in params.items():
        links[key] = f.dataset(np.asarray(val))
    assert len(links) + 1 <= 2 * LEAF_K
    attrs = [("version", np.int64(PREPARED_SAMPLE_FORMAT_VERSION)), ("approximation", approximation),
             ("gfffilename", gfffilename), ("gffhash", gffhash), ("fafilename", fafilename), ("fahash", fahash),
             ("date", date), ("args", args)]
    links["metadata"] = f.group({}, attrs)[0]
    data = f.finish(f.group(links))
    with open(path, "wb") as fh:
        fh.write(data)
    return len(data)
