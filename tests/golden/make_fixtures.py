#!/usr/bin/env python3
"""Convert the reference's two HDF5 test fixtures into flat little-endian binaries.

Run ONCE in the build container (where /root/reference exists); the outputs are
committed under tests/golden/ because /root/reference does not exist on the GPU box.

    python tests/golden/make_fixtures.py

Inputs (reference test data):
  test/dataset/mBr_M_6w_1.likelihood-matrix.h5   written by rnaseq_sample.jl:505-519
  test/dataset/mBr_M_6w_1.prep.h5                written by likelihood-approximation.jl:61-87

No HDF5 library exists in this image, so this is a ~150-line reader of exactly the
HDF5 features those two files use (superblock v0, v1 object headers, compact link
messages or a single fractal-heap direct block, contiguous / chunked+deflate layouts).
"""
import hashlib
import json
import os
import struct
import sys
import zlib

import numpy as np

REF = "/root/reference/test/dataset"
OUT = os.path.dirname(os.path.abspath(__file__))


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
                o, l = struct.unpack_from("<QQ", body, 0)
                blocks.append((o, l))
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
    raise ValueError("unsupported datatype class %d" % cls)


def parse_dataspace(body):
    ver, rank = body[0], body[1]
    off = 8 if ver == 1 else 4
    return [struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank)]


def read_chunks(buf, addr, ndims, out):
    """v1 B-tree (node type 1) over deflate-compressed chunks."""
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
            out.append((offs[0], zlib.decompress(buf[child:child + nbytes])))
        else:
            read_chunks(buf, child, ndims, out)
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
        read_chunks(buf, bt, nd, chunks)
        arr = np.empty(count, dtype)
        for off, raw in chunks:
            a = np.frombuffer(raw if filtered else raw, dtype)
            nn = min(len(a), count - off, cdims[0])
            arr[off:off + nn] = a[:nn]
        return arr
    raise ValueError("layout class %d" % cls)


def root_links(buf):
    root = struct.unpack_from("<Q", buf, 56 + 8)[0]
    links = {}
    for t, b in read_object_header(buf, root):
        if t == 0x06:
            name, addr = parse_link(b)
            links[name] = addr
    if links:
        return links
    # dense link storage: scan the single fractal-heap direct block
    p = buf.find(b"FHDB")
    assert p >= 0
    end = len(buf)
    q = p
    while True:
        q = buf.find(b"\x01\x10\x01", q + 1)
        if q < 0 or q > end:
            break
        nlen = buf[q + 3]
        name = buf[q + 4:q + 4 + nlen]
        if 0 < nlen < 64 and name.isascii() and name.replace(b"_", b"a").isalnum():
            addr = struct.unpack_from("<Q", buf, q + 4 + nlen)[0]
            if addr < len(buf):
                links[name.decode()] = addr
    return links


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    manifest = {}

    def dump(name, arr):
        arr = np.ascontiguousarray(arr)
        path = os.path.join(OUT, name)
        arr.tofile(path)
        manifest[name] = {"dtype": str(arr.dtype), "count": int(arr.size), "sha256": sha(arr)}

    lm = open(os.path.join(REF, "mBr_M_6w_1.likelihood-matrix.h5"), "rb").read()
    links = root_links(lm)
    m = int(read_dataset(lm, links["m"])[0])
    n = int(read_dataset(lm, links["n"])[0])
    colptr = read_dataset(lm, links["colptr"])
    rowval = read_dataset(lm, links["rowval"])
    nzval = read_dataset(lm, links["nzval"])
    efflens = read_dataset(lm, links["effective_lengths"])
    assert colptr.dtype == np.uint32 and rowval.dtype == np.uint32
    assert nzval.dtype == np.float32 and efflens.dtype == np.float32
    assert len(colptr) == n + 1 and colptr[0] == 1 and colptr[-1] == len(rowval) + 1
    assert len(nzval) == len(rowval) and rowval.min() == 1 and rowval.max() == m
    dump("fixture_colptr.u32", colptr)
    dump("fixture_rowval.u32", rowval)
    dump("fixture_nzval.f32", nzval)
    dump("fixture_efflens.f32", efflens)

    pp = open(os.path.join(REF, "mBr_M_6w_1.prep.h5"), "rb").read()
    links = root_links(pp)
    assert int(read_dataset(pp, links["m"])[0]) == m
    assert int(read_dataset(pp, links["n"])[0]) == n
    eff2 = read_dataset(pp, links["effective_lengths"])
    assert eff2.tobytes() == efflens.tobytes()
    for key, dt in (("mu", np.float32), ("omega", np.float32), ("alpha", np.float32),
                    ("node_parent_idxs", np.int32), ("node_js", np.int32)):
        a = read_dataset(pp, links[key])
        assert a.dtype == dt, (key, a.dtype)
        dump("fixture_prep_%s.%s" % (key, "f32" if dt == np.float32 else "i32"), a)
    manifest["_meta"] = {"m": m, "n": n, "nnz": int(len(rowval)),
                         "source": ["test/dataset/mBr_M_6w_1.likelihood-matrix.h5",
                                    "test/dataset/mBr_M_6w_1.prep.h5"]}
    with open(os.path.join(OUT, "fixture_manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(manifest["_meta"]))
    for k, v in sorted(manifest.items()):
        if k != "_meta":
            print(k, v["count"], v["sha256"][:16])


if __name__ == "__main__":
    sys.exit(main())
