"""ctypes access to the test-only oracle (oracle/liboracle.so) and to the real reference
(oracle/_ref: one-shot binaries + stage harness).  Test infrastructure, never product."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_u8p = C.POINTER(C.c_uint8)


def _ptr(a):
    return a.ctypes.data_as(_u8p)


def _as_u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    return np.ascontiguousarray(x, dtype=np.uint8)


class OrcTree(C.Structure):
    _fields_ = [("n_leaves", C.c_int), ("n_nodes", C.c_int), ("root", C.c_int),
                ("left", C.c_int * 511), ("right", C.c_int * 511), ("value", C.c_uint8 * 511)]


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)


_orc = None


def orc():
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_bwt.argtypes = [_u8p, C.c_size_t, _u8p, C.POINTER(C.c_uint64)]
        L.orc_ibwt.argtypes = [_u8p, C.c_size_t, C.c_uint64, _u8p]
        L.orc_mtf.argtypes = [_u8p, C.c_size_t, _u8p]
        L.orc_imtf.argtypes = [_u8p, C.c_size_t, _u8p]
        L.orc_huff_build.argtypes = [_u8p, C.c_size_t, C.POINTER(OrcTree)]
        L.orc_huff_build_from_hist.argtypes = [C.POINTER(C.c_uint64), _u8p, C.c_int, C.POINTER(OrcTree)]
        L.orc_tree_to_bytes.argtypes = [C.POINTER(OrcTree), _u8p]
        L.orc_tree_to_bytes.restype = C.c_size_t
        L.orc_bytes_to_tree.argtypes = [_u8p, C.c_size_t, C.POINTER(OrcTree)]
        L.orc_huff_encode.argtypes = [_u8p, C.c_size_t, C.POINTER(OrcTree), _u8p, C.c_size_t]
        L.orc_huff_encode.restype = C.c_size_t
        L.orc_huff_decode.argtypes = [_u8p, C.c_size_t, C.POINTER(OrcTree), C.c_size_t, _u8p]
        L.orc_compress.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_decompress.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_codes.argtypes = [C.POINTER(OrcTree), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        _orc = L
    return _orc


# ---- oracle wrappers (numpy in / numpy out) ----------------------------------------------
def o_bwt(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    p = C.c_uint64(0)
    rc = orc().orc_bwt(_ptr(a), a.size, _ptr(out), C.byref(p))
    assert rc == 0
    return out, int(p.value)


def o_ibwt(last, primary):
    a = _as_u8(last)
    out = np.empty_like(a)
    rc = orc().orc_ibwt(_ptr(a), a.size, primary, _ptr(out))
    assert rc == 0
    return out


def o_mtf(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    orc().orc_mtf(_ptr(a), a.size, _ptr(out))
    return out


def o_imtf(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    orc().orc_imtf(_ptr(a), a.size, _ptr(out))
    return out


def o_tree(mtf):
    a = _as_u8(mtf)
    t = OrcTree()
    rc = orc().orc_huff_build(_ptr(a), a.size, C.byref(t))
    assert rc == 0
    return t


def o_tree_from_hist(freq, order):
    f = np.ascontiguousarray(freq, dtype=np.uint64)
    o = _as_u8(order)
    t = OrcTree()
    rc = orc().orc_huff_build_from_hist(f.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(o), o.size, C.byref(t))
    assert rc == 0
    return t


def o_tree_bytes(t):
    buf = np.zeros(320, dtype=np.uint8)
    n = orc().orc_tree_to_bytes(C.byref(t), _ptr(buf))
    return buf[:n].copy()


def o_parse_tree(tree_bytes):
    a = _as_u8(tree_bytes)
    t = OrcTree()
    rc = orc().orc_bytes_to_tree(_ptr(a), a.size, C.byref(t))
    if rc != 0:
        raise ValueError("malformed tree")
    return t


def o_codes(t):
    lo = (C.c_uint64 * 256)()
    hi = (C.c_uint64 * 256)()
    ln = (C.c_int * 256)()
    orc().orc_codes(C.byref(t), lo, hi, ln)
    return np.array(lo, dtype=np.uint64), np.array(hi, dtype=np.uint64), np.array(ln, dtype=np.int32)


def o_encode(mtf, t):
    a = _as_u8(mtf)
    cap = a.size * 32 + 16
    out = np.zeros(cap, dtype=np.uint8)
    n = orc().orc_huff_encode(_ptr(a), a.size, C.byref(t), _ptr(out), cap)
    assert n > 0
    return out[:n].copy()


def o_decode(payload, t, n):
    a = _as_u8(payload)
    out = np.empty(n, dtype=np.uint8)
    rc = orc().orc_huff_decode(_ptr(a), a.size, C.byref(t), n, _ptr(out))
    if rc != 0:
        raise ValueError("decode failed")
    return out


def o_compress(data):
    a = _as_u8(data)
    cap = 24 + 320 + max(1, a.size) + 16
    out = np.zeros(cap, dtype=np.uint8)
    ln = C.c_size_t(0)
    rc = orc().orc_compress(_ptr(a), a.size, _ptr(out), cap, C.byref(ln))
    if rc != 0:
        raise ValueError("orc_compress rc=%d" % rc)
    return out[:ln.value].copy()


def o_decompress(blob):
    a = _as_u8(blob)
    n = int(np.frombuffer(a[8:16].tobytes(), dtype="<u8")[0])
    out = np.empty(max(n, 1), dtype=np.uint8)
    ln = C.c_size_t(0)
    rc = orc().orc_decompress(_ptr(a), a.size, _ptr(out), n, C.byref(ln))
    if rc != 0:
        raise ValueError("orc_decompress rc=%d" % rc)
    return out[:ln.value].copy()


# ---- the real reference -------------------------------------------------------------------
def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("ref_compress", "ref_decompress", "libref_stages.so"))


def ref_compress(data):
    """One fresh process per input: the byte-identity oracle."""
    a = _as_u8(data)
    with tempfile.TemporaryDirectory() as d:
        i, o = os.path.join(d, "in"), os.path.join(d, "out")
        a.tofile(i)
        subprocess.run([os.path.join(REF_DIR, "ref_compress"), i, o], check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(o, dtype=np.uint8)


def ref_decompress(blob):
    a = _as_u8(blob)
    with tempfile.TemporaryDirectory() as d:
        i, o = os.path.join(d, "in"), os.path.join(d, "out")
        a.tofile(i)
        subprocess.run([os.path.join(REF_DIR, "ref_decompress"), i, o], check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(o, dtype=np.uint8)


_ref = None


def refst():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_stages.so"))
        L.ref_bwt.argtypes = [_u8p, C.c_size_t, _u8p, C.POINTER(C.c_uint64)]
        L.ref_ibwt.argtypes = [_u8p, C.c_size_t, C.c_uint64, _u8p]
        L.ref_mtf.argtypes = [_u8p, C.c_size_t, _u8p]
        L.ref_imtf.argtypes = [_u8p, C.c_size_t, _u8p]
        L.ref_huff_decode.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_size_t, _u8p]
        L.ref_huff_encode_with_tree.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, _u8p, C.c_size_t]
        L.ref_huff_encode_with_tree.restype = C.c_size_t
        L.ref_tree_roundtrip.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t]
        L.ref_tree_roundtrip.restype = C.c_size_t
        _ref = L
    return _ref


def r_bwt(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    p = C.c_uint64(0)
    refst().ref_bwt(_ptr(a), a.size, _ptr(out), C.byref(p))
    return out, int(p.value)


def r_ibwt(last, primary):
    a = _as_u8(last)
    out = np.empty_like(a)
    refst().ref_ibwt(_ptr(a), a.size, primary, _ptr(out))
    return out


def r_mtf(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    refst().ref_mtf(_ptr(a), a.size, _ptr(out))
    return out


def r_imtf(data):
    a = _as_u8(data)
    out = np.empty_like(a)
    refst().ref_imtf(_ptr(a), a.size, _ptr(out))
    return out


def r_huff_decode(tree_bytes, payload, n):
    t, p = _as_u8(tree_bytes), _as_u8(payload)
    out = np.empty(n, dtype=np.uint8)
    refst().ref_huff_decode(_ptr(t), t.size, _ptr(p), p.size, n, _ptr(out))
    return out


def r_huff_encode_with_tree(tree_bytes, mtf):
    t, a = _as_u8(tree_bytes), _as_u8(mtf)
    cap = a.size * 32 + 16
    out = np.zeros(cap, dtype=np.uint8)
    n = refst().ref_huff_encode_with_tree(_ptr(t), t.size, _ptr(a), a.size, _ptr(out), cap)
    return out[:n].copy()


def r_tree_roundtrip(tree_bytes):
    t = _as_u8(tree_bytes)
    out = np.zeros(400, dtype=np.uint8)
    n = refst().ref_tree_roundtrip(_ptr(t), t.size, _ptr(out), 400)
    return out[:n].copy()


def split_container(blob):
    """(primary, n, tree_bytes, payload) of a reference-format file (io_utilities.h:44-50)."""
    a = _as_u8(blob)
    primary, n, tb = (int(x) for x in np.frombuffer(a[:24].tobytes(), dtype="<u8"))
    return primary, n, a[24:24 + tb].copy(), a[24 + tb:].copy()
