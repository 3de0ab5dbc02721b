"""B200-native BWT + move-to-front + Huffman codec: Python mirror of the reference interface.

The product is ``libbzap.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/bzap.h``).  This module is the thin ctypes binding a Python caller -- the tests and
``bench.py`` -- uses.  Function names and argument meaning mirror the reference's free functions
(``/root/reference/main.cpp``), so a parity test reads like the reference's own pipeline:

    compress(initial_file_name, encoded_file_name)         main.cpp:300-325
    decompress(encoded_file_name, decoded_file_name)       main.cpp:327-345
    bwt(data) -> (shift_position, encoded)                 main.cpp:77-91
    bwt_reverse(bwt_data, row_index)                       main.cpp:61-75
    move_to_front(data) / move_to_front_reverse(data)      main.cpp:93-130
    huffman(data) -> (encoded_data, tree)                  main.cpp:229-257
    encode_with_huffman(data, tree)                        main.cpp:158-172
    tree_to_bytes(tree) / bytes_to_tree(encoded_tree)      main.cpp:174-227
    huffman_reverse(data, tree, initial_data_size)         main.cpp:259-281

There is no CPU fallback: every call raises ``BzapError`` if the library is missing or no CUDA
device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbzap.so")

BZAP_OK = 0
ERR_IO, ERR_EMPTY, ERR_CORRUPT, ERR_CUDA, ERR_CAPACITY, ERR_ARG, ERR_TOO_LARGE, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7, -8


class BzapError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = "bzap error %d" % code
        try:
            msg += " (%s)" % lib().bzap_strerror(code).decode()
        except Exception:
            pass
        if detail:
            msg += ": " + detail
        super().__init__(msg)


class Tree(C.Structure):
    """bzap_tree of include/bzap.h (array form of the reference's BTree, main.cpp:13-26)."""
    _fields_ = [("n_leaves", C.c_int32), ("n_nodes", C.c_int32), ("root", C.c_int32),
                ("left", C.c_int32 * 511), ("right", C.c_int32 * 511), ("value", C.c_uint8 * 511)]


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_bwt", C.c_double), ("ms_mtf", C.c_double), ("ms_huffman", C.c_double),
                ("kernel_launches", C.c_uint64), ("bwt_rounds", C.c_uint32), ("bwt_sort_passes", C.c_uint32),
                ("decode_sync_iters", C.c_uint32), ("bwt_full_passes", C.c_uint32), ("payload_bytes", C.c_uint64),
                ("ms_sort", C.c_double), ("sort_bytes", C.c_uint64), ("sort_elems", C.c_uint64),
                ("ms_walk", C.c_double), ("walk_bytes", C.c_uint64)]


class DistStats(C.Structure):
    """bzap_dist_stats of include/bzap.h."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("rounds", C.c_uint32), ("peer_windows", C.c_uint32),
                ("own_rotations", C.c_uint64), ("exchanged_bytes", C.c_uint64), ("ms_total", C.c_double),
                ("ms_select_sort", C.c_double), ("ms_home", C.c_double), ("ms_rounds", C.c_double), ("ms_pull", C.c_double),
                ("ms_round_sort", C.c_double), ("ms_tail", C.c_double)]


COMM_ID_BYTES = 128
_u8p = C.POINTER(C.c_uint8)
_lib = None


def lib():
    """Loads libbzap.so; fails loudly when it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libbzap.so not built: run `make -C %s` (no CPU fallback exists)" % os.path.join(_HERE, "csrc"))
    L = C.CDLL(LIB_PATH)
    vp, sz, szp = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
    sig = {
        "bzap_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "bzap_ctx_destroy": (None, [vp]),
        "bzap_ctx_set_stream": (C.c_int, [vp, vp]),
        "bzap_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "bzap_strerror": (C.c_char_p, [C.c_int]),
        "bzap_last_error": (C.c_char_p, [vp]),
        "bzap_version": (C.c_char_p, []),
        "bzap_compress_file": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
        "bzap_decompress_file": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
        "bzap_compress_bound": (sz, [sz]),
        "bzap_compress": (C.c_int, [vp, vp, sz, vp, sz, szp]),
        "bzap_decompressed_size": (C.c_uint64, [vp, sz]),
        "bzap_decompress": (C.c_int, [vp, vp, sz, vp, sz, szp]),
        "bzap_compress_device": (C.c_int, [vp, vp, sz, vp, sz, szp]),
        "bzap_decompress_device": (C.c_int, [vp, vp, sz, vp, sz, szp]),
        "bzap_compress_batch": (C.c_int, [C.POINTER(vp), szp, C.POINTER(vp), szp, C.c_int, C.c_int]),
        "bzap_decompress_batch": (C.c_int, [C.POINTER(vp), szp, C.POINTER(vp), szp, szp, C.c_int, C.c_int]),
        "bzap_compress_batch_gpus": (C.c_int, [C.POINTER(vp), szp, C.POINTER(vp), szp, C.c_int, C.c_int, C.c_int]),
        "bzap_decompress_batch_gpus": (C.c_int, [C.POINTER(vp), szp, C.POINTER(vp), szp, szp, C.c_int, C.c_int, C.c_int]),
        "bzap_compress_files": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]),
        "bzap_decompress_files": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]),
        "bzap_bwt": (C.c_int, [vp, vp, sz, vp, C.POINTER(C.c_uint64)]),
        "bzap_ibwt": (C.c_int, [vp, vp, sz, C.c_uint64, vp]),
        "bzap_mtf": (C.c_int, [vp, vp, sz, vp]),
        "bzap_imtf": (C.c_int, [vp, vp, sz, vp]),
        "bzap_hist256": (C.c_int, [vp, vp, sz, C.POINTER(C.c_uint64), _u8p, C.POINTER(C.c_int)]),
        "bzap_huff_build": (C.c_int, [C.POINTER(C.c_uint64), _u8p, C.c_int, C.POINTER(Tree)]),
        "bzap_huff_codes": (C.c_int, [C.POINTER(Tree), C.POINTER(C.c_uint64), _u8p]),
        "bzap_tree_to_bytes": (C.c_int, [C.POINTER(Tree), _u8p, szp]),
        "bzap_bytes_to_tree": (C.c_int, [_u8p, sz, C.POINTER(Tree)]),
        "bzap_huff_encode": (C.c_int, [vp, vp, sz, C.POINTER(Tree), vp, sz, szp]),
        "bzap_huff_decode": (C.c_int, [vp, vp, sz, C.POINTER(Tree), sz, vp]),
        "bzap_comm_unique_id": (C.c_int, [_u8p]),
        "bzap_ctx_comm_init": (C.c_int, [vp, _u8p, C.c_int, C.c_int]),
        "bzap_compress_block_distributed": (C.c_int, [vp, vp, sz, vp, sz, szp]),
        "bzap_get_dist_stats": (C.c_int, [vp, C.POINTER(DistStats)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _u8(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    return np.frombuffer(bytes(x), dtype=np.uint8)


def _check(rc, ctx=None):
    if rc != BZAP_OK:
        detail = lib().bzap_last_error(ctx).decode(errors="replace") if rc != ERR_ARG or ctx else ""
        raise BzapError(rc, detail)


class Context:
    """A bzap_ctx: one CUDA stream plus scratch memory on one device (include/bzap.h)."""

    def __init__(self, device=-1):
        h = C.c_void_p()
        rc = lib().bzap_ctx_create(device, C.byref(h))
        if rc != BZAP_OK:
            raise BzapError(rc, "bzap_ctx_create: no usable CUDA device" if rc == ERR_CUDA else "")
        self.h = h

    def close(self):
        if self.h:
            lib().bzap_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(lib().bzap_ctx_set_stream(self.h, C.c_void_p(cuda_stream)), self.h)

    def stats(self):
        s = Stats()
        _check(lib().bzap_get_stats(self.h, C.byref(s)), self.h)
        return s

    # ---- one block over several GPUs (include/bzap.h: bzap_compress_block_distributed) ----
    def comm_init(self, comm_id, world, rank):
        """Collective: joins the NCCL communicator identified by the 128 bytes of comm_unique_id()."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(comm_id))
        _check(lib().bzap_ctx_comm_init(self.h, buf, world, rank), self.h)

    def compress_block_distributed(self, text_ptr, n, out_ptr, cap):
        """Collective: device pointers; the file lands in rank 0's out buffer (returns its length, 0 elsewhere)."""
        ln = C.c_size_t(0)
        _check(lib().bzap_compress_block_distributed(self.h, C.c_void_p(text_ptr), n, C.c_void_p(out_ptr), cap, C.byref(ln)), self.h)
        return ln.value

    def dist_stats(self):
        s = DistStats()
        _check(lib().bzap_get_dist_stats(self.h, C.byref(s)), self.h)
        return s

    # raw-pointer entry points used by bench.py (pinned host buffers / device buffers)
    def compress_ptr(self, in_ptr, n, out_ptr, cap, device=False):
        ln = C.c_size_t(0)
        f = lib().bzap_compress_device if device else lib().bzap_compress
        _check(f(self.h, C.c_void_p(in_ptr), n, C.c_void_p(out_ptr), cap, C.byref(ln)), self.h)
        return ln.value

    def decompress_ptr(self, in_ptr, n, out_ptr, cap, device=False):
        ln = C.c_size_t(0)
        f = lib().bzap_decompress_device if device else lib().bzap_decompress
        _check(f(self.h, C.c_void_p(in_ptr), n, C.c_void_p(out_ptr), cap, C.byref(ln)), self.h)
        return ln.value


def comm_unique_id():
    """128 bytes naming a new communicator (rank 0 creates it and hands it to the other ranks)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = lib().bzap_comm_unique_id(buf)
    if rc != BZAP_OK:
        raise BzapError(rc, "bzap_comm_unique_id")
    return bytes(buf)


_default_ctx = None


def _ctx(ctx=None):
    global _default_ctx
    if ctx is not None:
        return ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ---- file level ------------------------------------------------------------------------------------
def compress(initial_file_name, encoded_file_name, ctx=None):
    c = _ctx(ctx)
    _check(lib().bzap_compress_file(c.h, os.fsencode(initial_file_name), os.fsencode(encoded_file_name)), c.h)


def decompress(encoded_file_name, decoded_file_name, ctx=None):
    c = _ctx(ctx)
    _check(lib().bzap_decompress_file(c.h, os.fsencode(encoded_file_name), os.fsencode(decoded_file_name)), c.h)


# ---- buffer level ----------------------------------------------------------------------------------
def compress_bound(n):
    return lib().bzap_compress_bound(n)


def compress_bytes(data, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    out = np.empty(compress_bound(a.size), dtype=np.uint8)
    n = c.compress_ptr(a.ctypes.data, a.size, out.ctypes.data, out.size)
    return out[:n].copy()


def decompressed_size(blob):
    a = _u8(blob)
    return int(lib().bzap_decompressed_size(C.c_void_p(a.ctypes.data), a.size))


def decompress_bytes(blob, ctx=None):
    a = _u8(blob)
    n = decompressed_size(a)
    if n > (1 << 30):
        raise BzapError(ERR_TOO_LARGE, "header asks for %d bytes" % n)      # before allocating anything
    c = _ctx(ctx)
    out = np.empty(max(n, 1), dtype=np.uint8)
    got = c.decompress_ptr(a.ctypes.data, a.size, out.ctypes.data, n)
    return out[:got].copy()


def _batch(compress, blobs, sizes_out, n_streams, n_gpus):
    L = lib()
    k = len(blobs)
    arrs = [_u8(b) for b in blobs]
    outs = [np.empty(max(s, 1), dtype=np.uint8) for s in sizes_out]
    ins_p = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
    ns = (C.c_size_t * k)(*[a.size for a in arrs])
    outs_p = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
    lens = (C.c_size_t * k)()
    if compress:
        rc = L.bzap_compress_batch_gpus(ins_p, ns, outs_p, lens, k, n_gpus, n_streams)
    else:
        caps = (C.c_size_t * k)(*[int(s) for s in sizes_out])
        rc = L.bzap_decompress_batch_gpus(ins_p, ns, outs_p, caps, lens, k, n_gpus, n_streams)
    if rc != BZAP_OK:
        raise BzapError(rc, "bzap_%s_batch" % ("compress" if compress else "decompress"))
    return [o[:lens[i]].copy() for i, o in enumerate(outs)]


def compress_batch(datas, n_streams=0, n_gpus=0):
    """Independent inputs, one BWT block each (the reference's 14-file loop, main.cpp:424-437), over
    n_streams workers on each of n_gpus devices (0 = the current device)."""
    return _batch(True, datas, [compress_bound(len(_u8(d))) for d in datas], n_streams, n_gpus)


def decompress_batch(blobs, n_streams=0, n_gpus=0):
    sizes = [decompressed_size(b) for b in blobs]
    if any(s > (1 << 30) for s in sizes):
        raise BzapError(ERR_TOO_LARGE, "header asks for more than BZAP_MAX_BLOCK")     # before allocating
    return _batch(False, blobs, sizes, n_streams, n_gpus)


def batch_ptrs(compress, in_ptrs, in_lens, out_ptrs, out_caps, n_streams=0, n_gpus=0):
    """Raw-pointer form for pinned host buffers (bench.py): returns the output lengths."""
    k = len(in_ptrs)
    ins_p = (C.c_void_p * k)(*in_ptrs)
    ns = (C.c_size_t * k)(*in_lens)
    outs_p = (C.c_void_p * k)(*out_ptrs)
    lens = (C.c_size_t * k)()
    if compress:
        rc = lib().bzap_compress_batch_gpus(ins_p, ns, outs_p, lens, k, n_gpus, n_streams)
    else:
        caps = (C.c_size_t * k)(*out_caps)
        rc = lib().bzap_decompress_batch_gpus(ins_p, ns, outs_p, caps, lens, k, n_gpus, n_streams)
    if rc != BZAP_OK:
        raise BzapError(rc, "batch")
    return [lens[i] for i in range(k)]


def compress_files(in_paths, out_paths, n_gpus=0, n_streams=0):
    k = len(in_paths)
    a = (C.c_char_p * k)(*[os.fsencode(p) for p in in_paths])
    b = (C.c_char_p * k)(*[os.fsencode(p) for p in out_paths])
    rc = lib().bzap_compress_files(a, b, k, n_gpus, n_streams)
    if rc != BZAP_OK:
        raise BzapError(rc, "bzap_compress_files")


def decompress_files(in_paths, out_paths, n_gpus=0, n_streams=0):
    k = len(in_paths)
    a = (C.c_char_p * k)(*[os.fsencode(p) for p in in_paths])
    b = (C.c_char_p * k)(*[os.fsencode(p) for p in out_paths])
    rc = lib().bzap_decompress_files(a, b, k, n_gpus, n_streams)
    if rc != BZAP_OK:
        raise BzapError(rc, "bzap_decompress_files")


# ---- stage level: the reference's free functions -----------------------------------------------------
def bwt(data, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    out = np.empty(a.size, dtype=np.uint8)
    p = C.c_uint64(0)
    _check(lib().bzap_bwt(c.h, C.c_void_p(a.ctypes.data), a.size, C.c_void_p(out.ctypes.data), C.byref(p)), c.h)
    return int(p.value), out


def bwt_reverse(bwt_data, row_index, ctx=None):
    c = _ctx(ctx)
    a = _u8(bwt_data)
    out = np.empty(a.size, dtype=np.uint8)
    _check(lib().bzap_ibwt(c.h, C.c_void_p(a.ctypes.data), a.size, row_index, C.c_void_p(out.ctypes.data)), c.h)
    return out


def move_to_front(data, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    out = np.empty(a.size, dtype=np.uint8)
    _check(lib().bzap_mtf(c.h, C.c_void_p(a.ctypes.data), a.size, C.c_void_p(out.ctypes.data)), c.h)
    return out


def move_to_front_reverse(data, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    out = np.empty(a.size, dtype=np.uint8)
    _check(lib().bzap_imtf(c.h, C.c_void_p(a.ctypes.data), a.size, C.c_void_p(out.ctypes.data)), c.h)
    return out


def hist256(data, ctx=None):
    """(freq[256], first-appearance order of the present symbols): main.cpp:235-244."""
    c = _ctx(ctx)
    a = _u8(data)
    freq = (C.c_uint64 * 256)()
    order = (C.c_uint8 * 256)()
    nl = C.c_int(0)
    _check(lib().bzap_hist256(c.h, C.c_void_p(a.ctypes.data), a.size, freq, order, C.byref(nl)), c.h)
    return np.array(freq, dtype=np.uint64), np.array(order[:nl.value], dtype=np.uint8)


def huff_build(freq, order):
    f = np.ascontiguousarray(freq, dtype=np.uint64)
    o = _u8(order)
    t = Tree()
    _check(lib().bzap_huff_build(f.ctypes.data_as(C.POINTER(C.c_uint64)), o.ctypes.data_as(_u8p), o.size, C.byref(t)))
    return t


def huff_codes(tree):
    code = (C.c_uint64 * 256)()
    ln = (C.c_uint8 * 256)()
    _check(lib().bzap_huff_codes(C.byref(tree), code, ln))
    return np.array(code, dtype=np.uint64), np.array(ln, dtype=np.uint8)


def tree_to_bytes(tree):
    buf = (C.c_uint8 * 320)()
    ln = C.c_size_t(0)
    _check(lib().bzap_tree_to_bytes(C.byref(tree), buf, C.byref(ln)))
    return np.array(buf[:ln.value], dtype=np.uint8)


def bytes_to_tree(encoded_tree):
    a = _u8(encoded_tree)
    t = Tree()
    _check(lib().bzap_bytes_to_tree(a.ctypes.data_as(_u8p), a.size, C.byref(t)))
    return t


def encode_with_huffman(data, tree, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    cap = a.size * 8 + 64
    out = np.empty(cap, dtype=np.uint8)
    ln = C.c_size_t(0)
    _check(lib().bzap_huff_encode(c.h, C.c_void_p(a.ctypes.data), a.size, C.byref(tree), C.c_void_p(out.ctypes.data), cap,
                                  C.byref(ln)), c.h)
    return out[:ln.value].copy()


def huffman(data, ctx=None):
    """(encoded_data, tree) like the reference's huffman(): histogram -> tree -> bit packing."""
    freq, order = hist256(data, ctx)
    tree = huff_build(freq, order)
    return encode_with_huffman(data, tree, ctx), tree


def huffman_reverse(data, tree, initial_data_size, ctx=None):
    c = _ctx(ctx)
    a = _u8(data)
    out = np.empty(max(initial_data_size, 1), dtype=np.uint8)
    _check(lib().bzap_huff_decode(c.h, C.c_void_p(a.ctypes.data), a.size, C.byref(tree), initial_data_size,
                                  C.c_void_p(out.ctypes.data)), c.h)
    return out[:initial_data_size].copy()
