"""CPU-only checks of the shipped library: it loads, exports every symbol include/bzap.h declares,
its host-side Huffman model agrees with the oracle, and it FAILS LOUDLY without a CUDA device
(no CPU fallback).  No compute entry point is exercised here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import bwt_mtf_huffman_compressor_b200 as bz
import oracle_lib as O
import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bzap.h")).read()
    declared = sorted(set(re.findall(r"\b(bzap_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = bz.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.bzap_version()


def test_library_contains_sm100a_code():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", bz.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_product_never_references_the_oracle():
    # the oracle is test infrastructure: no product source may include / load it
    pkg = os.path.join(ROOT, "bwt_mtf_huffman_compressor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "bzap_oracle" not in txt and "orc_" not in txt, f


def _mtf_stream(name, calgary):
    return O.o_mtf(O.o_bwt(calgary[name])[0])


@pytest.mark.parametrize("name", ["bib", "geo", "obj1", "pic", "progc", "trans", "book1"])
def test_host_tree_builder_matches_oracle_and_golden(name, calgary, golden):
    m = _mtf_stream(name, calgary)
    freq = np.bincount(m, minlength=256).astype(np.uint64)
    _, first = np.unique(m, return_index=True)
    order = m[np.sort(first)]
    tree = bz.huff_build(freq, order)
    tb = bz.tree_to_bytes(tree)
    assert tb.tobytes().hex() == golden["calgary"][name]["tree_hex"]
    assert np.array_equal(tb, O.o_tree_bytes(O.o_tree(m)))
    # code table == oracle's
    code, ln = bz.huff_codes(tree)
    lo, hi, oln = O.o_codes(O.o_tree(m))
    present = oln >= 0
    assert np.array_equal(ln[present], oln[present].astype(np.uint8))
    assert np.array_equal(code[present], lo[present])
    # parse(serialise(tree)) serialises back to the same bytes
    assert np.array_equal(bz.tree_to_bytes(bz.bytes_to_tree(tb)), tb)


def test_host_tree_builder_tie_break_law_all_alphabet_sizes():
    for k in list(range(1, 20)) + [63, 64, 65, 66, 127, 128, 129, 130, 191, 192, 193, 194, 254, 255, 256]:
        m = np.random.default_rng(k).permutation(np.resize(np.arange(k, dtype=np.uint8), 30000))
        freq = np.bincount(m, minlength=256).astype(np.uint64)
        _, first = np.unique(m, return_index=True)
        order = m[np.sort(first)]
        assert np.array_equal(bz.tree_to_bytes(bz.huff_build(freq, order)), O.o_tree_bytes(O.o_tree(m))), k


def test_host_tree_random_weights_match_oracle():
    rng = np.random.default_rng(3)
    for trial in range(40):
        k = int(rng.integers(1, 257))
        syms = rng.permutation(256)[:k].astype(np.uint8)
        freq = np.zeros(256, dtype=np.uint64)
        freq[syms] = rng.integers(1, 6 if trial % 2 else 100000, k)
        t = bz.huff_build(freq, syms)
        assert np.array_equal(bz.tree_to_bytes(t), O.o_tree_bytes(O.o_tree_from_hist(freq, syms)))


def test_tree_bytes_kats():
    # SURVEY App. A worked vectors
    one = bz.huff_build(np.eye(1, 256, ord("a"), dtype=np.uint64)[0] * 1000, np.array([ord("a")], dtype=np.uint8))
    assert bz.tree_to_bytes(one).tobytes() == bytes([0x30, 0x80])
    t = bz.bytes_to_tree(bytes([0x30, 0x80]))
    assert t.n_leaves == 1 and t.value[t.root] == ord("a")
    code, ln = bz.huff_codes(t)
    assert ln[ord("a")] == 0


def test_bytes_to_tree_rejects_garbage():
    for bad in [b"", b"\xff", b"\xff\xff\xff\xff", b"\x80"]:
        with pytest.raises(bz.BzapError) as e:
            bz.bytes_to_tree(bad)
        assert e.value.code == bz.ERR_CORRUPT


def test_deep_tree_code_longer_than_64_bits_is_refused_for_encoding():
    # Fibonacci-like weights give a 70-deep comb; decoding supports it, the 64-bit code table does not
    k = 72
    freq = np.zeros(256, dtype=np.uint64)
    a, b = 1, 1
    for s in range(k):
        freq[s] = a
        a, b = b, a + b
    t = bz.huff_build(freq, np.arange(k, dtype=np.uint8))
    code = (C.c_uint64 * 256)()
    ln = (C.c_uint8 * 256)()
    assert bz.lib().bzap_huff_codes(C.byref(t), code, ln) == bz.ERR_TOO_LARGE
    assert np.array_equal(bz.tree_to_bytes(t), O.o_tree_bytes(O.o_tree_from_hist(freq, np.arange(k, dtype=np.uint8))))


def test_bounds_and_header_helpers():
    assert bz.compress_bound(0) == 24 + 320 + 1
    assert bz.compress_bound(1000) == 24 + 320 + 1000
    blob = O.o_compress(b"hello world, hello world")
    assert bz.decompressed_size(blob) == 24
    assert bz.decompressed_size(b"short") == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_cuda():
    h = C.c_void_p()
    assert bz.lib().bzap_ctx_create(0, C.byref(h)) == bz.ERR_CUDA
    with pytest.raises(bz.BzapError) as e:
        bz.compress_bytes(b"abc")
    assert e.value.code == bz.ERR_CUDA
    out = np.zeros(400, dtype=np.uint8)
    ln = C.c_size_t(0)
    a = np.frombuffer(b"abc", dtype=np.uint8)
    assert bz.lib().bzap_compress(None, a.ctypes.data, 3, out.ctypes.data, 400, C.byref(ln)) == bz.ERR_CUDA


def test_batch_and_distributed_entry_points_fail_loudly_without_cuda(tmp_path):
    """no device: the batch engine and the distributed block report BZAP_ERR_CUDA / BZAP_ERR_ARG, nothing falls
    back to the CPU; the communicator id needs NCCL but no GPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bz.BzapError) as e:
        bz.compress_batch([b"abc", b"defg"], n_streams=2)
    assert e.value.code == bz.ERR_CUDA
    with pytest.raises(bz.BzapError) as e:
        bz.decompress_batch([bytes(30)], n_streams=1)
    assert e.value.code == bz.ERR_CUDA
    src = tmp_path / "in"
    src.write_bytes(b"hello")
    with pytest.raises(bz.BzapError) as e:
        bz.compress_files([str(src)], [str(src) + ".bz"], n_gpus=1)
    assert e.value.code == bz.ERR_CUDA
    assert not (tmp_path / "in.bz").exists()
    ln = C.c_size_t(7)
    assert bz.lib().bzap_compress_block_distributed(None, None, 5, None, 0, C.byref(ln)) == bz.ERR_ARG
    try:
        cid = bz.comm_unique_id()
    except bz.BzapError as err:                    # a host without libnccl.so.2: the rest of the library still loads
        assert err.code == -9
    else:
        assert len(cid) == bz.COMM_ID_BYTES and any(cid)


def test_header_that_asks_for_too_much_is_refused_before_allocating():
    blob = np.zeros(40, dtype=np.uint8)
    blob[8:16] = np.frombuffer(np.array([1 << 40], dtype="<u8").tobytes(), dtype=np.uint8)
    with pytest.raises(bz.BzapError) as e:
        bz.decompress_bytes(blob)
    assert e.value.code == bz.ERR_TOO_LARGE
    with pytest.raises(bz.BzapError) as e:
        bz.decompress_batch([blob])
    assert e.value.code == bz.ERR_TOO_LARGE


def test_cli_argument_contract():
    # main.cpp:440-443: wrong argc prints the message without newline and returns 1
    for exe in ("bzap_compress", "bzap_decompress"):
        p = subprocess.run([os.path.join(ROOT, "bwt_mtf_huffman_compressor_b200", exe)], capture_output=True, text=True)
        assert p.returncode == 1
        assert p.stdout == "Wrong arguments. Pass only input and output file as parameters"
