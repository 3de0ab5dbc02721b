"""Pins the oracle restatement (oracle/bzap_oracle.c) against
 (a) the UNMODIFIED reference built into oracle/_ref (skipped where /root/reference was never
     available to build it) and
 (b) the committed golden vectors produced from that reference (tests/golden/golden.json).
CPU only."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
import workloads as W

need_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
SMALL = ["obj1", "progc", "paper1", "geo"]          # the reference's BWT is O(N^2 log N): keep it quick


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint8).tobytes()).hexdigest()


# ---- (b) golden vectors -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", W.CALGARY_FILES)
def test_oracle_compress_matches_calgary_golden(name, golden, calgary):
    g = golden["calgary"][name]
    blob = O.o_compress(calgary[name])
    assert blob.size == g["total"]
    assert sha(blob) == g["sha256"]
    primary, n, tree, payload = O.split_container(blob)
    assert (primary, n, tree.tobytes().hex(), payload.size) == (g["primary"], g["n"], g["tree_hex"], g["payload_bytes"])


def test_calgary_sizes_match_reference_readme(golden):
    # README.md:21-36 of the reference: encoded size and meta size (24 + tree bytes)
    readme = {"bib": (33205, 162), "book1": (267163, 161), "book2": (186994, 176), "geo": (69563, 344),
              "news": (133517, 178), "obj1": (11785, 344), "obj2": (88733, 344), "paper1": (18224, 167),
              "paper2": (28136, 173), "pic": (101508, 277), "progc": (13699, 172), "progl": (18745, 171),
              "progp": (12826, 169), "trans": (22400, 178)}
    for name, (total, meta) in readme.items():
        g = golden["calgary"][name]
        assert g["total"] == total and 24 + len(g["tree_hex"]) // 2 == meta


@pytest.mark.parametrize("name", W.CALGARY_FILES)
def test_oracle_decompress_roundtrip(name, calgary):
    data = calgary[name]
    assert O.o_decompress(O.o_compress(data)).tobytes() == data


def test_oracle_kats(golden):
    for k, g in golden["kat"].items():
        data = bytes.fromhex(g["input_hex"])
        ref_file = np.frombuffer(bytes.fromhex(g["file_hex"]), dtype=np.uint8)
        assert O.o_decompress(ref_file).tobytes() == data, k
        if g["oracle_matches_reference"]:
            assert O.o_compress(data).tobytes() == ref_file.tobytes(), k


def test_oracle_bwt_kats(golden):
    for k, g in golden["bwt_kat"].items():
        data = bytes.fromhex(g["input_hex"])
        last, p = O.o_bwt(data)
        assert p == g["primary"], k
        assert sha(last) == g["last_sha256"], k
        assert O.o_ibwt(last, p).tobytes() == data, k


@pytest.mark.parametrize("kind", W.DEGENERATE_KINDS)
def test_oracle_degenerate_16k(kind, golden):
    g = golden["degenerate_16k"][kind]
    blob = O.o_compress(W.degenerate(kind, 16384))
    assert g["oracle_matches_reference"]
    assert sha(blob) == g["sha256"]


def test_oracle_text_1m(golden):
    g = golden["text"][str(1 << 20)]
    d = W.synthetic_text(1 << 20)
    assert sha(d) == g["input_sha256"]          # the generator itself is pinned
    assert sha(O.o_compress(d)) == g["sha256"]


def test_bwt_against_naive_rotation_sort():
    rng = np.random.default_rng(1)
    cases = [b"banana", b"abab", b"aaaa", b"abcabcabc", b"mississippi", bytes(rng.integers(0, 3, 200, dtype=np.uint8)),
             bytes(rng.integers(0, 256, 500, dtype=np.uint8)), b"ab" * 50 + b"a"]
    for d in cases:
        n = len(d)
        rots = sorted(range(n), key=lambda i: (d[i:] + d[:i], i))     # ties by ascending start (stable_sort)
        last = bytes(d[(i + n - 1) % n] for i in rots)
        last_o, p = O.o_bwt(d)
        assert last_o.tobytes() == last and p == rots.index(0), d


# ---- (a) the real reference ------------------------------------------------------------------------------
@need_ref
@pytest.mark.parametrize("name", SMALL)
def test_oracle_stages_match_reference_functions(name, calgary):
    data = calgary[name]
    rl, rp = O.r_bwt(data)
    ol, op = O.o_bwt(data)
    assert rp == op and np.array_equal(rl, ol)
    rm = O.r_mtf(rl)
    assert np.array_equal(rm, O.o_mtf(ol))
    assert np.array_equal(O.r_imtf(rm), O.o_imtf(rm))
    assert np.array_equal(O.r_ibwt(rl, rp), O.o_ibwt(ol, op))
    t = O.o_tree(rm)
    tb = O.o_tree_bytes(t)
    assert np.array_equal(O.r_tree_roundtrip(tb), tb)
    enc = O.o_encode(rm, t)
    assert np.array_equal(O.r_huff_encode_with_tree(tb, rm), enc)
    assert np.array_equal(O.r_huff_decode(tb, enc, rm.size), O.o_decode(enc, O.o_parse_tree(tb), rm.size))


@need_ref
@pytest.mark.parametrize("name", SMALL)
def test_oracle_file_is_byte_identical_to_one_shot_reference(name, calgary):
    data = calgary[name]
    ref = O.ref_compress(data)
    assert np.array_equal(O.o_compress(data), ref)
    assert O.ref_decompress(O.o_compress(data)).tobytes() == data
    assert O.o_decompress(ref).tobytes() == data


@need_ref
def test_oracle_periodic_inputs_match_reference():
    for kind in ["a", "ab", "ba", "abc", "a_then_b", "bytes256", "rand256"]:
        d = W.degenerate(kind, 16384)
        assert np.array_equal(O.o_compress(d), O.ref_compress(d)), kind
    for d in [b"ba" * 1024, b"cab" * 300, b"zzzzy" * 50, (b"dcba" * 2000)[:4098]]:
        rl, rp = O.r_bwt(d)
        ol, op = O.o_bwt(d)
        assert rp == op and np.array_equal(rl, ol)


@need_ref
def test_oracle_tie_break_law_alphabet_sizes():
    # every alphabet size, all symbols in frequency ties, N = 30000 (valid window)
    for k in [1, 2, 3, 5, 17, 64, 65, 127, 128, 129, 130, 192, 193, 200, 255, 256]:
        d = np.resize(np.arange(k, dtype=np.uint8), 30000)
        d = np.random.default_rng(k).permutation(d)
        assert np.array_equal(O.o_compress(d), O.ref_compress(d)), k
