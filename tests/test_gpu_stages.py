"""Stage-level parity on the GPU: every SURVEY 8a row, CUDA path (through the C ABI) vs the
oracle on the same inputs.  Bit-exact (integer / byte work: tolerance 0)."""
import numpy as np
import pytest

import bwt_mtf_huffman_compressor_b200 as bz
import oracle_lib as O
import workloads as W
from gpu_util import assert_same

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(11)


def small_inputs():
    cases = {
        "one": b"x", "two_same": b"aa", "two": b"ab", "abc": b"abc", "banana": b"banana",
        "seven": b"abcdefg", "eight": b"abcdefgh", "nine": b"abcdefghi",
        "aaaa": b"aaaa", "abcabc": b"abcabc", "ba1024": b"ba" * 1024, "cab300": b"cab" * 300,
        "zzzzy50": b"zzzzy" * 50, "banana40x": b"banana" * 40 + b"x", "bab100": b"b" + b"ab" * 100,
        "dcba": (b"dcba" * 2000)[:4098], "a1000": b"a" * 1000, "zeros4097": b"\0" * 4097,
        "rand3_5000": bytes(RNG.integers(0, 3, 5000, dtype=np.uint8)),
        "rand256_70001": bytes(RNG.integers(0, 256, 70001, dtype=np.uint8)),
        "rand2_33000": bytes(RNG.integers(0, 2, 33000, dtype=np.uint8)),
        "period7_x3000": bytes(RNG.integers(0, 256, 7, dtype=np.uint8)) * 3000,
    }
    return cases


SMALL = small_inputs()


@pytest.mark.parametrize("name", list(SMALL))
def test_bwt_small(name):
    d = SMALL[name]
    p, last = bz.bwt(d)
    ol, op = O.o_bwt(d)
    assert p == op, "%s primary %d want %d" % (name, p, op)
    assert_same(last, ol, name)


def test_bwt_kats(golden):
    for k, g in golden["bwt_kat"].items():
        p, last = bz.bwt(bytes.fromhex(g["input_hex"]))
        assert p == g["primary"], k


@pytest.mark.parametrize("name", W.CALGARY_FILES)
def test_bwt_calgary(name, calgary):
    p, last = bz.bwt(calgary[name])
    ol, op = O.o_bwt(calgary[name])
    assert p == op
    assert_same(last, ol, name)


@pytest.mark.parametrize("kind", W.DEGENERATE_KINDS)
@pytest.mark.parametrize("n", [16384, 65536 + 3, 1 << 20])
def test_bwt_degenerate(kind, n):
    d = W.degenerate(kind, n)
    p, last = bz.bwt(d)
    ol, op = O.o_bwt(d)
    assert p == op, "%s/%d primary %d want %d" % (kind, n, p, op)
    assert_same(last, ol, kind)


def test_bwt_active_rounds():
    # inputs that leave the "all rotations" regime early and then need many rounds on a small
    # active set: long repeats inside random data, half-periodic blocks, a run at the end
    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, 300000, dtype=np.uint8)
    rep = base.copy()
    rep[200000:250000] = rep[50000:100000]              # one 50 KB repeat (LCP up to 50000)
    rep[260000:270000] = rep[50000:60000]
    cases = {
        "long_repeat": rep,
        "random_then_ab": np.concatenate([rng.integers(0, 256, 60000, dtype=np.uint8), W.degenerate("ab", 40000)]),
        "random_then_zeros": np.concatenate([rng.integers(0, 256, 70001, dtype=np.uint8), np.zeros(29999, dtype=np.uint8)]),
        "two_copies_plus_one": np.concatenate([base[:100000], base[:100000], base[:1]]),
        "text_with_dup": np.concatenate([W.synthetic_text(200000), W.synthetic_text(200000)[:150000]]),
        "small_alphabet": rng.integers(0, 2, 200000, dtype=np.uint8),
    }
    for name, d in cases.items():
        p, last = bz.bwt(d)
        ol, op = O.o_bwt(d)
        assert p == op, "%s primary %d want %d" % (name, p, op)
        assert_same(last, ol, name)


def mtf_inputs(calgary):
    out = {k: np.frombuffer(v, dtype=np.uint8) for k, v in SMALL.items()}
    for name in ["obj1", "geo", "book1", "pic"]:
        out["bwt_" + name] = O.o_bwt(calgary[name])[0]
    out["rand_1m"] = RNG.integers(0, 256, 1 << 20, dtype=np.uint8)
    out["rand16_300k"] = RNG.integers(0, 16, 300001, dtype=np.uint8)
    out["bytes256_x"] = W.degenerate("bytes256", 100000)
    out["desc"] = np.resize(np.arange(255, -1, -1, dtype=np.uint8), 77777)
    return out


def test_mtf_and_inverse(calgary):
    for name, d in mtf_inputs(calgary).items():
        want = O.o_mtf(d)
        got = bz.move_to_front(d)
        assert_same(got, want, "mtf " + name)
        back = bz.move_to_front_reverse(want)
        assert_same(back, d, "imtf " + name)
        # inverse MTF of arbitrary index streams (not only valid MTF outputs)
        assert_same(bz.move_to_front_reverse(d), O.o_imtf(d), "imtf-raw " + name)


def test_hist_and_first_appearance(calgary):
    for name, d in mtf_inputs(calgary).items():
        m = O.o_mtf(d)
        freq, order = bz.hist256(m)
        assert_same(freq, np.bincount(m, minlength=256).astype(np.uint64), "freq " + name)
        _, first = np.unique(m, return_index=True)
        assert_same(order, m[np.sort(first)], "order " + name)


def test_huffman_encode_decode(calgary):
    for name, d in mtf_inputs(calgary).items():
        m = O.o_mtf(d)
        t = O.o_tree(m)
        want = O.o_encode(m, t)
        payload, tree = bz.huffman(m)
        assert_same(bz.tree_to_bytes(tree), O.o_tree_bytes(t), "tree " + name)
        assert_same(payload, want, "encode " + name)
        dec = bz.huffman_reverse(want, bz.bytes_to_tree(O.o_tree_bytes(t)), m.size)
        assert_same(dec, m, "decode " + name)


def test_decode_arbitrary_trees():
    # the decoder must accept any tree a reference build could have written, not only ours:
    # permute the leaf order fed to the builder (different tie-breaks), skewed weights (deep codes)
    for trial in range(6):
        k = [2, 3, 40, 200, 256, 30][trial]
        syms = RNG.permutation(256)[:k].astype(np.uint8)
        w = (1.6 ** np.arange(k)) if trial == 5 else RNG.integers(1, 1000, k)
        w = np.maximum(1, w / w.sum() * 200000).astype(np.int64)
        m = RNG.permutation(np.repeat(syms, w))
        freq = np.bincount(m, minlength=256).astype(np.uint64)
        t = O.o_tree_from_hist(freq, RNG.permutation(syms))
        tb = O.o_tree_bytes(t)
        enc = O.o_encode(m, t)
        dec = bz.huffman_reverse(enc, bz.bytes_to_tree(tb), m.size)
        assert_same(dec, m, "trial %d" % trial)
        # and the encoder with a foreign tree
        assert_same(bz.encode_with_huffman(m, bz.bytes_to_tree(tb)), enc, "enc trial %d" % trial)


def test_decode_non_synchronising_stream():
    # all code words 8 bits except two 9-bit ones that appear once at the start: every later
    # subsequence starts off the true boundary and only the fixed-point iteration fixes it
    syms = np.arange(255, dtype=np.uint8)
    m = np.concatenate([np.array([254, 253], dtype=np.uint8), RNG.integers(0, 253, 40000, dtype=np.uint8)])
    t = O.o_tree(m)
    enc = O.o_encode(m, t)
    dec = bz.huffman_reverse(enc, bz.bytes_to_tree(O.o_tree_bytes(t)), m.size)
    assert_same(dec, m, "nosync")


def test_ibwt(calgary):
    for name in ["obj1", "geo", "paper1", "book1", "pic"]:
        ol, op = O.o_bwt(calgary[name])
        got = bz.bwt_reverse(ol, op)
        assert_same(got, np.frombuffer(calgary[name], dtype=np.uint8), name)
    for name, d in SMALL.items():
        ol, op = O.o_bwt(d)
        assert_same(bz.bwt_reverse(ol, op), np.frombuffer(d, dtype=np.uint8), name)
    for kind in W.DEGENERATE_KINDS:
        d = W.degenerate(kind, 100003)
        ol, op = O.o_bwt(d)
        assert_same(bz.bwt_reverse(ol, op), d, kind)


def test_ibwt_on_arbitrary_columns_matches_reference_walk():
    # not a BWT of anything: the walk of main.cpp:70-73 is still well defined (cycle of primary,
    # length not dividing N) and must be reproduced exactly
    for n, k in [(1000, 4), (4097, 256), (70000, 3), (5, 2)]:
        col = RNG.integers(0, k, n, dtype=np.uint8)
        for primary in [0, n // 3, n - 1]:
            assert_same(bz.bwt_reverse(col, primary), O.o_ibwt(col, primary), "n=%d k=%d p=%d" % (n, k, primary))


def test_errors():
    with pytest.raises(bz.BzapError) as e:
        bz.bwt(b"")
    assert e.value.code == bz.ERR_EMPTY
    with pytest.raises(bz.BzapError) as e:
        bz.bwt_reverse(b"abc", 3)
    assert e.value.code == bz.ERR_CORRUPT
