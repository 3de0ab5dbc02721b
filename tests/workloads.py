"""Deterministic inputs for the tests and the benchmark.

Nothing here is on the hot path: it only produces the bytes that are fed to it.

* Calgary corpus: the 14 files the reference ships as its test corpus
  (cmake-build-release/calgarycorpus/*), stored unmodified in tests/data/calgary.tar.xz.
* C3 text: whitespace-split words of book1, re-drawn with splitmix64 (seed 0x5EED0064).
* C4 degenerate / periodic blocks that stress rotation tie-breaking (main.cpp:46-59).
"""
import io
import lzma
import os
import tarfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/)
CALGARY_TAR = os.path.join(_ROOT, "tests", "data", "calgary.tar.xz")
CALGARY_FILES = ["bib", "book1", "book2", "geo", "news", "obj1", "obj2", "paper1", "paper2",
                 "pic", "progc", "progl", "progp", "trans"]  # order of main.cpp:418-419

_calgary_cache = None


def calgary():
    """dict name -> bytes of the 14 Calgary files."""
    global _calgary_cache
    if _calgary_cache is None:
        out = {}
        with tarfile.open(CALGARY_TAR, "r:xz") as tf:
            for m in tf.getmembers():
                out[m.name] = tf.extractfile(m).read()
        _calgary_cache = out
    return _calgary_cache


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64_stream(seed, count, start=0):
    """z_t for t = start .. start+count-1 of splitmix64 seeded with `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        t = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + t * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def synthetic_text(n, seed=0x5EED0064):
    """C3: English-like text of exactly n bytes (np.uint8 array).

    tokens = whitespace-split words of book1; emit token + b' ' with token index
    splitmix64(seed) mod n_tokens; truncate to n.
    """
    words = calgary()["book1"].split()
    n_tok = len(words)
    lens = np.array([len(w) + 1 for w in words], dtype=np.int64)
    offs = np.zeros(n_tok + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    flat = np.frombuffer(b"".join(w + b" " for w in words), dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    filled = 0
    t0 = 0
    batch = 1 << 20
    while filled < n:
        idx = (splitmix64_stream(seed, batch, t0) % np.uint64(n_tok)).astype(np.int64)
        t0 += batch
        l = lens[idx]
        ends = np.cumsum(l)
        total = int(ends[-1])
        starts = ends - l
        # gather: position p of the batch output belongs to token j = searchsorted(ends, p, 'right')
        src = np.repeat(offs[idx] - starts, l) + np.arange(total, dtype=np.int64)
        chunk = flat[src]
        take = min(total, n - filled)
        out[filled:filled + take] = chunk[:take]
        filled += take
    return out


def degenerate(kind, n):
    """C4 inputs: np.uint8 array of length n."""
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "a":
        return np.full(n, ord("a"), dtype=np.uint8)
    if kind == "ab":
        return np.resize(np.frombuffer(b"ab", dtype=np.uint8), n)
    if kind == "ba":
        return np.resize(np.frombuffer(b"ba", dtype=np.uint8), n)
    if kind == "abc":
        return np.resize(np.frombuffer(b"abc", dtype=np.uint8), n)
    if kind == "abcdefgh":
        return np.resize(np.frombuffer(b"abcdefgh", dtype=np.uint8), n)
    if kind == "a_then_b":
        x = np.full(n, ord("a"), dtype=np.uint8)
        x[-1] = ord("b")
        return x
    if kind == "rand256":
        blk = np.random.default_rng(4).integers(0, 256, 256, dtype=np.uint8)
        return np.resize(blk, n)
    if kind == "rand4k":
        blk = np.random.default_rng(5).integers(0, 256, 4096, dtype=np.uint8)
        return np.resize(blk, n)
    if kind == "bytes256":
        return np.resize(np.arange(256, dtype=np.uint8), n)
    if kind == "random":
        return np.random.default_rng(7).integers(0, 256, n, dtype=np.uint8)
    raise ValueError(kind)


DEGENERATE_KINDS = ["zeros", "a", "ab", "ba", "abc", "a_then_b", "abcdefgh", "rand256", "rand4k", "bytes256"]
