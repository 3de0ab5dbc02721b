"""(f-2) byte identity outside the windows where the closed-form tie-break law holds.

The reference breaks frequency ties by node ADDRESS.  bzap_heap_replay (csrc/heap_replay.c) predicts the address
order of the nodes for any (N, leaves) by replaying the reference's allocation script against this host's
allocator; with BZAP_HEAP_REPLAY=1 the host tree builder uses it for N < 64,600.  Checked here on CPU against
the unmodified reference binary: (1) the predicted order equals the order of the real BTree addresses (logged
through an LD_PRELOAD malloc tracer) in every failure window, (2) the tree bytes built by libbzap's host code
equal the tree bytes of the reference's file."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import bwt_mtf_huffman_compressor_b200 as bz
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(ROOT, "bwt_mtf_huffman_compressor_b200", "bzap_heap_replay")
TRACER = os.path.join(ROOT, "oracle", "_ref", "libmtrace.so")
# every failure window of App. B.3, their edges, and a few sizes where the law holds
SIZES = [1, 2, 3, 11, 100, 1000, 4097, 5000, 8000, 8192, 8193, 10200, 10300, 14400, 15000, 15300, 16384, 16500, 18000, 21100,
         25000, 30800, 31000, 31700, 38300, 40000, 43000, 50000, 63600, 64000, 64500, 64599]

pytestmark = pytest.mark.skipif(not (O.have_ref() and os.path.exists(TRACER) and os.path.exists(HELPER)),
                                reason="needs oracle/_ref (reference binary + malloc tracer) and the built helper")


def _data(n, alphabet, seed=1):
    return np.random.default_rng(seed + n).integers(0, alphabet, n, dtype=np.uint8)


def _reference_run(data, tmp):
    src = os.path.join(tmp, "in")
    data.tofile(src)
    env = dict(os.environ, LD_PRELOAD=TRACER, MTRACE_OUT=os.path.join(tmp, "addr"))
    subprocess.run([os.path.join(O.REF_DIR, "ref_compress"), src, src + ".bz"], env=env, check=True, stdout=subprocess.DEVNULL)
    addrs = [int(l, 16) for l in open(os.path.join(tmp, "addr"))]
    return [int(x) for x in np.argsort(np.argsort(addrs))], np.fromfile(src + ".bz", dtype=np.uint8)


@pytest.mark.parametrize("alphabet", [3, 100, 256])
def test_replay_predicts_the_reference_node_addresses(alphabet):
    with tempfile.TemporaryDirectory() as tmp:
        for n in SIZES:
            traced, _ = _reference_run(_data(n, alphabet), tmp)
            leaves = (len(traced) + 1) // 2
            out = subprocess.run([HELPER, str(n), str(leaves)], capture_output=True, text=True, check=True).stdout.split()
            assert [int(x) for x in out] == traced, (n, leaves)


def test_host_tree_builder_with_replay_equals_the_reference_in_every_window(monkeypatch):
    differs_without = 0
    with tempfile.TemporaryDirectory() as tmp:
        for n in SIZES:
            if n < 3:
                continue
            d = _data(n, 256)
            _, blob = _reference_run(d, tmp)
            _, _, ref_tree, _ = O.split_container(blob)
            last, _ = O.o_bwt(d)
            mtf = O.o_mtf(last)
            freq = np.bincount(mtf, minlength=256).astype(np.uint64)
            first = {}
            for i, s in enumerate(mtf.tolist()):
                first.setdefault(s, i)
            order = np.array(sorted(first, key=first.get), dtype=np.uint8)
            monkeypatch.setenv("BZAP_HEAP_REPLAY", "1")
            got = bz.tree_to_bytes(bz.huff_build(freq, order))
            assert np.array_equal(got, ref_tree), n
            monkeypatch.delenv("BZAP_HEAP_REPLAY")
            law = bz.tree_to_bytes(bz.huff_build(freq, order))
            differs_without += not np.array_equal(law, ref_tree)
    assert differs_without > 0        # the closed-form law alone does miss some of these sizes: the replay is what fixes them
