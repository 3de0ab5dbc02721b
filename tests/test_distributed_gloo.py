"""The distributed single-block path on CPU: world_size 2, 3 and 4 gloo processes run the executable
specification of csrc/dist_block.cu (tests/dist_model.py, numpy per rank) and the file must equal
the oracle's byte for byte.  Covers: key-only splitters (equal keys never straddle two ranks, also
when every key is equal and one rank owns the whole block), the request / response pull of
rank[(i+k) mod N], ranks going home, termination by fixed point, the MTF start-list hand-over, the
reduction of the Huffman statistics and the bit offsets of the payload pieces."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dist_model as M  # noqa: E402


def _with_repeats(rng):
    d = rng.integers(0, 256, 9000, dtype=np.uint8)
    d[4000:4300] = d[1000:1300]
    d[7000:7300] = d[1000:1300]
    return d


def _cases():
    rng = np.random.default_rng(3)
    return {
        "random": rng.integers(0, 256, 5000, dtype=np.uint8),
        "binary": rng.integers(0, 2, 4097, dtype=np.uint8),
        "all_a": np.full(3001, 97, dtype=np.uint8),                 # every key equal: one rank owns everything
        "abab": np.resize(np.frombuffer(b"ab", dtype=np.uint8), 2048),
        "abc_trunc": np.resize(np.frombuffer(b"abc", dtype=np.uint8), 1000),
        "a_then_b": np.concatenate([np.full(1500, 97, dtype=np.uint8), [98]]).astype(np.uint8),
        "tiny": np.frombuffer(b"banana", dtype=np.uint8).copy(),
        "one": np.frombuffer(b"x", dtype=np.uint8).copy(),
        "text": np.frombuffer((b"the quick brown fox jumps over the lazy dog " * 60), dtype=np.uint8).copy(),
        "repeats": _with_repeats(rng),
    }


def test_mtf_hand_over_composes():
    rng = np.random.default_rng(9)
    d = rng.integers(0, 40, 3000, dtype=np.uint8)
    import oracle_lib as O
    want = O.o_mtf(d)
    cuts = [0, 700, 701, 1900, 3000]
    lst = list(range(256))
    got = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        got.append(M.mtf_from(d[a:b], lst))
        lst = M.mtf_next_list(lst, M.mtf_summary(d[a:b]))
    assert np.array_equal(np.concatenate(got), want)


def test_splitters_are_deterministic_and_ordered():
    rng = np.random.default_rng(1)
    d = rng.integers(0, 256, 20000, dtype=np.uint8)
    for world in (1, 2, 5, 8):
        b = M.splitters(d, world)
        assert b == M.splitters(d.copy(), world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == 1 << 64
        assert all(x <= y for x, y in zip(b[:-1], b[1:]))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for name, d in _cases().items():
        st = {}
        blob = M.compress_block_distributed(d.copy(), None, st)
        out[name] = (None if blob is None else blob.tobytes(), st)
    dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_distributed_block_matches_oracle_over_gloo(world):
    import oracle_lib as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name, d in _cases().items():
        want = O.o_compress(d)
        blob, st = res[0][name]
        assert blob == want.tobytes(), name
        assert sum(st["counts"]) == d.size
        for r in range(1, world):
            assert res[r][name][0] is None
    # every key equal: key-only splitters leave the whole block with one rank (correct, not balanced)
    assert sorted(res[0]["all_a"][1]["counts"])[-1] == 3001
    # the random block is spread over all ranks
    assert min(res[0]["random"][1]["counts"]) > 0
