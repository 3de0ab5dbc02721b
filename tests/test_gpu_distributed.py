"""Distributed single-block path on real GPUs: GpuBackend (libbzap device-level ABI) + NCCL.
world_size 1 always runs (exercises every building block on one GPU); world_size 2 runs when the
box has two GPUs.  Results must equal the single-GPU path and the oracle, byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _inputs():
    from bwt_mtf_huffman_compressor_b200 import workloads as W
    rng = np.random.default_rng(8)
    return {
        "text1m": W.synthetic_text(1 << 20),
        "random300k": rng.integers(0, 256, 300007, dtype=np.uint8),
        "all_a": W.degenerate("a", 200000),
        "ab": W.degenerate("ab", 131072),
        "a_then_b": W.degenerate("a_then_b", 100001),
        "rand4k": W.degenerate("rand4k", 1 << 18),
        "book1": np.frombuffer(W.calgary()["book1"], dtype=np.uint8).copy(),
        "tiny": np.frombuffer(b"mississippi", dtype=np.uint8).copy(),
    }


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from bwt_mtf_huffman_compressor_b200 import distributed as D
    backend = D.GpuBackend()
    out = {}
    for name, d in _inputs().items():
        text = torch.from_numpy(d.copy()).cuda()
        blob, rounds = D.compress_block_distributed(text, None, backend)
        out[name] = None if blob is None else blob.cpu().numpy().tobytes()
    dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


def _run(world):
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
        p.join(timeout=120)
        assert p.exitcode == 0
    for name, d in _inputs().items():
        want = O.o_compress(d).tobytes()
        assert res[0][name] == want, name


def test_distributed_block_world1():
    _run(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_distributed_block_world2():
    _run(2)
