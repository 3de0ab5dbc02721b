"""One block compressed by several GPUs (bzap_compress_block_distributed, csrc/dist_block.cu) against the
oracle, byte for byte.  world_size 1 always runs (same code, exchanges become device copies); world_size
2 / 4 / 8 run when the box has that many GPUs: one process per GPU, the NCCL communicator is created and
owned by the library (bzap_comm_unique_id / bzap_ctx_comm_init); torch only carries the 128-byte id to
the other ranks (gloo) and owns the device buffers."""
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
    import workloads as W
    rng = np.random.default_rng(8)
    rep = rng.integers(0, 256, 90000, dtype=np.uint8)
    rep[40000:43000] = rep[10000:13000]
    rep[70000:73000] = rep[10000:13000]
    return {
        "text1m": W.synthetic_text(1 << 20),
        "text3m": W.synthetic_text(3 << 20, 0x5EED1024),
        "random300k": rng.integers(0, 256, 300007, dtype=np.uint8),
        "binary": rng.integers(0, 2, 70001, dtype=np.uint8),
        "repeats": rep,
        "all_a": W.degenerate("a", 200000),
        "zeros": W.degenerate("zeros", 5000),
        "ab": W.degenerate("ab", 131072),
        "abc": W.degenerate("abc", 100000),
        "a_then_b": W.degenerate("a_then_b", 100001),
        "rand4k": W.degenerate("rand4k", 1 << 18),
        "bytes256": W.degenerate("bytes256", 1 << 16),
        "book1": np.frombuffer(W.calgary()["book1"], dtype=np.uint8).copy(),
        "pic": np.frombuffer(W.calgary()["pic"], dtype=np.uint8).copy(),
        "tiny": np.frombuffer(b"mississippi", dtype=np.uint8).copy(),
        "one": np.frombuffer(b"x", dtype=np.uint8).copy(),
        "seven": np.frombuffer(b"abcabca", dtype=np.uint8).copy(),
    }


def make_comm(ctx, rank, world):
    """Creates the library-owned communicator; the id travels over the (gloo) process group."""
    import bwt_mtf_huffman_compressor_b200 as bz
    if world == 1:
        return
    box = [bz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], world, rank)


def _worker(rank, world, port, q, no_p2p=False):
    if no_p2p:
        os.environ["BZAP_DIST_NO_P2P"] = "1"      # bulk exchanges through ncclSend/ncclRecv groups instead of peer windows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bwt_mtf_huffman_compressor_b200 as bz
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = bz.Context(rank)
    make_comm(ctx, rank, world)
    out = {}
    for name, d in _inputs().items():
        text = torch.from_numpy(d.copy()).cuda()
        cap = bz.compress_bound(d.size)
        dst = torch.empty(cap if rank == 0 else 1, dtype=torch.uint8, device="cuda")
        ln = ctx.compress_block_distributed(text.data_ptr(), d.size, dst.data_ptr() if rank == 0 else 0, cap)
        s = ctx.dist_stats()
        assert s.world == world and s.rank == rank
        assert not (no_p2p and s.peer_windows)
        out[name] = dst[:ln].cpu().numpy().tobytes() if rank == 0 else None
    if world > 1:
        dist.barrier()
    q.put((rank, out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _run(world, no_p2p=False):
    import oracle_lib as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, no_p2p)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in procs)
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_distributed_block_world2_nccl_only():
    _run(2, no_p2p=True)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs four GPUs")
def test_distributed_block_world4():
    _run(4)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs eight GPUs")
def test_distributed_block_world8():
    _run(8)


def test_world1_in_process_matches_single_gpu_path():
    """same context type as every other entry point: no communicator = a world of one"""
    import bwt_mtf_huffman_compressor_b200 as bz
    import workloads as W
    d = W.synthetic_text(1 << 22)
    ctx = bz.Context(0)
    text = torch.from_numpy(d.copy()).cuda()
    cap = bz.compress_bound(d.size)
    dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ln = ctx.compress_block_distributed(text.data_ptr(), d.size, dst.data_ptr(), cap)
    ref = bz.compress_bytes(d, ctx)
    assert np.array_equal(dst[:ln].cpu().numpy(), ref)
    assert ctx.dist_stats().rounds >= 2
