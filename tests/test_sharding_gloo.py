"""N > 1 host logic on CPU: world_size-2 gloo processes shard the Calgary batch, process their
share (with the oracle standing in for the GPU, this is a test of the plumbing, not of the
product path), and agree on sizes and on the max-over-ranks timing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bwt_mtf_huffman_compressor_b200 import sharding  # noqa: E402
import workloads as W  # noqa: E402


def test_shard_files_is_a_balanced_partition():
    sizes = [len(W.calgary()[n]) for n in W.CALGARY_FILES]
    for world in (1, 2, 4, 8):
        shards = sharding.shard_files(sizes, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in s) for s in shards]
        assert max(loads) <= max(max(sizes), 1.34 * sum(sizes) / world)
    assert sharding.shard_files([], 3) == [[], [], []]
    assert sharding.shard_files([5], 2) == [[0], []]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    names = ["obj1", "progc", "paper1", "progp", "paper2"]
    cal = W.calgary()
    sizes = [len(cal[n]) for n in names]
    mine = sharding.shard_files(sizes, world)[rank]
    pairs = [(i, int(O.o_compress(cal[names[i]]).size)) for i in mine]
    all_sizes = sharding.gather_sizes(pairs, dist)
    slowest = sharding.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    q.put((rank, mine, all_sizes, slowest))
    dist.destroy_process_group()


def test_two_rank_file_batch_over_gloo(golden):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    names = ["obj1", "progc", "paper1", "progp", "paper2"]
    mine = {r: m for r, m, _, _ in res}
    assert sorted(mine[0] + mine[1]) == list(range(5)) and not set(mine[0]) & set(mine[1])
    for _, _, all_sizes, slowest in res:
        assert slowest == 2.0
        assert {names[i]: sz for i, sz in all_sizes.items()} == {n: golden["calgary"][n]["total"] for n in names}
