#!/usr/bin/env python
"""One large block over N GPUs (BASELINE config 5 ii) through bzap_compress_block_distributed.
Launch with torchrun (N > 1) or plain python (N = 1); prints one JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node N tools/bench_block.py --size 1073741824"""
import argparse, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
import bwt_mtf_huffman_compressor_b200 as bz
import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1 << 30)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--single", action="store_true", help="also time bzap_compress_device on rank 0")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = bz.Context(local)
if world > 1:
    box = [bz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], world, rank)
data = W.synthetic_text(a.size, 0x5EED1024)
text = torch.from_numpy(data).cuda()
cap = bz.compress_bound(a.size)
out = torch.empty(cap if rank == 0 else 1, dtype=torch.uint8, device="cuda")
best, stats = None, None
for it in range(a.reps + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ln = ctx.compress_block_distributed(text.data_ptr(), a.size, out.data_ptr() if rank == 0 else 0, cap)
    s = ctx.dist_stats()
    if it > 0 and (best is None or s.ms_total < best):
        best, stats = s.ms_total, s
t = torch.tensor([best], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = float(t[0]) * 1e-3
    line = {"workload": "one %d-byte text block (seed 0x5EED1024)" % a.size, "n_gpus": world, "rounds": stats.rounds,
            "seconds": round(sec, 4), "MBps": round(a.size / sec / 1e6, 1), "compressed_bytes": int(ln),
            "sha256": hashlib.sha256(out[:ln].cpu().numpy().tobytes()).hexdigest(),
            "rank0_ms": {k: round(getattr(stats, k), 2) for k in ("ms_select_sort", "ms_home", "ms_rounds", "ms_pull", "ms_round_sort", "ms_tail")},
            "rank0_own_rotations": int(stats.own_rotations), "rank0_sent_bytes": int(stats.exchanged_bytes)}
    g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden.json")))["text"].get(str(a.size))
    if g and g.get("seed") == "0x5EED1024":
        line["matches_golden"] = line["sha256"] == g["sha256"]
    if a.single:
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            l1 = ctx.compress_ptr(text.data_ptr(), a.size, d_out.data_ptr(), cap, device=True)
        line["single_gpu_ms"] = round(ctx.stats().ms_total, 2)
        line["equals_single_gpu_path"] = bool(l1 == ln and torch.equal(d_out[:l1], out[:ln]))
    print(json.dumps(line), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
