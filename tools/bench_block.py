#!/usr/bin/env python
"""One large block over N GPUs (BASELINE config 5 (ii)): distributed prefix-doubling BWT + the
remaining stages on rank 0.  Launch with torchrun; prints one JSON line on rank 0.
  python -m torch.distributed.run --nproc-per-node N tools/bench_block.py --size 268435456"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bwt_mtf_huffman_compressor_b200 as bz
from bwt_mtf_huffman_compressor_b200 import workloads as W, distributed as D

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1 << 28)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--check", action="store_true", help="compare with the single-GPU path on rank 0")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
data = W.synthetic_text(a.size, 0x5EED1024)
text = torch.from_numpy(data).cuda()
backend = D.GpuBackend()
best = None
for it in range(a.reps + 1):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    D.phase_times()
    blob, rounds = D.compress_block_distributed(text, None, backend)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if it > 0:
        best = dt if best is None else min(best, dt)
t = torch.tensor([best], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    line = {"workload": "one %d-byte text block, distributed prefix doubling" % a.size, "n_gpus": world, "rounds": rounds,
            "seconds": round(float(t[0]), 4), "MBps": round(a.size / float(t[0]) / 1e6, 1), "compressed_bytes": int(blob.numel()),
            "sha256": hashlib.sha256(blob.cpu().numpy().tobytes()).hexdigest()}
    if os.environ.get("BZAP_DIST_TIMING") == "1":
        line["phase_seconds_last_rep_rank0"] = {k: round(v, 4) for k, v in D.phase_times().items() if k}
    if a.check:
        ref = bz.compress_bytes(data)
        line["equals_single_gpu_path"] = bool(np.array_equal(ref, blob.cpu().numpy()))
    print(json.dumps(line), flush=True)
dist.destroy_process_group()
