#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}.json 2> gpurun_out/bench_${N}.err
echo "bench N=$N rc=$?"; cut -c1-300 gpurun_out/bench_${N}.json; grep -v "^\s*$" gpurun_out/bench_${N}.err | grep -i "error" | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_block.py --size 1073741824 --reps 1 > gpurun_out/block1g_${N}.json 2> gpurun_out/block1g_${N}.err
echo "block1g rc=$?"; cat gpurun_out/block1g_${N}.json; grep -v "^\s*$" gpurun_out/block1g_${N}.err | grep -i "error" | head -5
python - <<'PY' > gpurun_out/single1g.txt 2>&1
import sys, os, time, hashlib
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bwt_mtf_huffman_compressor_b200 as bz
from bwt_mtf_huffman_compressor_b200 import workloads as W
n = 1 << 30
d = W.synthetic_text(n, 0x5EED1024)
ctx = bz.Context(0)
x = torch.from_numpy(d).cuda()
out = torch.empty(bz.compress_bound(n), dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fl = ctx.compress_ptr(x.data_ptr(), n, out.data_ptr(), out.numel(), device=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.decompress_ptr(out.data_ptr(), fl, back.data_ptr(), n, device=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
print("single GPU 1 GiB: compress %.3f s (%.0f MB/s) decompress %.3f s (%.0f MB/s) roundtrip_ok=%s bytes=%d sha256=%s" % (
    t1 - t0, n / (t1 - t0) / 1e6, t2 - t1, n / (t2 - t1) / 1e6, bool(torch.equal(back, x)), fl,
    hashlib.sha256(out[:fl].cpu().numpy().tobytes()).hexdigest()))
PY
cat gpurun_out/single1g.txt | tail -3
