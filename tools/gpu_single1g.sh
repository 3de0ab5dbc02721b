#!/bin/bash
# one 1 GiB block on one GPU: compress / decompress time, round-trip check, sha256 of the file
mkdir -p gpurun_out
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
tail -3 gpurun_out/single1g.txt
