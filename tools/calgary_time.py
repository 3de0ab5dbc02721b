#!/usr/bin/env python
"""Per-file latency of the Calgary corpus on one stream (host buffers), plus launches per file."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bwt_mtf_huffman_compressor_b200 as bz
import workloads as W
ctx = bz.Context(0)
cal = W.calgary()
tot_c = tot_d = 0.0
for name in W.CALGARY_FILES:
    d = np.frombuffer(cal[name], dtype=np.uint8)
    h_in = torch.from_numpy(d.copy()).pin_memory()
    h_out = torch.empty(bz.compress_bound(d.size), dtype=torch.uint8).pin_memory()
    h_back = torch.empty(d.size, dtype=torch.uint8).pin_memory()
    best_c = best_d = 1e9
    for it in range(5):
        l0 = ctx.stats().kernel_launches
        t0 = time.perf_counter()
        fl = ctx.compress_ptr(h_in.data_ptr(), d.size, h_out.data_ptr(), h_out.numel())
        t1 = time.perf_counter()
        s = ctx.stats()
        lc = s.kernel_launches - l0
        ctx.decompress_ptr(h_out.data_ptr(), fl, h_back.data_ptr(), d.size)
        t2 = time.perf_counter()
        ld = ctx.stats().kernel_launches - l0 - lc
        best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    tot_c += best_c; tot_d += best_d
    print("%-7s n=%7d compress %.3f ms (%3d rounds, %4d launches) decompress %.3f ms (%3d launches)" % (
        name, d.size, best_c * 1e3, s.bwt_rounds, lc, best_d * 1e3, ld), flush=True)
print("sum: compress %.2f ms (%.0f MB/s) decompress %.2f ms (%.0f MB/s)" % (tot_c * 1e3, 3141622 / tot_c / 1e6, tot_d * 1e3, 3141622 / tot_d / 1e6))
