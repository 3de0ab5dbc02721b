#!/usr/bin/env python
"""Times the forward BWT / whole compress on a few inputs (device-resident) and prints the
library's own CUDA-event stats.  Used by tools/gpu_sweep.sh to compare kernel variants."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bwt_mtf_huffman_compressor_b200 as bz
import workloads as W

tag = sys.argv[1] if len(sys.argv) > 1 else ""
cases = {"text64m": W.synthetic_text(1 << 26)}
if "--more" in sys.argv:
    cases["random64m"] = W.degenerate("random", 1 << 26)
    cases["a_then_b16m"] = W.degenerate("a_then_b", 1 << 24)
    cases["rand4k16m"] = W.degenerate("rand4k", 1 << 24)
ctx = bz.Context(0)
for name, d in cases.items():
    n = d.size
    x = torch.from_numpy(d).cuda()
    out = torch.empty(bz.compress_bound(n), dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    best = None
    for it in range(4):
        fl = ctx.compress_ptr(x.data_ptr(), n, out.data_ptr(), out.numel(), device=True)
        s = ctx.stats()
        ctx.decompress_ptr(out.data_ptr(), fl, back.data_ptr(), n, device=True)
        sd = ctx.stats()
        row = (s.ms_total, s.ms_bwt, s.ms_sort, s.bwt_full_passes, s.sort_bytes, s.bwt_rounds, s.bwt_sort_passes, sd.ms_total, sd.ms_bwt, s.ms_mtf, sd.ms_mtf)
        if best is None or row[0] < best[0]:
            best = row
    ok = bool(torch.equal(back, x))
    gbs = best[4] / (best[2] * 1e-3) / 1e9 if best[2] > 0 else 0
    print("%-14s %-12s compress %.2f ms (bwt %.2f, full passes %d in %.2f ms = %.0f GB/s, rounds %d, passes %d) decompress %.2f ms (ibwt %.2f) mtf %.3f imtf %.3f roundtrip_ok=%s"
          % (tag, name, best[0], best[1], best[3], best[2], gbs, best[5], best[6], best[7], best[8], best[9], best[10], ok), flush=True)
