import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bwt_mtf_huffman_compressor_b200 as bz
from bwt_mtf_huffman_compressor_b200 import workloads as W
ctx = bz.Context(0)
for n in [(1 << 27) - 5, (1 << 27) + 12345, 1 << 28]:
    d = W.synthetic_text(n, 77)
    x = torch.from_numpy(d).cuda()
    out = torch.empty(bz.compress_bound(n), dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    try:
        t0 = time.time()
        fl = ctx.compress_ptr(x.data_ptr(), n, out.data_ptr(), out.numel(), device=True)
        s = ctx.stats()
        ctx.decompress_ptr(out.data_ptr(), fl, back.data_ptr(), n, device=True)
        torch.cuda.synchronize()
        print("n=%d compressed=%d rounds=%d roundtrip_ok=%s %.2fs" % (n, fl, s.bwt_rounds, bool(torch.equal(back, x)), time.time() - t0), flush=True)
    except Exception as e:
        print("n=%d FAILED %s" % (n, e), flush=True)
