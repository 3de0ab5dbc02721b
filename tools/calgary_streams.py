import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bwt_mtf_huffman_compressor_b200 as bz
import workloads as W
cal = W.calgary()
datas = [np.frombuffer(cal[n], dtype=np.uint8) for n in W.CALGARY_FILES]
for ns in (1, 2, 4, 8, 14):
    bc = bd = 1e9
    for it in range(6):
        t0 = time.perf_counter(); blobs = bz.compress_batch(datas, ns); t1 = time.perf_counter()
        outs = bz.decompress_batch(blobs, ns); t2 = time.perf_counter()
        if it: bc, bd = min(bc, t1 - t0), min(bd, t2 - t1)
    print("streams=%2d compress %.2f ms (%.0f MB/s) decompress %.2f ms (%.0f MB/s)" % (ns, bc * 1e3, 3141622 / bc / 1e6, bd * 1e3, 3141622 / bd / 1e6), flush=True)
