import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bwt_mtf_huffman_compressor_b200 as bz
from bwt_mtf_huffman_compressor_b200 import distributed as D, workloads as W
torch.cuda.set_device(0)
D._install_signatures()
be = D.GpuBackend()
for m in [(1 << 26) + 5, (1 << 27) + 7, 1 << 28]:
    text = torch.from_numpy(W.synthetic_text(m, 5)).cuda()
    keys = be.init_keys(text, m, 0, m)
    idx = torch.arange(0, m, dtype=torch.int32, device="cuda")
    ks, vs = be.sort_pairs(keys.clone(), idx.clone())
    torch.cuda.synchronize()
    bad_range = int(((vs < 0) | (vs >= m)).sum())
    # sortedness as unsigned: compare via (k ^ signbit)
    ku = ks ^ (-2 ** 63)
    unsorted = int((ku[1:] < ku[:-1]).sum())
    chk = int(vs.long().sum()) == m * (m - 1) // 2
    same_keys = bool(torch.equal(keys[vs.long().clamp(0, m - 1)], ks))
    print("m=%d out_of_range=%d unsorted=%d perm_sum_ok=%s keys_follow_vals=%s" % (m, bad_range, unsorted, chk, same_keys), flush=True)
    if bad_range:
        pos = torch.nonzero((vs < 0) | (vs >= m)).flatten()
        print("  first bad positions", pos[:5].tolist(), "count", pos.numel(), "min/max pos", int(pos.min()), int(pos.max()))
    rs, heads = be.rerank(ks, 0)
    print("  rerank heads", heads, "rs max", int(rs.max()), "rs[-1]", int(rs[-1]))
