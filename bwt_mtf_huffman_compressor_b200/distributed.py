"""ONE BWT block sorted over several GPUs: distributed cyclic prefix doubling.

Replaces bwt() (main.cpp:77-91) for a block that is spread over the GPUs of one box, one process
per GPU.  Same algorithm as csrc/bwt.cu -- sparse ranks, no index in any sort key, so equal
rotations keep equal rank and primary = rank[0] -- with the three exchange steps of a round done
as NCCL collectives over NVLink (torch.distributed) and every per-GPU step done by libbzap's
device-level C ABI (include/bzap.h, "device-level building blocks"):

  per round (prefix length k -> 2k), rank r owning text positions [lo_r, hi_r):
    1. shifted fetch   r2[i] = rank[(i + k) mod N]          all_to_all_single (<= 2 peers send)
    2. sample sort     keys (rank[i] << 32 | r2[i], i)      splitters from an all_gather of samples,
                       partition by (key, index) -- the index only balances the partition under
                       massive key duplication, it never enters the order of unequal keys --,
                       all_to_all_single of keys and indices, local onesweep sort
    3. re-rank         local sparse ranks + seam fix-up     all_gather of 5 words per rank
    4. ranks go home   (index, rank) to the owner of index  all_to_all_single, local scatter
  then L[j] = text[(SA[j] - 1) mod N] per rank, gathered on rank 0, which runs the remaining
  stages (MTF, Huffman, container) through bzap_compress_from_bwt_device.

The text is replicated on every GPU (a 1 GiB block is 0.6 % of a B200's HBM); only ranks, keys and
suffix-array slots are sharded.  The backend object hides the per-GPU steps so that the host logic
(plans, splitters, seams) is testable on CPU with gloo (tests/test_distributed_gloo.py supplies a
numpy backend; the product only ever uses GpuBackend).
"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import lib, Context, _check


class GpuBackend:
    """Per-GPU steps through the device-level C ABI; tensors live on the context's GPU."""

    def __init__(self, ctx=None, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.ctx = ctx or Context(self.device.index)
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def init_keys(self, text, n, lo, m):
        keys = torch.empty(m, dtype=torch.int64, device=self.device)
        if m:
            _check(lib().bzap_dev_init_keys(self.ctx.h, C.c_void_p(text.data_ptr()), n, lo, m, C.c_void_p(keys.data_ptr())), self.ctx.h)
        return keys

    def sort_pairs(self, keys, vals):
        m = keys.numel()
        if m == 0:
            return keys, vals
        kt, vt = torch.empty_like(keys), torch.empty_like(vals)
        flag = C.c_int(0)
        _check(lib().bzap_dev_sort_pairs(self.ctx.h, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), m,
                                         C.c_void_p(kt.data_ptr()), C.c_void_p(vt.data_ptr()), C.byref(flag)), self.ctx.h)
        return (kt, vt) if flag.value else (keys, vals)

    def rerank(self, keys_sorted, pos_base):
        m = keys_sorted.numel()
        rs = torch.empty(m, dtype=torch.int32, device=self.device)
        if m == 0:
            return rs, 0, 0
        counts = (C.c_uint32 * 2)()
        _check(lib().bzap_dev_rerank(self.ctx.h, C.c_void_p(keys_sorted.data_ptr()), m, pos_base, C.c_void_p(rs.data_ptr()), counts),
               self.ctx.h)
        return rs, int(counts[0]), int(counts[1])

    def partition_dest(self, keys, vals, split_keys, split_vals):
        m = keys.numel()
        dest = torch.empty(m, dtype=torch.uint8, device=self.device)
        if m == 0:
            return dest
        sk = np.ascontiguousarray(split_keys, dtype=np.uint64)
        sv = np.ascontiguousarray(split_vals, dtype=np.uint32)
        _check(lib().bzap_dev_partition_dest(self.ctx.h, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), m,
                                             C.c_void_p(sk.ctypes.data), C.c_void_p(sv.ctypes.data), sk.size,
                                             C.c_void_p(dest.data_ptr())), self.ctx.h)
        return dest

    def stable_perm_by_byte(self, dest):
        m = dest.numel()
        perm = torch.empty(m, dtype=torch.int32, device=self.device)
        cum = (C.c_uint32 * 257)()
        if m:
            _check(lib().bzap_dev_stable_perm_by_byte(self.ctx.h, C.c_void_p(dest.data_ptr()), m, C.c_void_p(perm.data_ptr()), cum),
                   self.ctx.h)
        return perm, np.array(cum, dtype=np.int64)

    def permute_pairs(self, keys, vals, perm):
        m = perm.numel()
        ok = torch.empty(m, dtype=torch.int64, device=self.device) if keys is not None else None
        ov = torch.empty(m, dtype=torch.int32, device=self.device)
        if m:
            _check(lib().bzap_dev_permute_pairs(self.ctx.h, C.c_void_p(keys.data_ptr() if keys is not None else None),
                                                C.c_void_p(vals.data_ptr()), C.c_void_p(perm.data_ptr()), m,
                                                C.c_void_p(ok.data_ptr() if ok is not None else None), C.c_void_p(ov.data_ptr())),
                   self.ctx.h)
        return ok, ov

    def bucket_by_index(self, idx, vals, shift):
        m = idx.numel()
        oi, ov = torch.empty_like(idx), torch.empty_like(vals)
        counts = (C.c_uint32 * 256)()
        if m:
            _check(lib().bzap_dev_bucket_by_index(self.ctx.h, C.c_void_p(idx.data_ptr()), C.c_void_p(vals.data_ptr()), m, shift,
                                                  C.c_void_p(oi.data_ptr()), C.c_void_p(ov.data_ptr()), counts), self.ctx.h)
        return oi, ov, np.array(counts, dtype=np.int64)

    def scatter(self, idx, vals, offset, out):
        if idx.numel():
            _check(lib().bzap_dev_scatter_u32(self.ctx.h, C.c_void_p(idx.data_ptr()), C.c_void_p(vals.data_ptr()), idx.numel(),
                                              offset, C.c_void_p(out.data_ptr())), self.ctx.h)

    def gather_last(self, text, n, sa):
        last = torch.empty(sa.numel(), dtype=torch.uint8, device=self.device)
        if sa.numel():
            _check(lib().bzap_dev_gather_last(self.ctx.h, C.c_void_p(text.data_ptr()), n, C.c_void_p(sa.data_ptr()), sa.numel(),
                                              C.c_void_p(last.data_ptr())), self.ctx.h)
        return last

    def finish_bwt(self, text, sa, rank, rs, k):
        """Active rounds to the end on this GPU (bzap_dev_bwt_finish): returns (last column, primary)."""
        n = text.numel()
        last = torch.empty(n, dtype=torch.uint8, device=self.device)
        prim = C.c_uint64(0)
        _check(lib().bzap_dev_bwt_finish(self.ctx.h, C.c_void_p(text.data_ptr()), n, C.c_void_p(sa.data_ptr()),
                                         C.c_void_p(rank.data_ptr()), C.c_void_p(rs.data_ptr()), k, C.c_void_p(last.data_ptr()),
                                         C.byref(prim)), self.ctx.h)
        return last, int(prim.value)

    def finish(self, last, n, primary):
        """MTF + Huffman + container on this GPU: the stages after bwt() (main.cpp:309-324)."""
        cap = lib().bzap_compress_bound(n)
        out = torch.empty(cap, dtype=torch.uint8, device=self.device)
        ln = C.c_size_t(0)
        _check(lib().bzap_compress_from_bwt_device(self.ctx.h, C.c_void_p(last.data_ptr()), n, primary, C.c_void_p(out.data_ptr()),
                                                   cap, C.byref(ln)), self.ctx.h)
        return out[:ln.value]


def _install_signatures():
    L = lib()
    vp, sz = C.c_void_p, C.c_size_t
    L.bzap_dev_init_keys.argtypes = [vp, vp, sz, sz, sz, vp]
    L.bzap_dev_sort_pairs.argtypes = [vp, vp, vp, sz, vp, vp, C.POINTER(C.c_int)]
    L.bzap_dev_rerank.argtypes = [vp, vp, sz, C.c_uint32, vp, C.POINTER(C.c_uint32)]
    L.bzap_dev_partition_dest.argtypes = [vp, vp, vp, sz, vp, vp, C.c_int, vp]
    L.bzap_dev_stable_perm_by_byte.argtypes = [vp, vp, sz, vp, C.POINTER(C.c_uint32)]
    L.bzap_compress_from_bwt_device.argtypes = [vp, vp, sz, C.c_uint64, vp, sz, C.POINTER(sz)]
    L.bzap_dev_permute_pairs.argtypes = [vp, vp, vp, vp, sz, vp, vp]
    L.bzap_dev_scatter_u32.argtypes = [vp, vp, vp, sz, C.c_uint32, vp]
    L.bzap_dev_gather_last.argtypes = [vp, vp, sz, vp, sz, vp]
    L.bzap_dev_bucket_by_index.argtypes = [vp, vp, vp, sz, C.c_int, vp, vp, C.POINTER(C.c_uint32)]
    L.bzap_dev_bucket_by_index.restype = C.c_int
    L.bzap_dev_bwt_finish.argtypes = [vp, vp, sz, vp, vp, vp, C.c_uint64, vp, C.POINTER(C.c_uint64)]
    L.bzap_dev_bwt_finish.restype = C.c_int
    for f in ("bzap_dev_permute_pairs", "bzap_dev_scatter_u32", "bzap_dev_gather_last","bzap_dev_init_keys", "bzap_dev_sort_pairs", "bzap_dev_rerank", "bzap_dev_partition_dest",
              "bzap_dev_stable_perm_by_byte", "bzap_compress_from_bwt_device"):
        getattr(L, f).restype = C.c_int


# ---- optional phase timing (BZAP_DIST_TIMING=1): wall clock with a device sync per phase ---------------
_TIMING = os.environ.get("BZAP_DIST_TIMING") == "1"
_phase = {}


def _tick(name, t0, dev):
    if not _TIMING:
        return 0.0
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    _phase[name] = _phase.get(name, 0.0) + (t1 - t0)
    return t1


def phase_times(reset=True):
    out = dict(_phase)
    if reset:
        _phase.clear()
    return out


# ---- host logic --------------------------------------------------------------------------------------
def index_shift(n):
    """Bucket digit of an index = (index >> shift) & 255 covers [0, n): shift = bits(n - 1) - 8."""
    return max(0, (max(n, 1) - 1).bit_length() - 8)


def shard_bounds(n, world):
    """Contiguous text shards of about n / world positions, cut at multiples of 2^index_shift(n) so
    that every 1/256 index bucket has exactly one owner (the last shards may be short or empty)."""
    unit = 1 << index_shift(n)
    shard = -(-(-(-n // world)) // unit) * unit
    return shard, [min(n, r * shard) for r in range(world + 1)]


def shift_plan(n, world, k):
    """plan[s][d] = list of (global_start, length): the pieces of rank s's shard that rank d needs
    for r2[i] = rank[(i + k) mod n], i in d's shard, in the order they appear along d's range."""
    _, b = shard_bounds(n, world)
    plan = [[[] for _ in range(world)] for _ in range(world)]
    for d in range(world):
        m = b[d + 1] - b[d]
        if m == 0:
            continue
        start = (b[d] + k) % n
        spans = [(start, min(m, n - start))]
        if spans[0][1] < m:
            spans.append((0, m - spans[0][1]))
        for a, ln in spans:
            for s in range(world):
                lo, hi = max(a, b[s]), min(a + ln, b[s + 1])
                if lo < hi:
                    plan[s][d].append((lo, hi - lo))
    return plan


def _all_to_all(tensors, send_counts, group):
    """all_to_all_single of several equally split tensors; returns (received tensors, recv_counts)."""
    world = dist.get_world_size(group)
    dev = tensors[0].device
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu()]
    outs = []
    for t in tensors:
        out = torch.empty(sum(recv_counts), dtype=t.dtype, device=dev)
        dist.all_to_all_single(out, t.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        outs.append(out)
    return outs, recv_counts


def _fetch_shifted(rank_local, n, k, group):
    world, me = dist.get_world_size(group), dist.get_rank(group)
    _, b = shard_bounds(n, world)
    plan = shift_plan(n, world, k % n)
    lo = b[me]
    send = [rank_local[a - lo:a - lo + ln] for d in range(world) for (a, ln) in plan[me][d]]
    send_counts = [sum(ln for _, ln in plan[me][d]) for d in range(world)]
    buf = torch.cat(send) if send else rank_local[:0]
    (recv,), _ = _all_to_all([buf], send_counts, group)
    m = b[me + 1] - b[me]
    r2 = torch.empty(m, dtype=rank_local.dtype, device=rank_local.device)
    start = (b[me] + k) % n if m else 0
    off = 0
    for s in range(world):
        for a, ln in plan[s][me]:
            dst = (a - start) % n
            r2[dst:dst + ln] = recv[off:off + ln]
            off += ln
    return r2


def _dist_sort(keys, vals, backend, group, samples_per_rank=256):
    """Sample sort of (u64 key bit patterns as int64, u32 index as int32) pairs over the group."""
    world = dist.get_world_size(group)
    m = keys.numel()
    dev = keys.device
    t = _tick("", 0, dev) if _TIMING else 0.0
    if world > 1:
        S = samples_per_rank
        if m:
            pos = (torch.arange(S, dtype=torch.int64, device=dev) * max(m - 1, 0)) // max(S - 1, 1)
            sk, sv = keys[pos], vals[pos]
            flag = torch.ones(S, dtype=torch.int64, device=dev)
        else:
            sk = torch.zeros(S, dtype=torch.int64, device=dev)
            sv = torch.zeros(S, dtype=torch.int32, device=dev)
            flag = torch.zeros(S, dtype=torch.int64, device=dev)
        pack = torch.stack([sk, sv.long(), flag])
        allp = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(allp, pack, group=group)
        allp = torch.cat(allp, dim=1).cpu().numpy()
        ok = allp[2] != 0
        k_np = allp[0][ok].view(np.uint64) if ok.any() else np.zeros(0, np.uint64)
        v_np = allp[1][ok].astype(np.uint32)
        order = np.lexsort((v_np, k_np))
        k_np, v_np = k_np[order], v_np[order]
        if k_np.size:
            cut = [(i * k_np.size) // world for i in range(1, world)]
            split_k, split_v = k_np[cut], v_np[cut]
        else:
            split_k, split_v = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        t = _tick("sort.splitters", t, dev)
        dest = backend.partition_dest(keys, vals, split_k, split_v)
        perm, cum = backend.stable_perm_by_byte(dest)
        send_counts = [int(cum[d + 1] - cum[d]) for d in range(world)] if m else [0] * world
        pk, pv = backend.permute_pairs(keys, vals, perm)
        t = _tick("sort.partition", t, dev)
        (keys, vals), _ = _all_to_all([pk, pv], send_counts, group)
        t = _tick("sort.all_to_all", t, dev)
    out = backend.sort_pairs(keys, vals)
    _tick("sort.local", t, dev)
    return out


def _rerank(keys_s, backend, group):
    """Global sparse ranks of the distributed sorted run; returns (rs, pos_base, n_groups)."""
    world, me = dist.get_world_size(group), dist.get_rank(group)
    dev = keys_s.device
    m = keys_s.numel()
    cnt = torch.tensor([m], dtype=torch.int64, device=dev)
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt, group=group)
    counts = [int(c) for c in torch.cat(allc).cpu()]
    pos_base = sum(counts[:me])
    rs, heads, singles = backend.rerank(keys_s, pos_base)
    info = torch.zeros(5, dtype=torch.int64, device=dev)
    if m:
        info[0], info[1], info[2] = keys_s[0], keys_s[-1], rs[-1].long()
    info[3] = heads
    info[4] = singles
    alli = [torch.empty_like(info) for _ in range(world)]
    dist.all_gather(alli, info, group=group)
    alli = torch.stack(alli).cpu().numpy()
    # seam fix-up: a group may continue from the previous non-empty rank (or span several ranks)
    bases = np.concatenate([[0], np.cumsum(counts)])
    prev_key, prev_val, have_prev = 0, 0, False
    carry_me, groups = None, 0
    settled = int(alli[:, 4].sum())          # singleton groups, counted per rank (a seam can only lower it)
    for r in range(world):
        if counts[r] == 0:
            continue
        first_key, last_key, last_rs, h = (int(x) for x in alli[r][:4])
        cont = have_prev and first_key == prev_key
        groups += h - (1 if cont else 0)
        carry = prev_val if cont else None
        if r == me:
            carry_me = carry
        # the last element sits in the run's first group iff its local rank is the run's base
        prev_val = carry if (cont and last_rs == int(bases[r])) else last_rs
        prev_key, have_prev = last_key, True
    if carry_me is not None and m:
        first_len = int(torch.searchsorted(rs, torch.tensor([pos_base], dtype=rs.dtype, device=dev), right=True))
        rs[:first_len] = carry_me
    return rs, pos_base, groups, settled


HANDOVER_DIVISOR = 2      # hand the block over to one GPU once at most half of the rotations are unsettled
                          # (the same switch point as the single-GPU path, csrc/bwt.cu)


def distributed_bwt(text, group=None, backend=None):
    """text: uint8 tensor holding the WHOLE block, replicated on every rank's device.
    Returns (last_column on rank 0 else None, primary index, rounds)."""
    group = group or dist.group.WORLD
    world, me = dist.get_world_size(group), dist.get_rank(group)
    backend = backend or GpuBackend()
    if isinstance(backend, GpuBackend):
        _install_signatures()
    n = text.numel()
    dev = text.device
    shard, b = shard_bounds(n, world)
    lo, hi = b[me], b[me + 1]
    m = hi - lo
    idx = torch.arange(lo, hi, dtype=torch.int32, device=dev)
    keys = backend.init_keys(text, n, lo, m)
    rank_local = torch.zeros(m, dtype=torch.int32, device=dev)
    k, prev_groups, rounds = 8, 0, 0
    while True:
        keys_s, idx_s = _dist_sort(keys, idx.clone(), backend, group)
        t = _tick("", 0, dev) if _TIMING else 0.0
        rs, pos_base, groups, settled = _rerank(keys_s, backend, group)
        t = _tick("rerank", t, dev)
        rounds += 1
        # ranks go home: (index, rank) to the owner of the index.  One stable bucketing pass by the
        # top 8 index bits groups the pairs by owner (shards are cut at bucket boundaries) and leaves
        # every owner's share ordered by 1/256 window, which keeps its local scatter cache friendly.
        if world > 1:
            shift = index_shift(n)
            pi, pr, counts = backend.bucket_by_index(idx_s, rs, shift)
            send_counts = [0] * world
            for bkt in range(256):
                if counts[bkt]:
                    send_counts[min(world - 1, (bkt << shift) // shard)] += int(counts[bkt])
            (ri, rr), _ = _all_to_all([pi, pr], send_counts, group)
        else:
            ri, rr = idx_s, rs
        backend.scatter(ri, rr, lo, rank_local)
        t = _tick("ranks_home", t, dev)
        if groups == n or k >= n or groups == prev_groups:
            break
        prev_groups = groups
        if world > 1 and n - settled <= n // HANDOVER_DIVISOR:
            # few rotations are still unsettled: further rounds would move all N pairs through the
            # sample sort for nothing.  Rank 0 collects the suffix array, the ranks in both orders
            # and finishes with the single-GPU active-set rounds (bzap_dev_bwt_finish).
            (sa_full, rs_full), _ = _all_to_all([idx_s, rs], [idx_s.numel()] + [0] * (world - 1), group)
            (rank_full,), _ = _all_to_all([rank_local], [m] + [0] * (world - 1), group)
            prim = torch.zeros(1, dtype=torch.int64, device=dev)
            last = None
            if me == 0:
                last, p0 = backend.finish_bwt(text, sa_full, rank_full, rs_full, k)
                prim[0] = p0
            dist.broadcast(prim, src=dist.get_global_rank(group, 0), group=group)
            return last, int(prim[0]), rounds
        r2 = _fetch_shifted(rank_local, n, k, group) if world > 1 else torch.roll(rank_local, -(k % n))
        keys = (rank_local.long() << 32) | r2.long()
        k *= 2
        _tick("shift+keys", t, dev)
    # last column of this rank's slot range, gathered on rank 0
    last_local = backend.gather_last(text, n, idx_s)
    if world > 1:
        (last,), _ = _all_to_all([last_local], [last_local.numel()] + [0] * (world - 1), group)
    else:
        last = last_local
    prim = torch.zeros(1, dtype=torch.int64, device=dev)
    if me == 0:
        prim[0] = rank_local[0].long()
    if world > 1:
        dist.broadcast(prim, src=dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0, group=group)
    return (last if me == 0 else None), int(prim[0]), rounds


def compress_block_distributed(text, group=None, backend=None):
    """Whole pipeline for one block over the group; the reference-format file (uint8 tensor) is
    returned on rank 0, None elsewhere."""
    backend = backend or GpuBackend()
    last, primary, rounds = distributed_bwt(text, group, backend)
    me = dist.get_rank(group or dist.group.WORLD)
    if me != 0:
        return None, rounds
    return backend.finish(last, text.numel(), primary), rounds
