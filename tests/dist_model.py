"""Executable specification of the distributed single-block path (csrc/dist_block.cu), numpy per
rank + torch.distributed (gloo) for the exchanges.  TEST INFRASTRUCTURE: the product is the C++/CUDA
implementation behind bzap_compress_block_distributed; this model pins its host logic -- splitters,
key-range ownership, the pull of rank[(i+k) mod N], "ranks go home", termination, the MTF state
hand-over, the Huffman statistics reduction and the bit offsets of the payload pieces -- on CPU
with world_size > 1 (tests/test_distributed_gloo.py).  Step names match the C++ (dist_block.cu).

One BWT block = whole input (main.cpp:77-91, README.md:40).  The text is replicated on every rank.
  0. splitters      8-byte keys of S hashed sample positions, the same on every rank (no exchange);
                    rank g owns the rotations whose 8-byte key lies in [spl[g], spl[g+1]): equal keys
                    never straddle two ranks, so groups and their sparse ranks are local for ever
  1. select + sort  own (key, start) pairs, local sort, sparse ranks rs = base_g + group head
  2. ranks go home  (start, rank) to the owner of text position `start` (contiguous shards, bucket aligned)
  3. rounds         active = rotations in a group > 1: pull r2 = rank[(start + k) mod N] from the
                    position owners (request / response all-to-all), sort (r1, r2) locally, new
                    sparse ranks inside the group's slot range, ranks go home, survivors stay
  4. last column    L[j] = text[(sa[j] - 1) mod N] per rank; primary = rank[0]
  5. MTF            per-rank "last occurrence" summary, all-gather, start list of every rank
  6. Huffman        per-rank histogram + first appearance, all-gather; tree on every rank; per-rank
                    bit counts -> bit offset of every rank's payload piece; pieces OR-merged on rank 0
"""
import numpy as np
import torch
import torch.distributed as dist

import oracle_lib as O

SAMPLES_PER_RANK = 512
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


# ---- exchanges ----------------------------------------------------------------------------------------
def allgather_obj(x, group):
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, x, group=group)
    return out


def alltoallv(parts, dtype, group):
    """parts[d] = numpy array for rank d; returns the list of arrays received from every rank."""
    world = dist.get_world_size(group)
    counts = [int(p.size) for p in parts]
    allc = allgather_obj(counts, group)
    me = dist.get_rank(group)
    recv_counts = [allc[s][me] for s in range(world)]
    send = torch.from_numpy(np.concatenate(parts).astype(dtype)) if sum(counts) else torch.zeros(0, dtype=torch.from_numpy(np.zeros(0, dtype)).dtype)
    recv = torch.empty(sum(recv_counts), dtype=send.dtype)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=counts, group=group)
    r = recv.numpy()
    off = np.concatenate([[0], np.cumsum(recv_counts)])
    return [r[off[s]:off[s + 1]] for s in range(world)]


# ---- step 0: splitters ----------------------------------------------------------------------------------
def sample_hash(j):
    """splitmix64 finaliser: position of sample j (the same arithmetic as dist_sample_pos in C++)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(j, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def window8(text, pos):
    n = text.size
    k = np.zeros(len(pos), dtype=np.uint64)
    for j in range(8):
        k = (k << np.uint64(8)) | text[(pos + j) % n].astype(np.uint64)
    return k


def splitters(text, world):
    """world + 1 bounds: bounds[0] = 0, bounds[world] = 2^64 (as python ints); rank g owns keys in
    [bounds[g], bounds[g+1])."""
    n = text.size
    S = SAMPLES_PER_RANK * world
    pos = (sample_hash(np.arange(S)) % np.uint64(n)).astype(np.int64)
    keys = np.sort(window8(text, pos))
    b = [0] + [int(keys[(g * S) // world]) for g in range(1, world)] + [1 << 64]
    return b


def shard_len(n, world):
    """Contiguous shards of text positions, aligned to the 256 top-bit buckets the exchange regroups by
    (Geometry in dist_block.cu): bucket = position >> shift, ceil(buckets / world) buckets per rank."""
    bits = max(int(n - 1).bit_length(), 0) if n > 1 else 0
    shift = max(bits - 8, 0)
    nb = -(-n // (1 << shift))
    return (-(-nb // world)) << shift


# ---- MTF pieces -------------------------------------------------------------------------------------------
def mtf_summary(seg):
    """last[s] = 1 + position of the last occurrence of s in the segment, 0 = absent."""
    last = np.zeros(256, dtype=np.int64)
    if seg.size:
        idx = np.arange(1, seg.size + 1)
        np.maximum.at(last, seg, idx)
    return last


def mtf_next_list(lst, last):
    """list after a segment with summary `last` that started from list `lst` (main.cpp:93-112)."""
    seen = [s for s in range(256) if last[s]]
    seen.sort(key=lambda s: -last[s])
    rest = [s for s in lst if not last[s]]
    return seen + rest


def mtf_from(seg, lst):
    lst = list(lst)
    out = np.empty(seg.size, dtype=np.uint8)
    for i, c in enumerate(seg.tolist()):
        p = lst.index(c)
        out[i] = p
        if p:
            lst.pop(p)
            lst.insert(0, c)
    return out


# ---- Huffman pieces -----------------------------------------------------------------------------------------
def pack_bits(mtf, codes, lens, bit_off):
    """bytes covering stream bits [bit_off, bit_off + sum(len)) of the packed payload, MSB first
    (io_utilities.h:87-94); returns (first_byte, bytes)."""
    bits = []
    for s in mtf.tolist():
        l = int(lens[s])
        c = int(codes[s])
        bits.extend((c >> (l - 1 - b)) & 1 for b in range(l))
    first = bit_off // 8
    total = bit_off % 8 + len(bits)
    arr = np.zeros((total + 7) // 8 * 8, dtype=np.uint8)
    arr[bit_off % 8: bit_off % 8 + len(bits)] = bits
    return first, np.packbits(arr)


# ---- the whole path -------------------------------------------------------------------------------------------
def compress_block_distributed(text, group=None, stats=None):
    """text: np.uint8 array, the same on every rank.  Returns the reference-format file (np.uint8)
    on rank 0, None elsewhere."""
    group = group or dist.group.WORLD
    world, me = dist.get_world_size(group), dist.get_rank(group)
    n = int(text.size)
    shard = shard_len(n, world)
    lo = min(n, me * shard)
    hi = min(n, lo + shard)

    # 0./1. own rotations, sorted, sparse ranks
    b = splitters(text, world)
    keys_all = window8(text, np.arange(n))
    owner = np.searchsorted(np.array(b[1:-1], dtype=np.uint64), keys_all, side="right") if world > 1 else np.zeros(n, np.int64)
    counts = np.bincount(owner, minlength=world)
    base = int(counts[:me].sum())
    mine = np.nonzero(owner == me)[0]
    order = np.argsort(keys_all[mine], kind="stable")
    sa = mine[order].astype(np.int64)
    ks = keys_all[mine][order]
    cnt = sa.size
    head = np.ones(cnt, dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    rs = base + np.maximum.accumulate(np.where(head, np.arange(cnt), 0)) if cnt else np.zeros(0, np.int64)

    rank_home = np.zeros(hi - lo, dtype=np.int64)

    def go_home(idx, val):
        d = idx // shard
        ri = alltoallv([idx[d == r] for r in range(world)], np.int64, group)
        rv = alltoallv([val[d == r] for r in range(world)], np.int64, group)
        for i, v in zip(ri, rv):
            rank_home[i - lo] = v

    go_home(sa, rs)
    single = head & np.append(head[1:], True) if cnt else np.zeros(0, bool)
    act = np.nonzero(~single)[0]
    act_idx, act_r1 = sa[act], rs[act]
    k, rounds = 8, 1
    tot = allgather_obj(int(act_idx.size), group)
    while sum(tot) and k < n:
        # pull r2 = rank[(start + k) mod n]
        pos = (act_idx + k) % n
        d = pos // shard
        q = np.argsort(d, kind="stable")                     # stable multisplit by owner
        req = alltoallv([pos[d == r] for r in range(world)], np.int64, group)
        resp = alltoallv([rank_home[p - lo] for p in req], np.int64, group)
        r2 = np.empty(act_idx.size, dtype=np.int64)
        r2[q] = np.concatenate(resp) if act_idx.size else np.zeros(0, np.int64)
        # local sort by (r1, r2); group g occupies suffix-array slots r1 .. r1 + size - 1
        o = np.lexsort((r2, act_r1))
        s_idx, s_r1, s_r2 = act_idx[o], act_r1[o], r2[o]
        m = s_idx.size
        gh = np.ones(m, dtype=bool)
        gh[1:] = s_r1[1:] != s_r1[:-1]
        gstart = np.maximum.accumulate(np.where(gh, np.arange(m), 0)) if m else np.zeros(0, np.int64)
        slot = s_r1 + (np.arange(m) - gstart)
        sh = gh.copy()
        sh[1:] |= s_r2[1:] != s_r2[:-1]
        newr = np.maximum.accumulate(np.where(sh, slot, 0)) if m else np.zeros(0, np.int64)
        sa[slot - base] = s_idx
        go_home(s_idx, newr)                                 # unchanged ranks are rewritten with the same value
        ssingle = sh & np.append(sh[1:], True) if m else np.zeros(0, bool)
        keep = ~ssingle
        act_idx, act_r1 = s_idx[keep], newr[keep]
        rounds += 1
        k *= 2
        info = allgather_obj((int(gh.sum()), int(sh.sum()), int(act_idx.size)), group)
        groups, subs = sum(i[0] for i in info), sum(i[1] for i in info)
        tot = [i[2] for i in info]
        if subs == groups:                                   # fixed point: no group was split
            break
    if stats is not None:
        stats["rounds"] = rounds
        stats["counts"] = counts.tolist()

    # 4. last column of this rank's slot range; primary = rank[0] (main.cpp:87-88)
    last = text[(sa - 1) % n]
    prim = allgather_obj(int(rank_home[0]) if lo == 0 and hi > 0 else -1, group)
    primary = max(prim)

    # 5. MTF with the start list handed over from the ranks before
    summ = allgather_obj(mtf_summary(last), group)
    lst = list(range(256))
    for r in range(me):
        lst = mtf_next_list(lst, summ[r])
    mtf = mtf_from(last, lst)

    # 6. Huffman statistics: frequencies add up, first appearances take the global minimum
    freq = np.bincount(mtf, minlength=256).astype(np.uint64)
    first = np.full(256, 1 << 62, dtype=np.int64)
    for s in range(256):
        w = np.nonzero(mtf == s)[0]
        if w.size:
            first[s] = base + int(w[0])
    allf = allgather_obj((freq, first), group)
    gfreq = sum(f for f, _ in allf)
    gfirst = np.min(np.stack([f for _, f in allf]), axis=0)
    syms = [s for s in range(256) if gfreq[s]]
    syms.sort(key=lambda s: gfirst[s])
    tree = O.o_tree_from_hist(gfreq, np.array(syms, dtype=np.uint8))
    tb = O.o_tree_bytes(tree)
    c_lo, c_hi, lens = O.o_codes(tree)
    codes = [(int(c_hi[s]) << 64) | int(c_lo[s]) for s in range(256)]
    bits_per_rank = [int(sum(int(f[s]) * int(lens[s]) for s in syms)) for f, _ in allf]
    head_bytes = 24 + tb.size
    bit_off = 8 * head_bytes + sum(bits_per_rank[:me])
    piece = pack_bits(mtf, codes, lens, bit_off)
    pieces = allgather_obj(piece, group)
    if me != 0:
        return None
    total_bits = sum(bits_per_rank)
    payload = max(1, (total_bits + 7) // 8)
    out = np.zeros(head_bytes + payload, dtype=np.uint8)
    out[:24] = np.frombuffer(np.array([primary, n, tb.size], dtype="<u8").tobytes(), dtype=np.uint8)
    out[24:head_bytes] = tb
    for fb, by in pieces:
        out[fb:fb + by.size] |= by
    return out
