// bwt.cu -- forward Burrows-Wheeler transform by cyclic prefix doubling on the GPU.
//
// Replaces bwt() (main.cpp:77-91) and its comparator bwt_cmp_straight (main.cpp:46-59).  The
// reference stable-sorts rotation start indices with an O(N) cyclic comparator; rotations that
// compare equal keep ascending start index.  Here:
//   round 0 : key[i] = first 8 bytes of rotation i (cyclic, big-endian)       -> onesweep sort
//   round r : key[i] = (rank[i] << 32) | rank[(i + k) mod N],  k = 8 * 2^(r-1) -> onesweep sort
//   re-rank : SPARSE ranks -- rank = sorted position of the first member of the equal-key group,
//             i.e. the number of rotations strictly smaller under the current prefix length.
// Equal rotations therefore keep equal rank for ever (no index ever enters a sort key), the last
// column is independent of the order inside an equal group, and because rotation 0 has the
// lowest start index of its group the reference's primary index (main.cpp:88) is rank[0].
// Termination: all ranks distinct, or k >= N, or a round that creates no new group (then the
// partition is a fixed point of doubling: every longer prefix induces the same partition).
//
// Two regimes.  FULL rounds sort all N rotations.  Once at most half of the rotations still sit
// in a group of size > 1, ACTIVE rounds take over: only those rotations are re-keyed (their group
// start r1 and a gathered rank[(i+k) mod N]), sorted, and written back inside their group's slot
// range of the suffix array; singletons are final and never touched again.
//
// Digit histograms for the sort passes are never computed from the full keys: digit j of an
// 8-byte window is a text byte, and digit j of either key half is a digit of a rank, so one byte
// histogram (round 0) or the rank-digit histogram fused into the re-rank kernel serves all
// eight passes.
#include "device_common.cuh"

#define RR_BLOCK 256
#define RR_ITEMS 8
#define RR_TILE (RR_BLOCK * RR_ITEMS)

// ---- round 0 keys --------------------------------------------------------------------------------
#define IK_BLOCK 256
#define IK_ITEMS 8
#define IK_TILE (IK_BLOCK * IK_ITEMS)
// keys[i] for rotations lo .. lo+m-1 of an n-byte text (lo = 0, m = n on a single GPU)
__global__ void __launch_bounds__(IK_BLOCK)
bwt_init_keys_kernel(const u8 *__restrict__ text, u32 n, u32 lo, u32 m, u64 *__restrict__ keys)
{
    __shared__ u8 s_b[IK_TILE + 8];
    const u32 tiles = (m + IK_TILE - 1) / IK_TILE;
    for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const u32 base = tile * IK_TILE;
        __syncthreads();
        for (u32 i = threadIdx.x; i < IK_TILE + 8; i += IK_BLOCK) {
            u64 p = (u64)lo + base + i;
            if (p >= n) p %= n;                   // cyclic window (main.cpp:38-44)
            s_b[i] = text[p];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IK_ITEMS; ++i) {
            u32 o = threadIdx.x + i * IK_BLOCK;
            if (base + o < m) {
                u64 k = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) k = (k << 8) | s_b[o + j];
                keys[base + o] = k;
            }
        }
    }
}

// ---- doubling keys (full rounds) -----------------------------------------------------------------
__global__ void __launch_bounds__(256) bwt_pair_keys_kernel(const u32 *__restrict__ rank, u32 n, u32 k, u64 *__restrict__ keys)
{
    const u32 kk = k % n;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u32 j = i + kk;                           // < 2n <= 2^31: no overflow
        if (j >= n) j -= n;
        keys[i] = ((u64)rank[i] << 32) | rank[j];
    }
}

// ---- re-rank (full rounds) ---------------------------------------------------------------------------
// sorted keys (+ payload = rotation start, nullptr = identity) -> rank[start] = sparse rank (text
// order), rs[j] = the same rank in sorted order, number of groups, number of singleton groups, and
// the 4 x 256 histogram of rank digits for the next round's passes.  A few hundred blocks each walk
// a contiguous range of tiles, so the shared-memory histogram is flushed once per block.
// Blocked accesses (thread t owns elements 8t..8t+7) would put 16 lanes on one bank pair: one pad
// slot per 8 elements makes the lane stride 9 (conflict free for 32- and 64-bit words)
#define RR_PAD(x) ((x) + ((x) >> 3))
struct RrSmem {
    u64 keys[RR_PAD(RR_TILE + 2) + 1];
    u32 r[RR_PAD(RR_TILE) + 1];
    u32 hist[4][256];
    u32 tmp[40];
    u32 ticket;
    u32 tile_prefix;
    u32 heads, singles;
};

// Work split: block b owns a CONTIGUOUS range of tiles and carries its running "last head" in
// shared memory, so the only cross-block dependency is the last head before the range.  Every block
// first finds the last head INSIDE its range by scanning backwards from the range end (the first
// chunk almost always contains one) and publishes it at once -- publishing never waits on another
// block, so there is no look-back chain -- then reads its predecessors' words until one holds a head.
#define RR_NONE (1ull << 62)
#define RR_FOUND (2ull << 62)
__global__ void __launch_bounds__(RR_BLOCK)
bwt_rerank_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ sa, u32 n, u32 ntiles, u32 pos_base,
                  u32 *__restrict__ rank,
                  u32 *__restrict__ rs, u32 *hist4, u32 *counters /* [0]=groups [1]=singletons [2]=range ticket */, u64 *status,
                  u32 *block_active /* per block: rotations of its range still in a group > 1 (may be null) */)
{
    __shared__ RrSmem S;
    const u32 tid = threadIdx.x;
    // ranges are handed out by ticket, not by blockIdx: phase 2 waits on lower ranges, which then
    // belong to blocks that already run (no assumption about the block dispatch order)
    const u32 bid = take_ticket(counters + 2, &S.ticket);
    const u32 tpb = (ntiles + gridDim.x - 1) / gridDim.x;          // tiles per block
    const u32 t0 = bid * tpb, t1 = min(ntiles, t0 + tpb);
    const u32 lo = t0 * RR_TILE, hi = min(n, t1 * RR_TILE);
    for (u32 i = tid; i < 4 * 256; i += RR_BLOCK) (&S.hist[0][0])[i] = 0;
    if (tid == 0) { S.heads = 0; S.singles = 0; S.tile_prefix = 0; }
    __syncthreads();
    // 1. last head inside [lo, hi): block-wide backward scan, RR_TILE positions per step
    if (t0 < t1) {
        u32 found = 0;                                             // pos_base + position + 1, 0 = none
        u32 end = hi;
        while (end > lo && !found) {
            const u32 beg = end - lo > RR_TILE ? end - RR_TILE : lo;
            u32 best = 0;
#pragma unroll
            for (int i = 0; i < RR_ITEMS; ++i) {
                u32 j = beg + tid + i * RR_BLOCK;
                if (j < end && (j == 0 || keys[j] != keys[j - 1])) best = pos_base + j + 1;
            }
            u32 tot;
            block_exclusive_max(best, S.tmp, &tot);
            found = tot;
            end = beg;
        }
        if (tid == 0) st_relaxed(&status[bid], found ? (RR_FOUND | (u64)(found - 1)) : RR_NONE);
    } else {
        if (tid == 0) {
            st_relaxed(&status[bid], RR_NONE);
            if (block_active) block_active[bid] = 0;
        }
        return;
    }
    // 2. last head before the range (block 0 owns position 0, which is always a head)
    if (tid < 32) {
        u32 pre = 0;
        int base = (int)bid - 1;
        while (base >= 0) {
            int t = base - (int)tid;
            u64 w = RR_NONE;
            if (t >= 0) do { w = ld_relaxed(&status[t]); } while (w == 0);
            u32 m = __ballot_sync(FULL_MASK, (w >> 62) == 2);
            if (m) {
                pre = __shfl_sync(FULL_MASK, (u32)(w & 0xffffffffu), __ffs(m) - 1);
                break;
            }
            base -= 32;
        }
        if (tid == 0) S.tile_prefix = pre;
    }
    // 3. the range, tile by tile; the next tile's keys are in flight while the current one is processed
    u64 kreg[RR_ITEMS], kprev = 0, knext = 0;
    auto fetch = [&](u32 t) {
        if (t >= t1) return;
        const u32 b = t * RR_TILE;
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) {
            u32 o = tid + i * RR_BLOCK;
            kreg[i] = b + o < n ? keys[b + o] : 0;
        }
        if (tid == 0) {
            kprev = b ? keys[b - 1] : 0;
            knext = b + RR_TILE < n ? keys[b + RR_TILE] : 0;
        }
    };
    fetch(t0);
    for (u32 tile = t0; tile < t1; ++tile) {
        const u32 base = tile * RR_TILE;
        __syncthreads();                                           // previous tile's readers are done; S.tile_prefix visible
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) S.keys[RR_PAD(tid + i * RR_BLOCK + 1)] = kreg[i];
        if (tid == 0) { S.keys[0] = kprev; S.keys[RR_PAD(RR_TILE + 1)] = knext; }
        __syncthreads();
        fetch(tile + 1);

        // blocked: thread owns RR_ITEMS consecutive sorted positions
        u32 loc[RR_ITEMS];
        u32 cur = 0, nheads = 0, nsingle = 0;
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) {
            u32 o = tid * RR_ITEMS + i;
            u32 p = base + o;
            const u64 k0 = S.keys[RR_PAD(o)], k1 = S.keys[RR_PAD(o + 1)], k2 = S.keys[RR_PAD(o + 2)];
            bool head = p < n && (p == 0 || k1 != k0);
            bool next_head = p + 1 >= n || k2 != k1;
            if (head) { cur = pos_base + p; ++nheads; nsingle += next_head; }
            loc[i] = cur;
        }
        u32 total;
        u32 tprefix = block_exclusive_max(cur, S.tmp, &total);
        if (nheads) atomicAdd(&S.heads, nheads);
        if (nsingle) atomicAdd(&S.singles, nsingle);
        const u32 pre = max(S.tile_prefix, tprefix);

        // ranks + run-aggregated digit histogram (ranks are non-decreasing along sorted order, so
        // the high digits are almost constant inside a thread's run)
        u32 run_d[4] = {0, 0, 0, 0}, run_c[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) {
            u32 o = tid * RR_ITEMS + i;
            if (base + o < n) {
                u32 r = max(pre, loc[i]);
                S.r[RR_PAD(o)] = r;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    u32 dig = (r >> (8 * d)) & 0xff;
                    if (run_c[d] && dig != run_d[d]) {
                        atomicAdd(&S.hist[d][run_d[d]], run_c[d]);
                        run_c[d] = 0;
                    }
                    run_d[d] = dig;
                    ++run_c[d];
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (run_c[d]) atomicAdd(&S.hist[d][run_d[d]], run_c[d]);
        __syncthreads();
        if (tid == 0) S.tile_prefix = max(S.tile_prefix, total);    // carried to the next tile

#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) {
            u32 o = tid + i * RR_BLOCK;
            u32 p = base + o;
            if (p < n) {
                u32 r = S.r[RR_PAD(o)];
                rs[p] = r;
                if (rank) rank[sa ? sa[p] : p] = r;          // nullptr: a bucketed scatter follows
            }
        }
    }
    __syncthreads();
    for (u32 i = tid; i < 4 * 256; i += RR_BLOCK) {
        u32 c = (&S.hist[0][0])[i];
        if (c) atomicAdd(&hist4[i], c);
    }
    if (tid == 0) {
        if (S.heads) atomicAdd(&counters[0], S.heads);
        if (S.singles) atomicAdd(&counters[1], S.singles);
        if (block_active) block_active[bid] = (hi - lo) - S.singles;
    }
}

// ---- active rounds -------------------------------------------------------------------------------------
#define AC_BLOCK 256
#define AC_ITEMS 4
#define AC_TILE (AC_BLOCK * AC_ITEMS)

// stable compaction helper: each thread brings `cnt` (0..AC_ITEMS) survivors; returns the global
// output offset of the thread's first survivor (sum look-back over tiles)
__device__ __forceinline__ u32 compact_offset(u32 cnt, u32 tile, u64 *status, u32 *s_tmp, u32 *s_base)
{
    u32 total;
    u32 ex = block_exclusive_sum(cnt, s_tmp, &total);
    if (threadIdx.x < 32) {
        u64 x = lookback_exclusive(status, tile, (u64)total, OpSum());
        if (threadIdx.x == 0) *s_base = (u32)x;
    }
    __syncthreads();
    u32 r = *s_base + ex;
    __syncthreads();
    return r;
}

// after the last full round: sorted position j is settled iff its group is a singleton, i.e.
// rs[j] == j and rs[j+1] == j+1.  Survivors keep their order: (start, group rank r1).
// Same contiguous ranges as bwt_rerank_kernel (same grid), whose per-block survivor counts give
// every block its output offset directly: no look-back, a block scan per tile is all it takes.
__global__ void __launch_bounds__(RR_BLOCK)
bwt_collect_active_kernel(const u32 *__restrict__ rs, const u32 *__restrict__ sa, u32 n, u32 ntiles, u32 pos_base,
                          const u32 *__restrict__ block_active, u32 *__restrict__ act_idx, u32 *__restrict__ act_r1)
{
    __shared__ u32 s_tmp[40];
    const u32 tid = threadIdx.x;
    const u32 tpb = (ntiles + gridDim.x - 1) / gridDim.x;
    const u32 t0 = blockIdx.x * tpb, t1 = min(ntiles, t0 + tpb);
    if (t0 >= t1) return;
    u32 part = 0;
    for (u32 b = tid; b < blockIdx.x; b += RR_BLOCK) part += block_active[b];
    u32 out;
    block_exclusive_sum(part, s_tmp, &out);                  // survivors of all earlier ranges
    for (u32 tile = t0; tile < t1; ++tile) {
        const u32 j0 = tile * RR_TILE + tid * RR_ITEMS;
        u32 r[RR_ITEMS + 1];
#pragma unroll
        for (int i = 0; i <= RR_ITEMS; ++i) r[i] = j0 + i < n ? rs[j0 + i] - pos_base : j0 + i;   // past the end counts as a head
        u32 keep = 0, cnt = 0;
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i) {
            u32 j = j0 + i;
            bool k = j < n && !(r[i] == j && r[i + 1] == j + 1);
            keep |= (u32)k << i;
            cnt += k;
        }
        u32 total;
        u32 o = out + block_exclusive_sum(cnt, s_tmp, &total);
        out += total;
#pragma unroll
        for (int i = 0; i < RR_ITEMS; ++i)
            if ((keep >> i) & 1u) {
                act_idx[o] = sa ? sa[j0 + i] : j0 + i;
                act_r1[o] = r[i] + pos_base;
                ++o;
            }
    }
}

// key[a] = (r1 << rshift) | rank[(start + k) mod N] with the eight digit histograms (M is small here).
// rshift = bits of a rank (BWT_ACTIVE_PACK) packs the pair into 2 x bits: one radix pass fewer when
// 2 x bits crosses a byte boundary less than 32 + bits does (26-bit ranks: 7 passes instead of 8).
#ifndef BWT_ACTIVE_PACK
#define BWT_ACTIVE_PACK 1
#endif
__global__ void __launch_bounds__(256)
bwt_active_keys_kernel(const u32 *__restrict__ act_idx, const u32 *__restrict__ act_r1, u32 m, const u32 *__restrict__ rank,
                       u32 n, u32 k, u32 rshift, u64 *__restrict__ keys, u32 *hist8)
{
    __shared__ u32 s_h[8 * 256];
    for (u32 i = threadIdx.x; i < 8 * 256; i += 256) s_h[i] = 0;
    __syncthreads();
    const u32 kk = k % n;
    for (u32 a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) {
        u32 j = act_idx[a] + kk;
        if (j >= n) j -= n;
        u64 key = ((u64)act_r1[a] << rshift) | rank[j];
        keys[a] = key;
#pragma unroll
        for (int p = 0; p < 8; ++p) atomicAdd(&s_h[p * 256 + ((u32)(key >> (8 * p)) & 0xffu)], 1u);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 8 * 256; i += 256) {
        u32 c = s_h[i];
        if (c) atomicAdd(&hist8[i], c);
    }
}

// sorted active (key, start) -> slot inside the group's range of the suffix array, new sparse rank.
//   gstart[a] = index of the first active element with the same r1     (max-scan of group heads)
//   pos[a]    = r1 + (a - gstart[a])                                     slot in the suffix array
//   newr[a]   = pos of the first element with the same (r1, r2)         (max-scan of sub-group heads)
// pos is strictly increasing in a, so both scans are max-scans of monotone values.
__global__ void __launch_bounds__(AC_BLOCK)
bwt_active_rerank_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ idx, u32 m, u32 rshift, u32 slot_base,
                         u32 *__restrict__ sa,
                         u32 *__restrict__ rank, u32 *__restrict__ newr, u32 *__restrict__ pos_out,
                         u32 *counters /* [0]=groups [1]=sub groups */, u64 *status_g, u64 *status_s, u32 *ticket)
{
    __shared__ u32 s_tmp[40];
    __shared__ u32 s_ticket, s_pg, s_ps;
    const u32 tile = take_ticket(ticket, &s_ticket);
    const u32 a0 = tile * AC_TILE + threadIdx.x * AC_ITEMS;
    u64 kcur[AC_ITEMS + 1];
    kcur[0] = (a0 > 0 && a0 - 1 < m) ? keys[a0 - 1] : 0;
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i) kcur[i + 1] = a0 + i < m ? keys[a0 + i] : 0;
    // pass 1: group heads (r1 changes) -> gstart; the scan carries (index + 1), 0 = none yet
    u32 gl[AC_ITEMS];
    u32 gcur = 0, ng = 0;
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i) {
        u32 a = a0 + i;
        bool gh = a < m && (a == 0 || (u32)(kcur[i + 1] >> rshift) != (u32)(kcur[i] >> rshift));
        if (gh) { gcur = a + 1; ++ng; }
        gl[i] = gcur;
    }
    u32 total;
    u32 tpre = block_exclusive_max(gcur, s_tmp, &total);
    if (threadIdx.x < 32) {
        u64 x = lookback_exclusive(status_g, tile, (u64)total, OpMax());
        if (threadIdx.x == 0) s_pg = (u32)x;
    }
    __syncthreads();
    const u32 gpre = max(s_pg, tpre);
    // pass 2: slots and sub-group heads -> new ranks; the scan carries (pos + 1)
    u32 pos[AC_ITEMS], sl[AC_ITEMS];
    u32 scur = 0, ns = 0;
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i) {
        u32 a = a0 + i;
        u32 gs = max(gpre, gl[i]) - 1u;                      // index of the group's first active element
        pos[i] = (u32)(kcur[i + 1] >> rshift) + (a - gs);
        bool sh = a < m && (a == 0 || kcur[i + 1] != kcur[i]);
        if (sh) { scur = pos[i] + 1; ++ns; }
        sl[i] = scur;
    }
    __syncthreads();
    u32 tpre2 = block_exclusive_max(scur, s_tmp, &total);
    if (threadIdx.x < 32) {
        u64 x = lookback_exclusive(status_s, tile, (u64)total, OpMax());
        if (threadIdx.x == 0) s_ps = (u32)x;
    }
    __syncthreads();
    const u32 spre = max(s_ps, tpre2);
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i) {
        u32 a = a0 + i;
        if (a < m) {
            u32 nr = max(spre, sl[i]) - 1u;
            u32 start = idx[a];
            sa[pos[i] - slot_base] = start;             // slot_base: first suffix-array slot this GPU holds
            if (rank) rank[start] = nr;                 // nullptr: the ranks travel home by exchange
            newr[a] = nr;
            pos_out[a] = pos[i];
        }
    }
    if (ng) atomicAdd(&counters[0], ng);
    if (ns) atomicAdd(&counters[1], ns);
}

// survivors of an active round: sub-groups that still have more than one member
__global__ void __launch_bounds__(AC_BLOCK)
bwt_active_compact_kernel(const u32 *__restrict__ newr, const u32 *__restrict__ pos, const u32 *__restrict__ idx, u32 m,
                          u32 *__restrict__ out_idx, u32 *__restrict__ out_r1, u32 *count, u64 *status, u32 *ticket)
{
    __shared__ u32 s_tmp[40];
    __shared__ u32 s_ticket, s_base;
    const u32 tile = take_ticket(ticket, &s_ticket);
    const u32 a0 = tile * AC_TILE + threadIdx.x * AC_ITEMS;
    bool head[AC_ITEMS + 1];
#pragma unroll
    for (int i = 0; i <= AC_ITEMS; ++i) head[i] = a0 + i >= m || newr[a0 + i] == pos[a0 + i];
    bool keep[AC_ITEMS];
    u32 cnt = 0;
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i) {
        keep[i] = a0 + i < m && !(head[i] && head[i + 1]);
        cnt += keep[i];
    }
    u32 o = compact_offset(cnt, tile, status, s_tmp, &s_base);
    if (cnt) atomicAdd(count, cnt);
#pragma unroll
    for (int i = 0; i < AC_ITEMS; ++i)
        if (keep[i]) {
            out_idx[o] = idx[a0 + i];
            out_r1[o] = newr[a0 + i];
            ++o;
        }
}

__global__ void iota_kernel(u32 *out, u32 n)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

// ---- last column --------------------------------------------------------------------------------------
// L[j] = text[(SA[j] + N - 1) mod N], j < m   (main.cpp:87); sa == nullptr means SA = identity
__global__ void __launch_bounds__(256) bwt_gather_kernel(const u8 *__restrict__ text, const u32 *__restrict__ sa, u32 n, u32 m,
                                                         u8 *__restrict__ last)
{
    const u32 nq = (m + 3) / 4;
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        u32 w = 0;
        u32 j0 = q * 4;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            u32 j = j0 + b;
            if (j < m) {
                u32 s = sa ? sa[j] : j;
                u32 c = text[s == 0 ? n - 1 : s - 1];
                w |= c << (8 * b);
            }
        }
        if (j0 + 3 < m) *reinterpret_cast<u32 *>(last + j0) = w;
        else
            for (u32 b = 0; j0 + b < m; ++b) last[j0 + b] = (u8)(w >> (8 * b));
    }
}

// ---- host driver -----------------------------------------------------------------------------------------
static inline u32 grid_for(size_t work_items, u32 per_block, u32 cap = 148u * 16u)
{
    size_t g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (u32)(g > cap ? cap : g);
}

// ---- active rounds, host loop ------------------------------------------------------------------------------
// One active round after its keys (r1 << rshift | r2), payloads (rotation starts) and the eight digit
// histograms are in ab->keys[0] / ab->vals[0] / w.d_hist8: sort, new sparse ranks inside every group's
// slot range of the suffix array (slot_base = first slot held by w.sa_buf; d_rank == nullptr when the
// ranks travel home by exchange, dist_block.cu), survivors compacted into next_idx / next_r1.
// Leaves w.d_counters = {groups, sub groups, survivors} on the device; no synchronisation.
int bwt_active_sort_rerank(bzap_ctx *ctx, const ActiveWork &w, SortBuffers *ab, u32 m, u32 rshift, u32 pass_mask, u32 slot_base,
                           u32 *d_rank, u32 *next_idx, u32 *next_r1, u64 **skeys, u32 **sidx, int *passes)
{
    const u32 mt = (m + AC_TILE - 1) / AC_TILE;
    ctx->arena_off = w.arena_mark;
    RET(dev_sort_pairs64(ctx, ab, m, pass_mask, w.d_hist8, 8, false, skeys, sidx, passes));
    LAUNCH(ctx, bwt_active_rerank_kernel, mt, AC_BLOCK, 0, *skeys, *sidx, m, rshift, slot_base, w.sa_buf, d_rank, w.newr, w.pos,
           w.d_counters, w.d_status, w.d_status + mt + 2, w.d_ticket);
    LAUNCH(ctx, bwt_active_compact_kernel, mt, AC_BLOCK, 0, w.newr, w.pos, *sidx, m, next_idx, next_r1, w.d_counters + 2,
           w.cstatus, w.d_ticket + 1);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

// w.ab.vals[0] / w.act_r1 hold the m unsettled rotations (start, group rank) in suffix-array order;
// ranks reflect prefixes of length *k
static int bwt_active_rounds(bzap_ctx *ctx, u32 n, u32 m, u64 *k_io, const ActiveWork &w, u32 *rounds, u32 *passes_total)
{
    u64 k = *k_io;
    u32 *h_cnt = (u32 *)(ctx->mailbox + 1024);
    SortBuffers ab = w.ab;
    u32 *act_r1 = w.act_r1, *next_idx = w.next_idx, *next_r1 = w.next_r1;
#if BWT_ACTIVE_PACK
    u32 rshift = 1;
    while (rshift < 32 && (1ull << rshift) < n) ++rshift;            // ranks are < n <= 2^rshift
    const u32 pass_mask = (1u << ((2 * rshift + 7) / 8)) - 1u;
#else
    const u32 rshift = 32, pass_mask = w.rank_mask;
#endif
    while (m) {
        CU(ctx, cudaMemsetAsync(w.zero_base, 0, w.zero_bytes, ctx->stream));
        LAUNCH(ctx, bwt_active_keys_kernel, grid_for(m, 256, 148 * 4), 256, 0, ab.vals[0], act_r1, m, w.d_rank, n,
               (u32)(k % n), rshift, ab.keys[0], w.d_hist8);
        int passes = 0;
        u64 *skeys = nullptr;
        u32 *sidx = nullptr;
        RET(bwt_active_sort_rerank(ctx, w, &ab, m, rshift, pass_mask, 0u, w.d_rank, next_idx, next_r1, &skeys, &sidx, &passes));
        *passes_total += (u32)passes;
        CU(ctx, cudaMemcpyAsync(h_cnt, w.d_counters, 3 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        ++*rounds;
        k *= 2;
        const u32 groups = h_cnt[0], subgroups = h_cnt[1], m2 = h_cnt[2];
        if (m2 == 0 || k >= n || subgroups == groups) break;
        // survivors become the next round's input: swap buffers instead of copying.  The sorted
        // payload may sit in either ab.vals[]; the one it does NOT sit in is free to become the
        // next survivor buffer, and the survivor buffer becomes the sort's payload buffer 0.
        {
            u32 *free_vals = sidx == ab.vals[0] ? ab.vals[1] : ab.vals[0];
            u32 *used_vals = sidx;
            ab.vals[0] = next_idx;
            ab.vals[1] = free_vals;
            next_idx = used_vals;
            u32 *t = act_r1; act_r1 = next_r1; next_r1 = t;
        }
        m = m2;
    }
    *k_io = k;
    return BZAP_OK;
}

int dev_bwt(bzap_ctx *ctx, const u8 *d_in, size_t n64, u8 *d_last, u64 *primary)
{
    if (n64 == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input");
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    ctx->sort_ev_used = 0;
    ctx->stats.sort_bytes = 0;
    ctx->stats.bwt_full_passes = 0;
    ctx->stats.sort_elems = n;
    ctx->stats.ms_sort = 0;
    SortBuffers sb;
    sb.keys[0] = arena_get<u64>(ctx, n);
    sb.keys[1] = arena_get<u64>(ctx, n);
    sb.vals[0] = arena_get<u32>(ctx, n);
    sb.vals[1] = arena_get<u32>(ctx, n);
    u32 *d_rank = arena_get<u32>(ctx, n);
    u32 *d_rs = arena_get<u32>(ctx, n);
    const u32 rr_tiles = (n + RR_TILE - 1) / RR_TILE;
    const u32 ac_tiles_max = (n / 2 + AC_TILE) / AC_TILE + 1;
    // control: hist8[8*256] | hist4[4*256] counters[4] ticket pad | status (u64)
    u32 *d_hist8 = arena_get<u32>(ctx, 8 * 256);
    const size_t status_u64 = (size_t)(rr_tiles > 2 * ac_tiles_max ? rr_tiles : 2 * ac_tiles_max) + 8;
    u32 *d_rrctl = arena_get<u32>(ctx, 4 * 256 + 8 + 2 * status_u64);
    // look-back words of the survivor compaction in the active rounds (<= n/2 elements in AC_TILE tiles)
    u64 *cstatus = arena_get<u64>(ctx, (size_t)ac_tiles_max + 4);
    u32 *d_bact = arena_get<u32>(ctx, 148 * 6 + 8);
    const u32 rr_grid = grid_for(rr_tiles, 1, 148 * 6);
    if (!sb.keys[0] || !sb.keys[1] || !sb.vals[0] || !sb.vals[1] || !d_rank || !d_rs || !d_hist8 || !d_rrctl || !cstatus || !d_bact)
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "bwt scratch");
    u32 *d_hist4 = d_rrctl, *d_counters = d_rrctl + 4 * 256, *d_ticket = d_counters + 4;
    u64 *d_status = (u64 *)(d_rrctl + 4 * 256 + 8);
    const size_t rrctl_bytes = (4 * 256 + 8 + 2 * status_u64) * sizeof(u32);
    const size_t arena_mark = ctx->arena_off;

    // round 0: byte histogram stands in for all eight digit histograms of the 8-byte windows
    CU(ctx, cudaMemsetAsync(d_hist4, 0, 256 * sizeof(u32), ctx->stream));
    RET(dev_byte_hist(ctx, d_in, n, d_hist4));
    // the keys themselves are built inside the first sort pass of every round (SortKeyGen)
    SortKeyGen gen{1, d_in, 0};

    // digits of (rank << 32 | rank) that can differ at all: ranks are < n
    u32 rank_bits = 1;
    while (rank_bits < 32 && (1ull << rank_bits) < n) ++rank_bits;
    const u32 nd = (rank_bits + 7) / 8;
    const u32 rank_mask = ((1u << nd) - 1u) | (((1u << nd) - 1u) << 4);
    u64 *keys = nullptr;
    u32 *sa = nullptr;
    u32 rounds = 0, passes_total = 0, prev_groups = 0;
    u64 k = 8;
    u32 *h_cnt = (u32 *)(ctx->mailbox + 1024);
    bool finished = false;
    u32 active = n;
    // ---- full rounds ----
    while (true) {
        int passes = 0;
        ctx->arena_off = arena_mark;             // sort control block is per round
        // round 0: one byte histogram serves all eight digits; later: four rank-digit histograms serve both key halves
        RET(dev_sort_pairs64(ctx, &sb, n, rounds == 0 ? 0xffu : rank_mask, d_hist4, rounds == 0 ? 1 : 4, true, &keys, &sa, &passes,
                             &gen));
        if (!keys) {
            // every digit is constant (all keys equal): no pass ran, write the keys out for the re-rank
            keys = sb.keys[0];
            if (gen.mode == 1)
                LAUNCH(ctx, bwt_init_keys_kernel, grid_for((n + IK_TILE - 1) / IK_TILE, 1, 148 * 8), IK_BLOCK, 0, d_in, n, 0u, n, keys);
            else
                LAUNCH(ctx, bwt_pair_keys_kernel, grid_for(n, 256 * 4), 256, 0, d_rank, n, gen.k, keys);
        }
        passes_total += (u32)passes;
        CU(ctx, cudaMemsetAsync(d_rrctl, 0, rrctl_bytes, ctx->stream));
        // rank[] is larger than L2 for big blocks: scatter it through a bucketing pass (radix_sort.cu)
        const bool bucketed = sa != nullptr && n > (12u << 20);
        LAUNCH(ctx, bwt_rerank_kernel, rr_grid, RR_BLOCK, 0, keys, sa, n, rr_tiles, 0u,
               bucketed ? (u32 *)nullptr : d_rank, d_rs, d_hist4, d_counters, d_status, d_bact);
        if (bucketed) RET(dev_scatter_perm(ctx, sa, d_rs, n, d_rank, (u32 *)keys, (u32 *)keys + n));
        CU(ctx, cudaMemcpyAsync(h_cnt, d_counters, 2 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        ++rounds;
        const u32 groups = h_cnt[0];
        active = n - h_cnt[1];
        if (groups == n || k >= n || groups == prev_groups) { finished = true; break; }
        prev_groups = groups;
        if (active <= n / 2) break;              // few rotations left unsettled: switch regime
        // next round's keys: (rank[i], rank[i + k]) in text order, payload = identity again
        gen = SortKeyGen{2, d_rank, (u32)(k % n)};
        k *= 2;
    }
    // ---- active rounds ----
    if (!finished) {
        // suffix array lives in the payload buffer the last sort ended in (materialise an identity one)
        u32 *sa_buf = sa;
        if (!sa_buf) {
            sa_buf = sb.vals[0];
            LAUNCH(ctx, iota_kernel, grid_for(n, 256 * 4), 256, 0, sa_buf, n);
        }
        u32 *other_vals = sa_buf == sb.vals[0] ? sb.vals[1] : sb.vals[0];
        // M <= n/2 elements: both key buffers fit into keys[0], both payload buffers into the free
        // payload array, and the four per-element u32 arrays into keys[1]
        const u32 half = n / 2;                                // m <= n/2
        SortBuffers ab;
        ab.keys[0] = sb.keys[0];
        ab.keys[1] = sb.keys[0] + half;
        ab.vals[0] = other_vals;
        ab.vals[1] = other_vals + half;
        u32 *scratch = (u32 *)sb.keys[1];
        u32 *act_r1 = scratch, *newr = scratch + half, *pos = scratch + 2 * (size_t)half;
        u32 *next_idx = d_rs;                                  // rs is dead once the survivors are collected
        u32 *next_r1 = d_rs + half;
        CU(ctx, cudaMemsetAsync(d_rrctl, 0, rrctl_bytes, ctx->stream));
        LAUNCH(ctx, bwt_collect_active_kernel, rr_grid, RR_BLOCK, 0, d_rs, sa, n, rr_tiles, 0u, d_bact, ab.vals[0], act_r1);
        ActiveWork w;
        w.ab = ab; w.act_r1 = act_r1; w.newr = newr; w.pos = pos; w.next_idx = next_idx; w.next_r1 = next_r1;
        w.sa_buf = sa_buf; w.d_rank = d_rank; w.d_hist8 = d_hist8; w.d_rrctl = d_rrctl; w.rrctl_bytes = rrctl_bytes;
        w.d_counters = d_counters; w.d_ticket = d_ticket; w.d_status = d_status; w.cstatus = cstatus;
        w.arena_mark = arena_mark; w.rank_mask = rank_mask;
        w.zero_base = (u8 *)d_hist8;
        w.zero_bytes = (size_t)((u8 *)(cstatus + ac_tiles_max + 4) - (u8 *)d_hist8);
        RET(bwt_active_rounds(ctx, n, active, &k, w, &rounds, &passes_total));
        sa = sa_buf;
    }
    LAUNCH(ctx, bwt_gather_kernel, grid_for((n + 3) / 4, 256), 256, 0, d_in, sa, n, n, d_last);
    u32 *h_primary = (u32 *)(ctx->mailbox + 1040);
    CU(ctx, cudaMemcpyAsync(h_primary, d_rank, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    *primary = *h_primary;
    ctx->stats.bwt_rounds = rounds;
    ctx->stats.bwt_sort_passes = passes_total;
    return BZAP_OK;
}

// ---- building blocks shared with the distributed single-block path (dist_block.cu) --------------------------
// sparse ranks of an already sorted key run whose first element sits at global slot pos_base,
// without a host round trip: d_ctl = rerank_ctl_words(m) zeroed words whose
// words [1024] / [1025] end up holding {groups, singleton groups}; d_bact (148 * 6 + 8 words) receives the
// per-range survivor counts that dev_collect_active consumes
size_t rerank_ctl_words(u32 m)
{
    const u32 rr_tiles = (m + RR_TILE - 1) / RR_TILE;
    return 4 * 256 + 8 + 2 * ((size_t)rr_tiles + 8);
}
int dev_rerank_sorted(bzap_ctx *ctx, const u64 *d_keys, u32 m, u32 pos_base, u32 *d_rs, u32 *d_ctl, u32 *d_bact)
{
    const u32 rr_tiles = (m + RR_TILE - 1) / RR_TILE;
    u32 *d_hist4 = d_ctl, *d_counters = d_ctl + 4 * 256;
    u64 *d_status = (u64 *)(d_ctl + 4 * 256 + 8);
    LAUNCH(ctx, bwt_rerank_kernel, grid_for(rr_tiles, 1, 148 * 6), RR_BLOCK, 0, d_keys, (const u32 *)nullptr, m, rr_tiles,
           pos_base, (u32 *)nullptr, d_rs, d_hist4, d_counters, d_status, d_bact);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
// rotations of a sorted run that still sit in a group > 1: (start, group rank), in slot order
int dev_collect_active(bzap_ctx *ctx, const u32 *d_rs, const u32 *d_sa, u32 m, u32 pos_base, const u32 *d_bact, u32 *act_idx,
                       u32 *act_r1)
{
    const u32 rr_tiles = (m + RR_TILE - 1) / RR_TILE;
    LAUNCH(ctx, bwt_collect_active_kernel, grid_for(rr_tiles, 1, 148 * 6), RR_BLOCK, 0, d_rs, d_sa, m, rr_tiles, pos_base, d_bact,
           act_idx, act_r1);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
size_t active_ctl_bytes(u32 m, ActiveWork *w, u8 *base)
{
    // hist8 | counters[4] ticket[4] | status x 2 | compaction status : one block, zeroed per round
    const u32 mt = (m + AC_TILE - 1) / AC_TILE;
    const size_t status_u64 = 2 * ((size_t)mt + 2) + 8;
    const size_t bytes = 8 * 256 * sizeof(u32) + 8 * sizeof(u32) + status_u64 * sizeof(u64) + ((size_t)mt + 4) * sizeof(u64);
    if (w && base) {
        w->d_hist8 = (u32 *)base;
        w->d_counters = w->d_hist8 + 8 * 256;
        w->d_ticket = w->d_counters + 4;
        w->d_status = (u64 *)(w->d_ticket + 4);
        w->cstatus = w->d_status + status_u64;
        w->zero_base = base;
        w->zero_bytes = bytes;
        w->d_rrctl = nullptr;
        w->rrctl_bytes = 0;
    }
    return bytes;
}

// last[j] = text[(sa[j] + n - 1) mod n] for the m suffix-array slots held by this GPU
__global__ void __launch_bounds__(256)
bwt_gather_slots_kernel(const u8 *__restrict__ text, const u32 *__restrict__ sa, u32 n, u32 m, u8 *__restrict__ last)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        u32 s = sa[j];
        last[j] = text[s == 0 ? n - 1 : s - 1];
    }
}
int dev_gather_slots(bzap_ctx *ctx, const u8 *d_text, const u32 *d_sa, u32 n, u32 m, u8 *d_last)
{
    if (((uintptr_t)d_last & 3) == 0) LAUNCH(ctx, bwt_gather_kernel, grid_for((m + 3) / 4, 256), 256, 0, d_text, d_sa, n, m, d_last);
    else LAUNCH(ctx, bwt_gather_slots_kernel, grid_for(m, 256), 256, 0, d_text, d_sa, n, m, d_last);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
