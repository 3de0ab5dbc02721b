// mtf.cu -- chunk-parallel move-to-front and its inverse.
//
// Forward (replaces move_to_front, main.cpp:93-112).  The list before a chunk is fully determined
// by WHEN each symbol was last seen: symbols ordered by last occurrence, most recent first,
// followed by the never-seen symbols in ascending order (the initial list of main.cpp:96-97 with
// every move-to-front preserving the relative order of the others).  So:
//   1. mtf_last_kernel    one thread per chunk walks it backwards and records the last position of
//                         every symbol it contains (per-chunk table, 256 x u32);
//   2. mtf_scan_*         element-wise exclusive MAX-scan of those tables over chunks (the
//                         associative merge of "last occurrence" states), two levels;
//   3. mtf_lists_kernel   per chunk: order the 256 symbols by that key -> the chunk's start list;
//   4. mtf_apply_kernel   one thread per chunk runs the exact sequential MTF from its start list.
//
// Inverse (replaces move_to_front_reverse, main.cpp:114-130).  Moves are position based, so the
// net effect of a chunk is a permutation of list POSITIONS, independent of the list contents:
//   1. imtf_walk_kernel   per chunk: run the moves on the IDENTITY list -> P_c (256 bytes), and emit for every
//                         input the position, in the chunk's (still unknown) start list, of the symbol it
//                         selects -- the walk itself does not depend on what the list holds;
//   2. imtf_scan_*        scan of the permutations under composition, two levels -> true start lists;
//   3. imtf_map_kernel    out[i] = start_list[chunk(i)][position emitted in 1]: one table look-up per byte
//                         (round 1 walked every chunk twice, once for P_c and once from its true list).
//
// A thread keeps its 256-entry list as 64 packed words in shared memory, laid out [word][thread]
// (conflict free); a move-to-front is a byte-wise funnel shift over the first pos/4+1 words.
#include "device_common.cuh"

#ifndef MTF_THREADS
#define MTF_THREADS 128          // threads (= chunks) per block in the per-chunk kernels
#endif
#ifndef MTF_TARGET_CHUNKS
#define MTF_TARGET_CHUNKS (148u * 512u)
#endif
#define MTF_GROUP 128            // chunks per scan group

__device__ __forceinline__ u32 low_bytes_mask(u32 r)   // low (r+1) bytes
{
    return 0xffffffffu >> (8u * (3u - r));
}

// forward step on a packed list: returns the position of c and moves it to the front
__device__ __forceinline__ u32 mtf_step(u32 *list /* [64][MTF_THREADS] */, u32 c)
{
    const u32 c4 = c * 0x01010101u;
    u32 carry = c;
    u32 k = 0;
    while (true) {
        u32 w = list[k * MTF_THREADS];
        u32 x = w ^ c4;
        u32 z = (x - 0x01010101u) & ~x & 0x80808080u;    // lowest flag marks the first zero byte
        u32 sh = (w << 8) | carry;
        if (z) {
            u32 r = ((u32)__ffs(z) - 1u) >> 3;
            u32 m = low_bytes_mask(r);
            list[k * MTF_THREADS] = (sh & m) | (w & ~m);
            return 4u * k + r;
        }
        list[k * MTF_THREADS] = sh;
        carry = w >> 24;
        ++k;
    }
}

// inverse step: returns the symbol at position v and moves it to the front
__device__ __forceinline__ u32 imtf_step(u32 *list, u32 v)
{
    const u32 kt = v >> 2, r = v & 3u;
    u32 wt = list[kt * MTF_THREADS];
    u32 s = (wt >> (8u * r)) & 0xffu;
    if (v) {
        u32 carry = s;
        for (u32 k = 0; k < kt; ++k) {
            u32 w = list[k * MTF_THREADS];
            list[k * MTF_THREADS] = (w << 8) | carry;
            carry = w >> 24;
        }
        u32 m = low_bytes_mask(r);
        list[kt * MTF_THREADS] = (((wt << 8) | carry) & m) | (wt & ~m);
    }
    return s;
}

// ---- forward ---------------------------------------------------------------------------------------
// last[chunk][sym] = MTF_KEY_BASE + position of the last occurrence of sym inside the chunk (0 = absent).
// The table must be zero on entry.  Keys 1..256 are left to a start list handed over from another GPU
// (dist_block.cu): position i of that list carries key 256 - i.
#define MTF_KEY_BASE 257u
__global__ void __launch_bounds__(MTF_THREADS)
mtf_last_kernel(const u8 *__restrict__ in, u32 n, u32 chunk, u32 nchunks, u32 *__restrict__ last)
{
    __shared__ u32 s_seen[8][MTF_THREADS];
    const u32 c = blockIdx.x * MTF_THREADS + threadIdx.x;
    if (c >= nchunks) return;
#pragma unroll
    for (int w = 0; w < 8; ++w) s_seen[w][threadIdx.x] = 0;
    const u32 beg = c * chunk;
    const u32 end = min(n, beg + chunk);
    u32 *row = last + (size_t)c * 256;
    u32 found = 0;
    u32 p = end;
    // tail not covering a whole 16-byte vector, then vectors backwards
    while (p > beg && (p & 15u)) {
        --p;
        u32 s = in[p];
        u32 w = s_seen[s >> 5][threadIdx.x], bit = 1u << (s & 31u);
        if (!(w & bit)) { s_seen[s >> 5][threadIdx.x] = w | bit; row[s] = p + MTF_KEY_BASE; ++found; }
    }
    while (p > beg && found < 256) {
        p -= 16;
        uint4 v = *reinterpret_cast<const uint4 *>(in + p);
        u32 ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 15; q >= 0; --q) {
            u32 s = (ws[q >> 2] >> (8 * (q & 3))) & 0xffu;
            u32 w = s_seen[s >> 5][threadIdx.x], bit = 1u << (s & 31u);
            if (!(w & bit)) { s_seen[s >> 5][threadIdx.x] = w | bit; row[s] = p + q + MTF_KEY_BASE; ++found; }
        }
    }
}

// exclusive max-scan inside each group of MTF_GROUP chunks; thread = symbol
__global__ void __launch_bounds__(256) mtf_scan_group_kernel(u32 *__restrict__ last, u32 nchunks, u32 *__restrict__ group_tot)
{
    const u32 g = blockIdx.x, s = threadIdx.x;
    const u32 c0 = g * MTF_GROUP, c1 = min(nchunks, c0 + MTF_GROUP);
    u32 run = 0;
    for (u32 c = c0; c < c1; c += 8) {
        u32 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = c + i < c1 ? last[(size_t)(c + i) * 256 + s] : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c + i < c1) { last[(size_t)(c + i) * 256 + s] = run; run = max(run, v[i]); }
    }
    group_tot[(size_t)g * 256 + s] = run;
}

// Exclusive max-scan over the group totals.  Block b scans groups b * seg_len .. in place; with more than one
// block the scan is two-level: seg_tot[b] receives the block's total, a second launch (one block) scans those,
// and mtf_lists_kernel takes the maximum of both levels (592 groups: 19 + 19 dependent steps instead of 592).
// total (may be null): the "last occurrence" summary of the whole input, i.e. what a following piece of the
// same block needs to know about this one.
#define MTF_SEG 32
__global__ void __launch_bounds__(256)
mtf_scan_top_kernel(u32 *__restrict__ group_tot, u32 ngroups, u32 seg_len, u32 *__restrict__ seg_tot, u32 *__restrict__ total)
{
    const u32 s = threadIdx.x;
    const u32 g0 = blockIdx.x * seg_len, g1 = min(ngroups, g0 + seg_len);
    u32 run = 0;
    for (u32 g = g0; g < g1; g += 8) {
        u32 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = g + i < g1 ? group_tot[(size_t)(g + i) * 256 + s] : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (g + i < g1) { group_tot[(size_t)(g + i) * 256 + s] = run; run = max(run, v[i]); }
    }
    if (seg_tot) seg_tot[(size_t)blockIdx.x * 256 + s] = run;
    if (total) total[s] = run;
}

// one warp per chunk.  key[s] = last occurrence of symbol s before the chunk (0 = never seen).
// List = seen symbols by decreasing key, then the unseen ones in ascending symbol order.  Only the
// seen symbols (a few dozen on text) need ranking: they are compacted first, each is ranked against
// the compacted keys, and an unseen symbol's slot follows from ballots alone.
__global__ void __launch_bounds__(256)
mtf_lists_kernel(const u32 *__restrict__ last, const u32 *__restrict__ group_tot, const u32 *__restrict__ seg_tot,
                 const u32 *__restrict__ init, u32 nchunks, u8 *__restrict__ lists)
{
    __shared__ __align__(16) u32 s_key[8][256];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const u32 c = blockIdx.x * 8 + warp;
    if (c >= nchunks) return;
    const u32 g = c / MTF_GROUP;
    u32 key[8], seen_below[8];
    u32 nseen = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u32 s = lane + 32u * i;                       // symbols in increasing order over (i, lane)
        u32 k = max(last[(size_t)c * 256 + s], group_tot[(size_t)g * 256 + s]);
        if (seg_tot) k = max(k, seg_tot[(size_t)(g / MTF_SEG) * 256 + s]);
        if (init) k = max(k, init[s]);
        key[i] = k;
        u32 m = __ballot_sync(FULL_MASK, k != 0);
        seen_below[i] = nseen + (u32)__popc(m & lanemask_lt());     // seen symbols smaller than s
        if (k) s_key[warp][seen_below[i]] = k;
        nseen += (u32)__popc(m);
    }
    for (u32 t = nseen + lane; t < ((nseen + 3u) & ~3u); t += 32) s_key[warp][t] = 0;   // pad to a multiple of 4
    __syncwarp();
    u32 pos[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pos[i] = key[i] ? 0u : nseen + (lane + 32u * i) - seen_below[i];
    for (u32 t = 0; t < nseen; t += 4) {
        uint4 o = *reinterpret_cast<const uint4 *>(&s_key[warp][t]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (key[i]) pos[i] += (o.x > key[i]) + (o.y > key[i]) + (o.z > key[i]) + (o.w > key[i]);
    }
    u8 *out = lists + (size_t)c * 256;
#pragma unroll
    for (int i = 0; i < 8; ++i) out[pos[i]] = (u8)(lane + 32u * i);
}

// HIST: also count output symbols (fused 256-bin histogram, privatised per block)
__global__ void __launch_bounds__(MTF_THREADS)
mtf_apply_kernel(const u8 *__restrict__ in, u32 n, u32 chunk, u32 nchunks, const u8 *__restrict__ lists,
                 u8 *__restrict__ out)
{
    __shared__ u32 s_list[64 * MTF_THREADS];
    const u32 c = blockIdx.x * MTF_THREADS + threadIdx.x;
    if (c >= nchunks) return;
    u32 *list = s_list + threadIdx.x;
    const u32 *src = reinterpret_cast<const u32 *>(lists + (size_t)c * 256);
#pragma unroll 8
    for (int k = 0; k < 64; ++k) list[k * MTF_THREADS] = src[k];
    const u32 beg = c * chunk;
    const u32 end = min(n, beg + chunk);
    u32 p = beg;
    for (; p + 16 <= end; p += 16) {
        uint4 v = *reinterpret_cast<const uint4 *>(in + p);
        u32 ws[4] = {v.x, v.y, v.z, v.w};
        u32 os[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            u32 s = (ws[q >> 2] >> (8 * (q & 3))) & 0xffu;
            os[q >> 2] |= mtf_step(list, s) << (8 * (q & 3));
        }
        *reinterpret_cast<uint4 *>(out + p) = make_uint4(os[0], os[1], os[2], os[3]);
    }
    for (; p < end; ++p) out[p] = (u8)mtf_step(list, in[p]);
}

// ---- inverse ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MTF_THREADS)
imtf_walk_kernel(const u8 *__restrict__ in, u32 n, u32 chunk, u32 nchunks, u8 *__restrict__ perms, u8 *__restrict__ out)
{
    __shared__ u32 s_list[64 * MTF_THREADS];
    const u32 c = blockIdx.x * MTF_THREADS + threadIdx.x;
    if (c >= nchunks) return;
    u32 *list = s_list + threadIdx.x;
#pragma unroll 8
    for (u32 k = 0; k < 64; ++k) list[k * MTF_THREADS] = 0x03020100u + k * 0x04040404u;
    const u32 beg = c * chunk;
    const u32 end = min(n, beg + chunk);
    u32 p = beg;
    for (; p + 16 <= end; p += 16) {
        uint4 v = *reinterpret_cast<const uint4 *>(in + p);
        u32 ws[4] = {v.x, v.y, v.z, v.w};
        u32 os[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 16; ++q)
            os[q >> 2] |= imtf_step(list, (ws[q >> 2] >> (8 * (q & 3))) & 0xffu) << (8 * (q & 3));
        *reinterpret_cast<uint4 *>(out + p) = make_uint4(os[0], os[1], os[2], os[3]);      // positions in the start list
    }
    for (; p < end; ++p) out[p] = (u8)imtf_step(list, in[p]);
    u32 *dst = reinterpret_cast<u32 *>(perms + (size_t)c * 256);
#pragma unroll 8
    for (int k = 0; k < 64; ++k) dst[k] = list[k * MTF_THREADS];
}

// in-group exclusive composition: pre_c[j] = (P_c0 o ... o P_{c-1})[j] starting from the identity,
// written over perms[c]; group_tot[g] = composition of the whole group.  thread = position j.
// state after chunk: A'[j] = A[P_c[j]]
__global__ void __launch_bounds__(256) imtf_scan_group_kernel(u8 *__restrict__ perms, u32 nchunks, u8 *__restrict__ group_tot)
{
    __shared__ u8 s_cur[2][256];
    const u32 g = blockIdx.x, j = threadIdx.x;
    const u32 c0 = g * MTF_GROUP, c1 = min(nchunks, c0 + MTF_GROUP);
    u32 cur = j;
    int buf = 0;
    for (u32 c = c0; c < c1; ++c) {
        u8 *row = perms + (size_t)c * 256;
        u32 pc = row[j];
        row[j] = (u8)cur;
        s_cur[buf][j] = (u8)cur;
        __syncthreads();
        cur = s_cur[buf][pc];
        buf ^= 1;
    }
    group_tot[(size_t)g * 256 + j] = (u8)cur;
}

// Exclusive composition over the group totals, in place; two-level like the forward scan: block b composes
// groups b * seg_len .., seg_tot[b] receives the block's composite, a second launch scans those, and
// imtf_lists_kernel composes both levels (composition is associative: Top_g = Seg_b o Local_g).
__global__ void __launch_bounds__(256)
imtf_scan_top_kernel(u8 *__restrict__ group_tot, u32 ngroups, u32 seg_len, u8 *__restrict__ seg_tot)
{
    __shared__ u8 s_cur[2][256];
    const u32 j = threadIdx.x;
    const u32 g0 = blockIdx.x * seg_len, g1 = min(ngroups, g0 + seg_len);
    u32 cur = j;
    int buf = 0;
    for (u32 g = g0; g < g1; ++g) {
        u8 *row = group_tot + (size_t)g * 256;
        u32 pg = row[j];
        row[j] = (u8)cur;
        s_cur[buf][j] = (u8)cur;
        __syncthreads();
        cur = s_cur[buf][pg];
        buf ^= 1;
    }
    if (seg_tot) seg_tot[(size_t)blockIdx.x * 256 + j] = (u8)cur;
}

// start list of chunk c: A_c[j] = Top_g[ pre_c[j] ], Top_g = Seg_b o Local_g when the top scan ran in two levels
// (the initial list is the identity, main.cpp:117-118)
__global__ void __launch_bounds__(256)
imtf_lists_kernel(const u8 *__restrict__ perms, const u8 *__restrict__ group_tot, const u8 *__restrict__ seg_tot, u32 nchunks,
                  u8 *__restrict__ lists)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;   // byte index into the list array
    if (i >= (size_t)nchunks * 256) return;
    const u32 c = (u32)(i >> 8);
    const u32 g = c / MTF_GROUP;
    u32 v = group_tot[(size_t)g * 256 + perms[i]];
    if (seg_tot) v = seg_tot[(size_t)(g / MTF_SEG) * 256 + v];
    lists[i] = (u8)v;
}

// out[i] = lists[chunk(i)][out[i]], in place; one warp per chunk, its 256-byte start list in shared memory
#define IMAP_WARPS 8
__global__ void __launch_bounds__(32 * IMAP_WARPS)
imtf_map_kernel(u8 *__restrict__ out, u32 n, u32 chunk, u32 nchunks, const u8 *__restrict__ lists)
{
    __shared__ __align__(16) u8 s_list[IMAP_WARPS][256];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const u32 c = blockIdx.x * IMAP_WARPS + warp;
    if (c >= nchunks) return;
    reinterpret_cast<u64 *>(s_list[warp])[lane] = reinterpret_cast<const u64 *>(lists + (size_t)c * 256)[lane];
    __syncwarp();
    const u8 *L = s_list[warp];
    const u32 beg = c * chunk;
    const u32 end = min(n, beg + chunk);
    u32 p = beg + lane * 16;
    for (; p + 16 <= end; p += 32 * 16) {
        uint4 v = *reinterpret_cast<const uint4 *>(out + p);
        u32 ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            ws[k] = (u32)L[ws[k] & 0xffu] | ((u32)L[(ws[k] >> 8) & 0xffu] << 8) | ((u32)L[(ws[k] >> 16) & 0xffu] << 16) |
                    ((u32)L[ws[k] >> 24] << 24);
        *reinterpret_cast<uint4 *>(out + p) = make_uint4(ws[0], ws[1], ws[2], ws[3]);
    }
    // ragged tail of the last chunk (fewer than 16 bytes): the lane whose turn it would have been
    if (p < end)
        for (u32 q = p; q < end; ++q) out[q] = L[out[q]];
}

// ---- host drivers ------------------------------------------------------------------------------------
// chunk length: multiple of 16, aims at >= ~64K chunks on big inputs and enough threads on small ones
static u32 pick_chunk(u32 n)
{
    u32 target = MTF_TARGET_CHUNKS;
    u32 c = (n + target - 1) / target;
    c = (c + 15u) & ~15u;
    if (c < 128) c = 128;
    if (c > 2048) c = 2048;
    return c;
}

size_t mtf_scratch_bytes(size_t n)
{
    const u32 chunk = pick_chunk((u32)n);
    const size_t nchunks = (n + chunk - 1) / chunk + 2;
    return nchunks * 1280 + (nchunks / MTF_GROUP + 2) * 1024 + (nchunks / MTF_GROUP / 32 + 4) * 1024 + 8192;
}

// phase 1: per-chunk "last occurrence" tables and their scans; plan->d_total[256] = summary of the whole input
int dev_mtf_begin(bzap_ctx *ctx, const u8 *d_in, size_t n64, MtfPlan *plan)
{
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    plan->n = n;
    plan->chunk = pick_chunk(n);
    plan->nchunks = (n + plan->chunk - 1) / plan->chunk;
    plan->ngroups = (plan->nchunks + MTF_GROUP - 1) / MTF_GROUP;
    plan->d_last = arena_get<u32>(ctx, (size_t)plan->nchunks * 256 + 256);
    plan->d_gtot = arena_get<u32>(ctx, (size_t)plan->ngroups * 256 + 256);
    plan->d_lists = arena_get<u8>(ctx, (size_t)plan->nchunks * 256 + 256);
    plan->d_total = arena_get<u32>(ctx, 256);
    plan->d_seg = arena_get<u32>(ctx, (size_t)(plan->ngroups / MTF_SEG + 2) * 256);
    plan->two_level = 0;
    if (!plan->d_last || !plan->d_gtot || !plan->d_lists || !plan->d_total || !plan->d_seg) return bzap_fail(ctx, BZAP_ERR_NOMEM, "mtf scratch");
    if (n == 0) {
        CU(ctx, cudaMemsetAsync(plan->d_total, 0, 256 * sizeof(u32), ctx->stream));
        return BZAP_OK;
    }
    const u32 cblocks = (plan->nchunks + MTF_THREADS - 1) / MTF_THREADS;
    CU(ctx, cudaMemsetAsync(plan->d_last, 0, (size_t)plan->nchunks * 256 * sizeof(u32), ctx->stream));
    LAUNCH(ctx, mtf_last_kernel, cblocks, MTF_THREADS, 0, d_in, n, plan->chunk, plan->nchunks, plan->d_last);
    LAUNCH(ctx, mtf_scan_group_kernel, plan->ngroups, 256, 0, plan->d_last, plan->nchunks, plan->d_gtot);
    if (plan->ngroups > 2 * MTF_SEG) {
        const u32 nseg = (plan->ngroups + MTF_SEG - 1) / MTF_SEG;
        LAUNCH(ctx, mtf_scan_top_kernel, nseg, 256, 0, plan->d_gtot, plan->ngroups, (u32)MTF_SEG, plan->d_seg, (u32 *)nullptr);
        LAUNCH(ctx, mtf_scan_top_kernel, 1, 256, 0, plan->d_seg, nseg, nseg, (u32 *)nullptr, plan->d_total);
        plan->two_level = 1;
    } else {
        LAUNCH(ctx, mtf_scan_top_kernel, 1, 256, 0, plan->d_gtot, plan->ngroups, plan->ngroups, (u32 *)nullptr, plan->d_total);
        plan->two_level = 0;
    }
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
// phase 2: start lists and the transform itself; d_init (may be null) = keys of the list the input starts
// from when it is not the identity of main.cpp:96-97 (a later piece of a block spread over several GPUs)
int dev_mtf_finish(bzap_ctx *ctx, const u8 *d_in, const MtfPlan *plan, const u32 *d_init, u8 *d_out)
{
    if (plan->n == 0) return BZAP_OK;
    const u32 cblocks = (plan->nchunks + MTF_THREADS - 1) / MTF_THREADS;
    LAUNCH(ctx, mtf_lists_kernel, (plan->nchunks + 7) / 8, 256, 0, plan->d_last, plan->d_gtot,
           plan->two_level ? (const u32 *)plan->d_seg : (const u32 *)nullptr, d_init, plan->nchunks, plan->d_lists);
    LAUNCH(ctx, mtf_apply_kernel, cblocks, MTF_THREADS, 0, d_in, plan->n, plan->chunk, plan->nchunks, plan->d_lists, d_out);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

int dev_mtf(bzap_ctx *ctx, const u8 *d_in, size_t n64, u8 *d_out)
{
    if (n64 == 0) return BZAP_OK;
    MtfPlan plan;
    RET(dev_mtf_begin(ctx, d_in, n64, &plan));
    return dev_mtf_finish(ctx, d_in, &plan, nullptr, d_out);
}

int dev_imtf(bzap_ctx *ctx, const u8 *d_in, size_t n64, u8 *d_out)
{
    if (n64 == 0) return BZAP_OK;
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    const u32 chunk = pick_chunk(n);
    const u32 nchunks = (n + chunk - 1) / chunk;
    const u32 ngroups = (nchunks + MTF_GROUP - 1) / MTF_GROUP;
    u8 *d_perms = arena_get<u8>(ctx, (size_t)nchunks * 256);
    u8 *d_gtot = arena_get<u8>(ctx, (size_t)ngroups * 256);
    u8 *d_lists = arena_get<u8>(ctx, (size_t)nchunks * 256);
    u8 *d_seg = arena_get<u8>(ctx, (size_t)(ngroups / MTF_SEG + 2) * 256);
    if (!d_perms || !d_gtot || !d_lists || !d_seg) return bzap_fail(ctx, BZAP_ERR_NOMEM, "imtf scratch");
    const u32 cblocks = (nchunks + MTF_THREADS - 1) / MTF_THREADS;
    LAUNCH(ctx, imtf_walk_kernel, cblocks, MTF_THREADS, 0, d_in, n, chunk, nchunks, d_perms, d_out);
    LAUNCH(ctx, imtf_scan_group_kernel, ngroups, 256, 0, d_perms, nchunks, d_gtot);
    const bool two_level = ngroups > 2 * MTF_SEG;
    if (two_level) {
        const u32 nseg = (ngroups + MTF_SEG - 1) / MTF_SEG;
        LAUNCH(ctx, imtf_scan_top_kernel, nseg, 256, 0, d_gtot, ngroups, (u32)MTF_SEG, d_seg);
        LAUNCH(ctx, imtf_scan_top_kernel, 1, 256, 0, d_seg, nseg, nseg, (u8 *)nullptr);
    } else {
        LAUNCH(ctx, imtf_scan_top_kernel, 1, 256, 0, d_gtot, ngroups, ngroups, (u8 *)nullptr);
    }
    LAUNCH(ctx, imtf_lists_kernel, (u32)(((size_t)nchunks * 256 + 255) / 256), 256, 0, d_perms, d_gtot,
           two_level ? (const u8 *)d_seg : (const u8 *)nullptr, nchunks, d_lists);
    LAUNCH(ctx, imtf_map_kernel, (nchunks + IMAP_WARPS - 1) / IMAP_WARPS, 32 * IMAP_WARPS, 0, d_out, n, chunk, nchunks, d_lists);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
