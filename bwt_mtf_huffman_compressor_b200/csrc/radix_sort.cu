// radix_sort.cu -- hand-written onesweep LSD radix sort (8-bit digits) for sm_100a.
//
// Replaces std::stable_sort in bwt() (main.cpp:82) and bwt_reverse() (main.cpp:67).  One pass =
// ONE kernel that reads every (key, payload) once and writes it once:
//   * digit counts for all passes are known before the first pass: they come from the kernel that
//     PRODUCES the keys (bwt.cu: one byte histogram, or the rank-digit histogram fused into the
//     re-rank kernel), from radix_hist_u8 below, or in closed form (permutation payloads);
//   * each 256-thread block takes a tile by atomic ticket, ranks its keys with a warp-level
//     multisplit (peer masks by MATCH.ANY or by eight votes, chosen per pass from the digit
//     histogram) into per-warp shared-memory digit counters and publishes the tile's 256
//     digit counts (flag+count in one 32-bit status word);
//   * keys and payloads are regrouped by digit in shared memory; only then is the tile's global
//     base per digit resolved by decoupled look-back over the previous tiles' status words, four
//     words per step, so that the wait overlaps the predecessors' own progress;
//   * the regrouped tile goes out as contiguous per-digit runs.
// Passes whose digit is constant over all keys are skipped: large sorts read the per-pass
// "trivial" flags back (that is what makes periodic inputs, all keys equal, cheap), small sorts
// take a host-computed mask of digits that can vary at all instead of a round trip.
// Measured variants of this kernel: profiles/r1_onesweep_variants.md.
#include "device_common.cuh"
#include <cstring>

#ifndef RS_BLOCK
#define RS_BLOCK 256
#endif
#ifndef RS_MINBLOCKS
#define RS_MINBLOCKS 3
#endif
#ifndef RS_LB_WIDE
#define RS_LB_WIDE 4
#endif
#ifndef RS_SMALL_N
#define RS_SMALL_N (1u << 20)
#endif
#ifndef RS_PREFETCH_TILES
#define RS_PREFETCH_TILES 148
#endif
#define RS_WARPS (RS_BLOCK / 32)
#define RS_FLAG_AGG (1u << 30)
#define RS_FLAG_INCL (2u << 30)
#define RS_VALUE_MASK ((1u << 30) - 1)

template <typename KeyT, int ITEMS> struct RsSmem {
    KeyT keys[RS_BLOCK * ITEMS];
    u32 vals[RS_BLOCK * ITEMS];
    u32 whist[RS_WARPS][256];
    u32 adj[256];
    u32 scan_tmp[40];
    u32 ticket;
};

// m &= bit B of d set ? vote : ~vote, spelled out as bit test, select, VOTE and one three-input
// LOP3 (m & ~(vote ^ s), s = all ones when the bit is set): the compiler's own lowering of the C
// expression spends six instructions per bit (shift, and, compare, select, vote, combine)
template <int B> __device__ __forceinline__ void vote_bit(u32 &m, u32 d)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, v, s;\n\t"
                 "and.b32 t, %1, %2;\n\t"
                 "setp.ne.u32 p, t, 0;\n\t"
                 "selp.b32 s, -1, 0, p;\n\t"
                 "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
                 "lop3.b32 %0, %0, v, s, 0x90;\n\t}"
                 : "+r"(m)
                 : "r"(d), "n"(1 << B));
}

template <typename KeyT> __device__ __forceinline__ u32 digit_of(KeyT k, int shift)
{
    return (u32)(k >> shift) & 0xffu;
}

// offsets = exclusive global digit offsets of this pass (256); status = tiles*256 zeroed words
//
// KEYGEN: where the keys of this pass come from.  0 = keys_in.  The first pass of a BWT sort builds
// its 64-bit keys on the fly instead of reading an array another kernel wrote (bwt.cu):
// 1 = the 8-byte cyclic window of the text at each position (gen_src = text bytes, main.cpp:38-44),
// 2 = (rank[i] << 32 | rank[(i + gen_k) % n]) (gen_src = ranks in text order).
template <typename KeyT, int ITEMS, bool IOTA_VALS, bool WRITE_KEYS, int KEYGEN = 0>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINBLOCKS)
onesweep_pass_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out, const u32 *__restrict__ vals_in,
                     u32 *__restrict__ vals_out, u32 n, int shift, const u32 *__restrict__ offsets, u32 *status,
                     u32 *ticket, const void *__restrict__ gen_src, u32 gen_k, const u32 *__restrict__ rank_mode)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    RsSmem<KeyT, ITEMS> &S = *reinterpret_cast<RsSmem<KeyT, ITEMS> *>(smem_raw);
    constexpr u32 TILE = RS_BLOCK * ITEMS;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 tile = take_ticket(ticket, &S.ticket);
    const u32 tile_base = tile * TILE;
    const u32 valid = min(TILE, n - tile_base);
    const u32 wbase = tile_base + warp * (32u * ITEMS);

    KeyT key[ITEMS];
    u32 rnk[ITEMS];
    const bool full_tile = valid == TILE;          // uniform: all but the last tile skip the bounds checks
    if (KEYGEN == 1) {
        // stage the tile's bytes (+7 of the next tile, cyclic) in the payload array, which is idle until the regroup
        u8 *sb = reinterpret_cast<u8 *>(S.vals);
        const u8 *text = static_cast<const u8 *>(gen_src);
        if ((u64)tile_base + TILE + 8 <= n && (reinterpret_cast<uintptr_t>(text) & 15u) == 0) {
            // interior tile of an aligned text: one 16-byte load per thread (TILE = 16 bytes x RS_BLOCK)
            static_assert(ITEMS == 16 || KEYGEN != 1, "text staging assumes 16 bytes per thread");
            reinterpret_cast<uint4 *>(sb)[tid] = reinterpret_cast<const uint4 *>(text + tile_base)[tid];
            if (tid < 8) sb[TILE + tid] = text[tile_base + TILE + tid];
        } else {
            for (u32 i = tid; i < TILE + 8; i += RS_BLOCK) {
                u64 p = (u64)tile_base + i;
                if (p >= n) p %= n;
                sb[i] = text[p];
            }
        }
        __syncthreads();
        // window at byte o: three aligned words, funnel-shifted, then byte-swapped to big-endian order
        const u32 *sw = reinterpret_cast<const u32 *>(sb);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 o = warp * (32u * ITEMS) + j * 32u + lane;
            u32 w0 = sw[o >> 2], w1 = sw[(o >> 2) + 1], w2 = sw[(o >> 2) + 2];
            u32 sh = (o & 3u) * 8u;
            u32 first = __funnelshift_r(w0, w1, sh), second = __funnelshift_r(w1, w2, sh);
            u64 k = ((u64)__byte_perm(first, 0, 0x0123) << 32) | __byte_perm(second, 0, 0x0123);
            key[j] = (full_tile || tile_base + o < n) ? (KeyT)k : (KeyT)~(KeyT)0;
        }
    } else if (KEYGEN == 2) {
        const u32 *rank = static_cast<const u32 *>(gen_src);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 idx = wbase + j * 32u + lane;
            if (full_tile || idx < n) {
                u32 i2 = idx + gen_k;             // < 2n <= 2^31
                if (i2 >= n) i2 -= n;
                key[j] = (KeyT)(((u64)rank[idx] << 32) | rank[i2]);
            } else {
                key[j] = (KeyT)~(KeyT)0;
            }
        }
    } else if (full_tile) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) key[j] = keys_in[wbase + j * 32u + lane];
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 idx = wbase + j * 32u + lane;
            key[j] = idx < n ? keys_in[idx] : (KeyT)~(KeyT)0;
        }
    }
#if RS_PREFETCH_TILES > 0
    // tiles are handed out in order, so the tile RS_PREFETCH_TILES ahead is started about one wave of
    // resident blocks from now: pull its keys and payloads into L2 (one 128-byte line per request)
    if (KEYGEN == 0) {
        const u64 ptile = (u64)tile + RS_PREFETCH_TILES;
        if ((ptile + 1) * TILE <= n) {
            const char *kp = reinterpret_cast<const char *>(keys_in + ptile * TILE);
            for (u32 o = tid * 128u; o < TILE * (u32)sizeof(KeyT); o += RS_BLOCK * 128u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + o));
            if (!IOTA_VALS) {
                const char *vp = reinterpret_cast<const char *>(vals_in + ptile * TILE);
                for (u32 o = tid * 128u; o < TILE * 4u; o += RS_BLOCK * 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + o));
            }
        }
    } else if (KEYGEN == 2) {
        const u64 ptile = (u64)tile + RS_PREFETCH_TILES;
        if ((ptile + 1) * TILE + gen_k <= n) {       // both rank windows without wrap-around
            const char *r1 = reinterpret_cast<const char *>(static_cast<const u32 *>(gen_src) + ptile * TILE);
            const char *r2 = r1 + (size_t)gen_k * 4u;
            for (u32 o = tid * 128u; o < TILE * 4u; o += RS_BLOCK * 128u) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + o));
            }
        }
    }
#endif
    for (u32 i = lane; i < 256; i += 32) S.whist[warp][i] = 0;
    __syncwarp();

    // warp-level multisplit: rank of each key among the warp's keys with the same digit.  The peer
    // mask comes either from MATCH.ANY, whose cost on the SM-wide ADU pipe grows with the number of
    // distinct digits in the warp (~30 cycles for text bytes, ~65 for uniform digits), or from eight
    // votes, one per digit bit (flat cost, ~24 more instructions): *rank_mode says which one the
    // digit distribution of this pass favours (radix_offsets_kernel).
    u32 *wh = S.whist[warp];
    const bool by_votes = rank_mode != nullptr && *rank_mode != 0;     // uniform over the grid
    auto rank_item = [&](int j, u32 d, u32 m) {
        u32 leader = (u32)__ffs(m) - 1u;
        u32 prev = 0;
        if (lane == leader) {
            prev = wh[d];
            wh[d] = prev + (u32)__popc(m);
        }
        prev = __shfl_sync(FULL_MASK, prev, leader);
        rnk[j] = prev + (u32)__popc(m & lanemask_lt());
        __syncwarp();
    };
    if (by_votes) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 d = digit_of(key[j], shift);
            u32 m = FULL_MASK;
            vote_bit<0>(m, d); vote_bit<1>(m, d); vote_bit<2>(m, d); vote_bit<3>(m, d);
            vote_bit<4>(m, d); vote_bit<5>(m, d); vote_bit<6>(m, d); vote_bit<7>(m, d);
            rank_item(j, d, m);
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 d = digit_of(key[j], shift);
            rank_item(j, d, __match_any_sync(FULL_MASK, d));
        }
    }
    __syncthreads();

    // thread b owns digit b: prefix over warps, publish the tile's count, scan over digits
    const u32 bin = tid & 255u;
    const bool owner = RS_BLOCK == 256 || tid < 256;    // blocks may be wider than the 256 digit values
    u32 *my_status = status + (size_t)tile * 256 + bin;
    u32 count = 0, tile_start;
    {
        if (owner) {
            u32 run = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                u32 t = S.whist[w][bin];
                S.whist[w][bin] = run;
                run += t;
            }
            count = run;
            if (tile == 0) st_relaxed(my_status, RS_FLAG_INCL | count);
            else st_relaxed(my_status, RS_FLAG_AGG | count);
        }
        u32 total;
        tile_start = block_exclusive_sum(count, S.scan_tmp, &total);
        if (owner) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) S.whist[w][bin] += tile_start;
        }
    }
    // global base of digit `bin`: decoupled look-back over the previous tiles.  It only feeds the
    // final write-out, so it runs after the shared-memory regroup, which gives
    // the predecessors time to publish their inclusive prefixes; status words are fetched
    // RS_LB_WIDE at a time so that a deep walk is not a chain of dependent L2 round trips.
    auto look_back = [&]() {
        if (!owner) return;
        u32 excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
                u32 sv[RS_LB_WIDE];
#pragma unroll
                for (int q = 0; q < RS_LB_WIDE; ++q)
                    sv[q] = t - q >= 0 ? ld_relaxed(status + (size_t)(t - q) * 256 + bin) : RS_FLAG_INCL;
#pragma unroll
                for (int q = 0; q < RS_LB_WIDE; ++q) {
                    if (!done) {
                        u32 v = sv[q];
                        while ((v >> 30) == 0) v = ld_relaxed(status + (size_t)(t - q) * 256 + bin);
                        excl += v & RS_VALUE_MASK;
                        done = (v >> 30) == 2;
                    }
                }
                t -= RS_LB_WIDE;
            }
            st_relaxed(my_status, RS_FLAG_INCL | (excl + count));
        }
        S.adj[bin] = offsets[bin] + excl - tile_start;
    };
    __syncthreads();

    // regroup keys and payloads by digit in shared memory
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        u32 pos = wh[digit_of(key[j], shift)] + rnk[j];
        S.keys[pos] = key[j];
        rnk[j] = pos;
    }
    // payload loads are issued only now: the key registers are dead
    u32 val[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        u32 idx = wbase + j * 32u + lane;
        if (IOTA_VALS) val[j] = idx;
        else val[j] = (full_tile || idx < n) ? vals_in[idx] : 0u;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) S.vals[rnk[j]] = val[j];
    look_back();
    __syncthreads();

    // contiguous per-digit runs go out; padding keys (digit 255, highest tile index) sit last
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        u32 idx = tid + i * RS_BLOCK;
        if (full_tile || idx < valid) {
            KeyT k = S.keys[idx];
            u32 g = S.adj[digit_of(k, shift)] + idx;
            if (WRITE_KEYS) keys_out[g] = k;
            vals_out[g] = S.vals[idx];
        }
    }
}

// ---- digit offsets -------------------------------------------------------------------------------
#ifndef RS_VOTE_DISTINCT
#define RS_VOTE_DISTINCT 8
#endif
// expected number of distinct digit values among the 32 keys of a warp, x 2^16, contribution of one
// digit value that holds c of n keys: 1 - (1 - c/n)^32
__device__ __forceinline__ u32 distinct32_term(u32 c, u32 n)
{
    float q = 1.0f - (float)c / (float)n;
    q *= q; q *= q; q *= q; q *= q; q *= q;
    return (u32)((1.0f - q) * 65536.0f);
}

// hist: hist_rows x 256 counts, pass p uses row p % hist_rows (several digits often share one
// histogram, see bwt.cu).  offsets: passes x 256 exclusive sums.  trivial[p] = 1 when one digit
// value holds all n keys (the pass is the identity permutation).  mode[p] = 1 when a warp is
// expected to see at least RS_VOTE_DISTINCT distinct digits (rank by votes, see the pass kernel).
__global__ void radix_offsets_kernel(const u32 *__restrict__ hist, int hist_rows, u32 *__restrict__ offsets, u32 *trivial,
                                     u32 *mode, u32 n)
{
    __shared__ u32 s_tmp[40];
    const int p = blockIdx.x;                    // one block per pass
    u32 c = hist[(p % hist_rows) * 256 + threadIdx.x];
    u32 total;
    u32 ex = block_exclusive_sum(c, s_tmp, &total);
    offsets[p * 256 + threadIdx.x] = ex;
    if (c == n) trivial[p] = 1;
    u32 distinct;
    block_exclusive_sum(distinct32_term(c, n), s_tmp, &distinct);
    if (threadIdx.x == 0) mode[p] = distinct >= ((u32)RS_VOTE_DISTINCT << 16);
}

// byte histogram for the inverse-BWT counting sort (u8 keys, one digit)
__global__ void __launch_bounds__(256) radix_hist_u8_kernel(const u8 *__restrict__ in, u32 n, u32 *hist)
{
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const u32 nvec = n / 16;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        uint4 v = in4[i];
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(&s_h[w[k] & 0xff], 1u);
            atomicAdd(&s_h[(w[k] >> 8) & 0xff], 1u);
            atomicAdd(&s_h[(w[k] >> 16) & 0xff], 1u);
            atomicAdd(&s_h[w[k] >> 24], 1u);
        }
    }
    for (u32 i = nvec * 16 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&s_h[in[i]], 1u);
    __syncthreads();
    u32 c = s_h[threadIdx.x];
    if (c) atomicAdd(&hist[threadIdx.x], c);
}

// cum[0..256]: exclusive byte counts (cum[256] = n); mode as in radix_offsets_kernel
__global__ void radix_cum_u8_kernel(const u32 *__restrict__ hist, u32 *__restrict__ cum, u32 *mode)
{
    __shared__ u32 s_tmp[40];
    u32 c = hist[threadIdx.x];
    u32 total;
    u32 ex = block_exclusive_sum(c, s_tmp, &total);
    cum[threadIdx.x] = ex;
    if (threadIdx.x == 255) cum[256] = ex + c;
    u32 distinct;
    block_exclusive_sum(distinct32_term(c, total), s_tmp, &distinct);
    if (threadIdx.x == 0) *mode = distinct >= ((u32)RS_VOTE_DISTINCT << 16);
}

// ---- host drivers ------------------------------------------------------------------------------
#ifndef RS_ITEMS_64
#define RS_ITEMS_64 16
#endif
#define RS_ITEMS_8 16
#ifndef RS_ITEMS_32
#define RS_ITEMS_32 20
#endif

size_t sort_scratch_bytes(u32 n)
{
    size_t tiles = ((size_t)n + RS_BLOCK * RS_ITEMS_64 - 1) / (RS_BLOCK * RS_ITEMS_64);
    // 8 passes of status words + tickets + offsets + flags
    return 8 * (tiles * 256 * sizeof(u32) + 256) + 64 * 1024;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of the kernel: every context
// (bound to one device, used by one thread at a time) raises it once for itself and remembers that
// in its own bit mask, so a second context on another GPU of the same process is served as well
enum { ATTR_SORT64 = 1u << 0, ATTR_SORT8 = 1u << 1, ATTR_SORT32 = 1u << 2, ATTR_SORT64B = 1u << 3 };
template <typename K> static int set_smem_attr(bzap_ctx *ctx, K kernel, size_t bytes)
{
    CU(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return BZAP_OK;
}

int dev_sort_pairs64(bzap_ctx *ctx, SortBuffers *b, u32 n, u32 pass_mask, const u32 *d_hist, int hist_rows, bool vals_are_iota,
                     u64 **out_keys,
                     u32 **out_vals, int *passes_run, const SortKeyGen *gen)
{
    constexpr int ITEMS = RS_ITEMS_64;
    const int passes = 8;
    const u32 tiles = (n + RS_BLOCK * ITEMS - 1) / (RS_BLOCK * ITEMS);
    const size_t status_words = (size_t)tiles * 256;
    // control block: offsets[8][256], trivial[8], tickets[8], ranking modes[8], then status per pass
    u32 *d_ctl = arena_get<u32>(ctx, 8 * 256 + 24 + passes * status_words);
    if (!d_ctl) return bzap_fail(ctx, BZAP_ERR_NOMEM, "sort scratch");
    u32 *d_offsets = d_ctl, *d_trivial = d_ctl + 8 * 256, *d_ticket = d_trivial + 8, *d_mode = d_ticket + 8,
        *d_status = d_mode + 8;
    CU(ctx, cudaMemsetAsync(d_trivial, 0, (24 + passes * status_words) * sizeof(u32), ctx->stream));
    LAUNCH(ctx, radix_offsets_kernel, passes, 256, 0, d_hist, hist_rows, d_offsets, d_trivial, d_mode, n);
    u32 *h_trivial = (u32 *)ctx->mailbox;
    if (n <= RS_SMALL_N) {
        // small inputs are launch / sync latency bound: a host round trip to learn which passes are
        // trivial costs more than simply running them (a trivial pass is the identity permutation)
        for (int p = 0; p < 8; ++p) h_trivial[p] = !((pass_mask >> p) & 1u);     // digits that can vary at all
    } else {
        CU(ctx, cudaMemcpyAsync(h_trivial, d_trivial, 8 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }

    auto k_iota = onesweep_pass_kernel<u64, ITEMS, true, true>;
    auto k_vals = onesweep_pass_kernel<u64, ITEMS, false, true>;
    auto k_text = onesweep_pass_kernel<u64, ITEMS, true, true, 1>;
    auto k_pair = onesweep_pass_kernel<u64, ITEMS, true, true, 2>;
    const size_t smem = sizeof(RsSmem<u64, ITEMS>);
    if (!(ctx->attr_mask & ATTR_SORT64)) {
        RET(set_smem_attr(ctx, k_iota, smem));
        RET(set_smem_attr(ctx, k_vals, smem));
        RET(set_smem_attr(ctx, k_text, smem));
        RET(set_smem_attr(ctx, k_pair, smem));
        ctx->attr_mask |= ATTR_SORT64;
    }
    if (gen && !vals_are_iota) return bzap_fail(ctx, BZAP_ERR_ARG, "key generation needs an identity payload");
    int cur = 0, run = 0;
    bool have_vals = !vals_are_iota;
    bool timed = false;
    for (int p = 0; p < passes; ++p) {
        if (h_trivial[p]) continue;
        const bool full = (u64)n == ctx->stats.sort_elems;   // roofline evidence: full-size passes only
        if (full && !timed && ctx->sort_ev_used + 2 <= 128) {
            cudaEvent_t &e0 = ctx->sort_ev[ctx->sort_ev_used], &e1 = ctx->sort_ev[ctx->sort_ev_used + 1];
            if (!e0) CU(ctx, cudaEventCreate(&e0));
            if (!e1) CU(ctx, cudaEventCreate(&e1));
            CU(ctx, cudaEventRecord(e0, ctx->stream));
            timed = true;
        }
        // algorithmic bytes: key 8 B read + 8 B written, payload 4 B written (+ 4 B read unless generated)
        // (a text-window pass reads 1 B per key instead of 8)
        if (full) {
            ctx->stats.sort_bytes += (u64)n * (have_vals ? 24 : (gen && gen->mode == 1) ? 13 : 20);
            ++ctx->stats.bwt_full_passes;
        }
        if (!have_vals && gen)
            LAUNCH(ctx, (gen->mode == 1 ? k_text : k_pair), tiles, RS_BLOCK, smem, (const u64 *)nullptr, b->keys[cur ^ 1],
                   (const u32 *)nullptr, b->vals[cur ^ 1], n, 8 * p, d_offsets + p * 256, d_status + (size_t)p * status_words,
                   d_ticket + p, gen->src, gen->k, (const u32 *)(d_mode + p));
        else if (!have_vals)
            LAUNCH(ctx, k_iota, tiles, RS_BLOCK, smem, b->keys[cur], b->keys[cur ^ 1], (const u32 *)nullptr,
                   b->vals[cur ^ 1], n, 8 * p, d_offsets + p * 256, d_status + (size_t)p * status_words, d_ticket + p,
                   (const void *)nullptr, 0u, (const u32 *)(d_mode + p));
        else
            LAUNCH(ctx, k_vals, tiles, RS_BLOCK, smem, b->keys[cur], b->keys[cur ^ 1], b->vals[cur], b->vals[cur ^ 1], n,
                   8 * p, d_offsets + p * 256, d_status + (size_t)p * status_words, d_ticket + p, (const void *)nullptr, 0u,
                   (const u32 *)(d_mode + p));
        have_vals = true;
        cur ^= 1;
        ++run;
    }
    if (timed) {
        CU(ctx, cudaEventRecord(ctx->sort_ev[ctx->sort_ev_used + 1], ctx->stream));
        ctx->sort_ev_used += 2;
    }
    CU(ctx, cudaGetLastError());
    *out_keys = (gen && run == 0) ? nullptr : b->keys[cur];
    *out_vals = have_vals ? b->vals[cur] : nullptr;
    if (passes_run) *passes_run = run;
    return BZAP_OK;
}

int dev_byte_hist(bzap_ctx *ctx, const u8 *d_bytes, u32 n, u32 *d_hist256)
{
    u32 grid = min((n / 16 + 255) / 256 + 1, 148u * 8u);
    LAUNCH(ctx, radix_hist_u8_kernel, grid, 256, 0, d_bytes, n, d_hist256);
    return BZAP_OK;
}

int dev_sort_positions_by_byte(bzap_ctx *ctx, const u8 *d_bytes, u32 n, u32 *d_T, u32 *d_cum)
{
    constexpr int ITEMS = RS_ITEMS_8;
    const u32 tiles = (n + RS_BLOCK * ITEMS - 1) / (RS_BLOCK * ITEMS);
    const size_t status_words = (size_t)tiles * 256;
    u32 *d_ctl = arena_get<u32>(ctx, 256 + 8 + status_words);
    if (!d_ctl) return bzap_fail(ctx, BZAP_ERR_NOMEM, "sort scratch");
    u32 *d_hist = d_ctl, *d_ticket = d_ctl + 256, *d_status = d_ctl + 264;
    CU(ctx, cudaMemsetAsync(d_ctl, 0, (264 + status_words) * sizeof(u32), ctx->stream));
    u32 grid = min((n / 16 + 255) / 256 + 1, 148u * 8u);
    LAUNCH(ctx, radix_hist_u8_kernel, grid, 256, 0, d_bytes, n, d_hist);
    LAUNCH(ctx, radix_cum_u8_kernel, 1, 256, 0, d_hist, d_cum, d_ticket + 1);
    auto k = onesweep_pass_kernel<u8, ITEMS, true, false>;
    const size_t smem = sizeof(RsSmem<u8, ITEMS>);
    if (!(ctx->attr_mask & ATTR_SORT8)) {
        RET(set_smem_attr(ctx, k, smem));
        ctx->attr_mask |= ATTR_SORT8;
    }
    LAUNCH(ctx, k, tiles, RS_BLOCK, smem, d_bytes, (u8 *)nullptr, (const u32 *)nullptr, d_T, n, 0, d_cum, d_status,
           d_ticket, (const void *)nullptr, 0u, (const u32 *)(d_ticket + 1));
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

// ---- permutation scatter: out[perm[j]] = vals[j] ------------------------------------------------------------
// A direct scatter of 4-byte words over an array larger than L2 costs a 32-byte sector read AND
// write per element.  Instead: one onesweep pass buckets (perm, val) by the TOP 8 bits of the
// target index (perm is a permutation of 0..n-1, so the digit histogram is known in closed form),
// after which consecutive elements target one n/256-entry window that stays L2 resident and every
// sector is written back exactly once.
__global__ void perm_offsets_kernel(u32 n, int shift, u32 *offsets, u32 *mode)
{
    u64 d = threadIdx.x;
    u64 lo = d << shift;
    offsets[threadIdx.x] = (u32)(lo < n ? lo : n);
    if (threadIdx.x == 0) *mode = 1;             // >= 128 equally likely digit values: rank by votes
}

#ifndef SC_BLOCKS_PER_SM
#define SC_BLOCKS_PER_SM 3
#endif
__global__ void __launch_bounds__(256)
scatter_u32_kernel(const u32 *__restrict__ idx, const u32 *__restrict__ vals, u32 n, u32 *__restrict__ out)
{
    // streaming loads: the inputs are read once and must not push half-written target sectors out of L2
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) out[__ldcs(idx + j)] = __ldcs(vals + j);
}

int dev_scatter_perm(bzap_ctx *ctx, const u32 *d_perm, const u32 *d_vals, u32 n, u32 *d_out, u32 *d_tmp_idx,
                     u32 *d_tmp_vals)
{
    // few resident blocks: the grid-stride loop then sweeps the bucketed array as a narrow front and
    // the 32-byte target sectors are completed in L2 before they are written back.  Measured on
    // B200, BWT of 64 MiB: 16 blocks/SM (two waves) 11.96 ms, 8: 11.39, 4: 10.94, 3: 10.87, 2: 11.03,
    // 1: 11.92 (ncu at 16/SM: 54 % of the 4-byte stores missed L2, DRAM writes 3.8x the array).
    const u32 sgrid = min((n + 1023) / 1024, 148u * (u32)SC_BLOCKS_PER_SM);
    if (n <= (12u << 20)) {                      // target array fits L2 comfortably: scatter directly
        LAUNCH(ctx, scatter_u32_kernel, sgrid, 256, 0, d_perm, d_vals, n, d_out);
        return BZAP_OK;
    }
    constexpr int ITEMS = RS_ITEMS_32;
    const u32 tiles = (n + RS_BLOCK * ITEMS - 1) / (RS_BLOCK * ITEMS);
    const size_t status_words = (size_t)tiles * 256;
    u32 *d_ctl = arena_get<u32>(ctx, 256 + 8 + status_words);
    if (!d_ctl) return bzap_fail(ctx, BZAP_ERR_NOMEM, "scatter scratch");
    u32 *d_offsets = d_ctl, *d_ticket = d_ctl + 256, *d_status = d_ctl + 264;
    CU(ctx, cudaMemsetAsync(d_ticket, 0, (8 + status_words) * sizeof(u32), ctx->stream));
    int bits = 0;
    while (((u64)1 << bits) < n) ++bits;          // n <= 2^bits
    const int shift = bits > 8 ? bits - 8 : 0;
    LAUNCH(ctx, perm_offsets_kernel, 1, 256, 0, n, shift, d_offsets, d_ticket + 1);
    auto k = onesweep_pass_kernel<u32, ITEMS, false, true>;
    const size_t smem = sizeof(RsSmem<u32, ITEMS>);
    if (!(ctx->attr_mask & ATTR_SORT32)) {
        RET(set_smem_attr(ctx, k, smem));
        ctx->attr_mask |= ATTR_SORT32;
    }
    LAUNCH(ctx, k, tiles, RS_BLOCK, smem, d_perm, d_tmp_idx, d_vals, d_tmp_vals, n, shift, d_offsets, d_status, d_ticket,
           (const void *)nullptr, 0u, (const u32 *)(d_ticket + 1));
    LAUNCH(ctx, scatter_u32_kernel, sgrid, 256, 0, d_tmp_idx, d_tmp_vals, n, d_out);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

// out[idx[j] - idx_offset] = vals[j]
__global__ void __launch_bounds__(256)
scatter_offset_kernel(const u32 *__restrict__ idx, const u32 *__restrict__ vals, u32 m, u32 off, u32 *__restrict__ out)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) out[idx[j] - off] = vals[j];
}
// ---- bucket (index, value) pairs by the top 8 bits of the index (distributed path: ranks go home) --------
__global__ void __launch_bounds__(256) radix_hist_u32_digit_kernel(const u32 *__restrict__ keys, u32 m, int shift, u32 *hist)
{
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        atomicAdd(&s_h[(keys[i] >> shift) & 0xffu], 1u);
    __syncthreads();
    u32 c = s_h[threadIdx.x];
    if (c) atomicAdd(&hist[threadIdx.x], c);
}

// ---- single bucketing passes without a host round trip (dist_block.cu) ------------------------------------
// Stable regroup of (key, payload) by the 8-bit digit (key >> shift) & 255.  d_hist256 holds the digit
// counts (device; for the u32 form they are counted here).  d_ctl: bucket_ctl_words(m) words of scratch.
size_t bucket_ctl_words(u32 m)
{
    const size_t t64 = ((size_t)m + RS_BLOCK * RS_ITEMS_64 - 1) / (RS_BLOCK * RS_ITEMS_64);
    const size_t t32 = ((size_t)m + RS_BLOCK * RS_ITEMS_32 - 1) / (RS_BLOCK * RS_ITEMS_32);
    return 264 + 8 + (t64 > t32 ? t64 : t32) * 256;
}
int dev_bucket_pass_u64(bzap_ctx *ctx, const u64 *d_keys, const u32 *d_vals, u32 m, int shift, u64 *d_keys_out, u32 *d_vals_out,
                        const u32 *d_hist256, u32 *d_ctl)
{
    constexpr int ITEMS = RS_ITEMS_64;
    const u32 tiles = (m + RS_BLOCK * ITEMS - 1) / (RS_BLOCK * ITEMS);
    u32 *d_cum = d_ctl, *d_ticket = d_ctl + 264, *d_status = d_ctl + 272;
    CU(ctx, cudaMemsetAsync(d_ticket, 0, (8 + (size_t)tiles * 256) * sizeof(u32), ctx->stream));
    LAUNCH(ctx, radix_cum_u8_kernel, 1, 256, 0, d_hist256, d_cum, d_ticket + 1);
    auto k = onesweep_pass_kernel<u64, ITEMS, false, true>;
    const size_t smem = sizeof(RsSmem<u64, ITEMS>);
    if (!(ctx->attr_mask & ATTR_SORT64B)) {
        RET(set_smem_attr(ctx, k, smem));
        ctx->attr_mask |= ATTR_SORT64B;
    }
    LAUNCH(ctx, k, tiles, RS_BLOCK, smem, d_keys, d_keys_out, d_vals, d_vals_out, m, shift, d_cum, d_status, d_ticket,
           (const void *)nullptr, 0u, (const u32 *)(d_ticket + 1));
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
int dev_bucket_pass_u32(bzap_ctx *ctx, const u32 *d_keys, const u32 *d_vals, u32 m, int shift, u32 *d_keys_out, u32 *d_vals_out,
                        u32 *d_hist256, u32 *d_ctl)
{
    constexpr int ITEMS = RS_ITEMS_32;
    const u32 tiles = (m + RS_BLOCK * ITEMS - 1) / (RS_BLOCK * ITEMS);
    u32 *d_cum = d_ctl, *d_ticket = d_ctl + 264, *d_status = d_ctl + 272;
    CU(ctx, cudaMemsetAsync(d_ticket, 0, (8 + (size_t)tiles * 256) * sizeof(u32), ctx->stream));
    CU(ctx, cudaMemsetAsync(d_hist256, 0, 256 * sizeof(u32), ctx->stream));
    LAUNCH(ctx, radix_hist_u32_digit_kernel, min((m + 255) / 256, 148u * 8u), 256, 0, d_keys, m, shift, d_hist256);
    LAUNCH(ctx, radix_cum_u8_kernel, 1, 256, 0, d_hist256, d_cum, d_ticket + 1);
    auto k = onesweep_pass_kernel<u32, ITEMS, false, true>;
    const size_t smem = sizeof(RsSmem<u32, ITEMS>);
    if (!(ctx->attr_mask & ATTR_SORT32)) {
        RET(set_smem_attr(ctx, k, smem));
        ctx->attr_mask |= ATTR_SORT32;
    }
    LAUNCH(ctx, k, tiles, RS_BLOCK, smem, d_keys, d_keys_out, d_vals, d_vals_out, m, shift, d_cum, d_status, d_ticket,
           (const void *)nullptr, 0u, (const u32 *)(d_ticket + 1));
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
// out[idx[j] - off] = vals[j], no synchronisation
int dev_scatter_offset_async(bzap_ctx *ctx, const u32 *d_idx, const u32 *d_vals, u32 m, u32 off, u32 *d_out)
{
    if (m == 0) return BZAP_OK;
    LAUNCH(ctx, scatter_offset_kernel, min((m + 255) / 256, 148u * (u32)SC_BLOCKS_PER_SM), 256, 0, d_idx, d_vals, m, off, d_out);
    return BZAP_OK;
}
