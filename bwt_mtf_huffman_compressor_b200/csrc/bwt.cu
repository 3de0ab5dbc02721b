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
// Digit histograms for the sort passes are never computed from the keys: digit j of an 8-byte
// window is a text byte, and digit j of either key half is a digit of a rank, so one byte
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
__global__ void __launch_bounds__(IK_BLOCK) bwt_init_keys_kernel(const u8 *__restrict__ text, u32 n, u64 *__restrict__ keys)
{
    __shared__ u8 s_b[IK_TILE + 8];
    const u32 tiles = (n + IK_TILE - 1) / IK_TILE;
    for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const u32 base = tile * IK_TILE;
        __syncthreads();
        for (u32 i = threadIdx.x; i < IK_TILE + 8; i += IK_BLOCK) {
            u64 p = (u64)base + i;
            if (p >= n) p %= n;                   // cyclic window (main.cpp:38-44)
            s_b[i] = text[p];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IK_ITEMS; ++i) {
            u32 o = threadIdx.x + i * IK_BLOCK;
            if (base + o < n) {
                u64 k = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) k = (k << 8) | s_b[o + j];
                keys[base + o] = k;
            }
        }
    }
}

// hist8[p][d] = src[(p % src_rows)][d] : every pass digit shares the same few histograms
__global__ void bwt_spread_hist_kernel(const u32 *__restrict__ src, int src_rows, u32 *__restrict__ hist8)
{
    for (int p = 0; p < 8; ++p) hist8[p * 256 + threadIdx.x] = src[(p % src_rows) * 256 + threadIdx.x];
}

// ---- doubling keys ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwt_pair_keys_kernel(const u32 *__restrict__ rank, u32 n, u32 k, u64 *__restrict__ keys)
{
    const u32 kk = k % n;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u32 j = i + kk;                           // < 2n <= 2^31: no overflow
        if (j >= n) j -= n;
        keys[i] = ((u64)rank[i] << 32) | rank[j];
    }
}

// ---- re-rank ---------------------------------------------------------------------------------------
// sorted keys (+ payload = rotation start, nullptr = identity) -> rank[start] = sparse rank,
// number of groups, and the 4 x 256 histogram of rank digits for the next round's passes.
struct RrSmem {
    u64 keys[RR_TILE + 1];
    u32 r[RR_TILE];
    u32 hist[4][256];
    u32 tmp[40];
    u32 ticket;
    u32 tile_prefix;
    u32 heads;
};

__global__ void __launch_bounds__(RR_BLOCK)
bwt_rerank_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ sa, u32 n, u32 *__restrict__ rank,
                  u32 *hist4, u32 *groups, u64 *status, u32 *ticket)
{
    __shared__ RrSmem S;
    const u32 tid = threadIdx.x;
    const u32 tile = take_ticket(ticket, &S.ticket);
    const u32 base = tile * RR_TILE;
    for (u32 i = tid; i < 4 * 256; i += RR_BLOCK) (&S.hist[0][0])[i] = 0;
    if (tid == 0) {
        S.keys[0] = base ? keys[base - 1] : 0;
        S.heads = 0;
    }
#pragma unroll
    for (int i = 0; i < RR_ITEMS; ++i) {
        u32 o = tid + i * RR_BLOCK;
        S.keys[o + 1] = base + o < n ? keys[base + o] : 0;
    }
    __syncthreads();

    // blocked: thread owns RR_ITEMS consecutive sorted positions
    u32 loc[RR_ITEMS];
    u32 cur = 0, nheads = 0;
#pragma unroll
    for (int i = 0; i < RR_ITEMS; ++i) {
        u32 o = tid * RR_ITEMS + i;
        u32 p = base + o;
        bool head = p < n && (p == 0 || S.keys[o + 1] != S.keys[o]);
        if (head) { cur = p; ++nheads; }
        loc[i] = cur;
    }
    u32 total;
    u32 tprefix = block_exclusive_max(cur, S.tmp, &total);
    if (tid < 32) {
        u64 x = lookback_exclusive(status, tile, (u64)total, OpMax());
        if (tid == 0) S.tile_prefix = (u32)x;
    }
    if (nheads) atomicAdd(&S.heads, nheads);
    __syncthreads();
    const u32 pre = max(S.tile_prefix, tprefix);

    // ranks + run-aggregated digit histogram (ranks are non-decreasing along sorted order, so the
    // high digits are almost constant inside a thread's run)
    u32 run_d[4] = {0, 0, 0, 0}, run_c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < RR_ITEMS; ++i) {
        u32 o = tid * RR_ITEMS + i;
        if (base + o < n) {
            u32 r = max(pre, loc[i]);
            S.r[o] = r;
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

#pragma unroll
    for (int i = 0; i < RR_ITEMS; ++i) {
        u32 o = tid + i * RR_BLOCK;
        u32 p = base + o;
        if (p < n) rank[sa ? sa[p] : p] = S.r[o];
    }
    for (u32 i = tid; i < 4 * 256; i += RR_BLOCK) {
        u32 c = (&S.hist[0][0])[i];
        if (c) atomicAdd(&hist4[i], c);
    }
    if (tid == 0 && S.heads) atomicAdd(groups, S.heads);
}

// ---- last column --------------------------------------------------------------------------------------
// L[j] = text[(SA[j] + N - 1) mod N]   (main.cpp:87); sa == nullptr means SA = identity
__global__ void __launch_bounds__(256) bwt_gather_kernel(const u8 *__restrict__ text, const u32 *__restrict__ sa, u32 n,
                                                         u8 *__restrict__ last)
{
    const u32 nq = (n + 3) / 4;
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        u32 w = 0;
        u32 j0 = q * 4;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            u32 j = j0 + b;
            if (j < n) {
                u32 s = sa ? sa[j] : j;
                u32 c = text[s == 0 ? n - 1 : s - 1];
                w |= c << (8 * b);
            }
        }
        if (j0 + 3 < n) *reinterpret_cast<u32 *>(last + j0) = w;
        else
            for (u32 b = 0; j0 + b < n; ++b) last[j0 + b] = (u8)(w >> (8 * b));
    }
}


// ---- host driver -----------------------------------------------------------------------------------------
static inline u32 grid_for(size_t work_items, u32 per_block, u32 cap = 148u * 16u)
{
    size_t g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (u32)(g > cap ? cap : g);
}

int dev_bwt(bzap_ctx *ctx, const u8 *d_in, size_t n64, u8 *d_last, u64 *primary)
{
    if (n64 == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input");
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    ctx->sort_ev_used = 0;
    ctx->stats.sort_bytes = 0;
    ctx->stats.sort_elems = n;
    ctx->stats.ms_sort = 0;
    SortBuffers sb;
    sb.keys[0] = arena_get<u64>(ctx, n);
    sb.keys[1] = arena_get<u64>(ctx, n);
    sb.vals[0] = arena_get<u32>(ctx, n);
    sb.vals[1] = arena_get<u32>(ctx, n);
    u32 *d_rank = arena_get<u32>(ctx, n);
    const u32 rr_tiles = (n + RR_TILE - 1) / RR_TILE;
    // control: hist8[8*256] | hist4[4*256] groups ticket pad | status[rr_tiles] (u64)
    u32 *d_hist8 = arena_get<u32>(ctx, 8 * 256);
    u32 *d_rrctl = arena_get<u32>(ctx, 4 * 256 + 8 + 2 * (size_t)rr_tiles);
    if (!sb.keys[0] || !sb.keys[1] || !sb.vals[0] || !sb.vals[1] || !d_rank || !d_hist8 || !d_rrctl)
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "bwt scratch");
    u32 *d_hist4 = d_rrctl, *d_groups = d_rrctl + 4 * 256, *d_ticket = d_groups + 1;
    u64 *d_status = (u64 *)(d_rrctl + 4 * 256 + 8);
    const size_t rrctl_bytes = (4 * 256 + 8 + 2 * (size_t)rr_tiles) * sizeof(u32);
    const size_t arena_mark = ctx->arena_off;

    // round 0: byte histogram stands in for all eight digit histograms of the 8-byte windows
    CU(ctx, cudaMemsetAsync(d_hist4, 0, 256 * sizeof(u32), ctx->stream));
    RET(dev_byte_hist(ctx, d_in, n, d_hist4));
    LAUNCH(ctx, bwt_spread_hist_kernel, 1, 256, 0, d_hist4, 1, d_hist8);
    LAUNCH(ctx, bwt_init_keys_kernel, grid_for((n + IK_TILE - 1) / IK_TILE, 1, 148 * 8), IK_BLOCK, 0, d_in, n, sb.keys[0]);

    u64 *keys = nullptr;
    u32 *sa = nullptr;
    u32 rounds = 0, passes_total = 0, prev_groups = 0;
    u64 k = 8;
    u32 *h_groups = (u32 *)(ctx->mailbox + 1024);
    while (true) {
        int passes = 0;
        ctx->arena_off = arena_mark;             // sort control block is per round
        RET(dev_sort_pairs64(ctx, &sb, n, 64, d_hist8, true, &keys, &sa, &passes));
        passes_total += (u32)passes;
        CU(ctx, cudaMemsetAsync(d_rrctl, 0, rrctl_bytes, ctx->stream));
        LAUNCH(ctx, bwt_rerank_kernel, rr_tiles, RR_BLOCK, 0, keys, sa, n, d_rank, d_hist4, d_groups, d_status, d_ticket);
        CU(ctx, cudaMemcpyAsync(h_groups, d_groups, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        ++rounds;
        const u32 groups = *h_groups;
        if (groups == n || k >= n || groups == prev_groups) break;
        prev_groups = groups;
        LAUNCH(ctx, bwt_spread_hist_kernel, 1, 256, 0, d_hist4, 4, d_hist8);
        // keys always rebuilt into buffer 0 in text order; payload = identity again
        LAUNCH(ctx, bwt_pair_keys_kernel, grid_for(n, 256 * 4), 256, 0, d_rank, n, (u32)(k % n), sb.keys[0]);
        k *= 2;
    }
    LAUNCH(ctx, bwt_gather_kernel, grid_for((n + 3) / 4, 256), 256, 0, d_in, sa, n, d_last);
    u32 *h_primary = (u32 *)(ctx->mailbox + 1032);
    CU(ctx, cudaMemcpyAsync(h_primary, d_rank, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    *primary = *h_primary;
    ctx->stats.bwt_rounds = rounds;
    ctx->stats.bwt_sort_passes = passes_total;
    return BZAP_OK;
}
