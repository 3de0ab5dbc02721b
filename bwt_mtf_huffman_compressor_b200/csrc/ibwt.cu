// ibwt.cu -- inverse Burrows-Wheeler transform: counting-sort map + parallel list ranking.
//
// Replaces bwt_reverse (main.cpp:61-75):
//     l_shift = stable_sort(iota(N)) by L[.]            (main.cpp:65-67)
//     repeat N times: out[i] = L[l_shift[row]]; row = l_shift[row]     (main.cpp:70-73)
// With T = l_shift (built by one onesweep counting-sort pass on the byte keys, radix_sort.cu) and
// F(r) = L[T[r]] = the byte whose cumulative-count interval contains r, the walk is
//     x_0 = primary,  x_{i+1} = T[x_i],  out[i] = F(x_i).
// T is a permutation and may have MANY cycles (periodic input: "abab" -> T = [2,3,0,1]); the walk
// stays on the cycle of `primary` and laps it N / cycle_length times.  It is parallelised by
// work-efficient list ranking with pseudo-random splitters:
//   1. one splitter row per bucket of 2^slog rows (hashed offset), plus `primary` itself.  Large
//      blocks use 64-row buckets; small ones (latency bound: the longest sub-list, ~ stride x ln(nodes)
//      dependent loads, is what the walk waits for) use 8-row buckets;
//   2. ibwt_walk_len_kernel : every splitter walks T until it lands on the next splitter ->
//      reduced list (next splitter, sublist length), emitting its bytes into a 256-byte slot on the way;
//      a sub-list that fills its slot continues as a NEW node appended to the reduced list, so no
//      sub-list is longer than a slot;
//   3. ibwt_wyllie_kernel   : pointer jumping on the reduced list, cut open at `primary`, gives
//      every splitter on primary's cycle its distance to the end, hence its output offset;
//   4. ibwt_copy_slots_kernel: every reachable node copies its slot to its place in the output;
//   5. ibwt_extend_kernel   : out[i] = out[i mod cycle_length] when the cycle is shorter than N.
// Nothing assumes a single list, so splitters on other cycles are simply never reached.
#include "device_common.cuh"

#define IB_STRIDE_LOG_LARGE 6
#define IB_STRIDE_LOG_SMALL 3
#define IB_SMALL_N (2u << 20)
#define IB_NIL 0xffffffffu

__device__ __forceinline__ u32 ib_hash(u32 x)
{
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
// row of the splitter of bucket b
__device__ __forceinline__ u32 ib_splitter_row(u32 b, u32 n, u32 slog)
{
    const u32 stride = 1u << slog;
    u32 lo = b << slog;
    u32 span = min(stride, n - lo);
    u32 h = ib_hash(b);
    return lo + (span == stride ? (h & (stride - 1u)) : h % span);
}
// node id of the splitter sitting on `row`, or IB_NIL.  Node `nb` (= number of buckets) is `primary`.
__device__ __forceinline__ u32 ib_node_of(u32 row, u32 n, u32 nb, u32 primary, u32 slog)
{
    if (row == primary) return nb;
    u32 b = row >> slog;
    return ib_splitter_row(b, n, slog) == row ? b : IB_NIL;
}

// F(r): byte c with cum[c] <= r < cum[c+1].  A 4096-entry coarse table gives the first candidate
// byte of r's slice; a short forward scan finishes (cum is monotone, 256 steps over N rows).
#define IB_COARSE_LOG 12
struct FirstCol {
    u32 cum[257];
    u8 coarse[1u << IB_COARSE_LOG];
    u32 shift;
};
__device__ __forceinline__ void ib_first_col_init(FirstCol &F, const u32 *__restrict__ cum, u32 n)
{
    for (u32 i = threadIdx.x; i < 257; i += blockDim.x) F.cum[i] = cum[i];
    u32 bits = 0;
    while (((u64)1 << bits) < n) ++bits;
    const u32 shift = bits > IB_COARSE_LOG ? bits - IB_COARSE_LOG : 0;
    if (threadIdx.x == 0) F.shift = shift;
    __syncthreads();
    for (u32 i = threadIdx.x; i < (1u << IB_COARSE_LOG); i += blockDim.x) {
        u64 r = (u64)i << shift;                       // first row of the slice
        u32 lo = 0, hi = 256;                          // largest c with cum[c] <= r (cum[0] = 0)
        if (r >= n) lo = 255;
        else
            while (hi - lo > 1) {
                u32 mid = (lo + hi) >> 1;
                if (F.cum[mid] <= r) lo = mid; else hi = mid;
            }
        F.coarse[i] = (u8)lo;
    }
    __syncthreads();
}
__device__ __forceinline__ u32 ib_first_col(const FirstCol &F, u32 r)
{
    u32 c = F.coarse[r >> F.shift];
    while (F.cum[c + 1] <= r) ++c;
    return c;
}

// node[j] = (next << 32) | len ; a regular splitter that coincides with `primary` is dropped
// (next = NIL, len = 0): node nb covers that row.
// Sub-list lengths are geometric, so one thread per sub-list would leave most lanes of a warp idle
// while the longest one finishes.  Instead every lane keeps pulling sub-lists from a global work
// counter (warp-aggregated, only when >= IB_REFILL lanes are idle): all lanes step until the list
// of sub-lists is exhausted.
// The walk already visits every row once, so it also emits the output bytes F(row) -- into a
// fixed slot of IB_SLOT bytes per sub-list, because the sub-list's place in the output is only
// known after the ranking.  A sub-list that fills its slot (geometric lengths, mean = the stride: ~2 % of
// them at a stride of 64) hands over to a fresh node taken from an atomic counter (counter[1]); that node
// is only ever reached through its predecessor's link, never looked up by row.
#define IB_REFILL 8
#define IB_SLOT 256
__global__ void __launch_bounds__(256)
ibwt_walk_len_kernel(const u32 *__restrict__ T, u32 n, u32 nb, u32 slog, u32 primary, const u32 *__restrict__ cum,
                     u64 *__restrict__ node, u8 *__restrict__ slots, u32 node_cap, u32 *counter)
{
    __shared__ FirstCol F;
    ib_first_col_init(F, cum, n);
    const u32 lane = lane_id();
    u32 j = IB_NIL, row = 0, len = 0;
    bool exhausted = false;
    while (true) {
        const u32 idle = __ballot_sync(FULL_MASK, j == IB_NIL);
        if (idle == FULL_MASK && exhausted) break;
        if (!exhausted && __popc(idle) >= IB_REFILL) {
            const u32 want = (u32)__popc(idle);
            u32 base = 0;
            if (lane == 0) base = atomicAdd(counter, want);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (j == IB_NIL) {
                u32 mine = base + (u32)__popc(idle & lanemask_lt());
                if (mine <= nb) {
                    row = mine == nb ? primary : ib_splitter_row(mine, n, slog);
                    if (mine < nb && row == primary) node[mine] = ((u64)IB_NIL << 32);
                    else { j = mine; len = 0; }
                }
            }
            exhausted = base + want > nb;
        }
        if (j != IB_NIL) {
            if (len == IB_SLOT) {                            // slot full: the rest of the sub-list becomes a node of its own
                const u32 nj = nb + 1u + atomicAdd(counter + 1, 1u);
                if (nj < node_cap) {
                    node[j] = ((u64)nj << 32) | IB_SLOT;
                    j = nj;
                    len = 0;
                } else {
                    atomicExch(counter + 2, 1u);             // cannot happen with node_cap = 5/4 of the splitters; reported
                    node[j] = ((u64)IB_NIL << 32) | IB_SLOT;
                    j = IB_NIL;
                    continue;
                }
            }
            slots[(size_t)j * IB_SLOT + len] = (u8)ib_first_col(F, row);   // out[i] = F(x_i), main.cpp:71
            row = T[row];
            ++len;
            u32 nx = ib_node_of(row, n, nb, primary, slog);
            if (nx != IB_NIL) {
                // the list is cut open in front of `primary`: the node that reaches it becomes the tail
                node[j] = ((u64)(nx == nb ? IB_NIL : nx) << 32) | len;
                j = IB_NIL;
            }
        }
    }
}

// one pointer-jumping round: dist[j] += dist[next[j]]; next[j] = next[next[j]] -- followed IB_HOPS times inside
// the same snapshot `in`, so a round multiplies every node's span by IB_HOPS + 1 (three dependent loads instead
// of one, but 11 launches instead of 21 for a million nodes: the rounds are launch bound at ~9 us each)
// (nodes = the splitters, `primary`, and however many continuation nodes the walk appended: counter[1])
#define IB_HOPS 3
__global__ void __launch_bounds__(256)
ibwt_wyllie_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u32 base_nodes, const u32 *__restrict__ counter)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= base_nodes + counter[1]) return;
    u64 a = in[j];
#pragma unroll
    for (int h = 0; h < IB_HOPS; ++h) {
        const u32 nx = (u32)(a >> 32);
        if (nx == IB_NIL) break;
        const u64 b = in[nx];
        a = (b & 0xffffffff00000000ull) | (u32)((u32)a + (u32)b);
    }
    out[j] = a;
}

// dist[j] (from the ranking) = bytes from splitter j to the end of the opened list; the list holds
// cycle_len = dist[nb] bytes, so sub-list j belongs at out[cycle_len - dist[j] ...].
// One warp per sub-list copies its slot; sub-lists on other cycles never reach `primary` and are skipped.
__global__ void __launch_bounds__(256)
ibwt_copy_slots_kernel(u32 nb, const u32 *__restrict__ counter, const u64 *__restrict__ len_node, const u64 *__restrict__ ranked,
                       const u8 *__restrict__ slots, u8 *__restrict__ out)
{
    const u32 j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (j > nb + counter[1]) return;
    const u64 rk = ranked[j];
    if ((u32)(rk >> 32) != IB_NIL) return;
    const u32 len = min((u32)len_node[j], (u32)IB_SLOT);
    const u32 o = (u32)ranked[nb] - (u32)rk;
    const u8 *src = slots + (size_t)j * IB_SLOT;
    for (u32 t = lane; t < len; t += 32) out[o + t] = src[t];
}

__global__ void __launch_bounds__(256) ibwt_extend_kernel(u8 *out, u32 n, u32 cycle_len)
{
    for (u32 i = cycle_len + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = out[i % cycle_len];
}

int dev_ibwt(bzap_ctx *ctx, const u8 *d_last, size_t n64, u64 primary64, u8 *d_out)
{
    if (n64 == 0) return BZAP_OK;
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    if (primary64 >= n64) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "primary index %llu >= N %zu",
                                           (unsigned long long)primary64, n64);
    const u32 n = (u32)n64, primary = (u32)primary64;
    const u32 slog = n <= IB_SMALL_N ? IB_STRIDE_LOG_SMALL : IB_STRIDE_LOG_LARGE;
    const u32 nb = (u32)(((u64)n + (1u << slog) - 1) >> slog);
    const u32 base_nodes = nb + 1;
    const u32 nodes = base_nodes + base_nodes / 4 + 1024;        // room for the continuation nodes of full slots
    u32 *d_T = arena_get<u32>(ctx, n);
    u32 *d_cum = arena_get<u32>(ctx, 260);
    u64 *d_len = arena_get<u64>(ctx, nodes);
    u64 *d_rank[2] = {arena_get<u64>(ctx, nodes), arena_get<u64>(ctx, nodes)};
    u32 *d_work = arena_get<u32>(ctx, 8);
    u8 *d_slots = arena_get<u8>(ctx, (size_t)nodes * IB_SLOT);
    if (!d_T || !d_cum || !d_len || !d_rank[0] || !d_rank[1] || !d_work || !d_slots)
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "ibwt scratch");
    CU(ctx, cudaMemsetAsync(d_work, 0, 8 * sizeof(u32), ctx->stream));
    const u32 wgrid = base_nodes / 256 + 1 < 148u * 8u ? base_nodes / 256 + 1 : 148u * 8u;
    RET(dev_sort_positions_by_byte(ctx, d_last, n, d_T, d_cum));
    const u32 nblk = (nodes + 255) / 256;
    CU(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    LAUNCH(ctx, ibwt_walk_len_kernel, wgrid, 256, 0, d_T, n, nb, slog, primary, d_cum, d_len, d_slots, nodes, d_work);
    CU(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->stats.walk_bytes = 5ull * n;            // T[row] (4 B) read + one slot byte written per row
    // pointer jumping: after r rounds every node has jumped (IB_HOPS + 1)^r links
    int cur = 0;
    const u64 *src = d_len;
    for (u64 span = 1; span < nodes; span *= (IB_HOPS + 1)) {
        LAUNCH(ctx, ibwt_wyllie_kernel, nblk, 256, 0, src, d_rank[cur], base_nodes, d_work);
        src = d_rank[cur];
        cur ^= 1;
    }
    if (src == d_len) {                                  // single node: ranking is the identity
        CU(ctx, cudaMemcpyAsync(d_rank[0], d_len, nodes * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        src = d_rank[0];
    }
    LAUNCH(ctx, ibwt_copy_slots_kernel, (u32)(((size_t)nodes * 32 + 255) / 256), 256, 0, nb, d_work, d_len, src, d_slots, d_out);
    u64 *h_cycle = (u64 *)(ctx->mailbox + 1056);
    CU(ctx, cudaMemcpyAsync(h_cycle, src + nb, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(h_cycle + 1, d_work, 4 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (((const u32 *)(h_cycle + 1))[2]) return bzap_fail(ctx, BZAP_ERR_CUDA, "list ranking ran out of continuation nodes");
    const u32 cycle_len = (u32)*h_cycle;
    if ((u32)(*h_cycle >> 32) != IB_NIL || cycle_len == 0 || cycle_len > n)
        return bzap_fail(ctx, BZAP_ERR_CUDA, "list ranking failed (cycle %u)", cycle_len);
    if (cycle_len < n) LAUNCH(ctx, ibwt_extend_kernel, 148 * 8, 256, 0, d_out, n, cycle_len);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
