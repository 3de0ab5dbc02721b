// huffman_enc.cu -- Huffman model statistics and the bit packer.
//
// hist_first_kernel   256-bin histogram of the MTF stream with block-privatised shared-memory
//                     counters, plus the position of each symbol's first appearance (the reference
//                     creates its leaves in first-appearance order, main.cpp:238-244, and that
//                     order feeds the tie-break).  Replaces the two loops of main.cpp:235-244.
// huff_encode_kernel  replaces encode_with_huffman (main.cpp:158-172) + append_bit
//                     (io_utilities.h:87-94): ONE pass.  Each thread owns 16 consecutive symbols;
//                     code lengths are summed, scanned inside the block, and the tile's bit base
//                     comes from a decoupled look-back over the previous tiles (single-pass
//                     exclusive scan of code lengths).  Codes are concatenated MSB-first into a
//                     shared-memory bit buffer and the buffer goes out with 128-bit stores; the two
//                     seam words a tile shares with its neighbours are OR-merged atomically.
//                     Stream bit b lands in byte b>>3 under mask 0x80>>(b&7) (io_utilities.h:92).
#include "device_common.cuh"

// ---- histogram + first appearance -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
hist_first_kernel(const u8 *__restrict__ in, u32 n, unsigned long long *freq, u32 *first_pos)
{
    __shared__ u32 s_h[256];
    __shared__ u32 s_f[256];
    s_h[threadIdx.x] = 0;
    s_f[threadIdx.x] = 0xffffffffu;
    __syncthreads();
    const u32 nvec = n / 16;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    u32 zeros = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        uint4 v = in4[i];
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            u32 s = (w[q >> 2] >> (8 * (q & 3))) & 0xffu;
            u32 p = i * 16 + q;
            if (p < s_f[s]) atomicMin(&s_f[s], p);
            if (s == 0) ++zeros;                      // rank 0 dominates an MTF'd BWT: keep it in a register
            else atomicAdd(&s_h[s], 1u);
        }
    }
    for (u32 p = nvec * 16 + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        u32 s = in[p];
        if (p < s_f[s]) atomicMin(&s_f[s], p);
        atomicAdd(&s_h[s], 1u);
    }
    if (zeros) atomicAdd(&s_h[0], zeros);
    __syncthreads();
    u32 c = s_h[threadIdx.x];
    if (c) {
        atomicAdd(&freq[threadIdx.x], (unsigned long long)c);
        atomicMin(&first_pos[threadIdx.x], s_f[threadIdx.x]);
    }
}

// device part only: d_stat = 256 x u64 counts followed by 256 x u32 first positions (0xffffffff = absent)
int dev_hist_launch(bzap_ctx *ctx, const u8 *d_in, u32 n, u64 *d_stat)
{
    u32 *d_first = (u32 *)(d_stat + 256);
    CU(ctx, cudaMemsetAsync(d_stat, 0, 256 * sizeof(u64), ctx->stream));
    CU(ctx, cudaMemsetAsync(d_first, 0xff, 256 * sizeof(u32), ctx->stream));
    if (n == 0) return BZAP_OK;
    u32 grid = (u32)((n / 16 + 255) / 256 + 1);
    if (grid > 148 * 8) grid = 148 * 8;
    LAUNCH(ctx, hist_first_kernel, grid, 256, 0, d_in, n, (unsigned long long *)d_stat, d_first);
    return BZAP_OK;
}

int dev_hist(bzap_ctx *ctx, const u8 *d_in, size_t n64, u64 freq[256], u8 order[256], int *n_leaves)
{
    if (n64 == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input");
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    u64 *d_freq = arena_get<u64>(ctx, 256 + 128);
    if (!d_freq) return bzap_fail(ctx, BZAP_ERR_NOMEM, "hist scratch");
    RET(dev_hist_launch(ctx, d_in, n, d_freq));
    u64 *h = (u64 *)(ctx->mailbox + 2048);
    CU(ctx, cudaMemcpyAsync(h, d_freq, 256 * sizeof(u64) + 256 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    const u32 *h_first = (const u32 *)(h + 256);
    int cnt = 0;
    u32 pos[256];
    for (int s = 0; s < 256; ++s) {
        freq[s] = h[s];
        if (h[s]) { order[cnt] = (u8)s; pos[cnt] = h_first[s]; ++cnt; }
    }
    // leaves in order of first appearance (insertion sort over <=256 entries)
    for (int i = 1; i < cnt; ++i) {
        u8 s = order[i];
        u32 p = pos[i];
        int j = i - 1;
        while (j >= 0 && pos[j] > p) { order[j + 1] = order[j]; pos[j + 1] = pos[j]; --j; }
        order[j + 1] = s;
        pos[j + 1] = p;
    }
    *n_leaves = cnt;
    return BZAP_OK;
}

// ---- bit packer ------------------------------------------------------------------------------------------
#define ENC_BLOCK 256
#define ENC_ITEMS 16
#define ENC_TILE (ENC_BLOCK * ENC_ITEMS)

struct EncTable {
    u64 code[256];
    u8 len[256];
};

// bits: shared-memory words, MSB-first; word 0 is the 128-bit aligned word holding the tile's first bit
struct BitSink {
    u32 *bits;
    u32 wi;        // next word to emit
    u64 acc;       // pending bits, right aligned
    u32 nacc;      // number of pending bits (< 32 between pushes)
    bool first;
    __device__ __forceinline__ void emit(u32 w)
    {
        if (first) { atomicOr(&bits[wi], w); first = false; }   // shared with the previous thread
        else bits[wi] = w;                                       // interior word: exclusively ours
        ++wi;
    }
    __device__ __forceinline__ void push(u32 v, u32 l)           // l <= 32
    {
        acc = (acc << l) | v;
        nacc += l;
        if (nacc >= 32) {
            nacc -= 32;
            emit((u32)(acc >> nacc));
            acc &= (1ull << nacc) - 1ull;
        }
    }
    __device__ __forceinline__ void flush()
    {
        if (nacc) atomicOr(&bits[wi], (u32)(acc << (32 - nacc)));  // shared with the next thread
    }
};

__global__ void __launch_bounds__(ENC_BLOCK)
huff_encode_kernel(const u8 *__restrict__ in, u32 n, const EncTable *__restrict__ table, u32 *__restrict__ out_words,
                   u64 bit_base, u64 *status, u32 *ticket, u32 smem_words)
{
    extern __shared__ __align__(16) u32 s_bits[];
    __shared__ u64 s_code[256];
    __shared__ u8 s_len[256];
    __shared__ u32 s_tmp[40];
    __shared__ u32 s_ticket;
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x;
    const u32 tile = take_ticket(ticket, &s_ticket);
    s_code[tid] = table->code[tid];
    s_len[tid] = table->len[tid];
    for (u32 i = tid; i < smem_words; i += ENC_BLOCK) s_bits[i] = 0;
    __syncthreads();

    const u32 p0 = tile * ENC_TILE + tid * ENC_ITEMS;
    u32 sym[4] = {0, 0, 0, 0};
    u32 cnt = 0;
    if (p0 + ENC_ITEMS <= n) {
        uint4 v = *reinterpret_cast<const uint4 *>(in + p0);
        sym[0] = v.x; sym[1] = v.y; sym[2] = v.z; sym[3] = v.w;
        cnt = ENC_ITEMS;
    } else if (p0 < n) {
        cnt = n - p0;
        for (u32 q = 0; q < cnt; ++q) sym[q >> 2] |= (u32)in[p0 + q] << (8 * (q & 3));
    }
    u32 my_bits = 0;
#pragma unroll
    for (int q = 0; q < ENC_ITEMS; ++q)
        if ((u32)q < cnt) my_bits += s_len[(sym[q >> 2] >> (8 * (q & 3))) & 0xffu];

    u32 tile_bits;
    const u32 my_off = block_exclusive_sum(my_bits, s_tmp, &tile_bits);
    if (tid < 32) {
        u64 x = lookback_exclusive(status, tile, (u64)tile_bits, OpSum());
        if (tid == 0) s_base = x + bit_base;
    }
    __syncthreads();
    const u64 g0 = s_base;                           // first bit of this tile in the output
    const u64 w0 = (g0 >> 5) & ~3ull;                // 128-bit aligned output word of buffer word 0
    const u32 lead = (u32)(g0 - (w0 << 5));          // 0..127

    if (my_bits) {
        BitSink sink;
        sink.bits = s_bits;
        u32 pos = lead + my_off;
        sink.wi = pos >> 5;
        sink.acc = 0;
        sink.nacc = pos & 31u;                       // leading zero bits: OR-merged with the neighbour
        sink.first = true;
#pragma unroll
        for (int q = 0; q < ENC_ITEMS; ++q) {
            if ((u32)q < cnt) {
                u32 s = (sym[q >> 2] >> (8 * (q & 3))) & 0xffu;
                u64 c = s_code[s];
                u32 l = s_len[s];
                if (l > 32) { sink.push((u32)(c >> 32), l - 32); sink.push((u32)c, 32); }
                else sink.push((u32)c, l);
            }
        }
        sink.flush();
    }
    __syncthreads();

    if (tile_bits == 0) return;
    const u64 g1 = g0 + tile_bits;                   // one past the last bit
    const u64 first_w = g0 >> 5, last_w = (g1 - 1) >> 5;
    // words fully owned by this tile
    const u64 own_lo = (g0 & 31) ? first_w + 1 : first_w;
    const u64 own_hi = (g1 & 31) ? last_w : last_w + 1;          // exclusive
    const u32 nvec = (u32)(((last_w - w0) >> 2) + 1);
    uint4 *out4 = reinterpret_cast<uint4 *>(out_words + w0);
    for (u32 v = tid; v < nvec; v += ENC_BLOCK) {
        uint4 x = *reinterpret_cast<const uint4 *>(&s_bits[4 * v]);
        x.x = __byte_perm(x.x, 0, 0x0123);           // MSB-first bit string -> byte order in memory
        x.y = __byte_perm(x.y, 0, 0x0123);
        x.z = __byte_perm(x.z, 0, 0x0123);
        x.w = __byte_perm(x.w, 0, 0x0123);
        const u64 gw = w0 + 4ull * v;
        if (gw >= own_lo && gw + 4 <= own_hi) {
            out4[v] = x;
        } else {
            u32 xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 w = gw + k;
                if (w < first_w || w > last_w) continue;
                if (w >= own_lo && w < own_hi) out_words[w] = xs[k];
                else if (xs[k]) atomicOr(&out_words[w], xs[k]);
            }
        }
    }
}

int dev_huff_encode(bzap_ctx *ctx, const u8 *d_in, size_t n64, const CodeTable *ct, u8 *d_file, u64 bit_base)
{
    if (n64 == 0) return BZAP_OK;
    const u32 n = (u32)n64;
    const u32 tiles = (n + ENC_TILE - 1) / ENC_TILE;
    EncTable *d_table = arena_get<EncTable>(ctx, 1);
    u64 *d_status = arena_get<u64>(ctx, (size_t)tiles + 2);
    if (!d_table || !d_status) return bzap_fail(ctx, BZAP_ERR_NOMEM, "encode scratch");
    u32 *d_ticket = (u32 *)(d_status + tiles);
    EncTable *h_table = (EncTable *)(ctx->mailbox + 8192);
    for (int s = 0; s < 256; ++s) { h_table->code[s] = ct->code[s]; h_table->len[s] = ct->len[s]; }
    CU(ctx, cudaMemcpyAsync(d_table, h_table, sizeof(EncTable), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_status, 0, ((size_t)tiles + 2) * sizeof(u64), ctx->stream));
    // worst-case bits of one tile + 128 bits of alignment lead, in 32-bit words, rounded to 16 bytes
    u32 words = (u32)(((u64)ENC_TILE * (u32)ct->max_len + 128 + 31) / 32 + 4);
    words = (words + 3u) & ~3u;
    const size_t smem = (size_t)words * sizeof(u32);
    if (smem > 200 * 1024) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "code length %d", ct->max_len);
    // the attribute is per function and device, not per launch: contexts on other threads (batch API)
    // may need a different size at the same time, so every context raises it once to the largest
    // size ever accepted
    if (smem > 40 * 1024 && !(ctx->attr_mask & (1u << 8))) {
        CU(ctx, cudaFuncSetAttribute(huff_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_mask |= 1u << 8;
    }
    LAUNCH(ctx, huff_encode_kernel, tiles, ENC_BLOCK, smem, d_in, n, d_table, (u32 *)d_file, bit_base, d_status, d_ticket,
           words);
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}
