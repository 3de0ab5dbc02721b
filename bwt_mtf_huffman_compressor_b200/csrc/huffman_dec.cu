// huffman_dec.cu -- chunk-parallel, speculative, self-synchronising Huffman decoder.
//
// Replaces huffman_reverse (main.cpp:259-281), which reads one bit at a time and looks a growing
// std::vector<bool> up in a hash map.  The reference format is ONE continuous MSB-first bit
// stream with no restart markers, so the stream is cut into fixed subsequences of DEC_SUBSEQ_BITS
// bits, one per thread, and the code-word boundaries are found by fixed-point iteration:
//   start[i]  = offset inside subsequence i of the first code word that STARTS in it (start[0]=0)
//   a thread decodes from start[i] until it crosses into subsequence i+1; where it lands is the
//   next subsequence's start.  All starts are guessed 0 and the sweep is repeated (only where the
//   input changed) until nothing moves.  Prefix codes self-synchronise after a few code words, so
//   in practice the second sweep already changes almost nothing; in the worst case (a stream that
//   never synchronises) the fixed point is still reached, one subsequence per sweep -- the slow
//   but correct on-GPU path.  Then: symbol counts -> exclusive scan -> final sweep that writes.
// Code words are resolved with a 12-bit primary LUT in shared memory and 8-bit sub tables in
// global memory for deeper trees (the file's tree is arbitrary, not canonical; any depth works).
// Decoding stops after exactly N symbols (main.cpp:268); padding bits are ignored; the empty code
// of a single-leaf tree consumes no bits (main.cpp:137-140, 271).
#include "device_common.cuh"

#define DEC_BLOCK 128
#define DEC_SUBSEQ_BITS 1024
#define DEC_SUBSEQ_WORDS (DEC_SUBSEQ_BITS / 32)
#define DEC_STRIDE (DEC_SUBSEQ_WORDS + 1)          // padded shared-memory stride (bank conflicts)
#define DEC_SLACK_WORDS 16                         // a code word may run <= 255 bits + refill past the end

struct DecSmem {
    u32 words[DEC_BLOCK * DEC_STRIDE + DEC_STRIDE];   // the block's subsequences + slack
    u16 lut[1 << DEC_PRIMARY_BITS];
};

// bit reader over the staged words of one block; `w` indexes payload words relative to the block
struct BitReader {
    const u32 *sw;
    u32 widx;        // next word (block relative) to load
    u64 win;         // MSB-aligned window
    u32 avail;       // valid bits in win
    __device__ __forceinline__ u32 word(u32 i) const { return sw[(i / DEC_SUBSEQ_WORDS) * DEC_STRIDE + (i % DEC_SUBSEQ_WORDS)]; }
    __device__ __forceinline__ void init(const u32 *s, u32 bitpos)
    {
        sw = s;
        widx = bitpos >> 5;
        u32 o = bitpos & 31u;
        win = (u64)word(widx) << (32 + o);
        avail = 32 - o;
        ++widx;
    }
    __device__ __forceinline__ void refill()
    {
        if (avail <= 32) {
            win |= (u64)word(widx) << (32 - avail);
            avail += 32;
            ++widx;
        }
    }
    __device__ __forceinline__ void skip(u32 nb) { win <<= nb; avail -= nb; }
};

// decodes one code word; returns the symbol and advances *pos by the code length
__device__ __forceinline__ u32 decode_one(BitReader &br, const u16 *lut, const u16 *__restrict__ sub, u32 &pos)
{
    br.refill();
    u32 e = lut[(u32)(br.win >> (64 - DEC_PRIMARY_BITS))];
    if (e & 0x8000u) {
        u32 nb = ((e >> 8) & 0xfu) + 1u;
        br.skip(nb);
        pos += nb;
        return e & 0xffu;
    }
    br.skip(DEC_PRIMARY_BITS);
    pos += DEC_PRIMARY_BITS;
    while (true) {
        br.refill();
        u32 f = sub[(e << DEC_SUB_BITS) + (u32)(br.win >> (64 - DEC_SUB_BITS))];
        if (f & 0x8000u) {
            u32 nb = ((f >> 8) & 0xfu) + 1u;
            br.skip(nb);
            pos += nb;
            return f & 0xffu;
        }
        br.skip(DEC_SUB_BITS);
        pos += DEC_SUB_BITS;
        e = f;
    }
}

__device__ __forceinline__ void dec_stage(DecSmem &S, const u32 *__restrict__ payload_words, u32 total_words,
                                          const u16 *__restrict__ tables, u32 block_first_word)
{
    const u32 nw = DEC_BLOCK * DEC_SUBSEQ_WORDS + DEC_SLACK_WORDS;
    for (u32 i = threadIdx.x; i < nw; i += DEC_BLOCK) {
        u32 g = block_first_word + i;
        u32 w = g < total_words ? payload_words[g] : 0u;
        S.words[(i / DEC_SUBSEQ_WORDS) * DEC_STRIDE + (i % DEC_SUBSEQ_WORDS)] = __byte_perm(w, 0, 0x0123);
    }
    for (u32 i = threadIdx.x; i < (1u << DEC_PRIMARY_BITS); i += DEC_BLOCK) S.lut[i] = tables[i];
    __syncthreads();
}

// One synchronisation sweep (Jacobi style: reads start_in/dirty_in, writes start_out/dirty_out).
// count[i] = number of code words starting inside subsequence i.
__global__ void __launch_bounds__(DEC_BLOCK)
huff_dec_sync_kernel(const u32 *__restrict__ payload_words, u32 total_words, u64 payload_bits,
                     const u16 *__restrict__ tables, u32 nsub, const u32 *__restrict__ start_in,
                     const u8 *__restrict__ dirty_in, u32 *__restrict__ start_out, u8 *__restrict__ dirty_out,
                     u32 *__restrict__ count, u32 *changed)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    DecSmem &S = *reinterpret_cast<DecSmem *>(smem_raw);
    const u32 i = blockIdx.x * DEC_BLOCK + threadIdx.x;
    // does any thread of the block have work?  (uniform: needed before the staging barrier)
    const bool dirty = i < nsub && dirty_in[i];
    if (i == 0) { start_out[0] = 0; dirty_out[0] = 0; }     // the stream starts at bit 0, for ever
    if (!__syncthreads_or(dirty)) {
        if (i < nsub && i + 1 < nsub) { start_out[i + 1] = start_in[i + 1]; dirty_out[i + 1] = 0; }
        return;
    }
    dec_stage(S, payload_words, total_words, tables, blockIdx.x * DEC_BLOCK * DEC_SUBSEQ_WORDS);
    if (i >= nsub) return;
    if (!dirty) {
        if (i + 1 < nsub) { start_out[i + 1] = start_in[i + 1]; dirty_out[i + 1] = 0; }
        return;
    }
    const u16 *sub = tables + (1u << DEC_PRIMARY_BITS);
    const u64 sub_begin = (u64)i * DEC_SUBSEQ_BITS;
    u64 limit = payload_bits - sub_begin;              // bits available from this subsequence on
    u32 pos = start_in[i];                             // relative to the subsequence
    u32 n_sym = 0;
    if (pos < DEC_SUBSEQ_BITS && pos < limit) {
        BitReader br;
        br.init(S.words, threadIdx.x * DEC_SUBSEQ_BITS + pos);
        const u32 stop = limit < DEC_SUBSEQ_BITS ? (u32)limit : DEC_SUBSEQ_BITS;
        while (pos < stop) {
            decode_one(br, S.lut, sub, pos);
            ++n_sym;
        }
    }
    count[i] = n_sym;
    if (i + 1 < nsub) {
        u32 land = pos >= DEC_SUBSEQ_BITS ? pos - DEC_SUBSEQ_BITS : 0u;
        u32 old = start_in[i + 1];
        start_out[i + 1] = land;
        u8 d = land != old;
        dirty_out[i + 1] = d;
        if (d) *changed = 1u;
    }
}

// final sweep: decode again from the settled starts and write out[prefix[i] + j], j < count
__global__ void __launch_bounds__(DEC_BLOCK)
huff_dec_write_kernel(const u32 *__restrict__ payload_words, u32 total_words, u64 payload_bits,
                      const u16 *__restrict__ tables, u32 nsub, const u32 *__restrict__ start,
                      const u64 *__restrict__ prefix, u32 n, u8 *__restrict__ out)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    DecSmem &S = *reinterpret_cast<DecSmem *>(smem_raw);
    const u32 i = blockIdx.x * DEC_BLOCK + threadIdx.x;
    dec_stage(S, payload_words, total_words, tables, blockIdx.x * DEC_BLOCK * DEC_SUBSEQ_WORDS);
    if (i >= nsub) return;
    const u16 *sub = tables + (1u << DEC_PRIMARY_BITS);
    const u64 sub_begin = (u64)i * DEC_SUBSEQ_BITS;
    const u64 limit = payload_bits - sub_begin;
    u32 pos = start[i];
    u64 o = prefix[i];
    if (o >= n || pos >= DEC_SUBSEQ_BITS || pos >= limit) return;
    BitReader br;
    br.init(S.words, threadIdx.x * DEC_SUBSEQ_BITS + pos);
    const u32 stop = limit < DEC_SUBSEQ_BITS ? (u32)limit : DEC_SUBSEQ_BITS;
    u32 oo = (u32)o;
    // bytes until the output index is 4-aligned, then packed 32-bit stores
    while (pos < stop && oo < n && (oo & 3u)) out[oo++] = (u8)decode_one(br, S.lut, sub, pos);
    while (pos < stop && oo < n) {
        u32 w = 0, k = 0;
        while (k < 4 && pos < stop && oo + k < n) { w |= decode_one(br, S.lut, sub, pos) << (8 * k); ++k; }
        if (k == 4) *reinterpret_cast<u32 *>(out + oo) = w;
        else for (u32 q = 0; q < k; ++q) out[oo + q] = (u8)(w >> (8 * q));
        oo += k;
    }
}

// exclusive sum scan u32 -> u64, single pass with look-back; out has count+1 entries (last = total)
#define SC_BLOCK 256
#define SC_ITEMS 8
__global__ void __launch_bounds__(SC_BLOCK)
scan_u32_u64_kernel(const u32 *__restrict__ in, u32 count, u64 *__restrict__ out, u64 *status, u32 *ticket)
{
    __shared__ u32 s_tmp[40];
    __shared__ u32 s_ticket;
    __shared__ u64 s_base;
    const u32 tile = take_ticket(ticket, &s_ticket);
    const u32 base = tile * SC_BLOCK * SC_ITEMS + threadIdx.x * SC_ITEMS;
    u32 v[SC_ITEMS];
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { v[k] = base + k < count ? in[base + k] : 0u; sum += v[k]; }
    u32 total;
    u32 ex = block_exclusive_sum(sum, s_tmp, &total);
    if (threadIdx.x < 32) {
        u64 x = lookback_exclusive(status, tile, (u64)total, OpSum());
        if (threadIdx.x == 0) s_base = x;
    }
    __syncthreads();
    u64 run = s_base + ex;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < count) out[base + k] = run;
        run += v[k];
        if (base + k + 1 == count) out[count] = run;
    }
}

__global__ void fill_kernel(u8 *out, u32 n, u8 value)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = value;
}

int dev_huff_decode(bzap_ctx *ctx, const u8 *d_payload, size_t payload_len, const DecodeTables *dt, size_t n64,
                    u8 *d_out)
{
    if (n64 == 0) return BZAP_OK;
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const u32 n = (u32)n64;
    ctx->stats.decode_sync_iters = 0;
    if (dt->single_leaf) {
        // empty code word: every symbol is the single leaf, no bit is read (main.cpp:271)
        LAUNCH(ctx, fill_kernel, 148 * 4, 256, 0, d_out, n, dt->single_value);
        CU(ctx, cudaGetLastError());
        return BZAP_OK;
    }
    if (payload_len == 0) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "empty payload");
    const u64 payload_bits = (u64)payload_len * 8;
    if (payload_bits < n64) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "payload shorter than N code words");
    const u64 nsub64 = (payload_bits + DEC_SUBSEQ_BITS - 1) / DEC_SUBSEQ_BITS;
    if (nsub64 > 0x7fffffffull / DEC_SUBSEQ_WORDS) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "payload of %zu bytes", payload_len);
    const u32 nsub = (u32)nsub64;
    const u32 total_words = (u32)((payload_len + 3) / 4);
    const u32 blocks = (nsub + DEC_BLOCK - 1) / DEC_BLOCK;
    const u32 sc_tiles = (nsub + SC_BLOCK * SC_ITEMS - 1) / (SC_BLOCK * SC_ITEMS);

    u16 *d_tables = arena_get<u16>(ctx, dt->n_entries);
    u32 *d_start[2] = {arena_get<u32>(ctx, nsub), arena_get<u32>(ctx, nsub)};
    u8 *d_dirty[2] = {arena_get<u8>(ctx, nsub), arena_get<u8>(ctx, nsub)};
    u32 *d_count = arena_get<u32>(ctx, nsub);
    u64 *d_prefix = arena_get<u64>(ctx, (size_t)nsub + 1);
    u64 *d_status = arena_get<u64>(ctx, (size_t)sc_tiles + 2);
    u32 *d_changed = arena_get<u32>(ctx, 4);
    if (!d_tables || !d_start[0] || !d_start[1] || !d_dirty[0] || !d_dirty[1] || !d_count || !d_prefix || !d_status ||
        !d_changed)
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "decode scratch");
    CU(ctx, cudaMemcpyAsync(d_tables, dt->entries, dt->n_entries * sizeof(u16), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_start[0], 0, nsub * sizeof(u32), ctx->stream));
    CU(ctx, cudaMemsetAsync(d_start[1], 0, nsub * sizeof(u32), ctx->stream));
    CU(ctx, cudaMemsetAsync(d_dirty[0], 1, nsub, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_dirty[1], 0, nsub, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_count, 0, nsub * sizeof(u32), ctx->stream));
    const size_t smem = sizeof(DecSmem);
    static_assert(sizeof(DecSmem) <= 48 * 1024, "decoder shared memory must fit the default dynamic limit");

    u32 *h_changed = (u32 *)(ctx->mailbox + 1040);
    int cur = 0;
    u32 iters = 0;
    // The first sweeps are issued back to back (typical streams settle in two; the third only
    // confirms, and touches no block without a dirty subsequence); the host looks at the flag of
    // the last one and keeps sweeping, one round trip each, only if something still moved.
    const u32 burst = nsub > 1 ? 3 : 1;
    while (true) {
        const u32 todo = iters == 0 ? burst : 1;
        for (u32 q = 0; q < todo; ++q) {
            CU(ctx, cudaMemsetAsync(d_changed, 0, sizeof(u32), ctx->stream));
            LAUNCH(ctx, huff_dec_sync_kernel, blocks, DEC_BLOCK, smem, (const u32 *)d_payload, total_words, payload_bits,
                   d_tables, nsub, d_start[cur], d_dirty[cur], d_start[cur ^ 1], d_dirty[cur ^ 1], d_count, d_changed);
            cur ^= 1;
            ++iters;
        }
        CU(ctx, cudaMemcpyAsync(h_changed, d_changed, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (!*h_changed) break;
        if (iters > nsub + 4) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "decoder did not reach a fixed point");
    }
    ctx->stats.decode_sync_iters = iters;
    CU(ctx, cudaMemsetAsync(d_status, 0, ((size_t)sc_tiles + 2) * sizeof(u64), ctx->stream));
    LAUNCH(ctx, scan_u32_u64_kernel, sc_tiles, SC_BLOCK, 0, d_count, nsub, d_prefix, d_status, (u32 *)(d_status + sc_tiles));
    u64 *h_total = (u64 *)(ctx->mailbox + 1048);
    CU(ctx, cudaMemcpyAsync(h_total, d_prefix + nsub, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    LAUNCH(ctx, huff_dec_write_kernel, blocks, DEC_BLOCK, smem, (const u32 *)d_payload, total_words, payload_bits, d_tables,
           nsub, d_start[cur], d_prefix, n, d_out);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    if (*h_total < n64) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "payload holds %llu code words, header says %zu",
                                          (unsigned long long)*h_total, n64);
    return BZAP_OK;
}
