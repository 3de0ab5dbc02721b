// bzap_internal.h -- context, scratch arena and the device-level stage entry points shared by
// the translation units of libbzap.so.  Not part of the public ABI (include/bzap.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include "../../include/bzap.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

struct bzap_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream all work is issued on
    cudaStream_t own_stream = nullptr;
    // scratch arena in HBM: one allocation, bump-allocated per top-level call
    u8 *arena = nullptr;
    size_t arena_cap = 0;
    size_t arena_off = 0;
    // small pinned host mailbox for flags / histograms coming back from the device
    u8 *mailbox = nullptr;               // 64 KiB pinned
    cudaEvent_t ev[8] = {};
    // event pairs bracketing each run of back-to-back onesweep passes (roofline evidence)
    cudaEvent_t sort_ev[128] = {};
    int sort_ev_used = 0;
    bzap_stats stats = {};
    u64 launches = 0;
    u32 attr_mask = 0;                  // per-device kernel attributes this context has already raised
    // one block spread over several GPUs (dist_block.cu): NCCL communicator of the ranks that share it
    void *comm = nullptr;               // ncclComm_t
    int world = 1, rank = 0;
    u8 *dist_host = nullptr;            // pinned staging for the small collectives
    u8 *dist_dev = nullptr;             // 1 MiB of device memory that is never exported (sample keys, counts, barrier words)
    u8 *dist_arena = nullptr;           // the arena of the distributed path: peers map it, so it is never the one the
    size_t dist_arena_cap = 0;          // other entry points reallocate, and it only grows under the protocol in dist_block.cu
    bool dist_had_peers = false;
    void *peers = nullptr;              // the other ranks' arenas mapped through CUDA IPC (dist_block.cu: PeerMap)
    bzap_dist_stats dstats = {};
    char err[256] = {0};
};

#define BZAP_MAILBOX_BYTES (64 * 1024)

// ---- error plumbing ------------------------------------------------------------------------
int bzap_fail(bzap_ctx *ctx, int code, const char *fmt, ...);
#define CU(ctx, call)                                                                             \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return bzap_fail(ctx, BZAP_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                             cudaGetErrorString(e_));                                             \
    } while (0)
#define RET(call)                                                                                 \
    do {                                                                                          \
        int r_ = (call);                                                                          \
        if (r_ != BZAP_OK) return r_;                                                             \
    } while (0)
// every kernel launch goes through this so that stats.kernel_launches is a true count
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                               \
    do {                                                                                          \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                          \
        ++(ctx)->launches;                                                                        \
    } while (0)

// ---- arena ------------------------------------------------------------------------------------
int arena_reserve(bzap_ctx *ctx, size_t bytes);     // grow (sync + realloc) if needed, reset offset
void *arena_alloc(bzap_ctx *ctx, size_t bytes);     // 256-byte aligned bump allocation, nullptr if full
template <typename T> static inline T *arena_get(bzap_ctx *ctx, size_t count)
{
    return (T *)arena_alloc(ctx, count * sizeof(T));
}
size_t scratch_bytes_compress(size_t n);
size_t scratch_bytes_decompress(size_t n, size_t payload);

// ---- host-side Huffman model (huffman_host.cpp) -------------------------------------------
struct CodeTable {
    u64 code[256];
    u8 len[256];
    int max_len;
};
// decode tables: primary LUT of 1<<DEC_PRIMARY_BITS u16 entries, then 256-entry sub tables
#define DEC_PRIMARY_BITS 12
#define DEC_SUB_BITS 8
struct DecodeTables {
    u16 *entries;      // malloc'ed, n_entries u16
    size_t n_entries;
    int max_len;       // depth of the tree
    int single_leaf;   // tree is one leaf: empty code (main.cpp:137-140)
    u8 single_value;
};
int huff_build_tree(const u64 freq[256], const u8 *order, int n_leaves, bzap_tree *t);
int huff_code_table(const bzap_tree *t, CodeTable *ct);      // BZAP_ERR_TOO_LARGE if a code > 64 bits
int huff_tree_to_bytes(const bzap_tree *t, u8 *out, size_t *len);
int huff_bytes_to_tree(const u8 *bytes, size_t len, bzap_tree *t);
int huff_decode_tables(const bzap_tree *t, DecodeTables *dt);
void huff_free_decode_tables(DecodeTables *dt);
u64 huff_total_bits(const u64 freq[256], const CodeTable *ct);

// ---- device-level stages (all pointers are device pointers, work on ctx->stream) ------------
// forward BWT (bwt.cu): writes last column, returns primary index on the host (syncs)
int dev_bwt(bzap_ctx *ctx, const u8 *d_in, size_t n, u8 *d_last, u64 *primary);
// forward / inverse MTF (mtf.cu)
int dev_mtf(bzap_ctx *ctx, const u8 *d_in, size_t n, u8 *d_out);
int dev_imtf(bzap_ctx *ctx, const u8 *d_in, size_t n, u8 *d_out);
// the forward transform in two phases, for a piece of a block that starts from a handed-over list (dist_block.cu)
struct MtfPlan {
    u32 n, chunk, nchunks, ngroups;
    u32 *d_last, *d_gtot, *d_total, *d_seg;
    u8 *d_lists;
    int two_level;
};
size_t mtf_scratch_bytes(size_t n);
int dev_mtf_begin(bzap_ctx *ctx, const u8 *d_in, size_t n, MtfPlan *plan);
int dev_mtf_finish(bzap_ctx *ctx, const u8 *d_in, const MtfPlan *plan, const u32 *d_init, u8 *d_out);
// histogram + first appearance order (huffman_enc.cu), syncs
int dev_hist(bzap_ctx *ctx, const u8 *d_in, size_t n, u64 freq[256], u8 order[256], int *n_leaves);
int dev_hist_launch(bzap_ctx *ctx, const u8 *d_in, u32 n, u64 *d_stat);   // no sync: 256 x u64 counts + 256 x u32 first positions
// bit packer (huffman_enc.cu): ORs the code stream into d_file starting at bit `bit_base`
// (d_file zeroed by the caller, 16-byte aligned, with >= 32 bytes of slack after the last bit)
int dev_huff_encode(bzap_ctx *ctx, const u8 *d_in, size_t n, const CodeTable *ct, u8 *d_file, u64 bit_base);
// decoder (huffman_dec.cu): d_payload must be 4-byte aligned with >= 64 zero bytes of slack
int dev_huff_decode(bzap_ctx *ctx, const u8 *d_payload, size_t payload_len, const DecodeTables *dt, size_t n,
                    u8 *d_out);
// inverse BWT (ibwt.cu)
int dev_ibwt(bzap_ctx *ctx, const u8 *d_last, size_t n, u64 primary, u8 *d_out);
void dist_comm_release(bzap_ctx *ctx);   // dist_block.cu: destroys the context's communicator, if any

// ---- radix sort building blocks (radix_sort.cu) ----------------------------------------------
// LSD onesweep over 64-bit keys with 32-bit payloads.  d_hist holds 8 x 256 digit counts already
// accumulated by the caller's key-producing kernel; pass_mask bit p = digit p can vary at all (used
// instead of a flag read-back for small inputs, where a host round trip costs more than a pass).
// vals_in == nullptr means "payload = element index" (materialised by the first executed pass).
// On return *out_keys / *out_vals point at the buffers holding the result (vals may be nullptr
// when no pass executed and the payload is still the identity).
struct SortBuffers {
    u64 *keys[2];
    u32 *vals[2];
};
// gen != nullptr (only with vals_are_iota): the keys are not in b->keys[0] yet; the first executed
// pass builds them from gen->src (mode 1: 8-byte text windows, mode 2: rank pairs at distance k).
// If no pass executes, *out_keys is nullptr and the caller materialises the keys itself.
struct SortKeyGen {
    int mode;
    const void *src;
    u32 k;
};
// state of the active-set rounds of the forward BWT (bwt.cu), shared with the distributed path
struct ActiveWork {
    SortBuffers ab;                       // key / payload ping-pong buffers for up to m elements
    u32 *act_r1, *newr, *pos;             // per active element: group rank, new rank, suffix-array slot
    u32 *next_idx, *next_r1;              // survivors of a round
    u32 *sa_buf, *d_rank;                 // the block's suffix array and text-order ranks (updated in place)
    u32 *d_hist8, *d_rrctl;
    size_t rrctl_bytes;
    u32 *d_counters, *d_ticket;
    u64 *d_status, *cstatus;
    size_t arena_mark;
    u32 rank_mask;
    u8 *zero_base;                        // hist8, rrctl and cstatus are allocated back to back:
    size_t zero_bytes;                    // one memset per round clears all three
};
int bwt_active_sort_rerank(bzap_ctx *ctx, const ActiveWork &w, SortBuffers *ab, u32 m, u32 rshift, u32 pass_mask, u32 slot_base,
                           u32 *d_rank, u32 *next_idx, u32 *next_r1, u64 **skeys, u32 **sidx, int *passes);
int dev_sort_pairs64(bzap_ctx *ctx, SortBuffers *b, u32 n, u32 pass_mask, const u32 *d_hist, int hist_rows, bool vals_are_iota,
                     u64 **out_keys, u32 **out_vals, int *passes_run, const SortKeyGen *gen = nullptr);
// stable counting sort of positions by byte value: T[r] = position of the r-th smallest (byte, pos)
// (main.cpp:67); d_cum receives the 257 exclusive byte counts
int dev_sort_positions_by_byte(bzap_ctx *ctx, const u8 *d_bytes, u32 n, u32 *d_T, u32 *d_cum);
size_t sort_scratch_bytes(u32 n);
// building blocks shared with the distributed single-block path (dist_block.cu)
size_t rerank_ctl_words(u32 m);
int dev_rerank_sorted(bzap_ctx *ctx, const u64 *d_keys, u32 m, u32 pos_base, u32 *d_rs, u32 *d_ctl, u32 *d_bact);
int dev_collect_active(bzap_ctx *ctx, const u32 *d_rs, const u32 *d_sa, u32 m, u32 pos_base, const u32 *d_bact, u32 *act_idx,
                       u32 *act_r1);
size_t active_ctl_bytes(u32 m, ActiveWork *w, u8 *base);   // lays the per-round control block of the active rounds out at base
int dev_gather_slots(bzap_ctx *ctx, const u8 *d_text, const u32 *d_sa, u32 n, u32 m, u8 *d_last);
size_t bucket_ctl_words(u32 m);
int dev_bucket_pass_u64(bzap_ctx *ctx, const u64 *d_keys, const u32 *d_vals, u32 m, int shift, u64 *d_keys_out, u32 *d_vals_out,
                        const u32 *d_hist256, u32 *d_ctl);
int dev_bucket_pass_u32(bzap_ctx *ctx, const u32 *d_keys, const u32 *d_vals, u32 m, int shift, u32 *d_keys_out, u32 *d_vals_out,
                        u32 *d_hist256, u32 *d_ctl);
int dev_scatter_offset_async(bzap_ctx *ctx, const u32 *d_idx, const u32 *d_vals, u32 m, u32 off, u32 *d_out);
// 256-bin byte histogram accumulated into d_hist256 (caller zeroes it)
int dev_byte_hist(bzap_ctx *ctx, const u8 *d_bytes, u32 n, u32 *d_hist256);
// out[perm[j]] = vals[j] for a permutation perm of 0..n-1; tmp buffers hold n u32 each
int dev_scatter_perm(bzap_ctx *ctx, const u32 *d_perm, const u32 *d_vals, u32 n, u32 *d_out, u32 *d_tmp_idx,
                     u32 *d_tmp_vals);
