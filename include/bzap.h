/*
 * bzap.h -- C ABI of the B200-native BWT + MTF + Huffman codec (libbzap.so).
 *
 * Drop-in boundary for the hot path of komour/bwt-mtf-huffman-compressor.  The reference has
 * no plugin/FFI layer: its boundary is the pair of free functions compress()/decompress()
 * (main.cpp:300-303, 327-330) plus the on-disk format (io_utilities.h:7-55).  Every entry point
 * below names the reference function (file:line under /root/reference) it replaces.  Files
 * written here are byte-identical to the reference's one-shot COMPRESS binary and decode with
 * its DECOMPRESS binary, and vice versa.
 *
 * All compute runs in hand-written sm_100a CUDA kernels; there is no CPU fallback.  Every call
 * fails with BZAP_ERR_CUDA when no CUDA device is usable.
 *
 * Conventions: plain pointers and sizes; return 0 (BZAP_OK) or a negative error; a NULL context
 * selects a lazily created process-wide default context on the current CUDA device.  A context
 * owns its stream and scratch memory and must not be used from two threads at once.
 */
#ifndef BZAP_H
#define BZAP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BZAP_OK               0
#define BZAP_ERR_IO          -1  /* cannot read / write a file                                  */
#define BZAP_ERR_EMPTY       -2  /* empty input: the reference segfaults (main.cpp:245-246)     */
#define BZAP_ERR_CORRUPT     -3  /* malformed header / tree / payload                           */
#define BZAP_ERR_CUDA        -4  /* no device, kernel or runtime failure                        */
#define BZAP_ERR_CAPACITY    -5  /* output buffer too small                                     */
#define BZAP_ERR_ARG         -6  /* NULL pointer / bad argument                                 */
#define BZAP_ERR_TOO_LARGE   -7  /* block larger than BZAP_MAX_BLOCK, or code word > 64 bits    */
#define BZAP_ERR_NOMEM       -8
#define BZAP_ERR_NCCL        -9  /* NCCL missing, or a collective failed                        */

#define BZAP_MAX_BLOCK   ((size_t)1 << 30)   /* one BWT block = whole input (README.md:40)      */
#define BZAP_HEADER_BYTES 24                 /* io_utilities.h:17-19                            */
#define BZAP_MAX_TREE_BYTES 320              /* ceil((10*256-1)/8), main.cpp:174-196            */

typedef struct bzap_ctx bzap_ctx;

/* Huffman tree in array form (replaces the BTree pointer graph, main.cpp:13-26).
 * Built trees number nodes by creation index: leaves 0..n_leaves-1 in order of first appearance
 * in the MTF stream (main.cpp:238-244), internal nodes after them in merge order (main.cpp:252).
 * Parsed trees (bzap_bytes_to_tree) number nodes in pre-order. */
typedef struct {
    int32_t n_leaves;
    int32_t n_nodes;
    int32_t root;
    int32_t left[511];   /* -1 for a leaf */
    int32_t right[511];
    uint8_t value[511];
} bzap_tree;

/* per-call device timings of the last pipeline call on a context (CUDA events), and the number
 * of kernels this library launched since the context was created */
typedef struct {
    double ms_total;
    double ms_bwt, ms_mtf, ms_huffman;      /* compress: bwt, mtf, hist+encode ; decompress: ibwt, imtf, decode */
    uint64_t kernel_launches;
    uint32_t bwt_rounds;                    /* prefix-doubling rounds of the last forward BWT   */
    uint32_t bwt_sort_passes;               /* onesweep passes executed in the last forward BWT */
    uint32_t decode_sync_iters;             /* self-synchronisation iterations of the last decode */
    uint32_t bwt_full_passes;               /* ... of which over all N rotations (timed: ms_sort / sort_bytes) */
    uint64_t payload_bytes;
    double ms_sort;                         /* CUDA-event time inside onesweep passes of the last forward BWT */
    uint64_t sort_bytes;                    /* algorithmic bytes those passes moved (key+payload read+write) */
    uint64_t sort_elems;                    /* elements per pass (N) */
    double ms_walk;                         /* CUDA-event time of the inverse BWT's list walk in the last decompress */
    uint64_t walk_bytes;                    /* its algorithmic bytes: one 4-byte T entry per row + one output byte */
} bzap_stats;

/* one block sorted over the GPUs of one box: phase timings (CUDA events on this rank's stream) */
typedef struct {
    int32_t world, rank;
    uint32_t rounds;                        /* prefix-doubling rounds, the 8-byte key sort included  */
    uint32_t peer_windows;                  /* 1 = bulk exchanges went through CUDA IPC peer windows, 0 = ncclSend/Recv */
    uint64_t own_rotations;                 /* rotations whose 8-byte key this rank owns             */
    uint64_t exchanged_bytes;               /* bytes this rank sent to peers (all phases)            */
    double ms_total, ms_select_sort, ms_home, ms_rounds, ms_pull, ms_round_sort, ms_tail;
} bzap_dist_stats;

/* ---- context ------------------------------------------------------------------------------ */
int  bzap_ctx_create(int device, bzap_ctx **out);
void bzap_ctx_destroy(bzap_ctx *ctx);
/* run subsequent calls on the caller's CUDA stream (cudaStream_t; NULL = the CUDA default stream);
 * (void*)-1 restores the context's own stream */
int  bzap_ctx_set_stream(bzap_ctx *ctx, void *cuda_stream);
int  bzap_get_stats(bzap_ctx *ctx, bzap_stats *out);
const char *bzap_strerror(int code);
const char *bzap_last_error(bzap_ctx *ctx);
const char *bzap_version(void);

/* ---- file level: the reference entry points ------------------------------------------------ */
/* replaces compress(in, out)   main.cpp:300-325  (read_bytes io_utilities.h:29-55, write_bytes :7-27) */
int bzap_compress_file(bzap_ctx *ctx, const char *in_path, const char *out_path);
/* replaces decompress(in, out) main.cpp:327-345 */
int bzap_decompress_file(bzap_ctx *ctx, const char *in_path, const char *out_path);

/* ---- buffer level (host memory; pinned memory makes the copies faster) ---------------------- */
/* 24 + 320 + max(1, n): an optimal prefix code never exceeds the fixed 8-bit code             */
size_t   bzap_compress_bound(size_t n);
int      bzap_compress(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);
/* header field 2 (io_utilities.h:46); 0 if the buffer is shorter than a header                */
uint64_t bzap_decompressed_size(const uint8_t *file, size_t len);
int      bzap_decompress(bzap_ctx *ctx, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *out_len);

/* ---- buffer level, device memory (input and output resident in HBM) ------------------------ */
int bzap_compress_device(bzap_ctx *ctx, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, size_t *out_len);
int bzap_decompress_device(bzap_ctx *ctx, const uint8_t *d_in, size_t in_len, uint8_t *d_out, size_t out_cap, size_t *out_len);

/* ---- batch: independent files, one BWT block each (the reference's 14-file loop, main.cpp:424-437) ----
 * Files are handed out largest first to a persistent pool of workers (one context = one stream + scratch
 * per worker), n_streams per device (0 = default 4), so that on every device the host-to-device copy of one
 * file and the read-back of another overlap the kernels of a third (pinned host buffers make the copies
 * asynchronous).  outs[i] needs bzap_compress_bound(ns[i]) bytes.  Returns the first error, if any.
 *   ..._batch       the current CUDA device
 *   ..._batch_gpus  devices 0 .. n_gpus-1 of this process, greedy largest-first over all their workers
 *   ..._files       the same for paths: in_paths[i] -> out_paths[i] (read_bytes / write_bytes,
 *                   io_utilities.h:7-55); n_gpus = 0 means the current device                          */
int bzap_compress_batch(const uint8_t *const *ins, const size_t *ns, uint8_t *const *outs, size_t *out_lens,
                        int count, int n_streams);
/* outs[i] holds out_caps[i] bytes (size it with bzap_decompressed_size); a header that asks for more
 * fails with BZAP_ERR_CAPACITY instead of overrunning the buffer                                       */
int bzap_decompress_batch(const uint8_t *const *ins, const size_t *in_lens, uint8_t *const *outs, const size_t *out_caps,
                          size_t *out_lens, int count, int n_streams);
int bzap_compress_batch_gpus(const uint8_t *const *ins, const size_t *ns, uint8_t *const *outs, size_t *out_lens,
                             int count, int n_gpus, int n_streams);
int bzap_decompress_batch_gpus(const uint8_t *const *ins, const size_t *in_lens, uint8_t *const *outs, const size_t *out_caps,
                               size_t *out_lens, int count, int n_gpus, int n_streams);
int bzap_compress_files(const char *const *in_paths, const char *const *out_paths, int count, int n_gpus, int n_streams);
int bzap_decompress_files(const char *const *in_paths, const char *const *out_paths, int count, int n_gpus, int n_streams);

/* ---- stage level (host pointers), one per SURVEY 8a row ------------------------------------- */
/* bwt()                   main.cpp:77-91 with bwt_cmp_straight :46-59                         */
int bzap_bwt(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *last_col, uint64_t *primary);
/* bwt_reverse()           main.cpp:61-75                                                     */
int bzap_ibwt(bzap_ctx *ctx, const uint8_t *last_col, size_t n, uint64_t primary, uint8_t *out);
/* move_to_front()         main.cpp:93-112                                                    */
int bzap_mtf(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out);
/* move_to_front_reverse() main.cpp:114-130                                                   */
int bzap_imtf(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out);
/* histogram + first-appearance order, main.cpp:235-244.  order[k] = symbol of leaf k          */
int bzap_hist256(bzap_ctx *ctx, const uint8_t *in, size_t n, uint64_t freq[256], uint8_t order[256], int *n_leaves);
/* merge loop, main.cpp:245-254, with the pointer-order tie-break of the one-shot reference
 * binary.  Host only.  With BZAP_HEAP_REPLAY=1 in the environment, inputs below
 * 64,600 bytes take the order from the allocator-replay helper (csrc/heap_replay.c)          
 * instead of the closed-form law: byte identity in the windows where the law does not hold.   */
int bzap_huff_build(const uint64_t freq[256], const uint8_t *order, int n_leaves, bzap_tree *tree);
/* traverse()/build_hashmap() main.cpp:132-156; code right-aligned in 64 bits, MSB-first       */
int bzap_huff_codes(const bzap_tree *tree, uint64_t code[256], uint8_t len[256]);
/* tree_to_bytes()         main.cpp:174-196; out needs BZAP_MAX_TREE_BYTES                     */
int bzap_tree_to_bytes(const bzap_tree *tree, uint8_t *out, size_t *len);
/* bytes_to_tree()         main.cpp:198-227                                                   */
int bzap_bytes_to_tree(const uint8_t *bytes, size_t len, bzap_tree *tree);
/* encode_with_huffman()   main.cpp:158-172; out_len = max(1, ceil(bits/8))                    */
int bzap_huff_encode(bzap_ctx *ctx, const uint8_t *in, size_t n, const bzap_tree *tree,
                     uint8_t *out, size_t out_cap, size_t *out_len);
/* huffman_reverse()       main.cpp:259-281; decodes exactly n symbols                         */
int bzap_huff_decode(bzap_ctx *ctx, const uint8_t *payload, size_t payload_len, const bzap_tree *tree,
                     size_t n, uint8_t *out);

/* ---- ONE block sorted over several GPUs (BASELINE config 5 ii) -----------------------
 * Replaces bwt() + move_to_front() + huffman() + write_bytes() (main.cpp:304-324) for a block that
 * is spread over the GPUs of one box: one process (or thread) per GPU, each with its own context;
 * the contexts share an NCCL communicator that the library creates and owns.
 *   rank 0:      bzap_comm_unique_id(id)  -> hand the 128 bytes to the other ranks (any transport)
 *   every rank:  bzap_ctx_comm_init(ctx, id, world, rank)          (collective)
 *   every rank:  bzap_compress_block_distributed(ctx, d_text, n, d_out, cap, &len)   (collective)
 * d_text: the whole block, resident on every GPU.  The file (byte-identical to bzap_compress) lands
 * in rank 0's d_out; the other ranks pass d_out = NULL and get *out_len = 0.
 * Distributed prefix doubling: rank g owns the rotations whose first 8 bytes fall into its key range
 * (splitters from a hashed sample), sorts them locally and keeps their suffix-array slots; per round
 * only unsettled rotations pull rank[(i+k) mod N] from the rank that owns text position i+k
 * (request / response all-to-all over NVLink) and send their new ranks home; MTF start lists, the
 * Huffman statistics and the bit offsets of the payload pieces are exchanged by all-gather.
 * A context without a communicator is a world of one (same code, no NCCL needed).
 * All three calls, and bzap_ctx_destroy of a context that holds a communicator, are collective: every rank
 * makes them in the same order.  An error on one rank (out of memory, a failed collective) leaves the others
 * waiting in their next exchange: treat it as fatal for the communicator.  The call works in an arena of its
 * own that the peers map through CUDA IPC; it grows only after every rank has closed its mappings.        */
#define BZAP_COMM_ID_BYTES 128
int bzap_comm_unique_id(uint8_t id[BZAP_COMM_ID_BYTES]);
int bzap_ctx_comm_init(bzap_ctx *ctx, const uint8_t id[BZAP_COMM_ID_BYTES], int world, int rank);
int bzap_compress_block_distributed(bzap_ctx *ctx, const uint8_t *d_text, size_t n, uint8_t *d_out, size_t out_cap,
                                    size_t *out_len);
int bzap_get_dist_stats(bzap_ctx *ctx, bzap_dist_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* BZAP_H */
