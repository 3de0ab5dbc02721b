// api.cu -- the C ABI (include/bzap.h): context, scratch arena, the compress / decompress
// pipelines and the container format.
//
// Reference boundary: compress() main.cpp:300-325, decompress() main.cpp:327-345, container
// write_bytes()/read_bytes() io_utilities.h:7-55:
//     u64 primary | u64 N | u64 tree_bytes | tree | payload          (native little endian)
// All compute is CUDA; there is no CPU path.  Host work is limited to the <=511-node Huffman
// model (huffman_host.cpp) and the 24-byte header.
#include "bzap_internal.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// ---- errors -------------------------------------------------------------------------------------------
int bzap_fail(bzap_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

extern "C" const char *bzap_strerror(int code)
{
    switch (code) {
    case BZAP_OK: return "ok";
    case BZAP_ERR_IO: return "file I/O error";
    case BZAP_ERR_EMPTY: return "empty input";
    case BZAP_ERR_CORRUPT: return "corrupt stream";
    case BZAP_ERR_CUDA: return "CUDA error (no device or kernel failure)";
    case BZAP_ERR_CAPACITY: return "output buffer too small";
    case BZAP_ERR_ARG: return "bad argument";
    case BZAP_ERR_TOO_LARGE: return "block too large";
    case BZAP_ERR_NOMEM: return "out of memory";
    case BZAP_ERR_NCCL: return "NCCL error (library missing or a collective failed)";
    default: return "unknown error";
    }
}
extern "C" const char *bzap_version(void) { return "bzap 0.1 (sm_100a)"; }

// ---- context ------------------------------------------------------------------------------------------
extern "C" int bzap_ctx_create(int device, bzap_ctx **out)
{
    if (!out) return BZAP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return BZAP_ERR_CUDA;   // fail loudly: no CPU path
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return BZAP_ERR_CUDA; }
    if (device >= count) return BZAP_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BZAP_ERR_CUDA;
    bzap_ctx *c = new (std::nothrow) bzap_ctx();
    if (!c) return BZAP_ERR_NOMEM;
    c->device = device;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&c->mailbox, BZAP_MAILBOX_BYTES) == cudaSuccess;
    for (int i = 0; ok && i < 8; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    if (!ok) { bzap_ctx_destroy(c); return BZAP_ERR_CUDA; }
    c->stream = c->own_stream;
    *out = c;
    return BZAP_OK;
}

extern "C" void bzap_ctx_destroy(bzap_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own_stream) { cudaStreamSynchronize(c->own_stream); }
    dist_comm_release(c);
    if (c->arena) cudaFree(c->arena);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 128; ++i) if (c->sort_ev[i]) cudaEventDestroy(c->sort_ev[i]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

static std::mutex g_default_mu;
static bzap_ctx *g_default = nullptr;
static bzap_ctx *resolve(bzap_ctx *ctx, int *rc)
{
    *rc = BZAP_OK;
    if (ctx) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { *rc = BZAP_ERR_CUDA; return nullptr; }
        return ctx;
    }
    std::lock_guard<std::mutex> g(g_default_mu);
    if (!g_default) *rc = bzap_ctx_create(-1, &g_default);
    else if (cudaSetDevice(g_default->device) != cudaSuccess) *rc = BZAP_ERR_CUDA;
    return *rc == BZAP_OK ? g_default : nullptr;
}
#define RESOLVE(ctx)                                                                              \
    int rc_resolve_;                                                                              \
    ctx = resolve(ctx, &rc_resolve_);                                                             \
    if (!ctx) return rc_resolve_;                                                                 \
    ctx->err[0] = 0

extern "C" int bzap_ctx_set_stream(bzap_ctx *ctx, void *stream)
{
    RESOLVE(ctx);
    // NULL is a real stream (CUDA's default stream, which is what torch.cuda.current_stream() is
    // unless the caller changed it): work must be ordered with the caller's, so it is honoured
    ctx->stream = stream == (void *)-1 ? ctx->own_stream : (cudaStream_t)stream;
    return BZAP_OK;
}
extern "C" int bzap_get_stats(bzap_ctx *ctx, bzap_stats *out)
{
    if (!out) return BZAP_ERR_ARG;
    RESOLVE(ctx);
    *out = ctx->stats;
    out->kernel_launches = ctx->launches;
    return BZAP_OK;
}
extern "C" const char *bzap_last_error(bzap_ctx *ctx)
{
    if (!ctx) ctx = g_default;
    return ctx ? ctx->err : "";
}

// ---- arena --------------------------------------------------------------------------------------------
int arena_reserve(bzap_ctx *ctx, size_t bytes)
{
    ctx->arena_off = 0;
    if (bytes <= ctx->arena_cap) return BZAP_OK;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) { cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_cap = 0; }
    size_t want = bytes + bytes / 16;
    cudaError_t e = cudaMalloc((void **)&ctx->arena, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    }
    ctx->arena_cap = want;
    return BZAP_OK;
}
void *arena_alloc(bzap_ctx *ctx, size_t bytes)
{
    size_t off = (ctx->arena_off + 255) & ~(size_t)255;
    if (off + bytes > ctx->arena_cap) return nullptr;
    ctx->arena_off = off + bytes;
    return ctx->arena + off;
}
size_t scratch_bytes_compress(size_t n)
{
    // text, last column, mtf, file image | 2 x u64 keys, 2 x u32 payload, rank | sort + mtf tables
    return 4 * (n + 1024) + 33 * n + sort_scratch_bytes((u32)n) + 2 * mtf_scratch_bytes(n) + (8u << 20);
}
size_t scratch_bytes_decompress(size_t n, size_t payload)
{
    // payload copy, mtf, last column, out | T | sort status | mtf tables | decode state
    return (payload + 4096) + 3 * (n + 1024) + 5 * n + sort_scratch_bytes((u32)n) + 2 * mtf_scratch_bytes(n) +
           (payload / 128 + 4096) * 32 + (8u << 20) +
           (n <= (2u << 20) ? 36 * n : 0);         // small blocks: 8-row splitter buckets in ibwt.cu (IB_SMALL_N)
}

// ---- header --------------------------------------------------------------------------------------------
static void put_u64(u8 *p, u64 v) { for (int i = 0; i < 8; ++i) p[i] = (u8)(v >> (8 * i)); }
static u64 get_u64(const u8 *p) { u64 v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | p[i]; return v; }

extern "C" size_t bzap_compress_bound(size_t n) { return BZAP_HEADER_BYTES + BZAP_MAX_TREE_BYTES + (n ? n : 1); }
extern "C" uint64_t bzap_decompressed_size(const uint8_t *file, size_t len)
{
    return (!file || len < BZAP_HEADER_BYTES) ? 0 : get_u64(file + 8);
}

static double ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// ---- compress pipeline: d_in (device) -> file image in the arena -------------------------------------
// main.cpp:304-324: bwt -> move_to_front -> huffman -> tree_to_bytes -> write_bytes
static int pipeline_after_bwt(bzap_ctx *ctx, const u8 *d_last, size_t n, u64 primary, u8 **d_file_out, size_t *file_len);

static int pipeline_compress(bzap_ctx *ctx, const u8 *d_in, size_t n, u8 **d_file_out, size_t *file_len)
{
    u8 *d_last = arena_get<u8>(ctx, n + 64);
    if (!d_last) return bzap_fail(ctx, BZAP_ERR_NOMEM, "pipeline scratch");
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    u64 primary = 0;
    RET(dev_bwt(ctx, d_in, n, d_last, &primary));
    return pipeline_after_bwt(ctx, d_last, n, primary, d_file_out, file_len);
}

// move_to_front -> huffman -> tree_to_bytes -> container (main.cpp:309-324), given the last column
static int pipeline_after_bwt(bzap_ctx *ctx, const u8 *d_last, size_t n, u64 primary, u8 **d_file_out, size_t *file_len)
{
    u8 *d_mtf = arena_get<u8>(ctx, n + 64);
    u8 *d_file = arena_get<u8>(ctx, bzap_compress_bound(n) + 64);
    if (!d_mtf || !d_file) return bzap_fail(ctx, BZAP_ERR_NOMEM, "pipeline scratch");
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    RET(dev_mtf(ctx, d_last, n, d_mtf));
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    u64 freq[256];
    u8 order[256];
    int n_leaves = 0;
    RET(dev_hist(ctx, d_mtf, n, freq, order, &n_leaves));
    bzap_tree tree;
    CodeTable ct;
    u8 *h_head = ctx->mailbox + 16384;                    // header + tree, pinned
    size_t tb = 0;
    int rc = huff_build_tree(freq, order, n_leaves, &tree);
    if (rc == BZAP_OK) rc = huff_code_table(&tree, &ct);
    if (rc == BZAP_OK) rc = huff_tree_to_bytes(&tree, h_head + BZAP_HEADER_BYTES, &tb);
    if (rc != BZAP_OK) return bzap_fail(ctx, rc, "huffman model");
    const u64 bits = huff_total_bits(freq, &ct);
    const size_t payload = bits ? (size_t)((bits + 7) / 8) : 1;   // max(1, ceil(bits/8)), main.cpp:162
    const size_t head = BZAP_HEADER_BYTES + tb;
    put_u64(h_head, primary);                                // io_utilities.h:17
    put_u64(h_head + 8, n);                                  // io_utilities.h:18
    put_u64(h_head + 16, tb);                                // io_utilities.h:19
    CU(ctx, cudaMemsetAsync(d_file, 0, ((head + payload + 63) & ~(size_t)31), ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_file, h_head, head, cudaMemcpyHostToDevice, ctx->stream));
    RET(dev_huff_encode(ctx, d_mtf, n, &ct, d_file, (u64)head * 8));
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    *d_file_out = d_file;
    *file_len = head + payload;
    ctx->stats.payload_bytes = payload;
    return BZAP_OK;
}

static void finish_stats(bzap_ctx *ctx)
{
    double ms_sort = 0;
    for (int i = 0; i + 1 < ctx->sort_ev_used; i += 2) ms_sort += ev_ms(ctx->sort_ev[i], ctx->sort_ev[i + 1]);
    ctx->stats.ms_sort = ms_sort;
    ctx->stats.ms_bwt = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_mtf = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_huffman = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
}

static int compress_any(bzap_ctx *ctx, const u8 *in, bool in_on_device, size_t n, u8 *out, bool out_on_device,
                        size_t out_cap, size_t *out_len)
{
    if (!in || !out || !out_len) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input (the reference crashes here, main.cpp:245)");
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    RET(arena_reserve(ctx, scratch_bytes_compress(n)));
    const u8 *d_in = in;
    if (!in_on_device) {
        u8 *d = arena_get<u8>(ctx, n + 64);
        if (!d) return bzap_fail(ctx, BZAP_ERR_NOMEM, "input staging");
        CU(ctx, cudaMemcpyAsync(d, in, n, cudaMemcpyHostToDevice, ctx->stream));
        d_in = d;
    } else if (((uintptr_t)in & 15) != 0) {
        u8 *d = arena_get<u8>(ctx, n + 64);                  // kernels use 16-byte vector loads
        if (!d) return bzap_fail(ctx, BZAP_ERR_NOMEM, "input staging");
        CU(ctx, cudaMemcpyAsync(d, in, n, cudaMemcpyDeviceToDevice, ctx->stream));
        d_in = d;
    }
    u8 *d_file = nullptr;
    size_t len = 0;
    RET(pipeline_compress(ctx, d_in, n, &d_file, &len));
    if (len > out_cap) return bzap_fail(ctx, BZAP_ERR_CAPACITY, "need %zu bytes, have %zu", len, out_cap);
    CU(ctx, cudaMemcpyAsync(out, d_file, len, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                            ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    finish_stats(ctx);
    *out_len = len;
    return BZAP_OK;
}

// ---- decompress pipeline ---------------------------------------------------------------------------------
// main.cpp:331-344: read_bytes(meta) -> bytes_to_tree -> huffman_reverse -> mtf_reverse -> bwt_reverse
static int decompress_any(bzap_ctx *ctx, const u8 *in, bool in_on_device, size_t in_len, u8 *out, bool out_on_device,
                          size_t out_cap, size_t *out_len)
{
    if (!in || !out_len) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (in_len < BZAP_HEADER_BYTES) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "file shorter than its header");
    u8 head[BZAP_HEADER_BYTES + BZAP_MAX_TREE_BYTES];
    const size_t head_avail = in_len < sizeof head ? in_len : sizeof head;
    if (in_on_device) {
        CU(ctx, cudaMemcpyAsync(ctx->mailbox + 16384, in, head_avail, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(head, ctx->mailbox + 16384, head_avail);
    } else {
        memcpy(head, in, head_avail);
    }
    const u64 primary = get_u64(head), n = get_u64(head + 8), tb = get_u64(head + 16);   // io_utilities.h:45-47
    if (tb > in_len - BZAP_HEADER_BYTES) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "tree_bytes %llu beyond file", (unsigned long long)tb);
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %llu bytes", (unsigned long long)n);
    if (n == 0) { *out_len = 0; return BZAP_OK; }
    if (!out || n > out_cap) return bzap_fail(ctx, BZAP_ERR_CAPACITY, "need %llu bytes, have %zu", (unsigned long long)n, out_cap);
    if (primary >= n) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "primary index %llu >= N", (unsigned long long)primary);
    bzap_tree tree;
    const size_t tree_avail = (size_t)tb < head_avail - BZAP_HEADER_BYTES ? (size_t)tb : head_avail - BZAP_HEADER_BYTES;
    int rc = huff_bytes_to_tree(head + BZAP_HEADER_BYTES, tree_avail, &tree);     // main.cpp:334
    if (rc != BZAP_OK) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "malformed Huffman tree");
    DecodeTables dt;
    rc = huff_decode_tables(&tree, &dt);
    if (rc != BZAP_OK) return bzap_fail(ctx, rc, "decode tables");
    const size_t payload_len = in_len - BZAP_HEADER_BYTES - (size_t)tb;
    const u8 *payload = in + BZAP_HEADER_BYTES + tb;

    rc = arena_reserve(ctx, scratch_bytes_decompress((size_t)n, payload_len));
    u8 *d_payload = nullptr, *d_mtf = nullptr, *d_last = nullptr, *d_out = nullptr;
    if (rc == BZAP_OK) {
        const size_t padded = (payload_len + 128 + 15) & ~(size_t)15;
        d_payload = arena_get<u8>(ctx, padded);
        d_mtf = arena_get<u8>(ctx, n + 64);
        d_last = arena_get<u8>(ctx, n + 64);
        d_out = (out_on_device && ((uintptr_t)out & 15) == 0) ? out : arena_get<u8>(ctx, n + 64);
        if (!d_payload || !d_mtf || !d_last || !d_out) rc = bzap_fail(ctx, BZAP_ERR_NOMEM, "pipeline scratch");
        // aligned, zero-padded copy of the payload (the file offset 24 + tree_bytes is arbitrary)
        if (rc == BZAP_OK && cudaMemsetAsync(d_payload + (payload_len & ~(size_t)15), 0, padded - (payload_len & ~(size_t)15), ctx->stream) != cudaSuccess) rc = BZAP_ERR_CUDA;
        if (rc == BZAP_OK && payload_len &&
            cudaMemcpyAsync(d_payload, payload, payload_len, in_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            rc = BZAP_ERR_CUDA;
    }
    if (rc == BZAP_OK) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        rc = dev_huff_decode(ctx, d_payload, payload_len, &dt, (size_t)n, d_mtf);      // main.cpp:338
        cudaEventRecord(ctx->ev[1], ctx->stream);
    }
    huff_free_decode_tables(&dt);
    if (rc == BZAP_OK) {
        rc = dev_imtf(ctx, d_mtf, (size_t)n, d_last);                                     // main.cpp:340
        cudaEventRecord(ctx->ev[2], ctx->stream);
    }
    if (rc == BZAP_OK) {
        rc = dev_ibwt(ctx, d_last, (size_t)n, primary, d_out);                            // main.cpp:342
        cudaEventRecord(ctx->ev[3], ctx->stream);
    }
    if (rc != BZAP_OK) return rc;
    if (d_out != out)
        CU(ctx, cudaMemcpyAsync(out, d_out, n, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    ctx->stats.ms_huffman = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_mtf = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_bwt = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    ctx->stats.ms_walk = ev_ms(ctx->ev[4], ctx->ev[5]);
    ctx->stats.payload_bytes = payload_len;
    *out_len = (size_t)n;
    return BZAP_OK;
}

extern "C" int bzap_compress(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len)
{
    RESOLVE(ctx);
    return compress_any(ctx, in, false, n, out, false, cap, out_len);
}
extern "C" int bzap_compress_device(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len)
{
    RESOLVE(ctx);
    return compress_any(ctx, in, true, n, out, true, cap, out_len);
}
extern "C" int bzap_decompress(bzap_ctx *ctx, const uint8_t *in, size_t in_len, uint8_t *out, size_t cap, size_t *out_len)
{
    RESOLVE(ctx);
    return decompress_any(ctx, in, false, in_len, out, false, cap, out_len);
}
extern "C" int bzap_decompress_device(bzap_ctx *ctx, const uint8_t *in, size_t in_len, uint8_t *out, size_t cap, size_t *out_len)
{
    RESOLVE(ctx);
    return decompress_any(ctx, in, true, in_len, out, true, cap, out_len);
}

// ---- files ------------------------------------------------------------------------------------------------
static int read_file(bzap_ctx *ctx, const char *path, std::vector<u8> &data)
{
    FILE *f = fopen(path, "rb");
    if (!f) return bzap_fail(ctx, BZAP_ERR_IO, "cannot open %s", path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return bzap_fail(ctx, BZAP_ERR_IO, "cannot size %s", path); }
    data.resize((size_t)sz);
    size_t got = sz ? fread(data.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return bzap_fail(ctx, BZAP_ERR_IO, "short read on %s", path);
    return BZAP_OK;
}
static int write_file(bzap_ctx *ctx, const char *path, const u8 *data, size_t len)
{
    FILE *f = fopen(path, "wb");
    if (!f) return bzap_fail(ctx, BZAP_ERR_IO, "cannot create %s", path);
    size_t put = len ? fwrite(data, 1, len, f) : 0;
    if (fclose(f) != 0 || put != len) return bzap_fail(ctx, BZAP_ERR_IO, "short write on %s", path);
    return BZAP_OK;
}

extern "C" int bzap_compress_file(bzap_ctx *ctx, const char *in_path, const char *out_path)
{
    if (!in_path || !out_path) return BZAP_ERR_ARG;
    RESOLVE(ctx);
    std::vector<u8> in, out;
    RET(read_file(ctx, in_path, in));                                   // read_bytes, io_utilities.h:29-55
    if (in.empty()) return bzap_fail(ctx, BZAP_ERR_EMPTY, "%s is empty", in_path);
    out.resize(bzap_compress_bound(in.size()));
    size_t len = 0;
    RET(compress_any(ctx, in.data(), false, in.size(), out.data(), false, out.size(), &len));
    return write_file(ctx, out_path, out.data(), len);                  // write_bytes, io_utilities.h:7-27
}

extern "C" int bzap_decompress_file(bzap_ctx *ctx, const char *in_path, const char *out_path)
{
    if (!in_path || !out_path) return BZAP_ERR_ARG;
    RESOLVE(ctx);
    std::vector<u8> in, out;
    RET(read_file(ctx, in_path, in));
    if (in.size() < BZAP_HEADER_BYTES) return bzap_fail(ctx, BZAP_ERR_CORRUPT, "%s is shorter than a header", in_path);
    u64 n = bzap_decompressed_size(in.data(), in.size());
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %llu bytes", (unsigned long long)n);
    out.resize((size_t)n + 1);
    size_t len = 0;
    RET(decompress_any(ctx, in.data(), false, in.size(), out.data(), false, (size_t)n, &len));
    return write_file(ctx, out_path, out.data(), len);
}

// ---- batch: independent files, one BWT block each, over the streams of one or several GPUs -------------------
// Replaces the reference's 14-file loop (main.cpp:424-437).  A persistent pool of worker threads, one per
// (device, stream slot), each with its own context (stream + scratch arena); a batch call hands the files
// out largest first (greedy: a free worker takes the largest remaining file), so that on every device the
// host-to-device copy of one file and the read-back of another overlap the kernels of a third -- the
// pinned-buffer I/O pipeline that replaces read_bytes / write_bytes (io_utilities.h:7-55) for batches.
namespace {
struct BatchJob {
    std::function<int(bzap_ctx *, int)> fn;
    std::vector<int> order;                     // file indices, largest first
    std::atomic<int> next{0};
    std::atomic<int> first_err{BZAP_OK};
    int dev_lo = 0, dev_hi = 1, n_streams = 1;  // devices dev_lo .. dev_hi-1, stream slots 0 .. n_streams-1 take part
    int pending = 0;                            // workers that still have to report (guarded by Engine::mu)
};
struct Worker {
    std::thread th;
    bzap_ctx *ctx = nullptr;
    int device = 0, slot = 0;
};
struct Engine {
    std::mutex api_mu;                          // one batch at a time
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<Worker *> workers;
    BatchJob *job = nullptr;
    unsigned long long generation = 0;
};
Engine *g_engine = nullptr;                     // never destroyed: worker threads outlive main()'s statics
std::once_flag g_engine_once;

void worker_main(Engine *E, Worker *w)
{
    cudaSetDevice(w->device);
    unsigned long long seen = 0;
    while (true) {
        BatchJob *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(E->mu);
            E->cv_work.wait(lk, [&] { return E->generation != seen; });
            seen = E->generation;
            job = E->job;
        }
        if (!job || w->device < job->dev_lo || w->device >= job->dev_hi || w->slot >= job->n_streams) continue;   // sits this one out
        const int count = (int)job->order.size();
        for (int i = job->next.fetch_add(1); i < count; i = job->next.fetch_add(1)) {
            w->ctx->err[0] = 0;
            int rc = job->fn(w->ctx, job->order[i]);
            int ok = BZAP_OK;
            if (rc != BZAP_OK) job->first_err.compare_exchange_strong(ok, rc);
        }
        std::lock_guard<std::mutex> lk(E->mu);
        if (--job->pending == 0) E->cv_done.notify_all();
    }
}

// runs fn(ctx, i) for i < count on n_streams workers of each of the devices 0 .. n_gpus-1 (n_gpus = 0: the
// current device only); sizes[i] orders the hand-out
int run_batch(int count, const size_t *sizes, int n_gpus, int n_streams, std::function<int(bzap_ctx *, int)> fn)
{
    if (count <= 0) return BZAP_OK;
    int have = 0, cur = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return BZAP_ERR_CUDA;      // fail loudly: no CPU path
    if (cudaGetDevice(&cur) != cudaSuccess) return BZAP_ERR_CUDA;
    if (n_streams <= 0) n_streams = 4;
    if (n_streams > 32) n_streams = 32;
    int dev_lo = 0, dev_hi = n_gpus;
    if (n_gpus <= 0) { dev_lo = cur; dev_hi = cur + 1; }
    if (dev_hi > have) return BZAP_ERR_ARG;
    std::call_once(g_engine_once, [] { g_engine = new Engine(); });
    Engine *E = g_engine;
    std::lock_guard<std::mutex> api(E->api_mu);
    // grow the pool: worker (device d, slot s) exists for every pair a batch has asked for so far
    for (int d = dev_lo; d < dev_hi; ++d)
        for (int sl = 0; sl < n_streams; ++sl) {
            bool found = false;
            for (Worker *w : E->workers) found = found || (w->device == d && w->slot == sl);
            if (found) continue;
            Worker *w = new Worker();
            w->device = d;
            w->slot = sl;
            int rc = bzap_ctx_create(d, &w->ctx);
            if (rc != BZAP_OK) { delete w; cudaSetDevice(cur); return rc; }
            w->th = std::thread(worker_main, E, w);
            w->th.detach();
            E->workers.push_back(w);
        }
    cudaSetDevice(cur);
    BatchJob job;
    job.fn = std::move(fn);
    job.order.resize(count);
    for (int i = 0; i < count; ++i) job.order[i] = i;
    if (sizes) std::stable_sort(job.order.begin(), job.order.end(), [&](int a, int b) { return sizes[a] > sizes[b]; });
    job.dev_lo = dev_lo;
    job.dev_hi = dev_hi;
    job.n_streams = n_streams;
    int part = 0;
    for (Worker *w : E->workers) part += (w->device >= dev_lo && w->device < dev_hi && w->slot < n_streams);
    {
        std::unique_lock<std::mutex> lk(E->mu);
        job.pending = part;
        E->job = &job;
        ++E->generation;
        E->cv_work.notify_all();
        E->cv_done.wait(lk, [&] { return job.pending == 0; });
        E->job = nullptr;
    }
    return job.first_err.load();
}
}   // namespace

extern "C" int bzap_compress_batch_gpus(const uint8_t *const *ins, const size_t *ns, uint8_t *const *outs, size_t *out_lens,
                                        int count, int n_gpus, int n_streams)
{
    if (!ins || !ns || !outs || !out_lens) return BZAP_ERR_ARG;
    return run_batch(count, ns, n_gpus, n_streams, [&](bzap_ctx *c, int i) {
        return compress_any(c, ins[i], false, ns[i], outs[i], false, bzap_compress_bound(ns[i]), &out_lens[i]);
    });
}
extern "C" int bzap_decompress_batch_gpus(const uint8_t *const *ins, const size_t *in_lens, uint8_t *const *outs,
                                          const size_t *out_caps, size_t *out_lens, int count, int n_gpus, int n_streams)
{
    if (!ins || !in_lens || !outs || !out_caps || !out_lens) return BZAP_ERR_ARG;
    return run_batch(count, in_lens, n_gpus, n_streams, [&](bzap_ctx *c, int i) {
        return decompress_any(c, ins[i], false, in_lens[i], outs[i], false, out_caps[i], &out_lens[i]);
    });
}
extern "C" int bzap_compress_batch(const uint8_t *const *ins, const size_t *ns, uint8_t *const *outs, size_t *out_lens,
                                   int count, int n_streams)
{
    return bzap_compress_batch_gpus(ins, ns, outs, out_lens, count, 0, n_streams);
}
extern "C" int bzap_decompress_batch(const uint8_t *const *ins, const size_t *in_lens, uint8_t *const *outs,
                                     const size_t *out_caps, size_t *out_lens, int count, int n_streams)
{
    return bzap_decompress_batch_gpus(ins, in_lens, outs, out_caps, out_lens, count, 0, n_streams);
}

// the reference's file loop (main.cpp:424-437) over n_gpus devices: in_paths[i] -> out_paths[i]
static int batch_files(const char *const *in_paths, const char *const *out_paths, int count, int n_gpus, int n_streams, bool comp)
{
    if (!in_paths || !out_paths) return BZAP_ERR_ARG;
    std::vector<size_t> sizes(count > 0 ? count : 0, 0);
    for (int i = 0; i < count; ++i) {
        if (!in_paths[i] || !out_paths[i]) return BZAP_ERR_ARG;
        FILE *f = fopen(in_paths[i], "rb");
        if (f) { fseek(f, 0, SEEK_END); long sz = ftell(f); fclose(f); sizes[i] = sz > 0 ? (size_t)sz : 0; }
    }
    return run_batch(count, sizes.data(), n_gpus, n_streams, [&](bzap_ctx *c, int i) {
        return comp ? bzap_compress_file(c, in_paths[i], out_paths[i]) : bzap_decompress_file(c, in_paths[i], out_paths[i]);
    });
}
extern "C" int bzap_compress_files(const char *const *in_paths, const char *const *out_paths, int count, int n_gpus, int n_streams)
{
    return batch_files(in_paths, out_paths, count, n_gpus, n_streams, true);
}
extern "C" int bzap_decompress_files(const char *const *in_paths, const char *const *out_paths, int count, int n_gpus, int n_streams)
{
    return batch_files(in_paths, out_paths, count, n_gpus, n_streams, false);
}

// ---- stage level ------------------------------------------------------------------------------------------
// Each stage stages its host input in the arena, runs the device stage and copies the result back.
static int stage_in(bzap_ctx *ctx, const u8 *h, size_t n, u8 **d)
{
    *d = arena_get<u8>(ctx, n + 128);
    if (!*d) return bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch");
    CU(ctx, cudaMemsetAsync(*d + (n & ~(size_t)15), 0, n + 128 - (n & ~(size_t)15), ctx->stream));
    if (n) CU(ctx, cudaMemcpyAsync(*d, h, n, cudaMemcpyHostToDevice, ctx->stream));
    return BZAP_OK;
}
static int stage_out(bzap_ctx *ctx, u8 *h, const u8 *d, size_t n)
{
    if (n) CU(ctx, cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

extern "C" int bzap_bwt(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *last_col, uint64_t *primary)
{
    RESOLVE(ctx);
    if (!in || !last_col || !primary) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input");
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    RET(arena_reserve(ctx, scratch_bytes_compress(n)));
    u8 *d_in, *d_out = arena_get<u8>(ctx, n + 64);
    RET(stage_in(ctx, in, n, &d_in));
    if (!d_out) return bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch");
    RET(dev_bwt(ctx, d_in, n, d_out, primary));
    return stage_out(ctx, last_col, d_out, n);
}
extern "C" int bzap_ibwt(bzap_ctx *ctx, const uint8_t *last_col, size_t n, uint64_t primary, uint8_t *out)
{
    RESOLVE(ctx);
    if (!last_col || !out) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n == 0) return BZAP_OK;
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    RET(arena_reserve(ctx, scratch_bytes_decompress(n, 0)));
    u8 *d_in, *d_out = arena_get<u8>(ctx, n + 64);
    RET(stage_in(ctx, last_col, n, &d_in));
    if (!d_out) return bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch");
    RET(dev_ibwt(ctx, d_in, n, primary, d_out));
    return stage_out(ctx, out, d_out, n);
}
static int mtf_stage(bzap_ctx *ctx, const u8 *in, size_t n, u8 *out, bool inverse)
{
    if (!in || !out) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n == 0) return BZAP_OK;
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    RET(arena_reserve(ctx, scratch_bytes_decompress(n, 0)));
    u8 *d_in, *d_out = arena_get<u8>(ctx, n + 64);
    RET(stage_in(ctx, in, n, &d_in));
    if (!d_out) return bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch");
    RET(inverse ? dev_imtf(ctx, d_in, n, d_out) : dev_mtf(ctx, d_in, n, d_out));
    return stage_out(ctx, out, d_out, n);
}
extern "C" int bzap_mtf(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out)
{
    RESOLVE(ctx);
    return mtf_stage(ctx, in, n, out, false);
}
extern "C" int bzap_imtf(bzap_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out)
{
    RESOLVE(ctx);
    return mtf_stage(ctx, in, n, out, true);
}
extern "C" int bzap_hist256(bzap_ctx *ctx, const uint8_t *in, size_t n, uint64_t freq[256], uint8_t order[256], int *n_leaves)
{
    RESOLVE(ctx);
    if (!in || !freq || !order || !n_leaves) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input");
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    RET(arena_reserve(ctx, n + (1u << 20)));
    u8 *d_in;
    RET(stage_in(ctx, in, n, &d_in));
    return dev_hist(ctx, d_in, n, freq, order, n_leaves);
}
extern "C" int bzap_huff_build(const uint64_t freq[256], const uint8_t *order, int n_leaves, bzap_tree *tree)
{
    return huff_build_tree(freq, order, n_leaves, tree);
}
extern "C" int bzap_huff_codes(const bzap_tree *tree, uint64_t code[256], uint8_t len[256])
{
    if (!tree || !code || !len) return BZAP_ERR_ARG;
    CodeTable ct;
    int rc = huff_code_table(tree, &ct);
    if (rc != BZAP_OK) return rc;
    memcpy(code, ct.code, sizeof ct.code);
    memcpy(len, ct.len, sizeof ct.len);
    return BZAP_OK;
}
extern "C" int bzap_tree_to_bytes(const bzap_tree *tree, uint8_t *out, size_t *len) { return huff_tree_to_bytes(tree, out, len); }
extern "C" int bzap_bytes_to_tree(const uint8_t *bytes, size_t len, bzap_tree *tree) { return huff_bytes_to_tree(bytes, len, tree); }

extern "C" int bzap_huff_encode(bzap_ctx *ctx, const uint8_t *in, size_t n, const bzap_tree *tree, uint8_t *out,
                                size_t out_cap, size_t *out_len)
{
    RESOLVE(ctx);
    if (!in || !tree || !out || !out_len) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    CodeTable ct;
    int rc = huff_code_table(tree, &ct);
    if (rc != BZAP_OK) return bzap_fail(ctx, rc, "code table");
    // payload size needs the histogram of the input under this code: count on the device
    RET(arena_reserve(ctx, (size_t)ct.max_len * n / 8 + 2 * n + (4u << 20)));
    u8 *d_in;
    RET(stage_in(ctx, in, n, &d_in));
    u64 bits = 0;
    if (n) {
        u64 freq[256];
        u8 order[256];
        int leaves;
        RET(dev_hist(ctx, d_in, n, freq, order, &leaves));
        for (int s = 0; s < 256; ++s)
            if (freq[s] && ct.len[s] == 0 && tree->n_leaves > 1) return bzap_fail(ctx, BZAP_ERR_ARG, "symbol %d has no code", s);
        bits = huff_total_bits(freq, &ct);
    }
    const size_t payload = bits ? (size_t)((bits + 7) / 8) : 1;
    if (payload > out_cap) return bzap_fail(ctx, BZAP_ERR_CAPACITY, "need %zu bytes", payload);
    u8 *d_out = arena_get<u8>(ctx, payload + 64);
    if (!d_out) return bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch");
    CU(ctx, cudaMemsetAsync(d_out, 0, (payload + 63) & ~(size_t)31, ctx->stream));
    RET(dev_huff_encode(ctx, d_in, n, &ct, d_out, 0));
    *out_len = payload;
    return stage_out(ctx, out, d_out, payload);
}

extern "C" int bzap_huff_decode(bzap_ctx *ctx, const uint8_t *payload, size_t payload_len, const bzap_tree *tree, size_t n,
                                uint8_t *out)
{
    RESOLVE(ctx);
    if (!payload || !tree || !out) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n);
    if (n == 0) return BZAP_OK;
    DecodeTables dt;
    int rc = huff_decode_tables(tree, &dt);
    if (rc != BZAP_OK) return bzap_fail(ctx, rc, "decode tables");
    rc = arena_reserve(ctx, scratch_bytes_decompress(n, payload_len));
    u8 *d_in = nullptr, *d_out = nullptr;
    if (rc == BZAP_OK) rc = stage_in(ctx, payload, payload_len, &d_in);
    if (rc == BZAP_OK) { d_out = arena_get<u8>(ctx, n + 64); if (!d_out) rc = bzap_fail(ctx, BZAP_ERR_NOMEM, "stage scratch"); }
    if (rc == BZAP_OK) rc = dev_huff_decode(ctx, d_in, payload_len, &dt, n, d_out);
    huff_free_decode_tables(&dt);
    if (rc != BZAP_OK) return rc;
    return stage_out(ctx, out, d_out, n);
}
