// dist_block.cu -- ONE BWT block compressed by several GPUs of one box (BASELINE config 5 ii).
//
// Replaces bwt() + move_to_front() + huffman() + write_bytes() (main.cpp:304-324) for a block that is
// spread over G GPUs: one rank (process or thread) per GPU, NCCL send/recv groups over NVLink for
// the exchanges, every per-GPU step a kernel of this library.  The text is resident on every GPU;
// ranks, keys and suffix-array slots are sharded.  tests/dist_model.py is the executable
// specification of the host logic (same step names), run on CPU over gloo.
//
//   0. splitters      8-byte keys of hashed sample positions, identical on every rank (no exchange).
//                     Rank g owns the rotations whose 8-byte key lies in [bound[g], bound[g+1]): equal
//                     keys never straddle two ranks, so groups -- and their sparse ranks -- stay local
//   1. select + sort  own (key, start) pairs compacted from the text, onesweep sort, sparse ranks
//                     rs = base_g + index of the group head (base_g = rotations owned by lower ranks)
//   2. ranks go home  (start, rank) pairs bucketed by the owner of text position `start`
//                     (contiguous shards aligned to the top-8-bit buckets), all-to-all, scatter
//   3. rounds         only rotations still in a group > 1: pull r2 = rank[(start + k) mod N] from the
//                     position owners (request / response all-to-all), sort (r1, r2) locally, new
//                     sparse ranks inside the group's slot range, ranks go home, survivors stay
//   4. last column    L[j] = text[(sa[j] - 1) mod N] for the slots this rank holds; primary = rank[0]
//   5. MTF            "last occurrence" summary of every rank's piece, all-gather, start list of each piece
//   6. Huffman        per-rank histogram + first appearance, all-gather; the tree on every rank; per-rank
//                     bit counts give the bit offset of every rank's payload piece; pieces are sent
//                     to rank 0 and OR-merged at their (non byte-aligned) seams
//
// A context without a communicator is a world of one: the same code with device-to-device copies in
// place of the exchanges.
#include "device_common.cuh"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <vector>

#define DIST_MAX_WORLD 16
#define DIST_SAMPLES_PER_RANK 512
#define DIST_HOST_BYTES (256 * 1024)

// ---- NCCL, bound at run time --------------------------------------------------------------------------------
// libnccl.so.2 is resolved with dlopen so that libbzap.so loads on hosts without NCCL (every other entry
// point works there); a process that already holds NCCL (PyTorch) shares that copy.
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load()
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    bool ok = true;
#define NCCL_SYM(field, name)                                                                     \
    do {                                                                                          \
        *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                                   \
        ok = ok && g_nccl.field != nullptr;                                                       \
    } while (0)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(Send, "ncclSend");
    NCCL_SYM(Recv, "ncclRecv");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(Broadcast, "ncclBroadcast");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.ok = ok;
}
const NcclApi *nccl()
{
    std::call_once(g_nccl_once, nccl_load);
    return g_nccl.ok ? &g_nccl : nullptr;
}
}   // namespace

#define NC(ctx, call)                                                                             \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess)                                                                    \
            return bzap_fail(ctx, BZAP_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                             g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");            \
    } while (0)

extern "C" int bzap_comm_unique_id(uint8_t id[BZAP_COMM_ID_BYTES])
{
    if (!id) return BZAP_ERR_ARG;
    const NcclApi *N = nccl();
    if (!N) return BZAP_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == BZAP_COMM_ID_BYTES, "NCCL unique id size");
    ncclUniqueId u;
    if (N->GetUniqueId(&u) != ncclSuccess) return BZAP_ERR_NCCL;
    memcpy(id, &u, sizeof u);
    return BZAP_OK;
}

// Peer windows: every rank's scratch arena (one cudaMalloc) is mapped into the other ranks' address spaces
// through CUDA IPC, so that the bulk exchanges are plain device-to-device copies over NVLink into the
// receiver's buffer (measured on this pool: 760 GB/s per copy against 150-270 GB/s for the same segments
// through ncclSend/ncclRecv groups).  NCCL keeps the small collectives, which double as the barriers.
struct PeerMap {
    cudaIpcMemHandle_t handle[DIST_MAX_WORLD];
    u8 *base[DIST_MAX_WORLD];
    bool open[DIST_MAX_WORLD];
};
static void peers_close(bzap_ctx *ctx)
{
    PeerMap *pm = (PeerMap *)ctx->peers;
    if (!pm) return;
    for (int p = 0; p < DIST_MAX_WORLD; ++p)
        if (pm->open[p]) cudaIpcCloseMemHandle(pm->base[p]);
    delete pm;
    ctx->peers = nullptr;
}

// Collective when the ranks have mapped each other's arenas: everybody closes its mappings, and only after
// everybody has done so is any arena freed (freeing exported memory that is still mapped elsewhere is undefined).
void dist_comm_release(bzap_ctx *ctx)
{
    peers_close(ctx);
    if (ctx->comm && g_nccl.ok) {
        if (ctx->dist_had_peers && ctx->dist_dev) {
            u32 *w = (u32 *)ctx->dist_dev;
            if (g_nccl.AllReduce(w, w + 4, 1, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream) == ncclSuccess) {
                // bounded: a rank that died must not keep the survivors from shutting down
                const auto t0 = std::chrono::steady_clock::now();
                while (cudaStreamQuery(ctx->stream) == cudaErrorNotReady &&
                       std::chrono::steady_clock::now() - t0 < std::chrono::seconds(5)) {}
            }
        }
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    ctx->dist_had_peers = false;
    if (ctx->dist_arena) { cudaFree(ctx->dist_arena); ctx->dist_arena = nullptr; ctx->dist_arena_cap = 0; }
    if (ctx->dist_dev) { cudaFree(ctx->dist_dev); ctx->dist_dev = nullptr; }
    if (ctx->dist_host) { cudaFreeHost(ctx->dist_host); ctx->dist_host = nullptr; }
}

extern "C" int bzap_ctx_comm_init(bzap_ctx *ctx, const uint8_t id[BZAP_COMM_ID_BYTES], int world, int rank)
{
    if (!ctx || !id || world < 1 || world > DIST_MAX_WORLD || rank < 0 || rank >= world) return BZAP_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return BZAP_ERR_CUDA;
    ctx->err[0] = 0;
    if (ctx->comm) dist_comm_release(ctx);
    if (world == 1) return BZAP_OK;
    const NcclApi *N = nccl();
    if (!N) return bzap_fail(ctx, BZAP_ERR_NCCL, "libnccl.so.2 not found (dlopen)");
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    NC(ctx, N->CommInitRank(&c, world, u, rank));
    ctx->comm = c;
    ctx->world = world;
    ctx->rank = rank;
    return BZAP_OK;
}

extern "C" int bzap_get_dist_stats(bzap_ctx *ctx, bzap_dist_stats *out)
{
    if (!ctx || !out) return BZAP_ERR_ARG;
    *out = ctx->dstats;
    return BZAP_OK;
}

// ---- kernels --------------------------------------------------------------------------------------------------
#define DB_BLOCK 256
#define DB_ITEMS 8
#define DB_TILE (DB_BLOCK * DB_ITEMS)

// position of sample j: splitmix64 finaliser (tests/dist_model.py sample_hash)
__host__ __device__ static inline u64 dist_sample_hash(u64 j)
{
    u64 z = (j + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void dist_sample_keys_kernel(const u8 *__restrict__ text, u32 n, u32 samples, u64 *__restrict__ keys)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= samples) return;
    u64 p = dist_sample_hash(j) % n;
    u64 k = 0;
    for (int b = 0; b < 8; ++b) {                  // cyclic window (main.cpp:38-44)
        k = (k << 8) | text[p];
        if (++p == n) p = 0;
    }
    keys[j] = k;
}

// ---- streaming the text through shared memory: 1-D bulk copies (TMA engine) on an mbarrier pipeline ----------
// Both text sweeps give every block a contiguous range of 2 KiB tiles.  One elected thread issues
// cp.async.bulk for tile t+1 into the other of two buffers while the block computes the windows of tile t;
// the copy completes on that buffer's mbarrier (expect_tx / try_wait.parity), so nobody spends instructions
// on the staging and the load latency hides behind the previous tile.  Tiles that wrap around the end of
// the text, or a text that is not 16-byte aligned, are staged by hand.
struct TileStream {
    __align__(16) u8 buf[2][DB_TILE + 16];
    unsigned long long bar[2];
};
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 phase)
{
    u32 done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(phase)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ bool tile_is_bulk(const u8 *text, u32 n, u32 tile)
{
    return (u64)tile * DB_TILE + DB_TILE + 16 <= n && (reinterpret_cast<uintptr_t>(text) & 15u) == 0;
}
// called by ONE thread, after a block barrier that retired every read of `buf`
__device__ __forceinline__ void tile_issue(TileStream &TS, u32 b, const u8 *text, u32 tile)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy accesses of the buffer first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&TS.bar[b])), "r"((u32)(DB_TILE + 16)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(TS.buf[b])),
                 "l"(text + (size_t)tile * DB_TILE), "r"((u32)(DB_TILE + 16)), "r"(smem_addr(&TS.bar[b]))
                 : "memory");
}
__device__ __forceinline__ void tile_stream_begin(TileStream &TS, const u8 *text, u32 n, u32 t0, u32 t1)
{
    if (threadIdx.x == 0) {
        mbar_init(&TS.bar[0], 1);
        mbar_init(&TS.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && t0 < t1 && tile_is_bulk(text, n, t0)) tile_issue(TS, 0, text, t0);
}
// makes tile `tile` available in TS.buf[b] (b = parity of tile - t0) and starts the copy of the next one;
// `phases` (bit b = phase of barrier b) is uniform over the block.  The caller ends the iteration with
// __syncthreads() so that the buffer can be refilled.
__device__ __forceinline__ const u32 *tile_stream_get(TileStream &TS, const u8 *text, u32 n, u32 tile, u32 t0, u32 t1, u32 &phases)
{
    const u32 b = (tile - t0) & 1u;
    if (threadIdx.x == 0 && tile + 1 < t1 && tile_is_bulk(text, n, tile + 1)) tile_issue(TS, b ^ 1u, text, tile + 1);
    if (tile_is_bulk(text, n, tile)) {
        mbar_wait(&TS.bar[b], (phases >> b) & 1u);
        phases ^= 1u << b;
    } else {
        const u32 tile_base = tile * DB_TILE;
        for (u32 i = threadIdx.x; i < DB_TILE + 8; i += DB_BLOCK) {
            u64 p = (u64)tile_base + i;
            if (p >= n) p %= n;                    // cyclic window (main.cpp:38-44)
            TS.buf[b][i] = text[p];
        }
        __syncthreads();
    }
    return reinterpret_cast<const u32 *>(TS.buf[b]);
}
// big-endian 8-byte window starting at byte o of the staged tile
__device__ __forceinline__ u64 dist_window(const u32 *sw, u32 o)
{
    u32 w0 = sw[o >> 2], w1 = sw[(o >> 2) + 1], w2 = sw[(o >> 2) + 2];
    u32 sh = (o & 3u) * 8u;
    u32 first = __funnelshift_r(w0, w1, sh), second = __funnelshift_r(w1, w2, sh);
    return ((u64)__byte_perm(first, 0, 0x0123) << 32) | __byte_perm(second, 0, 0x0123);
}

// Both text sweeps give every block a CONTIGUOUS range of tiles: the count sweep records how many own
// rotations each range holds, so the select sweep knows every range's output offset up front -- no
// look-back between tiles (a tile does so little work here that a look-back degenerates into a chain:
// 5.9 ms for 1 GiB of text with one ticketed tile per block).
// counts[0] = rotations with key < lo_key (owned by lower ranks), counts[1] = rotations in [lo_key, hi_key)
__global__ void __launch_bounds__(DB_BLOCK)
dist_owner_count_kernel(const u8 *__restrict__ text, u32 n, u64 lo_key, u64 hi_key, int hi_open, u32 tiles, u32 *counts,
                        u32 *__restrict__ range_cnt)
{
    __shared__ TileStream TS;
    __shared__ u32 s_tmp[40];
    const u32 tpb = (tiles + gridDim.x - 1) / gridDim.x;
    const u32 t0 = min(tiles, blockIdx.x * tpb), t1 = min(tiles, t0 + tpb);
    u32 below = 0, mine = 0, phases = 0;
    tile_stream_begin(TS, text, n, t0, t1);
    for (u32 tile = t0; tile < t1; ++tile) {
        const u32 base = tile * DB_TILE;
        const u32 *sw = tile_stream_get(TS, text, n, tile, t0, t1, phases);
#pragma unroll
        for (int i = 0; i < DB_ITEMS; ++i) {
            u32 o = threadIdx.x + i * DB_BLOCK;
            if (base + o < n) {
                u64 k = dist_window(sw, o);
                below += k < lo_key;
                mine += k >= lo_key && (hi_open || k < hi_key);
            }
        }
        __syncthreads();
    }
    u32 tb, tm;
    block_exclusive_sum(below, s_tmp, &tb);
    block_exclusive_sum(mine, s_tmp, &tm);
    if (threadIdx.x == 0) {
        range_cnt[blockIdx.x] = tm;
        if (tb) atomicAdd(&counts[0], tb);
        if (tm) atomicAdd(&counts[1], tm);
    }
}

// own rotations, in text order: keys[], starts[] and the eight digit histograms of the keys (same grid as the count sweep)
__global__ void __launch_bounds__(DB_BLOCK)
dist_select_kernel(const u8 *__restrict__ text, u32 n, u64 lo_key, u64 hi_key, int hi_open, u32 tiles,
                   const u32 *__restrict__ range_cnt, u64 *__restrict__ keys, u32 *__restrict__ starts, u32 *hist8)
{
    __shared__ TileStream TS;
    __shared__ u32 s_h[8 * 256];
    __shared__ u32 s_tmp[40];
    for (u32 i = threadIdx.x; i < 8 * 256; i += DB_BLOCK) s_h[i] = 0;
    const u32 tpb = (tiles + gridDim.x - 1) / gridDim.x;
    const u32 t0 = min(tiles, blockIdx.x * tpb), t1 = min(tiles, t0 + tpb);
    u32 phases = 0;
    tile_stream_begin(TS, text, n, t0, t1);
    u32 part = 0;
    for (u32 b = threadIdx.x; b < blockIdx.x; b += DB_BLOCK) part += range_cnt[b];
    u32 out;
    block_exclusive_sum(part, s_tmp, &out);               // own rotations of all earlier ranges
    for (u32 tile = t0; tile < t1; ++tile) {
        const u32 base = tile * DB_TILE;
        const u32 *sw = tile_stream_get(TS, text, n, tile, t0, t1, phases);
        u64 k[DB_ITEMS];
        u32 keep = 0, cnt = 0;
#pragma unroll
        for (int i = 0; i < DB_ITEMS; ++i) {
            u32 o = threadIdx.x * DB_ITEMS + i;        // blocked: a thread owns consecutive positions, output stays in text order
            k[i] = dist_window(sw, o);
            bool own = base + o < n && k[i] >= lo_key && (hi_open || k[i] < hi_key);
            keep |= (u32)own << i;
            cnt += own;
        }
        u32 total;
        u32 o = out + block_exclusive_sum(cnt, s_tmp, &total);
        out += total;
#pragma unroll
        for (int i = 0; i < DB_ITEMS; ++i)
            if ((keep >> i) & 1u) {
                keys[o] = k[i];
                starts[o] = base + threadIdx.x * DB_ITEMS + i;
                ++o;
#pragma unroll
                for (int p = 0; p < 8; ++p) atomicAdd(&s_h[p * 256 + ((u32)(k[i] >> (8 * p)) & 0xffu)], 1u);
            }
        __syncthreads();                            // the tile's buffer may be refilled
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 8 * 256; i += DB_BLOCK) {
        u32 c = s_h[i];
        if (c) atomicAdd(&hist8[i], c);
    }
}

// request for rank[(start + k) mod n]: (position << 32) | r1, plus the histogram of the position's bucket
// (bucket = position >> shift; whole buckets belong to one owner)
__global__ void __launch_bounds__(256)
dist_request_kernel(const u32 *__restrict__ act_idx, const u32 *__restrict__ act_r1, u32 m, u32 n, u32 kk, u32 shift,
                    u64 *__restrict__ req, u32 *hist256)
{
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (u32 a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) {
        u32 p = act_idx[a] + kk;                   // < 2n <= 2^31
        if (p >= n) p -= n;
        req[a] = ((u64)p << 32) | act_r1[a];
        atomicAdd(&s_h[p >> shift], 1u);
    }
    __syncthreads();
    u32 c = s_h[threadIdx.x];
    if (c) atomicAdd(&hist256[threadIdx.x], c);
}

// the owner's side of the pull: resp[j] = rank_home[position - lo]
__global__ void __launch_bounds__(256)
dist_respond_kernel(const u64 *__restrict__ req, u32 cnt, const u32 *__restrict__ rank_home, u32 lo, u32 *__restrict__ resp)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x)
        resp[j] = rank_home[(u32)(req[j] >> 32) - lo];
}
// the same for requests that travelled as bare positions (only the position crosses NVLink: r1 stays home)
__global__ void __launch_bounds__(256)
dist_respond_pos_kernel(const u32 *__restrict__ pos, u32 cnt, const u32 *__restrict__ rank_home, u32 lo, u32 *__restrict__ resp)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) resp[j] = rank_home[pos[j] - lo];
}
__global__ void __launch_bounds__(256) dist_request_pos_kernel(const u64 *__restrict__ req, u32 m, u32 *__restrict__ pos)
{
    for (u32 a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) pos[a] = (u32)(req[a] >> 32);
}

// sort keys of a round: (r1 << rshift) | r2, with their eight digit histograms
__global__ void __launch_bounds__(256)
dist_pair_keys_kernel(const u64 *__restrict__ req, const u32 *__restrict__ r2, u32 m, u32 rshift, u64 *__restrict__ keys,
                      u32 *hist8)
{
    __shared__ u32 s_h[8 * 256];
    for (u32 i = threadIdx.x; i < 8 * 256; i += 256) s_h[i] = 0;
    __syncthreads();
    for (u32 a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) {
        u64 key = ((u64)(u32)req[a] << rshift) | r2[a];
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

// file[first + i] |= piece[i]: a payload piece lands at its byte offset; the seam bytes it shares with its
// neighbours are OR-merged (bit offsets of the pieces are not byte aligned)
__global__ void __launch_bounds__(256) dist_or_merge_kernel(u8 *__restrict__ file, const u8 *__restrict__ piece, size_t bytes)
{
    // both sides are 16-byte aligned by construction
    const size_t nvec = bytes / 16;
    uint4 *f4 = reinterpret_cast<uint4 *>(file);
    const uint4 *p4 = reinterpret_cast<const uint4 *>(piece);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 a = f4[i], b = p4[i];
        a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
        f4[i] = a;
    }
    for (size_t i = nvec * 16 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x)
        file[i] |= piece[i];
}

// ---- host side ---------------------------------------------------------------------------------------------------
namespace {
struct Geometry {
    int G, me;
    u32 n;
    u32 shift;        // bucket = position >> shift (at most 256 buckets)
    u32 nb;           // buckets in use
    u32 bpr;          // buckets per rank
    u32 lo(int g) const { u64 v = ((u64)g * bpr) << shift; return (u32)(v < n ? v : n); }
    u32 hi(int g) const { return lo(g + 1); }
    int owner_of_bucket(u32 b) const { u32 o = b / bpr; return (int)(o < (u32)G ? o : G - 1); }
};

// BZAP_DIST_TIMING=1: every named phase is bracketed by stream synchronisations and timed on the host
// (profiling only: the synchronisations cost a little); the table goes to stderr at the end of the call
struct PhaseTable {
    bool on = false;
    cudaStream_t stream = nullptr;
    const char *names[32];
    double ms[32];
    int count = 0;
    std::chrono::steady_clock::time_point t0;
    void start() { if (on) { cudaStreamSynchronize(stream); t0 = std::chrono::steady_clock::now(); } }
    void stop(const char *name)
    {
        if (!on) return;
        cudaStreamSynchronize(stream);
        double d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        int i = 0;
        while (i < count && strcmp(names[i], name) != 0) ++i;
        if (i == count) { if (count == 32) return; names[count] = name; ms[count] = 0; ++count; }
        ms[i] += d;
        t0 = std::chrono::steady_clock::now();
    }
    double get(const char *name) const
    {
        for (int i = 0; i < count; ++i) if (strcmp(names[i], name) == 0) return ms[i];
        return 0;
    }
};

struct Xchg {
    PhaseTable pt;
    bzap_ctx *ctx;
    const NcclApi *N;
    ncclComm_t comm;
    Geometry geo;
    u32 *d_small;     // device scratch for the small collectives: G x 1024 words
    u8 *h;            // pinned host, DIST_HOST_BYTES
    u64 sent_bytes = 0;
    // peer windows (see PeerMap): base of every rank's arena in this address space, and the offsets of
    // the buffers that peers push into
    bool p2p = false;
    u8 *peer_base[DIST_MAX_WORLD] = {};
    u64 peer_off[DIST_MAX_WORLD][4] = {};
};
enum { BUF_HOME_IDX = 0, BUF_HOME_VAL = 1, BUF_REQ_IN = 2, BUF_R2 = 3 };
struct CountMatrix {
    u64 c[DIST_MAX_WORLD][DIST_MAX_WORLD];       // c[s][d] = elements rank s holds for rank d
};

static inline u32 grid_1d(u64 items, u32 per_block, u32 cap = 148u * 8u)
{
    u64 g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (u32)(g > cap ? cap : g);
}

// all ranks' copies of `words` u32 at d_mine -> host array [G][words]  (synchronises the stream)
static int gather_words(Xchg &X, const u32 *d_mine, u32 words, u32 **h_out)
{
    bzap_ctx *ctx = X.ctx;
    const int G = X.geo.G;
    u32 *h = (u32 *)X.h;
    if ((size_t)G * words * sizeof(u32) > DIST_HOST_BYTES) return bzap_fail(ctx, BZAP_ERR_ARG, "gather_words: too large");
    if (G == 1) {
        CU(ctx, cudaMemcpyAsync(h, d_mine, words * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        NC(ctx, X.N->AllGather(d_mine, X.d_small, words, ncclUint32, X.comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(h, X.d_small, (size_t)G * words * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *h_out = h;
    return BZAP_OK;
}

// element-wise sum of `words` u32 over the ranks -> host (synchronises); local copy of the input -> h_local
static int reduce_words(Xchg &X, const u32 *d_mine, u32 words, u32 *h_local, u32 *h_sum)
{
    bzap_ctx *ctx = X.ctx;
    u32 *h = (u32 *)X.h;
    CU(ctx, cudaMemcpyAsync(h, d_mine, words * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    if (X.geo.G > 1) {
        NC(ctx, X.N->AllReduce(d_mine, X.d_small, words, ncclUint32, ncclSum, X.comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(h + words, X.d_small, words * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_local, h, words * sizeof(u32));
    memcpy(h_sum, X.geo.G > 1 ? h + words : h, words * sizeof(u32));
    return BZAP_OK;
}

// all-to-all of variable segments: send holds the segments for ranks 0..G-1 back to back, recv receives the
// segments from ranks 0..G-1 back to back; C (read transposed when `back`: an answer travels the way its
// request came) says how many elements of `es` bytes everybody holds for everybody.  send2 / recv2 (may be
// null): a second array with the same segmentation.  `buf` / `buf2` name the receiving buffers in the
// peers' arenas.  With peer windows the segments are copied straight into the receivers' buffers and a
// one-word all-reduce tells everybody that all copies have landed; otherwise one ncclSend/ncclRecv group.
static int alltoallv(Xchg &X, const void *send, void *recv, size_t es, const CountMatrix &C, bool back, int buf,
                     const void *send2 = nullptr, void *recv2 = nullptr, int buf2 = 0)
{
    bzap_ctx *ctx = X.ctx;
    const int G = X.geo.G, me = X.geo.me;
    auto cnt = [&](int s_, int d_) { return back ? C.c[d_][s_] : C.c[s_][d_]; };
    u64 my_s = 0, my_r = 0;
    for (int p = 0; p < me; ++p) { my_s += cnt(me, p); my_r += cnt(p, me); }
    if (cnt(me, me)) {                             // own segment
        CU(ctx, cudaMemcpyAsync((u8 *)recv + my_r * es, (const u8 *)send + my_s * es, cnt(me, me) * es, cudaMemcpyDeviceToDevice,
                                ctx->stream));
        if (send2)
            CU(ctx, cudaMemcpyAsync((u8 *)recv2 + my_r * es, (const u8 *)send2 + my_s * es, cnt(me, me) * es,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (G == 1) return BZAP_OK;
    if (X.p2p) {
        u64 soffs[DIST_MAX_WORLD], run = 0;
        for (int p = 0; p < G; ++p) { soffs[p] = run; run += cnt(me, p); }
        // destinations in the order me+1, me+2, ...: at every step each GPU is written by one sender only
        // (everybody starting with rank 0 would make the receiver's NVLink ingress the bottleneck)
        for (int i = 1; i < G; ++i) {
            const int p = (me + i) % G;
            const u64 c = cnt(me, p);
            if (!c) continue;
            u64 dst = 0;                           // where my segment starts in p's buffer: after the lower ranks' segments
            for (int s_ = 0; s_ < me; ++s_) dst += cnt(s_, p);
            CU(ctx, cudaMemcpyAsync(X.peer_base[p] + X.peer_off[p][buf] + dst * es, (const u8 *)send + soffs[p] * es, c * es,
                                    cudaMemcpyDefault, ctx->stream));
            if (send2)
                CU(ctx, cudaMemcpyAsync(X.peer_base[p] + X.peer_off[p][buf2] + dst * es, (const u8 *)send2 + soffs[p] * es, c * es,
                                        cudaMemcpyDefault, ctx->stream));
            X.sent_bytes += c * es * (send2 ? 2 : 1);
        }
        // every rank's copies precede its contribution in stream order: when the reduction completes here,
        // all segments addressed to this rank have landed
        u32 *d_flag = X.d_small + (size_t)G * 1024 - 8;
        NC(ctx, X.N->AllReduce(d_flag, d_flag + 4, 1, ncclUint32, ncclSum, X.comm, ctx->stream));
        return BZAP_OK;
    }
    u64 soff = 0, roff = 0;
    NC(ctx, X.N->GroupStart());
    for (int p = 0; p < G; ++p) {
        const u64 sc = cnt(me, p), rc = cnt(p, me);
        if (p != me) {
            if (sc) {
                NC(ctx, X.N->Send((const u8 *)send + soff * es, sc * es, ncclUint8, p, X.comm, ctx->stream));
                if (send2) NC(ctx, X.N->Send((const u8 *)send2 + soff * es, sc * es, ncclUint8, p, X.comm, ctx->stream));
                X.sent_bytes += sc * es * (send2 ? 2 : 1);
            }
            if (rc) {
                NC(ctx, X.N->Recv((u8 *)recv + roff * es, rc * es, ncclUint8, p, X.comm, ctx->stream));
                if (recv2) NC(ctx, X.N->Recv((u8 *)recv2 + roff * es, rc * es, ncclUint8, p, X.comm, ctx->stream));
            }
        }
        soff += sc;
        roff += rc;
    }
    NC(ctx, X.N->GroupEnd());
    return BZAP_OK;
}

// who holds how much for whom, from everybody's bucket histograms (whole buckets belong to one owner)
static void matrix_from_hists(const Geometry &geo, const u32 *h_all /* [G][256] */, CountMatrix *C)
{
    for (int s = 0; s < geo.G; ++s) {
        for (int d = 0; d < geo.G; ++d) C->c[s][d] = 0;
        for (u32 b = 0; b < 256; ++b) C->c[s][geo.owner_of_bucket(b)] += h_all[(size_t)s * 256 + b];
    }
}

// maps the peers' arenas (or learns that it cannot): collective.  offs = offsets of this rank's receiving
// buffers inside its arena
static int peers_setup(Xchg &X, const u64 offs[4])
{
    bzap_ctx *ctx = X.ctx;
    const int G = X.geo.G, me = X.geo.me;
    X.p2p = false;
    if (G == 1) return BZAP_OK;
    const bool want = getenv("BZAP_DIST_NO_P2P") == nullptr;
    struct Hello {
        cudaIpcMemHandle_t handle;
        u64 off[4];
        u32 ok, pad[7];
    };
    static_assert(sizeof(Hello) % 4 == 0 && sizeof(Hello) / 4 <= 64, "hello size");
    const u32 words = sizeof(Hello) / 4;
    Hello mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = want && cudaIpcGetMemHandle(&mine.handle, ctx->arena) == cudaSuccess;
    if (!mine.ok) cudaGetLastError();
    for (int i = 0; i < 4; ++i) mine.off[i] = offs[i];
    u32 *d_mine = X.d_small + (size_t)G * 1024 - 128;
    u8 *h_stage = X.h + DIST_HOST_BYTES - 1024;
    memcpy(h_stage, &mine, sizeof mine);
    CU(ctx, cudaMemcpyAsync(d_mine, h_stage, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    u32 *h_all = nullptr;
    RET(gather_words(X, d_mine, words, &h_all));
    if (!ctx->peers) {
        PeerMap *pm = new (std::nothrow) PeerMap();
        if (!pm) return bzap_fail(ctx, BZAP_ERR_NOMEM, "peer map");
        memset(pm, 0, sizeof *pm);
        ctx->peers = pm;
    }
    PeerMap *pm = (PeerMap *)ctx->peers;
    u32 all_ok = 1;
    std::vector<Hello> hello(G);
    for (int p = 0; p < G; ++p) {
        memcpy(&hello[p], h_all + (size_t)p * words, sizeof(Hello));
        all_ok &= hello[p].ok;
    }
    u32 my_ok = all_ok;
    if (all_ok) {
        for (int p = 0; p < G && my_ok; ++p) {
            if (p == me) continue;
            if (pm->open[p] && memcmp(&pm->handle[p], &hello[p].handle, sizeof(cudaIpcMemHandle_t)) == 0) continue;
            if (pm->open[p]) { cudaIpcCloseMemHandle(pm->base[p]); pm->open[p] = false; }
            void *b = nullptr;
            if (cudaIpcOpenMemHandle(&b, hello[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                my_ok = 0;
                break;
            }
            pm->base[p] = (u8 *)b;
            pm->handle[p] = hello[p].handle;
            pm->open[p] = true;
        }
    }
    // everybody must have mapped everybody, or nobody uses the windows
    u32 *h_ok = (u32 *)h_stage;
    *h_ok = my_ok;
    CU(ctx, cudaMemcpyAsync(d_mine, h_ok, sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    RET(gather_words(X, d_mine, 1, &h_all));
    bool ok = true;
    for (int p = 0; p < G; ++p) ok = ok && h_all[p] != 0;
    if (ok)
        for (int p = 0; p < G; ++p) {
            X.peer_base[p] = p == me ? ctx->arena : pm->base[p];
            for (int i = 0; i < 4; ++i) X.peer_off[p][i] = hello[p].off[i];
        }
    X.p2p = ok;
    if (ok) ctx->dist_had_peers = true;
    return BZAP_OK;
}

struct HomeBufs {
    u32 *b_idx, *b_val;       // bucketed pairs (capacity: pairs sent)
    u32 *x_idx, *x_val;       // received pairs regrouped by shard window (capacity: shard)
    u32 *r_idx, *r_val;       // received pairs (capacity: shard)
    u32 *d_hist, *d_bctl;     // 256 words, bucket_ctl_words
};

// (start, rank) pairs travel to the owner of text position `start` and are scattered into its shard
static int ranks_go_home(Xchg &X, const u32 *d_idx, const u32 *d_val, u32 cnt, const HomeBufs &B, u32 *rank_home)
{
    bzap_ctx *ctx = X.ctx;
    const Geometry &geo = X.geo;
    X.pt.start();
    const u32 *s_idx = B.b_idx, *s_val = B.b_val;                // what is sent: pairs grouped by owner
    if (cnt) RET(dev_bucket_pass_u32(ctx, d_idx, d_val, cnt, (int)geo.shift, B.b_idx, B.b_val, B.d_hist, B.d_bctl));
    else CU(ctx, cudaMemsetAsync(B.d_hist, 0, 256 * sizeof(u32), ctx->stream));
    X.pt.stop("home.bucket");
    u32 *h_all = nullptr;
    RET(gather_words(X, B.d_hist, 256, &h_all));
    X.pt.stop("home.counts");
    CountMatrix C;
    matrix_from_hists(geo, h_all, &C);
    u64 total = 0;
    for (int s = 0; s < geo.G; ++s) total += C.c[s][geo.me];
    if (total > (u64)(geo.hi(geo.me) - geo.lo(geo.me))) return bzap_fail(ctx, BZAP_ERR_CUDA, "ranks_go_home: %llu pairs for a shard of %u",
                                                                          (unsigned long long)total, geo.hi(geo.me) - geo.lo(geo.me));
    RET(alltoallv(X, s_idx, B.r_idx, sizeof(u32), C, false, BUF_HOME_IDX, s_val, B.r_val, BUF_HOME_VAL));
    X.pt.stop("home.alltoall");
    // The pairs arrive as one run per sender.  Scattered as they are, every 32-byte sector of the shard
    // would be touched once per sender, far apart in time, and be read-modified-written each time (measured
    // at 4 GPUs: 31 G pairs/s against 112 G/s for one sender).  So the owner regroups what it received by
    // the top bits of the shard offset first: all contributions to a window of <= 2^20 entries (4 MiB) sit
    // together and every sector is completed in L2.  (idx >> s) & 255 is injective over the shard because
    // the shard starts at a multiple of 2^shift >= 2^s and spans at most 256 windows.
    const u32 *x_idx = B.r_idx, *x_val = B.r_val;
    if (total > (4u << 20)) {
        u32 lbits = 0;
        while (((u64)1 << lbits) < (u64)(geo.hi(geo.me) - geo.lo(geo.me))) ++lbits;
        if (lbits > 28) {                                             // two LSD passes: windows of 2^(lbits-16) entries
            RET(dev_bucket_pass_u32(ctx, B.r_idx, B.r_val, (u32)total, (int)lbits - 16, B.x_idx, B.x_val, B.d_hist, B.d_bctl));
            RET(dev_bucket_pass_u32(ctx, B.x_idx, B.x_val, (u32)total, (int)lbits - 8, B.r_idx, B.r_val, B.d_hist, B.d_bctl));
        } else if (lbits > 20) {
            RET(dev_bucket_pass_u32(ctx, B.r_idx, B.r_val, (u32)total, (int)lbits - 8, B.x_idx, B.x_val, B.d_hist, B.d_bctl));
            x_idx = B.x_idx;
            x_val = B.x_val;
        }
    }
    X.pt.stop("home.regroup");
    RET(dev_scatter_offset_async(ctx, x_idx, x_val, (u32)total, geo.lo(geo.me), rank_home));
    X.pt.stop("home.scatter");
    CU(ctx, cudaGetLastError());
    return BZAP_OK;
}

static double ev_ms2(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
static void put_u64le(u8 *p, u64 v) { for (int i = 0; i < 8; ++i) p[i] = (u8)(v >> (8 * i)); }
}   // namespace

namespace {
// the distributed path works in its own arena (ctx->dist_arena): swapped in for the duration of the call
struct ArenaSwap {
    bzap_ctx *ctx;
    u8 *arena;
    size_t cap, off;
    explicit ArenaSwap(bzap_ctx *c) : ctx(c), arena(c->arena), cap(c->arena_cap), off(c->arena_off)
    {
        c->arena = c->dist_arena;
        c->arena_cap = c->dist_arena_cap;
        c->arena_off = 0;
    }
    ~ArenaSwap()
    {
        ctx->dist_arena = ctx->arena;
        ctx->dist_arena_cap = ctx->arena_cap;
        ctx->arena = arena;
        ctx->arena_cap = cap;
        ctx->arena_off = off;
    }
};

// Grows the arena if needed.  Peers may have it mapped: if ANY rank has to reallocate, every rank first closes
// its mappings, and nobody frees before everybody has closed.
static int dist_reserve(bzap_ctx *ctx, const NcclApi *N, int G, size_t bytes)
{
    u32 *w = (u32 *)ctx->dist_dev;                 // [0] flag in, [4] flag out, [8] barrier in, [12] barrier out
    u32 grow = bytes > ctx->arena_cap ? 1u : 0u;
    if (G > 1) {
        u32 *h = (u32 *)(ctx->dist_host + DIST_HOST_BYTES - 2048);
        *h = grow;
        CU(ctx, cudaMemcpyAsync(w, h, sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
        NC(ctx, N->AllReduce(w, w + 4, 1, ncclUint32, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(h + 1, w + 4, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (h[1]) {
            peers_close(ctx);
            NC(ctx, N->AllReduce(w + 8, w + 12, 1, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return arena_reserve(ctx, bytes);
}
}   // namespace

extern "C" int bzap_compress_block_distributed(bzap_ctx *ctx, const uint8_t *d_text, size_t n64, uint8_t *d_out, size_t out_cap,
                                               size_t *out_len)
{
    if (!ctx) return BZAP_ERR_ARG;                                   // a communicator belongs to an explicit context
    if (cudaSetDevice(ctx->device) != cudaSuccess) return BZAP_ERR_CUDA;
    ctx->err[0] = 0;
    const int G = ctx->comm ? ctx->world : 1, me = ctx->comm ? ctx->rank : 0;
    if (!d_text || !out_len || (me == 0 && !d_out)) return bzap_fail(ctx, BZAP_ERR_ARG, "null pointer");
    if (n64 == 0) return bzap_fail(ctx, BZAP_ERR_EMPTY, "empty input (the reference crashes here, main.cpp:245)");
    if (n64 > BZAP_MAX_BLOCK) return bzap_fail(ctx, BZAP_ERR_TOO_LARGE, "block of %zu bytes", n64);
    const NcclApi *N = G > 1 ? nccl() : nullptr;
    if (G > 1 && !N) return bzap_fail(ctx, BZAP_ERR_NCCL, "libnccl.so.2 not found");
    const u32 n = (u32)n64;
    *out_len = 0;
    if (!ctx->dist_host) CU(ctx, cudaMallocHost((void **)&ctx->dist_host, DIST_HOST_BYTES));
    if (!ctx->dist_dev) {
        CU(ctx, cudaMalloc((void **)&ctx->dist_dev, 1u << 20));
        CU(ctx, cudaMemset(ctx->dist_dev, 0, 1u << 20));
    }
    ArenaSwap arena_swap(ctx);
    bzap_dist_stats &st = ctx->dstats;
    st = bzap_dist_stats{};
    st.world = G;
    st.rank = me;
    Xchg X;
    X.pt.on = getenv("BZAP_DIST_TIMING") != nullptr;
    X.pt.stream = ctx->stream;
    PhaseTable &pt = X.pt;
    X.ctx = ctx;
    X.N = N;
    X.comm = (ncclComm_t)ctx->comm;
    X.h = ctx->dist_host;
    Geometry &geo = X.geo;
    geo.G = G;
    geo.me = me;
    geo.n = n;
    u32 bits = 0;
    while (((u64)1 << bits) < n) ++bits;
    geo.shift = bits > 8 ? bits - 8 : 0;
    geo.nb = (u32)(((u64)n + ((u64)1 << geo.shift) - 1) >> geo.shift);
    geo.bpr = (geo.nb + G - 1) / G;
    const u32 lo = geo.lo(me), shard = geo.hi(me) - lo;
    u32 rshift = 1;
    while (rshift < 32 && (1ull << rshift) < n) ++rshift;            // ranks are < n <= 2^rshift
    const u32 pair_mask = (1u << ((2 * rshift + 7) / 8)) - 1u;

    // ---- 0. splitters + how many rotations are mine (first, small reservation) ---------------------------------
    const u32 S = DIST_SAMPLES_PER_RANK * (u32)G;
    // scratch of this phase lives in ctx->dist_dev (1 MiB: 64 B of collective words, counts, range counts, sample keys)
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    u64 bound_lo = 0, bound_hi = 0;
    int hi_open = 1;
    if (G > 1) {
        u64 *d_sk = (u64 *)(ctx->dist_dev + 512 * 1024);
        static_assert((size_t)DIST_SAMPLES_PER_RANK * DIST_MAX_WORLD * 8 <= 512 * 1024, "sample keys must fit their half of dist_dev");
        LAUNCH(ctx, dist_sample_keys_kernel, (S + 255) / 256, 256, 0, d_text, n, S, d_sk);
        std::vector<u64> sk(S);
        CU(ctx, cudaMemcpyAsync(ctx->dist_host, d_sk, (size_t)S * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(sk.data(), ctx->dist_host, (size_t)S * 8);
        std::sort(sk.begin(), sk.end());
        // bounds[g] = sk[g * S / G] for g = 1..G-1; rank g owns [bounds[g], bounds[g+1])
        bound_lo = me == 0 ? 0 : sk[((size_t)me * S) / G];
        if (me + 1 < G) { bound_hi = sk[((size_t)(me + 1) * S) / G]; hi_open = 0; }
    }
    u32 *d_cnt = (u32 *)(ctx->dist_dev + 1024);
    CU(ctx, cudaMemsetAsync(d_cnt, 0, 8 * sizeof(u32), ctx->stream));
    const u32 text_tiles = (n + DB_TILE - 1) / DB_TILE;
    const u32 sel_grid = grid_1d(text_tiles, 1, 148 * 8);
    u32 base = 0, m = n;
    std::vector<u32> range_cnt(sel_grid);
    {
        u32 *d_rc = (u32 *)(ctx->dist_dev + 4096);
        static_assert(148 * 8 * 4 + 4096 <= 512 * 1024, "range counts must fit dist_dev");
        LAUNCH(ctx, dist_owner_count_kernel, sel_grid, DB_BLOCK, 0, d_text, n, bound_lo, bound_hi, hi_open, text_tiles, d_cnt, d_rc);
        CU(ctx, cudaMemcpyAsync(ctx->dist_host, d_cnt, 2 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->dist_host + 64, d_rc, (size_t)sel_grid * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        base = ((u32 *)ctx->dist_host)[0];
        m = ((u32 *)ctx->dist_host)[1];
        memcpy(range_cnt.data(), ctx->dist_host + 64, (size_t)sel_grid * 4);
    }
    st.own_rotations = m;

    // ---- reservation: everything below is carved from one arena ------------------------------------------------
    const size_t mm = (size_t)m + 64, sh = (size_t)shard + 64;
    const size_t ctl_words = 8 * 256 + 64 + rerank_ctl_words(m) + bucket_ctl_words(m > shard ? m : shard) + 1024 + (size_t)G * 1024 +
                             (148 * 6 + 8) + 4096;
    const size_t need = 16 * mm + 8 * mm /* sort keys, payloads */ + 4 * mm /* rs */ + 12 * mm /* V1..V3 */ +
                        16 * mm /* act_r1 next_r1 newr pos */ + 4 * mm /* r2 */ + 4 * sh /* rank_home */ + 8 * sh /* requests in */ +
                        4 * sh /* responses out */ + 16 * sh /* pairs in, regrouped */ + ctl_words * 4 + active_ctl_bytes(m, nullptr, nullptr) +
                        sort_scratch_bytes(m) + 4 * (mm + 1024) /* last column, mtf, piece */ +
                        (me == 0 ? 2 * bzap_compress_bound(n) : 0) /* file image, staged pieces */ + mtf_scratch_bytes(m) +
                        (size_t)m / 16 + (16u << 20);
    RET(dist_reserve(ctx, N, G, need));
    u64 *K0 = arena_get<u64>(ctx, mm), *K1 = arena_get<u64>(ctx, mm);
    u32 *SV0 = arena_get<u32>(ctx, mm), *SV1 = arena_get<u32>(ctx, mm);
    u32 *d_rs = arena_get<u32>(ctx, mm);
    u32 *V1 = arena_get<u32>(ctx, mm), *V2 = arena_get<u32>(ctx, mm), *V3 = arena_get<u32>(ctx, mm);
    u32 *act_r1 = arena_get<u32>(ctx, mm), *next_r1 = arena_get<u32>(ctx, mm), *newr = arena_get<u32>(ctx, mm),
        *posb = arena_get<u32>(ctx, mm);
    u32 *d_r2 = arena_get<u32>(ctx, mm);
    u32 *rank_home = arena_get<u32>(ctx, sh);
    u64 *req_in = arena_get<u64>(ctx, sh);
    u32 *resp_out = arena_get<u32>(ctx, sh);
    u32 *home_idx = arena_get<u32>(ctx, sh), *home_val = arena_get<u32>(ctx, sh);
    u32 *home_xi = arena_get<u32>(ctx, sh), *home_xv = arena_get<u32>(ctx, sh);
    u32 *d_hist8 = arena_get<u32>(ctx, 8 * 256 + 64);
    u32 *d_rrctl = arena_get<u32>(ctx, rerank_ctl_words(m));
    u32 *d_bctl = arena_get<u32>(ctx, bucket_ctl_words(m > shard ? m : shard));
    u32 *d_h256 = arena_get<u32>(ctx, 1024);
    X.d_small = arena_get<u32>(ctx, (size_t)G * 1024);
    u32 *d_bact = arena_get<u32>(ctx, 148 * 6 + 8);
    u32 *d_range_cnt = arena_get<u32>(ctx, sel_grid + 8);
    u8 *d_actctl = arena_get<u8>(ctx, active_ctl_bytes(m, nullptr, nullptr));
    d_cnt = arena_get<u32>(ctx, 16);
    if (!K0 || !K1 || !SV0 || !SV1 || !d_rs || !V1 || !V2 || !V3 || !act_r1 || !next_r1 || !newr || !posb || !d_r2 || !rank_home ||
        !req_in || !resp_out || !home_idx || !home_val || !home_xi || !home_xv || !d_hist8 || !d_rrctl || !d_bctl || !d_h256 || !X.d_small || !d_bact ||
        !d_range_cnt || !d_actctl || !d_cnt)
        return bzap_fail(ctx, BZAP_ERR_NOMEM, "distributed block scratch");
    const size_t arena_mark = ctx->arena_off;
    {
        const u64 offs[4] = {(u64)((u8 *)home_idx - ctx->arena), (u64)((u8 *)home_val - ctx->arena), (u64)((u8 *)req_in - ctx->arena),
                             (u64)((u8 *)d_r2 - ctx->arena)};
        CU(ctx, cudaMemsetAsync(X.d_small + (size_t)G * 1024 - 8, 0, 8 * sizeof(u32), ctx->stream));
        RET(peers_setup(X, offs));
        st.peer_windows = X.p2p ? 1u : 0u;
    }

    // ---- 1. select + sort + sparse ranks -------------------------------------------------------------------------
    u64 *keys = nullptr;
    u32 *sa = nullptr;
    CU(ctx, cudaMemsetAsync(d_hist8, 0, (8 * 256 + 64) * sizeof(u32), ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_range_cnt, range_cnt.data(), (size_t)sel_grid * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_rrctl, 0, rerank_ctl_words(m) * sizeof(u32), ctx->stream));
    pt.start();
    if (m) {
        LAUNCH(ctx, dist_select_kernel, sel_grid, DB_BLOCK, 0, d_text, n, bound_lo, bound_hi, hi_open, text_tiles, d_range_cnt, K0, SV0,
               d_hist8);
        pt.stop("first.select");
        SortBuffers sb;
        sb.keys[0] = K0; sb.keys[1] = K1; sb.vals[0] = SV0; sb.vals[1] = SV1;
        int passes = 0;
        RET(dev_sort_pairs64(ctx, &sb, m, 0xffu, d_hist8, 8, false, &keys, &sa, &passes));
        ctx->arena_off = arena_mark;
        pt.stop("first.sort");
        RET(dev_rerank_sorted(ctx, keys, m, base, d_rs, d_rrctl, d_bact));
        pt.stop("first.rerank");
    } else {
        sa = SV0;
    }
    u32 *V0 = sa == SV0 ? SV1 : SV0;                                 // the payload buffer the suffix array does not live in
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    // ---- 2. ranks go home ------------------------------------------------------------------------------------------
    HomeBufs HB;
    HB.r_idx = home_idx; HB.r_val = home_val; HB.x_idx = home_xi; HB.x_val = home_xv; HB.d_hist = d_h256; HB.d_bctl = d_bctl;
    HB.b_idx = V1; HB.b_val = V2;
    RET(ranks_go_home(X, sa, d_rs, m, HB, rank_home));
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));

    // ---- 3. rounds ---------------------------------------------------------------------------------------------------
    // {groups, singleton groups} of the first sort -> who is still unsettled
    u32 h_loc[4] = {0, 0, 0, 0}, h_sum[4] = {0, 0, 0, 0};
    CU(ctx, cudaMemsetAsync(d_cnt, 0, 4 * sizeof(u32), ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_cnt, d_rrctl + 4 * 256, 2 * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    RET(reduce_words(X, d_cnt, 4, h_loc, h_sum));
    u32 M = m - h_loc[1];
    u64 total_active = (u64)n - h_sum[1];            // every rotation is owned by exactly one rank
    pt.start();
    if (M) RET(dev_collect_active(ctx, d_rs, sa, m, base, d_bact, V0, act_r1));
    pt.stop("first.collect");
    ActiveWork w;
    active_ctl_bytes(m, &w, d_actctl);
    w.newr = newr;
    w.pos = posb;
    w.sa_buf = sa;
    w.d_rank = nullptr;
    w.arena_mark = arena_mark;
    w.rank_mask = 0;
    u64 k = 8;
    u32 rounds = 1;
    u32 *act_idx = V0, *next_idx = V3;
    while (total_active > 0 && k < n) {
        const u32 kk = (u32)(k % n);
        CU(ctx, cudaMemsetAsync(w.zero_base, 0, w.zero_bytes, ctx->stream));
        CU(ctx, cudaMemsetAsync(d_h256, 0, 256 * sizeof(u32), ctx->stream));
        // pull r2 = rank[(start + k) mod n]
        pt.start();
        if (M) {
            LAUNCH(ctx, dist_request_kernel, grid_1d(M, 256 * 4), 256, 0, act_idx, act_r1, M, n, kk, geo.shift, K0, d_h256);
            if (G > 1) RET(dev_bucket_pass_u64(ctx, K0, act_idx, M, 32 + (int)geo.shift, K1, V1, d_h256, d_bctl));
        }
        pt.stop("pull.request+bucket");
        u64 *bk = G > 1 ? K1 : K0;                   // requests grouped by owner (one owner: nothing to regroup)
        u32 *bidx = G > 1 ? V1 : act_idx;
        u64 total_req = M;
        if (G > 1) {
            u32 *h_all = nullptr;
            RET(gather_words(X, d_h256, 256, &h_all));
            pt.stop("pull.counts");
            CountMatrix C;
            matrix_from_hists(geo, h_all, &C);
            total_req = 0;
            for (int s = 0; s < G; ++s) total_req += C.c[s][me];
            if (total_req > shard) return bzap_fail(ctx, BZAP_ERR_CUDA, "pull: %llu requests for a shard of %u", (unsigned long long)total_req, shard);
            // only the positions travel (4 of the 8 request bytes); d_r2 is free until the answers arrive
            if (M) LAUNCH(ctx, dist_request_pos_kernel, grid_1d(M, 256 * 4), 256, 0, bk, M, d_r2);
            RET(alltoallv(X, d_r2, req_in, sizeof(u32), C, false, BUF_REQ_IN));
            pt.stop("pull.alltoall requests");
            if (total_req) LAUNCH(ctx, dist_respond_pos_kernel, grid_1d(total_req, 256 * 4), 256, 0, (const u32 *)req_in, (u32)total_req, rank_home, lo, resp_out);
            pt.stop("pull.respond");
            RET(alltoallv(X, resp_out, d_r2, sizeof(u32), C, true, BUF_R2));
            pt.stop("pull.alltoall responses");
        } else if (M) {
            LAUNCH(ctx, dist_respond_kernel, grid_1d(M, 256 * 4), 256, 0, bk, M, rank_home, lo, d_r2);
            pt.stop("pull.respond");
        }
        // local sort by (r1, r2), new sparse ranks inside every group's slot range, survivors
        u64 *skeys = nullptr;
        u32 *sidx = nullptr;
        if (M) {
            u64 *kin = bk == K0 ? K1 : K0;           // the key buffer the requests do NOT live in
            LAUNCH(ctx, dist_pair_keys_kernel, grid_1d(M, 256 * 4), 256, 0, bk, d_r2, M, rshift, kin, w.d_hist8);
            SortBuffers ab;
            ab.keys[0] = kin; ab.keys[1] = bk;
            ab.vals[0] = bidx; ab.vals[1] = V2;
            int passes = 0;
            RET(bwt_active_sort_rerank(ctx, w, &ab, M, rshift, pair_mask, base, nullptr, next_idx, next_r1, &skeys, &sidx, &passes));
            ctx->arena_off = arena_mark;
        }
        pt.stop("round.sort+rerank");
        // new ranks go home (all of this round's rotations: the unchanged ones are rewritten with the same value)
        HB.b_idx = G > 1 ? act_idx : V1;             // dead buffers of this round
        HB.b_val = posb;
        RET(ranks_go_home(X, sidx, newr, M, HB, rank_home));
        ++rounds;
        k *= 2;
        pt.start();
        RET(reduce_words(X, w.d_counters, 4, h_loc, h_sum));
        pt.stop("round.termination");
        const u64 groups = h_sum[0], subs = h_sum[1];
        total_active = h_sum[2];
        M = h_loc[2];
        { u32 *t = act_idx; act_idx = next_idx; next_idx = t; }
        { u32 *t = act_r1; act_r1 = next_r1; next_r1 = t; }
        if (subs == groups) break;                   // fixed point: no group was split
    }
    st.rounds = rounds;
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));

    // ---- 4. last column, primary index ------------------------------------------------------------------------------
    u8 *d_last = arena_get<u8>(ctx, mm + 64), *d_mtf = arena_get<u8>(ctx, mm + 64);
    u64 *d_stat = arena_get<u64>(ctx, 256 + 128 + 8);
    u32 *d_init = arena_get<u32>(ctx, 256);
    if (!d_last || !d_mtf || !d_stat || !d_init) return bzap_fail(ctx, BZAP_ERR_NOMEM, "tail scratch");
    pt.start();
    if (m) RET(dev_gather_slots(ctx, d_text, sa, n, m, d_last));
    pt.stop("tail.last column");
    u64 primary = 0;
    {
        u32 *d_p = d_cnt + 8;
        if (me == 0) CU(ctx, cudaMemcpyAsync(d_p, rank_home, sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));   // rank 0 owns position 0
        if (G > 1) NC(ctx, N->Broadcast(d_p, d_p, 1, ncclUint32, 0, X.comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->dist_host, d_p, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        primary = *(u32 *)ctx->dist_host;
    }

    // ---- 5. MTF: start list of this piece from the summaries of the pieces before it ----------------------------------
    MtfPlan plan;
    RET(dev_mtf_begin(ctx, d_last, m, &plan));
    const u32 *d_init_use = nullptr;
    if (G > 1) {
        u32 *h_tot = nullptr;
        RET(gather_words(X, plan.d_total, 256, &h_tot));
        if (me > 0) {
            int list[256], nlist = 256;
            bool seen_any[256];
            for (int s = 0; s < 256; ++s) { list[s] = s; seen_any[s] = false; }
            for (int r = 0; r < me; ++r) {
                const u32 *t = h_tot + (size_t)r * 256;
                int seen[256], ns = 0;
                for (int s = 0; s < 256; ++s) if (t[s]) { seen[ns++] = s; seen_any[s] = true; }
                std::sort(seen, seen + ns, [&](int a, int b) { return t[a] > t[b]; });     // most recent first (keys are distinct)
                int merged[256], q = 0;
                for (int i = 0; i < ns; ++i) merged[q++] = seen[i];
                for (int i = 0; i < nlist; ++i) if (!t[list[i]]) merged[q++] = list[i];
                memcpy(list, merged, sizeof list);
            }
            u32 *h_init = (u32 *)(ctx->dist_host + 128 * 1024);
            for (int s = 0; s < 256; ++s) h_init[s] = 0;
            for (int i = 0; i < 256; ++i) if (seen_any[list[i]]) h_init[list[i]] = 256u - (u32)i;
            CU(ctx, cudaMemcpyAsync(d_init, h_init, 256 * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
            d_init_use = d_init;
        }
    }
    RET(dev_mtf_finish(ctx, d_last, &plan, d_init_use, d_mtf));
    pt.stop("tail.mtf");

    // ---- 6. Huffman: global statistics, the tree everywhere, pieces at their bit offsets ------------------------------
    RET(dev_hist_launch(ctx, d_mtf, m, d_stat));
    const u32 stat_words = 256 * 2 + 256;
    u32 *h_stat = nullptr;
    RET(gather_words(X, (const u32 *)d_stat, stat_words, &h_stat));
    u64 gfreq[256];
    u64 rfreq[DIST_MAX_WORLD][256];
    u64 first_key[256];                              // (rank << 32) | first position inside that rank's piece
    for (int s = 0; s < 256; ++s) { gfreq[s] = 0; first_key[s] = ~0ull; }
    for (int r = 0; r < G; ++r) {
        const u64 *f = (const u64 *)(h_stat + (size_t)r * stat_words);
        const u32 *fp = h_stat + (size_t)r * stat_words + 512;
        for (int s = 0; s < 256; ++s) {
            rfreq[r][s] = f[s];
            gfreq[s] += f[s];
            if (f[s] && first_key[s] == ~0ull) first_key[s] = ((u64)r << 32) | fp[s];
        }
    }
    u8 order[256];
    int n_leaves = 0;
    for (int s = 0; s < 256; ++s) if (gfreq[s]) order[n_leaves++] = (u8)s;
    std::sort(order, order + n_leaves, [&](u8 a, u8 b) { return first_key[a] < first_key[b]; });   // first appearance, main.cpp:238-244
    bzap_tree tree;
    CodeTable ct;
    u8 head_bytes[BZAP_HEADER_BYTES + BZAP_MAX_TREE_BYTES];
    size_t tb = 0;
    int rc = huff_build_tree(gfreq, order, n_leaves, &tree);
    if (rc == BZAP_OK) rc = huff_code_table(&tree, &ct);
    if (rc == BZAP_OK) rc = huff_tree_to_bytes(&tree, head_bytes + BZAP_HEADER_BYTES, &tb);
    if (rc != BZAP_OK) return bzap_fail(ctx, rc, "huffman model");
    const size_t head = BZAP_HEADER_BYTES + tb;
    u64 bits_of[DIST_MAX_WORLD], bit_off[DIST_MAX_WORLD], total_bits = 0;
    for (int r = 0; r < G; ++r) {
        bits_of[r] = huff_total_bits(rfreq[r], &ct);
        bit_off[r] = 8ull * head + total_bits;
        total_bits += bits_of[r];
    }
    const size_t payload = total_bits ? (size_t)((total_bits + 7) / 8) : 1;     // max(1, ceil(bits/8)), main.cpp:162
    const size_t file_len = head + payload;
    auto piece_first = [&](int r) { return (size_t)((bit_off[r] / 8) & ~15ull); };
    auto piece_bytes = [&](int r) { return bits_of[r] ? (size_t)((bit_off[r] - 8ull * piece_first(r) + bits_of[r] + 7) / 8) : (size_t)0; };
    if (me == 0) {
        if (file_len > out_cap) return bzap_fail(ctx, BZAP_ERR_CAPACITY, "need %zu bytes, have %zu", file_len, out_cap);
        u8 *d_file = arena_get<u8>(ctx, file_len + 128);
        size_t stage_bytes = 0;
        for (int r = 1; r < G; ++r) stage_bytes += (piece_bytes(r) + 15) & ~(size_t)15;
        u8 *d_stage = arena_get<u8>(ctx, stage_bytes + 64);
        if (!d_file || !d_stage) return bzap_fail(ctx, BZAP_ERR_NOMEM, "file image");
        put_u64le(head_bytes, primary);              // io_utilities.h:17
        put_u64le(head_bytes + 8, n);                // io_utilities.h:18
        put_u64le(head_bytes + 16, tb);              // io_utilities.h:19
        u8 *h_head = ctx->dist_host + 192 * 1024;
        memcpy(h_head, head_bytes, head);
        CU(ctx, cudaMemsetAsync(d_file, 0, (file_len + 63) & ~(size_t)31, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_file, h_head, head, cudaMemcpyHostToDevice, ctx->stream));
        RET(dev_huff_encode(ctx, d_mtf, m, &ct, d_file, bit_off[0]));
        if (G > 1) {
            NC(ctx, N->GroupStart());
            size_t off = 0;
            for (int r = 1; r < G; ++r) {
                const size_t pb = piece_bytes(r);
                if (pb) NC(ctx, N->Recv(d_stage + off, pb, ncclUint8, r, X.comm, ctx->stream));
                off += (pb + 15) & ~(size_t)15;
            }
            NC(ctx, N->GroupEnd());
            off = 0;
            for (int r = 1; r < G; ++r) {
                const size_t pb = piece_bytes(r);
                if (pb) LAUNCH(ctx, dist_or_merge_kernel, grid_1d(pb / 16 + 1, 256), 256, 0, d_file + piece_first(r), d_stage + off, pb);
                off += (pb + 15) & ~(size_t)15;
            }
        }
        CU(ctx, cudaMemcpyAsync(d_out, d_file, file_len, cudaMemcpyDeviceToDevice, ctx->stream));
        *out_len = file_len;
    } else {
        const size_t pb = piece_bytes(me);
        u8 *d_piece = arena_get<u8>(ctx, pb + 128);
        if (!d_piece) return bzap_fail(ctx, BZAP_ERR_NOMEM, "payload piece");
        CU(ctx, cudaMemsetAsync(d_piece, 0, (pb + 95) & ~(size_t)31, ctx->stream));
        RET(dev_huff_encode(ctx, d_mtf, m, &ct, d_piece, bit_off[me] - 8ull * piece_first(me)));
        if (pb) {
            NC(ctx, N->Send(d_piece, pb, ncclUint8, 0, X.comm, ctx->stream));
            X.sent_bytes += pb;
        }
    }
    pt.stop("tail.huffman+gather");
    CU(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    st.exchanged_bytes = X.sent_bytes;
    st.ms_total = ev_ms2(ctx->ev[0], ctx->ev[7]);
    st.ms_select_sort = ev_ms2(ctx->ev[0], ctx->ev[1]);
    st.ms_home = ev_ms2(ctx->ev[1], ctx->ev[2]);
    st.ms_pull = pt.get("pull.request+bucket") + pt.get("pull.counts") + pt.get("pull.alltoall requests") + pt.get("pull.respond") +
                 pt.get("pull.alltoall responses");
    st.ms_round_sort = pt.get("round.sort+rerank");
    if (pt.on && me == 0) {
        fprintf(stderr, "[bzap dist] world %d n %u rounds %u total %.2f ms (with profiling synchronisations)\n", G, n, rounds, st.ms_total);
        for (int i = 0; i < pt.count; ++i) fprintf(stderr, "[bzap dist]   %-28s %9.2f ms\n", pt.names[i], pt.ms[i]);
    }
    st.ms_rounds = ev_ms2(ctx->ev[2], ctx->ev[3]);
    st.ms_tail = ev_ms2(ctx->ev[3], ctx->ev[7]);
    ctx->stats.bwt_rounds = rounds;
    ctx->stats.payload_bytes = payload;
    return BZAP_OK;
}
