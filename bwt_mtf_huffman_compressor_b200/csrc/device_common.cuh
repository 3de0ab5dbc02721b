// device_common.cuh -- warp/block primitives shared by the kernels: relaxed status-word access,
// single-pass "decoupled look-back" prefix over tiles, block scans, tile tickets.
#pragma once
#include "bzap_internal.h"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Status words carry flag and value in ONE naturally aligned word, so a relaxed load observes
// both atomically and no fence is needed between "value" and "flag".
__device__ __forceinline__ u32 ld_relaxed(const u32 *p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u32 *p, u32 v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Tiles are handed out by an atomic ticket, not blockIdx: a tile can only wait on tiles whose
// blocks already run, which makes the look-back spin deadlock free under any block schedule.
__device__ __forceinline__ u32 take_ticket(u32 *counter, u32 *s_slot)
{
    if (threadIdx.x == 0) *s_slot = atomicAdd(counter, 1u);
    __syncthreads();
    u32 t = *s_slot;
    __syncthreads();
    return t;
}

struct OpSum {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a + b; }
    static __device__ __forceinline__ u64 identity() { return 0; }
};
struct OpMax {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; }
    static __device__ __forceinline__ u64 identity() { return 0; }
};

#define LB_FLAG_SHIFT 62
#define LB_AGG (1ull << LB_FLAG_SHIFT)
#define LB_INCL (2ull << LB_FLAG_SHIFT)
#define LB_VALUE_MASK ((1ull << LB_FLAG_SHIFT) - 1)

// Exclusive prefix of `aggregate` over tiles 0..tile-1 (commutative Op).  Called by all 32 lanes
// of ONE warp with the same arguments; status[] must be zero before the kernel starts.
template <class Op>
__device__ __forceinline__ u64 lookback_exclusive(u64 *status, u32 tile, u64 aggregate, Op op)
{
    const u32 lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed(&status[0], LB_INCL | aggregate);
        return Op::identity();
    }
    if (lane == 0) st_relaxed(&status[tile], LB_AGG | aggregate);
    u64 excl = Op::identity();
    int base = (int)tile - 1;
    while (true) {
        int t = base - (int)lane;
        u64 s = LB_INCL;   // virtual tile before tile 0: inclusive identity
        if (t >= 0) {
            do { s = ld_relaxed(&status[t]); } while ((s >> LB_FLAG_SHIFT) == 0);
        }
        u32 incl_mask = __ballot_sync(FULL_MASK, (s >> LB_FLAG_SHIFT) == 2);
        u32 first = incl_mask ? (u32)(__ffs(incl_mask) - 1) : 32u;
        u64 v = (lane <= first) ? (s & LB_VALUE_MASK) : Op::identity();
        if (t < 0) v = Op::identity();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL_MASK, v, o));
        excl = op(excl, v);
        if (incl_mask) break;
        base -= 32;
    }
    if (lane == 0) st_relaxed(&status[tile], LB_INCL | op(excl, aggregate));
    return excl;
}

// Block-wide exclusive scan of one u32 per thread.  s_tmp needs (blockDim/32 + 1) u32.
// Returns the exclusive prefix; *total receives the block sum (same for all threads).
__device__ __forceinline__ u32 block_exclusive_sum(u32 v, u32 *s_tmp, u32 *total)
{
    const u32 lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nwarps ? s_tmp[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(FULL_MASK, wi, o);
            if (lane >= (u32)o) wi += t;
        }
        if (lane < nwarps) s_tmp[lane] = wi - w;
        if (lane == 31) s_tmp[32] = wi;
    }
    __syncthreads();
    u32 r = s_tmp[warp] + incl - v;
    *total = s_tmp[32];
    __syncthreads();
    return r;
}

// Block-wide exclusive max-scan of one u32 per thread (identity 0).  s_tmp: 33 u32.
__device__ __forceinline__ u32 block_exclusive_max(u32 v, u32 *s_tmp, u32 *total)
{
    const u32 lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= (u32)o) incl = max(incl, t);
    }
    u32 excl = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane == 0) excl = 0;
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nwarps ? s_tmp[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(FULL_MASK, wi, o);
            if (lane >= (u32)o) wi = max(wi, t);
        }
        u32 we = __shfl_up_sync(FULL_MASK, wi, 1);
        if (lane == 0) we = 0;
        if (lane < nwarps) s_tmp[lane] = we;
        if (lane == 31) s_tmp[32] = wi;
    }
    __syncthreads();
    u32 r = max(s_tmp[warp], excl);
    *total = s_tmp[32];
    __syncthreads();
    return r;
}
