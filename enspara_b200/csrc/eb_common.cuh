// Shared device/host helpers for libenspara_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/enspara_b200.h"

namespace eb {

// ------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns a status; text is kept per thread
// ------------------------------------------------------------------------------------------
extern thread_local char g_err[512];

inline int fail(int code, const char *fmt, const char *a = "", long b = 0, long c = 0)
{
    snprintf(g_err, sizeof g_err, fmt, a, b, c);
    return code;
}

#define EB_CHECK_ARG(cond, msg)                                                                \
    do {                                                                                       \
        if (!(cond)) return eb::fail(EB_ERR_INVALID, "%s", msg);                               \
    } while (0)

#define EB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return eb::fail(EB_ERR_CUDA, "CUDA error: %s (at line %ld)",                       \
                            cudaGetErrorString(e__), (long)__LINE__);                          \
    } while (0)

#define EB_LAUNCH_CHECK() EB_CUDA(cudaGetLastError())

int sm_count();

// ------------------------------------------------------------------------------------------
// candidate record (see include/enspara_b200.h)
// ------------------------------------------------------------------------------------------
struct RecHeader {
    double dist;
    int64_t index;
    double trace;
    int64_t reserved;
};
static_assert(sizeof(RecHeader) == 32, "record header is 32 bytes");

// Padded atom count.  Rows are always 32-byte (sector) aligned; when rounding up to a multiple
// of 32 atoms costs at most 3% extra bytes the rows become 128-byte (cache line) aligned, which
// makes every 8-lane request exactly one line (otherwise L1 pulls both straddled lines from L2).
__host__ __device__ inline int rmsd_apad(int n_atoms)
{
    const int a8 = (n_atoms + 7) & ~7;
    const int a32 = (n_atoms + 31) & ~31;
    return (a32 * 100 <= n_atoms * 103) ? a32 : a8;
}
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// per-block arg-max partial
struct Partial {
    double dist;
    int64_t index;
};

constexpr int kMaxGrid = 4096;  // upper bound on blocks of any step kernel (partials size)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// streaming 128-bit load: read-only path, do not allocate in L1 (each frame byte is used once)
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// same, asking L2 to fetch the 256-byte pair of lines (sequential streams use the neighbour next)
__device__ __forceinline__ float4 ldg_stream_256(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ double shfl_xor_d(double v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// (dist, index) ordering of the reference's np.argmax: larger dist wins, ties -> lower index.
__device__ __forceinline__ bool better(double d, int64_t i, double bd, int64_t bi)
{
    return (d > bd) || (d == bd && i < bi);
}

__device__ __forceinline__ void warp_argmax(double &d, int64_t &i)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, m);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, i, m);
        if (better(od, oi, d, i)) {
            d = od;
            i = oi;
        }
    }
}

// Block-wide arg-max; result valid in thread 0.  `sh` needs 32 Partial entries.
__device__ __forceinline__ void block_argmax(double &d, int64_t &i, Partial *sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (blockDim.x + 31) >> 5;
    warp_argmax(d, i);
    if (lane == 0) {
        sh[warp].dist = d;
        sh[warp].index = i;
    }
    __syncthreads();
    if (warp == 0) {
        d = (lane < n_warps) ? sh[lane].dist : -2.0;
        i = (lane < n_warps) ? sh[lane].index : INT64_MAX;
        warp_argmax(d, i);
    }
    __syncthreads();
}

// Prologue shared by all step kernels: choose the winning candidate record.
// Every thread gets the same answer.  Records with index < 0 are empty shards.
__device__ __forceinline__ int pick_candidate(const unsigned char *cand, int n_cand,
                                              size_t rec_bytes, double &best_d, int64_t &best_i)
{
    int best_r = -1;
    best_d = -1.0;
    best_i = INT64_MAX;
    for (int r = 0; r < n_cand; ++r) {
        const RecHeader *h = reinterpret_cast<const RecHeader *>(cand + (size_t)r * rec_bytes);
        const double d = __ldcg(&h->dist);
        const int64_t i = __ldcg(reinterpret_cast<const long long *>(&h->index));
        if (i >= 0 && (best_r < 0 || better(d, i, best_d, best_i))) {
            best_r = r;
            best_d = d;
            best_i = i;
        }
    }
    return best_r;
}

// Stop rule of kcenters.py:217.  Returns true when the step must run.
__device__ __forceinline__ bool step_active(const eb_kc_state *st, int n_clusters_limit,
                                            double maxdist, double cutoff, int &k)
{
    k = *reinterpret_cast<const volatile int32_t *>(&st->n_centers);
    const int done = *reinterpret_cast<const volatile int32_t *>(&st->done);
    if (done) return false;
    return (k < n_clusters_limit) && (maxdist > cutoff);
}

// Epilogue shared by all step kernels.  Each block contributes its (dist,index) arg-max (valid
// in thread 0); the last block to arrive reduces all partials and returns true in ALL of its
// threads with (d,i) = shard arg-max (local index).  Other blocks return false.
__device__ __forceinline__ bool grid_argmax_last_block(double &d, int64_t &i, Partial *partials,
                                                       eb_kc_state *st, Partial *sh,
                                                       int *sh_flag)
{
    block_argmax(d, i, sh);
    if (threadIdx.x == 0) {
        partials[blockIdx.x].dist = d;
        partials[blockIdx.x].index = i;
        __threadfence();
        const unsigned ticket = atomicAdd(&st->blocks_done, 1u);
        *sh_flag = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!*sh_flag) return false;
    __threadfence();
    d = -2.0;
    i = INT64_MAX;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        const double od = __ldcg(&partials[b].dist);
        const int64_t oi = __ldcg(reinterpret_cast<const long long *>(&partials[b].index));
        if (better(od, oi, d, i)) {
            d = od;
            i = oi;
        }
    }
    block_argmax(d, i, sh);
    if (threadIdx.x == 0) {
        sh[0].dist = d;
        sh[0].index = i;
    }
    __syncthreads();
    d = sh[0].dist;
    i = sh[0].index;
    __syncthreads();
    return true;
}

#endif  // __CUDACC__

}  // namespace eb
