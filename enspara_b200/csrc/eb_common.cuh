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

// Row parts of the TMA-staged step kernel: a coordinate row of A_pad = 32 m floats is cut into
// P parts of whole 128-byte lines, each at most 256 floats (the TMA box limit) and at least 160
// floats (640 bytes: smaller stages lose to the LDG kernel -- measured at 1.25M frames, rows of
// 288 floats cut into 3 x 96: TMA 0.846 ms, LDG 0.806 ms; rows of 384 = 2 x 192: TMA 0.891 ms,
// LDG 0.994 ms).  Returns the smallest such P in 1..4, or 0 when there is none.
__host__ __device__ inline int rmsd_tma_parts(int A_pad)
{
    if (A_pad <= 0 || (A_pad & 31)) return 0;
    const int m = A_pad >> 5;
    for (int P = 1; P <= 4; ++P)
        if (m % P == 0 && m / P <= 8 && m / P >= 5) return P;
    return 0;
}

// Padded atom count.  Rows are always 32-byte (sector) aligned.  When a multiple of 32 atoms
// that the TMA-staged kernel can cut into parts is within 12 % it is taken (large shards then
// stream at 88-103 % of the copy peak instead of the LDG kernel's 75-81 %: 350 -> 384 atoms
// 80.9 % -> 90.2 % algorithmic, 240 -> 256 81.6 %); otherwise any multiple of 32 within 3 %
// (whole 128-byte lines per 8-lane request); otherwise a multiple of 8 -- e.g. 264 atoms stay
// 264: padded to 288 (3 x 96 floats) both kernels were slower than the unpadded LDG kernel
// (75 % / 72 % vs 78 %).  Padding atoms are zeros: they change no sum, so results do not
// depend on A_pad.
__host__ __device__ inline int rmsd_apad(int n_atoms)
{
    const int a8 = (n_atoms + 7) & ~7;
    const int a32 = (n_atoms + 31) & ~31;
    for (int a = a32; a * 100 <= n_atoms * 112; a += 32)
        if (rmsd_tma_parts(a)) return a;
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

// ------------------------------------------------------------------------------------------
// Candidate exchange over peer memory (NVLink P2P), fused into the step kernel.
//
// The reference exchanges the new centre with two allgathers + a Bcast + a Barrier + an
// allreduce per iteration (cluster/kcenters.py:332-348, mpi/ops.py:137-138).  Here every rank
// owns one symmetric buffer (same layout on every GPU, mapped into every peer):
//     [  0) u64 seq                  number of records this rank has published so far
//     [128) u64 flags[2][8]          flags[p][r] = seq of the last record rank r put in parity p
//     [256) records[2][size][stride] candidate records, double buffered by seq parity
// The last block of a step writes its shard's candidate record straight into slot `rank` of
// EVERY peer's buffer (plain stores over NVLink), fences at system scope and then releases
// flags[p][rank] = seq on every peer.  The next step's prologue acquires the `size` flags of
// its own buffer and reads the records locally: no collective launch, no host involvement.
// Double buffering is enough: a rank can be at most one publish ahead of any peer, because
// publishing record q+1 requires having consumed every peer's record q.
// ------------------------------------------------------------------------------------------
struct Exch {
    const long long *peers;   // device array: base address of every rank's buffer (nullptr: off)
    int size;
    int rank;
    unsigned rec_stride;      // bytes between records (>= record bytes, multiple of 16)
    unsigned long long timeout_ns;  // bounded spin: give up (state->error = 1) after this long
};
constexpr int kExchMaxRanks = 8;
constexpr int kExchFlagsOff = 128;
constexpr int kExchRecordsOff = 256;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Consumer side (all threads of the block): wait until every rank's record of the current
// sequence number has arrived in THIS rank's buffer; returns the local record array.
// The spin is bounded: if a peer's flag does not arrive within e.timeout_ns (a rank died or
// queued a different launch sequence) every thread of the block gets nullptr, and the caller
// marks the run failed instead of hanging the GPU.  `wait_ns` (thread 0) receives the time
// spent waiting, for the per-step breakdown in eb_kc_state.
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ const unsigned char *exch_wait(const Exch &e,
                                                          unsigned long long *wait_ns = nullptr)
{
    unsigned char *me = reinterpret_cast<unsigned char *>(e.peers[e.rank]);
    const unsigned long long q = *reinterpret_cast<volatile unsigned long long *>(me);
    const int p = (int)(q & 1ull);
    int bad = 0;
    unsigned long long t_in = 0;
    if (wait_ns && threadIdx.x == 0) t_in = globaltimer_ns();
    if ((int)threadIdx.x < e.size) {
        const unsigned long long *flag =
            reinterpret_cast<const unsigned long long *>(me + kExchFlagsOff) +
            p * kExchMaxRanks + threadIdx.x;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_sys(flag) != q) {
            if ((++spins & 0x3ffu) == 0u) {
                const unsigned long long t = globaltimer_ns();
                if (t0 == 0) {
                    t0 = t;
                } else if (t - t0 > e.timeout_ns) {
                    bad = 1;
                    break;
                }
            }
        }
    }
    if (__syncthreads_or(bad)) return nullptr;
    if (wait_ns && threadIdx.x == 0) *wait_ns = globaltimer_ns() - t_in;
    return me + kExchRecordsOff + (size_t)p * e.size * e.rec_stride;
}

// Publisher side (all threads of ONE block): `rec` (rec_bytes, a multiple of 16, in local global
// memory, written by this block) goes to slot `rank` of every peer, then the flags are released.
__device__ __forceinline__ void exch_publish(const Exch &e, const unsigned char *rec,
                                             size_t rec_bytes)
{
    unsigned char *me = reinterpret_cast<unsigned char *>(e.peers[e.rank]);
    const unsigned long long q = *reinterpret_cast<volatile unsigned long long *>(me) + 1ull;
    const int p = (int)(q & 1ull);
    __syncthreads();   // the record is complete
    const int n4 = (int)(rec_bytes >> 4);
    const float4 *src = reinterpret_cast<const float4 *>(rec);
    for (int d = 0; d < e.size; ++d) {
        float4 *dst = reinterpret_cast<float4 *>(
            reinterpret_cast<unsigned char *>(e.peers[d]) + kExchRecordsOff +
            ((size_t)p * e.size + e.rank) * e.rec_stride);
        for (int t = threadIdx.x; t < n4; t += blockDim.x) dst[t] = __ldcg(src + t);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(me) = q;
    if ((int)threadIdx.x < e.size) {
        unsigned long long *flag =
            reinterpret_cast<unsigned long long *>(
                reinterpret_cast<unsigned char *>(e.peers[threadIdx.x]) + kExchFlagsOff) +
            p * kExchMaxRanks + e.rank;
        st_release_sys(flag, q);
    }
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
