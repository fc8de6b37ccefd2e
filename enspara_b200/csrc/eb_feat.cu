// K2: feature-vector distances (euclidean / manhattan / squared euclidean) replacing enspara's
// Cython libdist, fused with the k-centers bookkeeping, plus the many-centres assignment.
//
// Reference behaviour reproduced bit for bit (paths under /root/reference/enspara/):
//   geometry/libdist.pyx:122-145  _euclidean: out[i] += (X[i,j]-y[j])**2 in j order, then sqrt
//   geometry/libdist.pyx:100-117  _manhattan: out[i] += fabs(X[i,j]-y[j]) in j order
//   (generated C: the difference and the square are computed in the array's own C type --
//    float for float32, promoted int for int8/16/32, long for int64 -- and added to a float64
//    accumulator; gcc folds powf(d, 2) to d*d.)
//   cluster/kcenters.py:243-311, 314-378, 217   the iteration and stop rule, as in K1
//   cluster/util.py:159-205                      assign_to_nearest_center
//
// The float64 accumulation must keep the reference's j order to be bit-identical, so one
// lane owns one row and walks it sequentially; coalescing comes from staging: a warp pulls a
// tile of 32 rows x 256 bytes into shared memory with 16-byte vector loads (two rows per
// request), padded so the per-lane 16-byte reads are bank-conflict free.  HBM bound: each row
// byte is read once; dist (float64) is a coalesced read-modify-write.
#include <stdlib.h>

#include "eb_tma.cuh"

namespace eb {

constexpr int kFeatThreads = 128;
constexpr int kFeatWarps = kFeatThreads / 32;
constexpr int kTileBytes = 256;                 // bytes of one row held per tile
constexpr int kTileStride = kTileBytes + 16;    // 17 x 16 B: odd -> conflict-free LDS.128

template <typename T> struct Wide { using type = T; };
template <> struct Wide<int8_t> { using type = int; };
template <> struct Wide<int16_t> { using type = int; };
template <> struct Wide<int32_t> { using type = int; };
template <> struct Wide<int64_t> { using type = long long; };

// one term of the row sum, exactly as the reference's generated C computes it
template <int METRIC> __device__ __forceinline__ double term_f32(float x, float y)
{
    const float d = __fsub_rn(x, y);
    if (METRIC == EB_METRIC_MANHATTAN) return fabs((double)d);
    return (double)__fmul_rn(d, d);
}
template <int METRIC> __device__ __forceinline__ double term_f64(double x, double y)
{
    const double d = __dsub_rn(x, y);
    if (METRIC == EB_METRIC_MANHATTAN) return fabs(d);
    return __dmul_rn(d, d);
}
template <typename T, int METRIC> __device__ __forceinline__ double term_int(T x, T y)
{
    using W = typename Wide<T>::type;
    const W d = (W)x - (W)y;
    if (METRIC == EB_METRIC_MANHATTAN) return fabs((double)d);
    return (double)(W)(d * d);
}
template <typename T, int METRIC> struct Term {
    static __device__ __forceinline__ double f(T x, T y) { return term_int<T, METRIC>(x, y); }
};
template <int METRIC> struct Term<float, METRIC> {
    static __device__ __forceinline__ double f(float x, float y) { return term_f32<METRIC>(x, y); }
};
template <int METRIC> struct Term<double, METRIC> {
    static __device__ __forceinline__ double f(double x, double y)
    {
        return term_f64<METRIC>(x, y);
    }
};

template <int METRIC> __device__ __forceinline__ double finish(double acc)
{
    return METRIC == EB_METRIC_EUCLIDEAN ? sqrt(acc) : acc;
}

// Warp-cooperative: distances of rows [base, base+32) to the point y (shared memory).
// Lane l returns the distance of row base+l (garbage when that row is >= n).
template <typename T, int METRIC>
__device__ __forceinline__ double warp_rows_distance(const T *__restrict__ X, long n, long F,
                                                     long base, const T *y_sh,
                                                     unsigned char *tile, bool vec_ok)
{
    const int lane = threadIdx.x & 31;
    constexpr int EPT = kTileBytes / (int)sizeof(T);  // elements per row per tile
    double acc = 0.0;
    const long rows_here = min(32L, n - base);
    for (long j0 = 0; j0 < F; j0 += EPT) {
        const int fe = (int)min((long)EPT, F - j0);  // elements of this tile
        if (vec_ok) {
            // 16 lanes cover one row's 256 bytes; the warp covers two rows per request
            const int nvec = (fe * (int)sizeof(T) + 15) >> 4;
            const int r0 = lane >> 4, c16 = lane & 15;
#pragma unroll 4
            for (int r = r0; r < 32; r += 2) {
                if (r < rows_here && c16 < nvec) {
                    const int4 v = __ldg(reinterpret_cast<const int4 *>(
                        reinterpret_cast<const unsigned char *>(X + (base + r) * F + j0) +
                        16 * c16));
                    *reinterpret_cast<int4 *>(tile + r * kTileStride + 16 * c16) = v;
                }
            }
        } else {
            for (int e = lane; e < 32 * fe; e += 32) {
                const int r = e / fe, c = e - r * fe;
                if (r < rows_here)
                    reinterpret_cast<T *>(tile + r * kTileStride)[c] =
                        __ldg(X + (base + r) * F + j0 + c);
            }
        }
        __syncwarp();
        const T *row = reinterpret_cast<const T *>(tile + lane * kTileStride);
        constexpr int VE = 16 / (int)sizeof(T);  // elements per 16-byte read
        int c = 0;
        for (; c + VE <= fe; c += VE) {
            const int4 raw = *reinterpret_cast<const int4 *>(row + c);
            const T *v = reinterpret_cast<const T *>(&raw);
#pragma unroll
            for (int u = 0; u < VE; ++u)
                acc = __dadd_rn(acc, Term<T, METRIC>::f(v[u], y_sh[j0 + c + u]));
        }
        for (; c < fe; ++c) acc = __dadd_rn(acc, Term<T, METRIC>::f(row[c], y_sh[j0 + c]));
        __syncwarp();
    }
    return finish<METRIC>(acc);
}


// ------------------------------------------------------------------------------------------
// TMA staging: a warp's tile (32 rows x 256 bytes) arrives as two 32x128-byte boxes of a 2-D
// tensor map with the 128-byte hardware swizzle, so lane l can walk ITS row with conflict-free
// 16-byte reads (chunk c of row r sits at chunk c ^ (r & 7)); one elected lane issues the
// copies, the LSU never touches the staging traffic, and tile q+1 is in flight while tile q is
// being walked.
// ------------------------------------------------------------------------------------------
constexpr int kBoxBytes = 32 * 128;           // one box: 32 rows x 128 bytes
constexpr int kMultiMaxThreads = 384;         // multi-iteration kernel: up to 12 warps per block
constexpr int kKeepMB = 64;                   // L2-resident part of the shard (multi-iteration kernel)
constexpr int kTmaTileBytes = 2 * kBoxBytes;  // 256 bytes of each of 32 rows

// A full tile (256 bytes of each of the 32 rows), straight-line: row l of a box starts at
// l * 128 and its 16-byte chunk c sits at ((c ^ (l & 7)) << 4); with the 128-byte aligned row
// base that is ONE xor per load instead of the shift/mask/add sequence of the general loop
// (the body was issue-bound, not HBM-bound: DESIGN.md K2).  Same terms, same order.
template <typename T, int METRIC>
__device__ __forceinline__ double tile_terms_full(double acc, const unsigned char *tile, int lane,
                                                  const T *y_sh, long j0)
{
    constexpr int VE = 16 / (int)sizeof(T);
    const uint32_t base = smem_u32(tile) + (uint32_t)lane * 128u + (uint32_t)((lane & 7) << 4);
    const int4 *yq = reinterpret_cast<const int4 *>(y_sh + j0);
#pragma unroll
    for (int q = 0; q < kTileBytes / 16; ++q) {
        int4 raw;
        const uint32_t addr = (base ^ (uint32_t)((q & 7) << 4)) + (uint32_t)((q >> 3) * kBoxBytes);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                     : "r"(addr)
                     : "memory");
        const int4 yraw = yq[q];    // broadcast read
        const T *v = reinterpret_cast<const T *>(&raw);
        const T *yv = reinterpret_cast<const T *>(&yraw);
#pragma unroll
        for (int u = 0; u < VE; ++u) acc = __dadd_rn(acc, Term<T, METRIC>::f(v[u], yv[u]));
    }
    return acc;
}

template <typename T, int METRIC>
__device__ __forceinline__ double tile_terms_swz(double acc, const unsigned char *tile, int lane,
                                                 const T *y_sh, long j0, int fe)
{
    constexpr int VE = 16 / (int)sizeof(T);
#ifndef EB_K2_DIAG_SPLIT
    if (fe == kTileBytes / (int)sizeof(T)) return tile_terms_full<T, METRIC>(acc, tile, lane, y_sh, j0);
#endif
    const unsigned char *row = tile + lane * 128;
    const int sw = lane & 7;
#ifdef EB_K2_DIAG_SPLIT
    double acc2 = 0.0;
#endif
#pragma unroll 4
    for (int c = 0; c < fe; c += VE) {
        const int byteoff = c * (int)sizeof(T);
        const int box = byteoff >> 7, chunk = (byteoff >> 4) & 7;
        const int4 raw = *reinterpret_cast<const int4 *>(row + box * kBoxBytes +
                                                         ((chunk ^ sw) << 4));
        const int4 yraw = *reinterpret_cast<const int4 *>(y_sh + j0 + c);  // broadcast read
        const T *v = reinterpret_cast<const T *>(&raw);
        const T *yv = reinterpret_cast<const T *>(&yraw);
#ifdef EB_K2_DIAG_SPLIT   // diagnostic only (breaks the summation order): two chains per row
#pragma unroll
        for (int u = 0; u < VE; u += 2) {
            acc = __dadd_rn(acc, Term<T, METRIC>::f(v[u], yv[u]));
            acc2 = __dadd_rn(acc2, Term<T, METRIC>::f(v[u + 1], yv[u + 1]));
        }
#else
#pragma unroll
        for (int u = 0; u < VE; ++u) acc = __dadd_rn(acc, Term<T, METRIC>::f(v[u], yv[u]));
#endif
    }
#ifdef EB_K2_DIAG_SPLIT
    acc = __dadd_rn(acc, acc2);
#endif
    return acc;
}

// one staged tile: add this tile's terms to the lane's running row sum, in feature order
template <typename T, int METRIC>
__device__ __forceinline__ double tile_terms(double acc, const unsigned char *tile, int lane,
                                             const T *y_sh, long j0, int fe)
{
    const T *row = reinterpret_cast<const T *>(tile + lane * kTileStride);
    constexpr int VE = 16 / (int)sizeof(T);
    int c = 0;
    for (; c + VE <= fe; c += VE) {
        const int4 raw = *reinterpret_cast<const int4 *>(row + c);
        const int4 yraw = *reinterpret_cast<const int4 *>(y_sh + j0 + c);  // broadcast read
        const T *v = reinterpret_cast<const T *>(&raw);
        const T *yv = reinterpret_cast<const T *>(&yraw);
#pragma unroll
        for (int u = 0; u < VE; ++u) acc = __dadd_rn(acc, Term<T, METRIC>::f(v[u], yv[u]));
    }
    for (; c < fe; ++c) acc = __dadd_rn(acc, Term<T, METRIC>::f(row[c], y_sh[j0 + c]));
    return acc;
}

enum FeatMode { kFStep = 0, kFSeed = 1, kFDistOnly = 2 };

struct FeatSmem {
    Partial red[32];
    int flag;
    double maxdist;
    int64_t center_index;
};

template <typename T, int METRIC, int MODE>
__global__ void __launch_bounds__(kFeatThreads)
k_kcenters_step_feat(const T *__restrict__ X, long n, long F, long frame_offset,
                     const unsigned char *cand_in, int n_cand, size_t rec_bytes, double *dist,
                     int *assign, int n_clusters_limit, double cutoff, eb_kc_state *state,
                     int64_t *center_list, Partial *partials, unsigned char *cand_out,
                     const T *y_direct, double *out_only, int vec_ok,
                     const __grid_constant__ CUtensorMap tmap)
{
    // layout: [per-warp tiles: 2 x 8 KB, 1024-byte aligned for the swizzle] [FeatSmem] [bars] [y]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    FeatSmem *ss = reinterpret_cast<FeatSmem *>(tiles + (size_t)kFeatWarps * 2 * kTmaTileBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(ss) +
                                                  align16(sizeof(FeatSmem)));
    T *y_sh = reinterpret_cast<T *>(bars + 2 * kFeatWarps);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int k = 0;
    if (lane == 0) {
        mbar_init(bars + 2 * warp, 1);
        mbar_init(bars + 2 * warp + 1, 1);
        fence_mbar_init();
        if (vec_ok) tma_prefetch_desc(&tmap);
    }
    __syncwarp();

    if (MODE == kFStep) {
        double cd;
        int64_t ci;
        const int r = pick_candidate(cand_in, n_cand, rec_bytes, cd, ci);
        const bool active = (r >= 0) && step_active(state, n_clusters_limit, cd, cutoff, k);
        if (!active) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (!state->done) {
                    state->done = 1;
                    state->maxdist = cd;
                } else {
                    state->n_noop += 1;
                }
            }
            return;
        }
        const T *src = reinterpret_cast<const T *>(cand_in + (size_t)r * rec_bytes +
                                                   sizeof(RecHeader));
        for (long j = threadIdx.x; j < F; j += blockDim.x) y_sh[j] = src[j];
        if (threadIdx.x == 0) {
            ss->center_index = ci;
            ss->maxdist = cd;
        }
        __syncthreads();
    } else if (MODE == kFDistOnly) {
        for (long j = threadIdx.x; j < F; j += blockDim.x) y_sh[j] = y_direct[j];
        __syncthreads();
    }

    double best_d = -2.0;
    int64_t best_i = INT64_MAX;
    const long n_chunks = (n + 31) >> 5;
    const long warps_total = (long)gridDim.x * kFeatWarps;
    unsigned char *tile = tiles + (size_t)warp * 2 * kTmaTileBytes;

    // commit one finished row (strict '<' update, arg-max tracking)
    // `old` = the row's current distance; the pipelined path requests it before it waits for
    // the row's first tile so that its latency hides behind the row sum
    auto commit = [&](long row, double d, double old) {
        if (row < n) {
            double cur;
            if (MODE == kFSeed) {
                cur = dist[row];
            } else if (MODE == kFDistOnly) {
                out_only[row] = d;
                cur = 0.0;
            } else {
                if (d < old) {  // strict '<', kcenters.py:304
                    dist[row] = d;
                    assign[row] = k;
                }
                cur = (d < old) ? d : old;
            }
            if (cur > best_d) {
                best_d = cur;
                best_i = row;
            }
        }
    };

    const long first_chunk = (long)blockIdx.x * kFeatWarps + warp;
    if (MODE != kFSeed && vec_ok) {
        // ---- pipelined TMA path ------------------------------------------------------------
        constexpr int EPT = kTileBytes / (int)sizeof(T);   // elements of a row per tile
        constexpr int EPB = 128 / (int)sizeof(T);          // elements of a row per box
        const int nt = (int)((F + EPT - 1) / EPT);         // tiles per 32-row chunk
        uint64_t *bar = bars + 2 * warp;
        // producer cursor (next tile to request) and consumer cursor (tile being walked)
        long p_chunk = first_chunk, c_chunk = first_chunk;
        int p_jt = 0, c_jt = 0, p_buf = 0;
        auto issue = [&]() {
            if (lane == 0) {
                const long j0 = (long)p_jt * EPT;
                const int fe = (int)min((long)EPT, F - j0);
                const int n_box = (fe + EPB - 1) / EPB;
                unsigned char *dst = tile + p_buf * kTmaTileBytes;
                mbar_expect_tx(bar + p_buf, (uint32_t)n_box * kBoxBytes);
                for (int b = 0; b < n_box; ++b)
                    tma_load_2d(dst + b * kBoxBytes, &tmap, (int)(j0 + (long)b * EPB),
                                (int)(p_chunk << 5), bar + p_buf);
            }
            p_buf ^= 1;
            if (++p_jt == nt) {
                p_jt = 0;
                p_chunk += warps_total;
            }
        };
        uint32_t phase0 = 0, phase1 = 0;
        double acc = 0.0, old_pre = 0.0;
        int c_buf = 0;
        if (p_chunk < n_chunks) issue();
        while (c_chunk < n_chunks) {
            if (p_chunk < n_chunks) issue();
            if (MODE == kFStep && c_jt == 0 && (c_chunk << 5) + lane < n)
                old_pre = dist[(c_chunk << 5) + lane];
            mbar_wait(bar + c_buf, c_buf ? phase1 : phase0);
            if (c_buf) phase1 ^= 1; else phase0 ^= 1;
            const long j0 = (long)c_jt * EPT;
            const int fe = (int)min((long)EPT, F - j0);
            acc = tile_terms_swz<T, METRIC>(acc, tile + c_buf * kTmaTileBytes, lane, y_sh, j0, fe);
            if (c_jt == nt - 1) {
                commit((c_chunk << 5) + lane, finish<METRIC>(acc), old_pre);
                acc = 0.0;
            }
            c_buf ^= 1;
            if (++c_jt == nt) {
                c_jt = 0;
                c_chunk += warps_total;
            }
            __syncwarp();  // everyone is done with this buffer before it is refilled
        }
    } else {
        for (long chunk = first_chunk; chunk < n_chunks; chunk += warps_total) {
            const long base = chunk << 5;
            double d = 0.0;
            if (MODE != kFSeed)
                d = warp_rows_distance<T, METRIC>(X, n, F, base, y_sh, tile, false);
            commit(base + lane, d, (MODE == kFStep && base + lane < n) ? dist[base + lane] : 0.0);
        }
    }
    if (MODE == kFDistOnly) return;

    if (!grid_argmax_last_block(best_d, best_i, partials, state, ss->red, &ss->flag)) return;

    RecHeader *out = reinterpret_cast<RecHeader *>(cand_out);
    const bool empty = (best_i == INT64_MAX);
    if (!empty) {
        const T *src = X + best_i * F;
        T *dst = reinterpret_cast<T *>(cand_out + sizeof(RecHeader));
        for (long j = threadIdx.x; j < F; j += blockDim.x) dst[j] = src[j];
    }
    if (threadIdx.x == 0) {
        out->dist = empty ? -1.0 : best_d;
        out->index = empty ? -1 : frame_offset + best_i;
        out->trace = 0.0;
        out->reserved = 0;
        if (MODE == kFStep) {
            center_list[k] = ss->center_index;
            state->n_centers = k + 1;
            state->last_center = ss->center_index;
            state->maxdist = ss->maxdist;
        } else {
            state->n_centers = n_clusters_limit;  // seed mode: carries first_center_id
            state->done = 0;
            state->n_noop = 0;
            state->maxdist = 0.0;
            state->last_center = -1;
        }
        state->local_maxdist = empty ? -1.0 : best_d;
        state->blocks_done = 0;
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------
// Persistent multi-iteration step (single shard, TMA path): up to n_steps iterations of
// kcenters.py:217-226 in ONE cooperative launch.  A launch per iteration costs ~18 us of fixed
// overhead at 1M x 64 (launch gap, prologue round trips, DRAM ramp, the single-CTA arg-max
// tail) on top of a 41 us HBM-bound body; here the grid stays resident, the shard arg-max is
// a grid barrier after which EVERY block reduces the per-block partials itself (no serial
// last-block tail), the winner's row is read straight from X, and the TMA ring keeps running
// across the barrier (the first tiles of iteration i+1 are in flight while iteration i's
// arg-max is being agreed on -- X does not change).  Same arithmetic, same tie rules, same
// state protocol as k_kcenters_step_feat<T, METRIC, kFStep>: on exit cand holds the next
// candidate record and *state the counters, so single launches and multi launches mix freely.
// gbar: monotonically increasing arrival counter (a multiple of gridDim.x between launches).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(kMultiMaxThreads)
k_kcenters_multi_feat(const T *__restrict__ X, long n, long F, long frame_offset,
                      unsigned char *cand, size_t rec_bytes, double *dist, int *assign,
                      int n_clusters_limit, double cutoff, eb_kc_state *state,
                      int64_t *center_list, Partial *partials, unsigned long long *gbar,
                      int n_steps, const __grid_constant__ CUtensorMap tmap, unsigned int *dyn,
                      long keep_chunks)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    const int W = (int)(blockDim.x >> 5);      // warps per block: 12, 8 or 4 (launcher)
    FeatSmem *ss = reinterpret_cast<FeatSmem *>(tiles + (size_t)W * 2 * kTmaTileBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(ss) +
                                                  align16(sizeof(FeatSmem)));
    T *y_sh = reinterpret_cast<T *>(bars + 2 * W);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned G = gridDim.x;
    if (lane == 0) {
        mbar_init(bars + 2 * warp, 1);
        mbar_init(bars + 2 * warp + 1, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }
    __syncwarp();

    // nobody writes *state / *gbar's base before the first barrier, so these reads agree
    const int k0 = *reinterpret_cast<const volatile int32_t *>(&state->n_centers);
    const int done0 = *reinterpret_cast<const volatile int32_t *>(&state->done);
    const unsigned long long gbase = (ld_acquire_gpu_u64(gbar) / G) * G;

    constexpr int EPT = kTileBytes / (int)sizeof(T);
    constexpr int EPB = 128 / (int)sizeof(T);
    const int nt = (int)((F + EPT - 1) / EPT);
    const long n_chunks = (n + 31) >> 5;
    const long warps_total = (long)G * W;
    const long first_chunk = (long)blockIdx.x * W + warp;
    const bool has_work = first_chunk < n_chunks;
    unsigned char *tile = tiles + (size_t)warp * 2 * kTmaTileBytes;
    uint64_t *bar = bars + 2 * warp;

    // Work distribution.  The body is HBM-fair only to a few per cent: with a purely static
    // split block 0 waited 7-8 us per iteration for the slowest SM (measured).  So the first
    // 7/8 of every warp's chunks are static (stride warps_total: the TMA ring can run ahead,
    // also across the arg-max barrier) and the last eighth of the shard is handed out
    // dynamically, one chunk per ticket of a global counter (dyn[iteration & 1]; block 0 resets
    // the other one while nobody uses it).  Tickets only grow, so every lane still meets its
    // rows in increasing order and the first-occurrence arg-max rule holds.
    const long per_warp_static = (n_chunks * 7 / 8) / warps_total;
    const bool dynamic_tail = per_warp_static >= 4;
    const long n_static = dynamic_tail ? per_warp_static * warps_total : 0;
    // producer cursor: (iteration, chunk, tile); runs at most two tiles ahead of the consumer
    int p_iter = 0, p_jt = 0, p_buf = 0, inflight = 0;
    long p_slot = 0;                    // static chunks already issued in iteration p_iter
    long p_chunk = first_chunk;         // chunk being issued (valid while p_iter < n_steps)
    long ticket = -1;                   // prefetched dynamic ticket (-1: none)
    long buf_chunk0 = 0, buf_chunk1 = 0;     // scalars, not arrays: dynamic indexing would put
    int buf_iter0 = -1, buf_iter1 = -1;      // them in local memory, on the consumer's path
    auto fetch_ticket = [&](int iter) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(&dyn[iter & 1], 1u);
        return n_static + (long)__shfl_sync(0xffffffffu, t, 0);
    };
    // chunk after the one just issued; moves on to the next iteration when this one is done
    auto advance_chunk = [&]() {
        if (!dynamic_tail) {
            p_chunk += warps_total;
            if (p_chunk >= n_chunks) {
                p_chunk = first_chunk;
                ++p_iter;
            }
            return;
        }
        if (++p_slot < per_warp_static) {
            p_chunk = first_chunk + p_slot * warps_total;
            return;
        }
        const long t = ticket >= 0 ? ticket : fetch_ticket(p_iter);
        if (t < n_chunks) {
            p_chunk = t;
            ticket = fetch_ticket(p_iter);      // needed at the next advance: latency hidden
        } else {
            ticket = -1;
            p_slot = 0;
            p_chunk = first_chunk;
            ++p_iter;
        }
    };
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    auto try_issue = [&]() {
        if (!has_work || p_iter >= n_steps || inflight >= 2) return;
        if (lane == 0) {
            const long j0 = (long)p_jt * EPT;
            const int fe = (int)min((long)EPT, F - j0);
            const int n_box = (fe + EPB - 1) / EPB;
            unsigned char *dst = tile + p_buf * kTmaTileBytes;
            mbar_expect_tx(bar + p_buf, (uint32_t)n_box * kBoxBytes);
            // the first keep_chunks chunks of the shard stay in L2 from one iteration to the
            // next (evict_last); the rest streams through (evict_first)
            const uint64_t pol = p_chunk < keep_chunks ? pol_keep : pol_stream;
            for (int b = 0; b < n_box; ++b)
                tma_load_2d_hint(dst + b * kBoxBytes, &tmap, (int)(j0 + (long)b * EPB),
                                 (int)(p_chunk << 5), bar + p_buf, pol);
        }
        if (p_buf) {
            buf_chunk1 = p_chunk;
            buf_iter1 = p_iter;
        } else {
            buf_chunk0 = p_chunk;
            buf_iter0 = p_iter;
        }
        p_buf ^= 1;
        ++inflight;
        if (++p_jt == nt) {
            p_jt = 0;
            advance_chunk();
        }
    };
    uint32_t phase0 = 0, phase1 = 0;
    int c_buf = 0;
    auto wait_tile = [&]() {
        mbar_wait(bar + c_buf, c_buf ? phase1 : phase0);
        if (c_buf) phase1 ^= 1; else phase0 ^= 1;
    };
    try_issue();     // the first tile does not depend on the centre: start it right away

    int it = 0;
    bool stopped = false;
    unsigned long long sync_ns = 0, t_body_end = 0;
    double win_d = 0.0;        // distance / global index of the winner examined last
    int64_t win_i = -1;
    for (; it < n_steps; ++it) {
        // ---- the centre of this iteration: candidate record (first) or the partials ---------
        const T *src;
        bool have;
        if (it == 0) {
            const RecHeader *h = reinterpret_cast<const RecHeader *>(cand);
            win_d = __ldcg(&h->dist);
            win_i = __ldcg(reinterpret_cast<const long long *>(&h->index));
            have = win_i >= 0;
            src = reinterpret_cast<const T *>(cand + sizeof(RecHeader));
        } else {
            const Partial *pp = partials + (size_t)((it - 1) & 1) * G;
            double d = -2.0;
            int64_t i = INT64_MAX;
            for (unsigned b = threadIdx.x; b < G; b += blockDim.x) {
                const double od = __ldcg(&pp[b].dist);
                const int64_t oi = __ldcg(reinterpret_cast<const long long *>(&pp[b].index));
                if (better(od, oi, d, i)) {
                    d = od;
                    i = oi;
                }
            }
            block_argmax(d, i, ss->red);
            if (threadIdx.x == 0) {
                ss->red[0].dist = d;
                ss->red[0].index = i;
            }
            __syncthreads();
            d = ss->red[0].dist;
            i = ss->red[0].index;
            __syncthreads();
            have = i != INT64_MAX;
            win_d = have ? d : -1.0;
            win_i = have ? frame_offset + i : -1;
            src = X + (have ? i : 0) * F;
        }
        const int k = k0 + it;
        if (!(have && !done0 && k < n_clusters_limit && win_d > cutoff)) {   // kcenters.py:217
            stopped = true;
            break;
        }
        for (long j = threadIdx.x; j < F; j += blockDim.x) y_sh[j] = src[j];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            center_list[k] = win_i;
            // diagnostic: time block 0 spends between two bodies (arg-max barrier + prologue)
            if (it > 0) sync_ns += globaltimer_ns() - t_body_end;
        }
        __syncthreads();

        // ---- body: every row once, strict '<' update, arg-max --------------------------------
        double best_d = -2.0;
        int64_t best_i = INT64_MAX;
        double acc = 0.0, old_pre = 0.0;
        if (dynamic_tail && blockIdx.x == 0 && threadIdx.x == 0) dyn[(it + 1) & 1] = 0u;
        int c_jt = 0;
        for (;;) {
            try_issue();
            // this iteration is over for the warp when the next tile in the ring belongs to
            // the following iteration, or nothing is left to request
            if (inflight == 0 || (c_buf ? buf_iter1 : buf_iter0) != it) break;
            const long c_chunk = c_buf ? buf_chunk1 : buf_chunk0;
            const long row = (c_chunk << 5) + lane;
            // the row's current distance is requested before the tile is waited for, so its
            // latency hides behind the row sum.  L2 read: a dynamically assigned chunk may have
            // been updated by another SM in the previous iteration (L1 is not coherent across
            // SMs); the grid barrier ordered that write before this read.
            if (c_jt == 0 && row < n) old_pre = __ldcg(dist + row);
            wait_tile();
            const long j0 = (long)c_jt * EPT;
            const int fe = (int)min((long)EPT, F - j0);
            acc = tile_terms_swz<T, METRIC>(acc, tile + c_buf * kTmaTileBytes, lane, y_sh, j0,
                                            fe);
            if (c_jt == nt - 1) {
                if (row < n) {
                    const double d = finish<METRIC>(acc);
                    const double old = old_pre;
                    if (d < old) {  // strict '<', kcenters.py:304
                        dist[row] = d;
                        assign[row] = k;
                    }
                    const double cur = (d < old) ? d : old;
                    if (cur > best_d) {
                        best_d = cur;
                        best_i = row;
                    }
                }
                acc = 0.0;
            }
            if (++c_jt == nt) c_jt = 0;
            c_buf ^= 1;
            --inflight;
            __syncwarp();  // everyone is done with this buffer before it is refilled
        }
        // ---- shard arg-max: per-block partial, grid barrier ------------------------------------
        if (blockIdx.x == 0 && threadIdx.x == 0) t_body_end = globaltimer_ns();
        block_argmax(best_d, best_i, ss->red);
        if (threadIdx.x == 0) {
            Partial *pp = partials + (size_t)(it & 1) * G;
            pp[blockIdx.x].dist = best_d;
            pp[blockIdx.x].index = best_i;
            __threadfence();
            atomicAdd(gbar, 1ull);
            const unsigned long long target = gbase + (unsigned long long)(it + 1) * G;
            while (ld_acquire_gpu_u64(gbar) < target) {
                __nanosleep(32);   // back off: hundreds of pollers on one line slow the arrivals
            }
        }
        __syncthreads();
    }

    // tiles requested for an iteration that will not run must land before the CTA exits
    while (inflight > 0) {
        wait_tile();
        c_buf ^= 1;
        --inflight;
    }
    if (blockIdx.x != 0) return;

    // ---- block 0: leave *state and the candidate record as a chain of single steps would ----
    if (it > 0) {
        const Partial *pp = partials + (size_t)((it - 1) & 1) * G;
        double d = -2.0;
        int64_t i = INT64_MAX;
        for (unsigned b = threadIdx.x; b < G; b += blockDim.x) {
            const double od = __ldcg(&pp[b].dist);
            const int64_t oi = __ldcg(reinterpret_cast<const long long *>(&pp[b].index));
            if (better(od, oi, d, i)) {
                d = od;
                i = oi;
            }
        }
        block_argmax(d, i, ss->red);
        if (threadIdx.x == 0) {
            ss->red[0].dist = d;
            ss->red[0].index = i;
        }
        __syncthreads();
        d = ss->red[0].dist;
        i = ss->red[0].index;
        const bool empty = (i == INT64_MAX);
        RecHeader *out = reinterpret_cast<RecHeader *>(cand);
        if (!empty) {
            const T *srow = X + i * F;
            T *dst = reinterpret_cast<T *>(cand + sizeof(RecHeader));
            for (long j = threadIdx.x; j < F; j += blockDim.x) dst[j] = srow[j];
        }
        if (threadIdx.x == 0) {
            out->dist = empty ? -1.0 : d;
            out->index = empty ? -1 : frame_offset + i;
            out->trace = 0.0;
            out->reserved = 0;
            state->n_centers = k0 + it;
            state->last_center = center_list[k0 + it - 1];
            state->local_maxdist = empty ? -1.0 : d;
        }
    }
    if (threadIdx.x == 0) {
        if (stopped) {
            if (!done0) {
                state->done = 1;
                state->maxdist = win_d;
                state->n_noop += n_steps - it - 1;
            } else {
                state->n_noop += n_steps;
            }
        } else {
            // the distance of the LAST chosen centre's candidate (what the last prologue saw)
            state->maxdist = win_d;
        }
        state->blocks_done = 0;
        state->wait_ns += (long long)sync_ns;
        __threadfence();
    }
    (void)rec_bytes;
}

// ------------------------------------------------------------------------------------------
// many centres: every row against k centres in order, strict '<'.  A warp owns 32 rows and
// re-stages them per centre from L1/L2 (k is small whenever this path matters for features).
// ------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void __launch_bounds__(kFeatThreads)
k_feat_assign(const T *__restrict__ X, long n, long F, const T *__restrict__ centers, int k,
              const int64_t *__restrict__ frame_idx, long m, double *out_dist, int *out_assign,
              int accumulate, int scatter)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    T *y_all = reinterpret_cast<T *>(tiles + (size_t)kFeatWarps * 32 * kTileStride);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *tile = tiles + (size_t)warp * 32 * kTileStride;
    T *y_sh = y_all + (size_t)warp * F;  // one staged centre per warp
    const long n_chunks = (m + 31) >> 5;
    const long warps_total = (long)gridDim.x * kFeatWarps;
    constexpr int EPT = kTileBytes / (int)sizeof(T);

    for (long chunk = (long)blockIdx.x * kFeatWarps + warp; chunk < n_chunks;
         chunk += warps_total) {
        const long item = (chunk << 5) + lane;
        const bool valid = item < m;
        const long row = valid ? (frame_idx ? frame_idx[item] : item) : 0;
        double best_d = INFINITY;
        int best_c = 0;
        const long opos = scatter ? row : item;  // where this row's result lives
        if (valid && accumulate) {
            best_d = out_dist[opos];
            best_c = out_assign[opos];
        }
        const bool single_tile = (F <= EPT);  // the whole row fits one staged tile
        if (single_tile) {
            for (int e = lane; e < 32 * (int)F; e += 32) {
                const int r = e / (int)F, cc = e - r * (int)F;
                const long it = (chunk << 5) + r;
                if (it < m) {
                    const long rr = frame_idx ? frame_idx[it] : it;
                    reinterpret_cast<T *>(tile + r * kTileStride)[cc] = __ldg(X + rr * F + cc);
                }
            }
        }
        for (int c = 0; c < k; ++c) {
            __syncwarp();
            for (long j = lane; j < F; j += 32) y_sh[j] = __ldg(centers + (size_t)c * F + j);
            __syncwarp();
            double acc = 0.0;
            const T *rowp = reinterpret_cast<const T *>(tile + lane * kTileStride);
            if (single_tile) {
                for (int cc = 0; cc < (int)F; ++cc)
                    acc = __dadd_rn(acc, Term<T, METRIC>::f(rowp[cc], y_sh[cc]));
            } else {
                // gather-stage the warp's rows tile by tile (rows need not be contiguous)
                for (long j0 = 0; j0 < F; j0 += EPT) {
                    const int fe = (int)min((long)EPT, F - j0);
                    for (int e = lane; e < 32 * fe; e += 32) {
                        const int r = e / fe, cc = e - r * fe;
                        const long it = (chunk << 5) + r;
                        if (it < m) {
                            const long rr = frame_idx ? frame_idx[it] : it;
                            reinterpret_cast<T *>(tile + r * kTileStride)[cc] =
                                __ldg(X + rr * F + j0 + cc);
                        }
                    }
                    __syncwarp();
                    for (int cc = 0; cc < fe; ++cc)
                        acc = __dadd_rn(acc, Term<T, METRIC>::f(rowp[cc], y_sh[j0 + cc]));
                    __syncwarp();
                }
            }
            const double d = finish<METRIC>(acc);
            if (valid && d < best_d) {
                best_d = d;
                best_c = c;
            }
        }
        if (valid) {
            out_dist[opos] = best_d;
            out_assign[opos] = best_c;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
template <typename T> static CUtensorMapDataType tmap_dtype();
template <> CUtensorMapDataType tmap_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <> CUtensorMapDataType tmap_dtype<double>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT64; }
template <> CUtensorMapDataType tmap_dtype<int8_t>() { return CU_TENSOR_MAP_DATA_TYPE_UINT8; }
template <> CUtensorMapDataType tmap_dtype<int16_t>() { return CU_TENSOR_MAP_DATA_TYPE_UINT16; }
template <> CUtensorMapDataType tmap_dtype<int32_t>() { return CU_TENSOR_MAP_DATA_TYPE_INT32; }
template <> CUtensorMapDataType tmap_dtype<int64_t>() { return CU_TENSOR_MAP_DATA_TYPE_INT64; }

static size_t elem_size(int dtype)
{
    switch (dtype) {
        case EB_DT_F32: return 4;
        case EB_DT_F64: return 8;
        case EB_DT_I8: return 1;
        case EB_DT_I16: return 2;
        case EB_DT_I32: return 4;
        case EB_DT_I64: return 8;
    }
    return 0;
}

// the multi-iteration kernel: one block of up to 12 warps per SM (148 partials and pollers at
// the arg-max barrier instead of 444)
static size_t feat_multi_smem(long F, size_t es, int warps)
{
    return (size_t)warps * 2 * kTmaTileBytes + align16(sizeof(FeatSmem)) + 16 * warps +
           align16((size_t)F * es) + 1024;
}

static size_t feat_smem(long F, size_t es, int n_y)
{
    return (size_t)kFeatWarps * 2 * kTmaTileBytes + align16(sizeof(FeatSmem)) +
           16 * kFeatWarps + align16((size_t)F * es) * n_y + 1024;
}

static int feat_grid(long n, size_t smem)
{
    const long chunks = (n + 31) / 32;
    long blocks = (chunks + kFeatWarps - 1) / kFeatWarps;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    const long cap = (long)per_sm * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T, int METRIC, int MODE>
static int launch_feat(const void *X, long n, long F, long frame_offset, const void *cand_in,
                       int n_cand, double *dist, int *assign, int limit, double cutoff,
                       eb_kc_state *state, int64_t *center_list, void *partials, void *cand_out,
                       const void *y_direct, double *out_only, cudaStream_t stream)
{
    const size_t smem = feat_smem(F, sizeof(T), 1);
    if (smem > 227 * 1024)
        return fail(EB_ERR_LIMIT, "%s: n_features=%ld needs %ld bytes of shared memory",
                    "feature kernel", F, (long)smem);
    auto kern = k_kcenters_step_feat<T, METRIC, MODE>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        configured = 227 * 1024;
    }
    const size_t rec_bytes = sizeof(RecHeader) + align16((size_t)F * sizeof(T));
    int vec_ok = (((size_t)F * sizeof(T)) % 16 == 0) && (((uintptr_t)X) % 16 == 0) &&
                 MODE != kFSeed && n > 0 && n < (1L << 31) && F < (1L << 31);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (vec_ok) {
        const uint32_t box_cols = (uint32_t)min((long)(128 / sizeof(T)), F);
        // a box narrower than 128 bytes cannot use the 128-byte swizzle: fall back
        if (box_cols * sizeof(T) != 128) {
            vec_ok = 0;
        } else {
            const int rc = make_tmap_2d(&tmap, X, tmap_dtype<T>(), sizeof(T), (uint64_t)n,
                                        (uint64_t)F, (uint64_t)F * sizeof(T), 32, box_cols,
                                        CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc != EB_OK) return rc;
        }
    }
    kern<<<feat_grid(n, smem), kFeatThreads, smem, stream>>>(
        (const T *)X, n, F, frame_offset, (const unsigned char *)cand_in, n_cand, rec_bytes, dist,
        assign, limit, cutoff, state, center_list, (Partial *)partials,
        (unsigned char *)cand_out, (const T *)y_direct, out_only, vec_ok, tmap);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

// Persistent multi-iteration launch (single shard).  Returns EB_OK with *used = 0 when the
// shape is not eligible (no 128-byte TMA boxes, too few rows, cooperative launch unsupported):
// the caller then queues single steps.
template <typename T, int METRIC>
static int launch_feat_multi(const void *X, long n, long F, long frame_offset, void *cand,
                             double *dist, int *assign, int limit, double cutoff,
                             eb_kc_state *state, int64_t *center_list, void *partials,
                             int n_steps, cudaStream_t stream, int *used)
{
    *used = 0;
    int warps = 12;
    while (warps > 4 && feat_multi_smem(F, sizeof(T), warps) > 227 * 1024) warps -= 4;
    const size_t smem = feat_multi_smem(F, sizeof(T), warps);
    if (smem > 227 * 1024) return EB_OK;
    const int threads = 32 * warps;
    const bool vec_ok = (((size_t)F * sizeof(T)) % 16 == 0) && (((uintptr_t)X) % 16 == 0) &&
                        n > 0 && n < (1L << 31) && F < (1L << 31) &&
                        (size_t)min((long)(128 / sizeof(T)), F) * sizeof(T) == 128;
    if (!vec_ok) return EB_OK;
    static int coop = -1;
    if (coop < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
        coop = v;
    }
    if (!coop) return EB_OK;
    auto kern = k_kcenters_multi_feat<T, METRIC>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        configured = 227 * 1024;
    }
    int per_sm = 0;
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return EB_OK;
    const long chunks = (n + 31) / 32;
    long grid_l = (chunks + warps - 1) / warps;
    const long cap = (long)per_sm * sm_count();
    if (grid_l > cap) grid_l = cap;
    int grid = (int)grid_l;
    if (2 * grid + 1 > kMaxGrid) return EB_OK;   // two partial buffers of `grid` entries + tickets
    CUtensorMap tmap;
    const int rc = make_tmap_2d(&tmap, X, tmap_dtype<T>(), sizeof(T), (uint64_t)n, (uint64_t)F,
                                (uint64_t)F * sizeof(T), 32, (uint32_t)(128 / sizeof(T)),
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != EB_OK) return rc;
    const T *Xp = (const T *)X;
    unsigned char *candp = (unsigned char *)cand;
    size_t rec_bytes = sizeof(RecHeader) + align16((size_t)F * sizeof(T));
    Partial *pp = (Partial *)partials;
    unsigned long long *gbar = reinterpret_cast<unsigned long long *>(&state->reserved);
    // two ticket counters of the dynamic tail live behind the two partial buffers
    unsigned int *dyn = reinterpret_cast<unsigned int *>(pp + 2 * (size_t)grid);
    EB_CUDA(cudaMemsetAsync(dyn, 0, 2 * sizeof(unsigned int), stream));
    // rows kept L2-resident across iterations (EB_K2_L2_MB, default kKeepMB); nothing is pinned
    // when the whole shard fits anyway
    static const long keep_mb = [] {
        const char *e = getenv("EB_K2_L2_MB");
        return e ? atol(e) : (long)kKeepMB;
    }();
    const size_t chunk_bytes = (size_t)32 * (size_t)F * sizeof(T);
    long keep_chunks = (long)(((size_t)keep_mb << 20) / chunk_bytes);
    void *args[] = {&Xp, &n, &F, &frame_offset, &candp, &rec_bytes, &dist, &assign, &limit,
                    &cutoff, &state, &center_list, &pp, &gbar, &n_steps, &tmap, &dyn,
                    &keep_chunks};
    EB_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(threads), args,
                                        smem, stream));
    *used = 1;
    return EB_OK;
}

template <typename T, typename... Args> static int dispatch_multi_metric(int metric, Args... args)
{
    switch (metric) {
        case EB_METRIC_EUCLIDEAN: return launch_feat_multi<T, EB_METRIC_EUCLIDEAN>(args...);
        case EB_METRIC_MANHATTAN: return launch_feat_multi<T, EB_METRIC_MANHATTAN>(args...);
        case EB_METRIC_SQEUCLIDEAN: return launch_feat_multi<T, EB_METRIC_SQEUCLIDEAN>(args...);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown metric");
}

template <typename... Args> static int dispatch_multi(int dtype, int metric, Args... args)
{
    switch (dtype) {
        case EB_DT_F32: return dispatch_multi_metric<float>(metric, args...);
        case EB_DT_F64: return dispatch_multi_metric<double>(metric, args...);
        case EB_DT_I8: return dispatch_multi_metric<int8_t>(metric, args...);
        case EB_DT_I16: return dispatch_multi_metric<int16_t>(metric, args...);
        case EB_DT_I32: return dispatch_multi_metric<int32_t>(metric, args...);
        case EB_DT_I64: return dispatch_multi_metric<int64_t>(metric, args...);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown dtype");
}

template <typename T, int MODE, typename... Args>
static int dispatch_metric(int metric, Args... args)
{
    switch (metric) {
        case EB_METRIC_EUCLIDEAN: return launch_feat<T, EB_METRIC_EUCLIDEAN, MODE>(args...);
        case EB_METRIC_MANHATTAN: return launch_feat<T, EB_METRIC_MANHATTAN, MODE>(args...);
        case EB_METRIC_SQEUCLIDEAN: return launch_feat<T, EB_METRIC_SQEUCLIDEAN, MODE>(args...);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown metric");
}

template <int MODE, typename... Args> static int dispatch(int dtype, int metric, Args... args)
{
    switch (dtype) {
        case EB_DT_F32: return dispatch_metric<float, MODE>(metric, args...);
        case EB_DT_F64: return dispatch_metric<double, MODE>(metric, args...);
        case EB_DT_I8: return dispatch_metric<int8_t, MODE>(metric, args...);
        case EB_DT_I16: return dispatch_metric<int16_t, MODE>(metric, args...);
        case EB_DT_I32: return dispatch_metric<int32_t, MODE>(metric, args...);
        case EB_DT_I64: return dispatch_metric<int64_t, MODE>(metric, args...);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown dtype");
}

template <typename T, int METRIC>
static int launch_assign(const void *X, long n, long F, const void *centers, int k,
                         const int64_t *frame_idx, long m, double *out_dist, int *out_assign,
                         int accumulate, int scatter, cudaStream_t stream)
{
    const size_t smem = (size_t)kFeatWarps * 32 * kTileStride +
                        align16((size_t)F * sizeof(T)) * kFeatWarps;
    if (smem > 227 * 1024)
        return fail(EB_ERR_LIMIT, "%s: n_features=%ld needs %ld bytes of shared memory",
                    "feat_assign", F, (long)smem);
    auto kern = k_feat_assign<T, METRIC>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        configured = 227 * 1024;
    }
    kern<<<feat_grid(m, smem), kFeatThreads, smem, stream>>>(
        (const T *)X, n, F, (const T *)centers, k, frame_idx, m, out_dist, out_assign, accumulate,
        scatter);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

template <typename T, typename... Args> static int dispatch_assign_metric(int metric, Args... args)
{
    switch (metric) {
        case EB_METRIC_EUCLIDEAN: return launch_assign<T, EB_METRIC_EUCLIDEAN>(args...);
        case EB_METRIC_MANHATTAN: return launch_assign<T, EB_METRIC_MANHATTAN>(args...);
        case EB_METRIC_SQEUCLIDEAN: return launch_assign<T, EB_METRIC_SQEUCLIDEAN>(args...);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown metric");
}

}  // namespace eb

using namespace eb;

extern "C" {

size_t eb_feat_record_bytes(int64_t n_features, int dtype)
{
    return sizeof(RecHeader) + align16((size_t)n_features * elem_size(dtype));
}

int eb_kcenters_step_feat(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                          int64_t frame_offset, const void *cand_in, int n_cand, double *dist,
                          int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                          eb_kc_state *state, int64_t *center_list, void *partials,
                          void *cand_out, int n_steps, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_features > 0, "kcenters_step_feat: bad shape");
    EB_CHECK_ARG(n_steps >= 1, "kcenters_step_feat: n_steps < 1");
    EB_CHECK_ARG(n_steps == 1 || (n_cand == 1 && cand_in == cand_out),
                 "kcenters_step_feat: n_steps > 1 needs a single shard (cand_in == cand_out)");
    EB_CHECK_ARG(elem_size(dtype) != 0, "kcenters_step_feat: unknown dtype");
    EB_CHECK_ARG(n_cand >= 1 && cand_in && cand_out && state && partials && center_list,
                 "kcenters_step_feat: null pointer / n_cand < 1");
    // EB_K2_MULTI=0 forces one launch per iteration (developer A/B switch)
    static const int multi_on = [] {
        const char *e = getenv("EB_K2_MULTI");
        return e ? atoi(e) : 1;
    }();
    if (n_steps > 1 && multi_on) {
        // single shard: all iterations in one persistent cooperative launch when eligible
        int used = 0;
        const int rc = dispatch_multi(dtype, metric, X, (long)n, (long)n_features,
                                      (long)frame_offset, cand_out, dist, assign,
                                      (int)n_clusters_limit, dist_cutoff, state, center_list,
                                      partials, n_steps, (cudaStream_t)stream, &used);
        if (rc != EB_OK) return rc;
        if (used) return EB_OK;
    }
    for (int it = 0; it < n_steps; ++it) {
        const int rc = dispatch<kFStep>(dtype, metric, X, (long)n, (long)n_features,
                                        (long)frame_offset, cand_in, n_cand, dist, assign,
                                        (int)n_clusters_limit, dist_cutoff, state, center_list,
                                        partials, cand_out, (const void *)nullptr,
                                        (double *)nullptr, (cudaStream_t)stream);
        if (rc != EB_OK) return rc;
    }
    return EB_OK;
}

int eb_kcenters_seed_feat(const void *X, int64_t n, int64_t n_features, int dtype,
                          int64_t frame_offset, const double *dist, int32_t first_center_id,
                          eb_kc_state *state, void *partials, void *cand_out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_features > 0, "kcenters_seed_feat: bad shape");
    EB_CHECK_ARG(elem_size(dtype) != 0, "kcenters_seed_feat: unknown dtype");
    EB_CHECK_ARG(state && partials && cand_out, "kcenters_seed_feat: null pointer");
    EB_CUDA(cudaMemsetAsync(state, 0, sizeof(eb_kc_state), (cudaStream_t)stream));
    return dispatch<kFSeed>(dtype, EB_METRIC_EUCLIDEAN, X, (long)n, (long)n_features,
                            (long)frame_offset, (const void *)nullptr, 0,
                            const_cast<double *>(dist), (int *)nullptr, (int)first_center_id,
                            0.0, state, (int64_t *)nullptr, partials, cand_out,
                            (const void *)nullptr, (double *)nullptr, (cudaStream_t)stream);
}

int eb_feat_one_to_all(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                       const void *y, double *out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_features > 0, "feat_one_to_all: bad shape");
    EB_CHECK_ARG(elem_size(dtype) != 0, "feat_one_to_all: unknown dtype");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(X && y && out, "feat_one_to_all: null pointer");
    return dispatch<kFDistOnly>(dtype, metric, X, (long)n, (long)n_features, 0L,
                                (const void *)nullptr, 0, (double *)nullptr, (int *)nullptr, 0,
                                0.0, (eb_kc_state *)nullptr, (int64_t *)nullptr,
                                (void *)nullptr, (void *)nullptr, y, out, (cudaStream_t)stream);
}

int eb_feat_assign(const void *X, int64_t n, int64_t n_features, int dtype, int metric,
                   const void *centers, int32_t k, const int64_t *frame_idx, int64_t n_idx,
                   double *out_dist, int32_t *out_assign, int accumulate, int scatter,
                   void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_features > 0 && k >= 0 && n_idx >= 0, "feat_assign: bad shape");
    EB_CHECK_ARG(elem_size(dtype) != 0, "feat_assign: unknown dtype");
    const long m = frame_idx ? n_idx : n;
    if (m == 0 || k == 0) return EB_OK;
    EB_CHECK_ARG(X && centers && out_dist && out_assign, "feat_assign: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case EB_DT_F32:
            return dispatch_assign_metric<float>(metric, X, (long)n, (long)n_features, centers,
                                                 (int)k, frame_idx, m, out_dist, out_assign,
                                                 accumulate, scatter, s);
        case EB_DT_F64:
            return dispatch_assign_metric<double>(metric, X, (long)n, (long)n_features, centers,
                                                  (int)k, frame_idx, m, out_dist, out_assign,
                                                  accumulate, scatter, s);
        case EB_DT_I8:
            return dispatch_assign_metric<int8_t>(metric, X, (long)n, (long)n_features, centers,
                                                  (int)k, frame_idx, m, out_dist, out_assign,
                                                  accumulate, scatter, s);
        case EB_DT_I16:
            return dispatch_assign_metric<int16_t>(metric, X, (long)n, (long)n_features, centers,
                                                   (int)k, frame_idx, m, out_dist, out_assign,
                                                   accumulate, scatter, s);
        case EB_DT_I32:
            return dispatch_assign_metric<int32_t>(metric, X, (long)n, (long)n_features, centers,
                                                   (int)k, frame_idx, m, out_dist, out_assign,
                                                   accumulate, scatter, s);
        case EB_DT_I64:
            return dispatch_assign_metric<int64_t>(metric, X, (long)n, (long)n_features, centers,
                                                   (int)k, frame_idx, m, out_dist, out_assign,
                                                   accumulate, scatter, s);
    }
    return fail(EB_ERR_INVALID, "%s", "unknown dtype");
}

}  // extern "C"
