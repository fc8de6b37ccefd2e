// K5 (centring + trace, AoS -> SoA) and K1 (fused RMSD k-centers step) for sm_100a.
//
// Reference behaviour reproduced (paths under /root/reference/enspara/):
//   cluster/kcenters.py:243-311  _kcenters_iteration      (argmax, distance, strict-< update)
//   cluster/kcenters.py:314-378  _kcenters_iteration_mpi  (cross-shard argmax, centre bcast)
//   cluster/kcenters.py:217      stop rule
//   mdtraj.rmsd (third party)    centring, traces, Theobald QCP  
//
// Design: the step is HBM bound -- 12*A_pad + 12 bytes per frame are read
// exactly once.  A warp owns 32 consecutive frames; each group of 8 lanes streams one frame
// with 128-byte, sector-aligned float4 requests (4 frames in flight per warp-instruction),
// accumulates the 3x3 inner-product matrix in float64 against the centre staged in shared
// memory as float64, butterflies the 9 sums over the 8 lanes and parks them in shared memory;
// after 8 such rounds every lane solves the QCP quartic for ONE of the warp's 32 frames, so
// the Newton iterations run on full warps and the dist/assign update is a coalesced 128-byte
// read-modify-write.  The shard arg-max (first occurrence) is folded into the same launch
// through a last-block reduction that also publishes the candidate record for the next step.
#include <stdlib.h>

#include "eb_rmsd.cuh"
#include "eb_tma.cuh"

namespace eb {

thread_local char g_err[512] = "";

int sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return 148;
        cached = n > 0 ? n : 148;
    }
    return cached;
}

// ------------------------------------------------------------------------------------------
// K5: one warp per frame.  The AoS frame (12*A bytes, contiguous) is pulled into shared
// memory with coalesced loads, the centroid is accumulated in float64, and the centred
// float32 rows x|y|z are written out with coalesced stores; trace from the ROUNDED values.
// ------------------------------------------------------------------------------------------
constexpr int kCtrWarps = 8;

__global__ void __launch_bounds__(kCtrWarps * 32)
k_center_and_trace(const float *__restrict__ aos, long n, int A, int A_pad, int precentered,
                   float *__restrict__ soa, double *__restrict__ traces)
{
    extern __shared__ float sh_frames[];  // kCtrWarps * 3*A floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = sh_frames + (size_t)warp * 3 * A;
    const long warps_total = (long)gridDim.x * kCtrWarps;
    for (long f = (long)blockIdx.x * kCtrWarps + warp; f < n; f += warps_total) {
        const float *src = aos + (size_t)f * 3 * A;
        for (int t = lane; t < 3 * A; t += 32) buf[t] = __ldg(src + t);
        __syncwarp();
        double sx = 0, sy = 0, sz = 0;
        for (int a = lane; a < A; a += 32) {
            sx += (double)buf[3 * a];
            sy += (double)buf[3 * a + 1];
            sz += (double)buf[3 * a + 2];
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            sx += shfl_xor_d(sx, m);
            sy += shfl_xor_d(sy, m);
            sz += shfl_xor_d(sz, m);
        }
        const double mx = precentered ? 0.0 : sx / A;
        const double my = precentered ? 0.0 : sy / A;
        const double mz = precentered ? 0.0 : sz / A;
        float *ox = soa + (size_t)f * 3 * A_pad;
        float *oy = ox + A_pad, *oz = oy + A_pad;
        double g = 0;
        for (int a = lane; a < A_pad; a += 32) {
            float x = 0.f, y = 0.f, z = 0.f;
            if (a < A) {
                x = (float)((double)buf[3 * a] - mx);
                y = (float)((double)buf[3 * a + 1] - my);
                z = (float)((double)buf[3 * a + 2] - mz);
                g += (double)x * x + (double)y * y + (double)z * z;
            }
            ox[a] = x;
            oy[a] = y;
            oz[a] = z;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) g += shfl_xor_d(g, m);
        if (lane == 0) traces[f] = g;
        __syncwarp();
    }
}

__global__ void k_soa_to_aos(const float *__restrict__ soa, long n, int A, int A_pad,
                             float *__restrict__ aos)
{
    const long total = n * (long)A * 3;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long)gridDim.x * blockDim.x) {
        const long f = t / (3L * A);
        const int r = (int)(t - f * 3L * A);
        const int a = r / 3, c = r - 3 * a;
        aos[t] = soa[(size_t)f * 3 * A_pad + (size_t)c * A_pad + a];
    }
}

__global__ void k_gather_frames(const float *__restrict__ soa, const double *__restrict__ traces,
                                int A_pad, const int64_t *__restrict__ idx, long m,
                                float *__restrict__ out, double *__restrict__ out_tr)
{
    const int row = 3 * A_pad / 4;  // float4 per frame
    for (long j = blockIdx.x; j < m; j += gridDim.x) {
        const long f = idx[j];
        const float4 *s = reinterpret_cast<const float4 *>(soa + (size_t)f * 3 * A_pad);
        float4 *d = reinterpret_cast<float4 *>(out + (size_t)j * 3 * A_pad);
        for (int t = threadIdx.x; t < row; t += blockDim.x) d[t] = s[t];
        if (threadIdx.x == 0 && out_tr) out_tr[j] = traces[f];
    }
}

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
constexpr int kStepThreads = 256;
constexpr int kStepWarps = kStepThreads / 32;
constexpr int kSumStride = 9;  // doubles per frame slot; odd -> conflict-free 64-bit reads

enum StepMode { kModeStep = 0, kModeSeed = 1, kModeDistOnly = 2, kModeCC = 3 };
// VAR: 0 default, 1/2 developer load variants, 3 = triangle-inequality pruning
// (kcenters.py:287-296): frames with dist <= d(new centre, their centre) / 2 are not streamed.
constexpr int kVarTri = 3;

struct StepSmem {
    Partial red[32];
    int flag;
    int winner;
    double center_trace;
    double maxdist;
    int64_t center_index;
};

// ROUNDS: 4-frame rounds per warp chunk (a chunk is 4*ROUNDS consecutive frames).  8 for large
// shards (32 frames: coalesced 128-byte dist/assign updates); fewer when there are not enough
// frames to give every resident warp a full chunk (small trajectories, the k stored centres of
// the triangle mode, PAM's proposal-to-medoids distances), so the per-launch latency is one or
// two rounds instead of eight.
template <bool EXACT, int MODE, int VAR, int ROUNDS = 8>
__global__ void __launch_bounds__(kStepThreads, 2)
k_kcenters_step_rmsd(const float *__restrict__ xyz, const double *__restrict__ traces, long n,
                     int A, int A_pad, long frame_offset, const unsigned char *cand_in,
                     int n_cand, size_t rec_bytes, float *dist, int *assign,
                     int n_clusters_limit, double cutoff, eb_kc_state *state,
                     int64_t *center_list, Partial *partials, unsigned char *cand_out,
                     const float *center_direct, const double *center_trace_direct, float *out_only,
                     float *cstore, double *cstore_traces, const float *__restrict__ cc, Exch exch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem *ss = reinterpret_cast<StepSmem *>(smem_raw);
    double *sums = reinterpret_cast<double *>(smem_raw + align16(sizeof(StepSmem)));
    double *center_base = sums + kStepWarps * 32 * kSumStride;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, l8 = lane & 7;
    const int A4 = A_pad >> 2;

    int k = 0;
    double Gb = 0.0;
    CenterSmem cs = center_smem_carve(center_base, A4);

    const size_t rec_len = rec_bytes;   // bytes of one record (rec_bytes becomes the stride)
    if (MODE == kModeStep || MODE == kModeCC) {
        // ---- prologue: winner among the gathered candidates, stop rule -------------------
        if (exch.peers) {
            // peer-memory exchange: the records arrive in this rank's symmetric buffer; a run
            // that already stopped publishes nothing, so do not wait for anything either
            if (*reinterpret_cast<const volatile int32_t *>(&state->done)) {
                if (MODE == kModeStep && blockIdx.x == 0 && threadIdx.x == 0) state->n_noop += 1;
                return;
            }
            unsigned long long waited = 0;
            cand_in = exch_wait(exch, &waited);
            if (!cand_in) {   // a peer never published: fail the run instead of hanging
                if (threadIdx.x == 0) {
                    state->error = 1;
                    state->done = 1;
                }
                return;
            }
            if (blockIdx.x == 0 && threadIdx.x == 0) state->wait_ns += (long long)waited;
            n_cand = exch.size;
            rec_bytes = exch.rec_stride;
        }
        double cd;
        int64_t ci;
        const int r = pick_candidate(cand_in, n_cand, rec_bytes, cd, ci);
        const bool active = (r >= 0) && step_active(state, n_clusters_limit, cd, cutoff, k);
        if (!active) {
            if (MODE == kModeStep && blockIdx.x == 0 && threadIdx.x == 0) {
                if (!state->done) {
                    state->done = 1;
                    state->maxdist = cd;
                } else {
                    state->n_noop += 1;
                }
            }
            return;
        }
        const unsigned char *rec = cand_in + (size_t)r * rec_bytes;
        Gb = __ldcg(&reinterpret_cast<const RecHeader *>(rec)->trace);
        center_smem_fill(cs, reinterpret_cast<const float *>(rec + sizeof(RecHeader)), A_pad);
        if (threadIdx.x == 0) {
            ss->center_index = ci;
            ss->maxdist = cd;
        }
        if (MODE == kModeCC) {
            // distances from the new centre to the k centres chosen so far (the stored copies
            // are this launch's "frames"), and the new centre joins the store at slot k
            n = k;
            if (blockIdx.x == 0) {
                const float4 *src = reinterpret_cast<const float4 *>(rec + sizeof(RecHeader));
                float4 *dst = reinterpret_cast<float4 *>(cstore + (size_t)k * 3 * A_pad);
                for (int t = threadIdx.x; t < 3 * A4; t += blockDim.x) dst[t] = __ldcg(src + t);
                if (threadIdx.x == 0) cstore_traces[k] = Gb;
            }
        }
        __syncthreads();
    } else if (MODE == kModeDistOnly) {
        Gb = __ldg(center_trace_direct);
        center_smem_fill(cs, center_direct, A_pad);
        __syncthreads();
    }

    // ---- body -----------------------------------------------------------------------------
    double best_d = -2.0;
    int64_t best_i = INT64_MAX;
    constexpr int FPC = 4 * ROUNDS;                     // frames per chunk
    const long n_chunks = (n + FPC - 1) / FPC;
    const long warps_total = (long)gridDim.x * kStepWarps;
    double *my_sums = sums + (size_t)warp * 32 * kSumStride;

    for (long chunk = (long)blockIdx.x * kStepWarps + warp; chunk < n_chunks;
         chunk += warps_total) {
        const long base = chunk * FPC;
        // lanes beyond the chunk own no frame (ROUNDS < 8)
        const long f = (lane < FPC) ? base + lane : n;
        // triangle-inequality pruning: lane l decides for frame base + l, the warp shares the
        // decisions as a bit mask; rounds whose four frames are all pruned load nothing
        unsigned need_mask = 0xffffffffu;
        float tri_old = 0.f;
        if (VAR == kVarTri && MODE == kModeStep) {
            bool need = false;
            if (f < n) {
                tri_old = dist[f];
                const int a = assign[f];
                need = (a < 0) || (tri_old > 0.5f * __ldg(cc + a));   // kcenters.py:289
            }
            need_mask = __ballot_sync(0xffffffffu, need);
        }
        if (VAR == kVarTri && MODE == kModeDistOnly) {
            // PAM full pass: a frame of another cluster whose distance to its medoid is at most
            // (1 - 1e-5)/2 of that medoid's distance to the proposal cannot get closer to the
            // proposal than it is to its medoid (triangle inequality; the margin keeps the
            // float32 rounding of both distances, ~1e-7 relative, out of the decision), so the
            // split of kmedoids.py:644-658 leaves it untouched: report +inf, do not read it.
            // n_clusters_limit carries the proposal's cluster id.
            bool need = false;
            if (f < n) {
                const int a = assign[f];
                need = (a < 0) || (a == n_clusters_limit) ||
                       !(dist[f] <= 0.499995f * __ldg(cc + a));
            }
            need_mask = __ballot_sync(0xffffffffu, need);
        }
        if (MODE != kModeSeed) {
#pragma unroll 1
            for (int s = 0; s < ROUNDS; ++s) {
                if (VAR == kVarTri && ((need_mask >> (4 * s)) & 0xfu) == 0u) continue;
                const long fs = base + 4 * s + g;
                double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                if (VAR == 2) {
                    // software prefetch of the frame this group streams in the next round
                    // (or the first round of the warp's next chunk) into L2
                    const long fn = (s < ROUNDS - 1) ? fs + 4 : (chunk + warps_total) * FPC + g;
                    if (fn < n) {
                        const char *pn = reinterpret_cast<const char *>(xyz + (size_t)fn * 3 * A_pad);
                        const int n_lines = (12 * A_pad + 127) >> 7;
                        for (int l = l8; l < n_lines; l += 8) prefetch_l2(pn + ((size_t)l << 7));
                    }
                }
                if (fs < n && (VAR != kVarTri || ((need_mask >> (4 * s + g)) & 1u)))
                    frame_inner_products<EXACT, (VAR == 1 ? 1 : 0)>(
                        m, xyz + (size_t)fs * 3 * A_pad, A4, l8, cs);
                group8_reduce(m);
                if (l8 == 0) {
                    double *dst = my_sums + (4 * s + g) * kSumStride;
#pragma unroll
                    for (int e = 0; e < 9; ++e) dst[e] = m[e];
                }
            }
            __syncwarp();
        }
        if (f < n) {
            double cur;
            if (MODE == kModeSeed) {
                cur = (double)dist[f];
            } else if (VAR == kVarTri && MODE == kModeStep && !((need_mask >> lane) & 1u)) {
                cur = (double)tri_old;     // pruned: distance and assignment stay
            } else if (VAR == kVarTri && MODE == kModeDistOnly && !((need_mask >> lane) & 1u)) {
                out_only[f] = INFINITY;    // pruned: provably not closer to the proposal
                cur = 0.0;
            } else {
                double m[9];
                const double *src = my_sums + lane * kSumStride;
#pragma unroll
                for (int e = 0; e < 9; ++e) m[e] = src[e];
                const float d = rmsd_from_msd(qcp_msd(m, traces[f], Gb, A));
                if (MODE == kModeDistOnly || MODE == kModeCC) {
                    out_only[f] = d;
                    cur = 0.0;
                } else {
                    const float old = dist[f];
                    if (d < old) {  // strict '<', kcenters.py:304
                        dist[f] = d;
                        assign[f] = k;
                    }
                    cur = (double)((d < old) ? d : old);
                }
            }
            if (cur > best_d) {  // frames arrive in increasing f per lane: first max kept
                best_d = cur;
                best_i = f;
            }
        }
        if (MODE != kModeSeed) __syncwarp();
    }
    if (MODE == kModeDistOnly || MODE == kModeCC) return;

    // ---- epilogue: shard arg-max, candidate record, centre list ----------------------------
    if (!grid_argmax_last_block(best_d, best_i, partials, state, ss->red, &ss->flag)) return;

    RecHeader *out = reinterpret_cast<RecHeader *>(cand_out);
    const bool empty = (best_i == INT64_MAX);
    if (!empty) {
        const float4 *src = reinterpret_cast<const float4 *>(xyz + (size_t)best_i * 3 * A_pad);
        float4 *dst = reinterpret_cast<float4 *>(cand_out + sizeof(RecHeader));
        for (int t = threadIdx.x; t < 3 * A4; t += blockDim.x) dst[t] = __ldcg(src + t);
    }
    if (threadIdx.x == 0) {
        out->dist = empty ? -1.0 : best_d;
        out->index = empty ? -1 : frame_offset + best_i;
        out->trace = empty ? 0.0 : traces[best_i];
        out->reserved = 0;
        if (MODE == kModeStep) {
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
    if (exch.peers) exch_publish(exch, cand_out, rec_len);
}

// ------------------------------------------------------------------------------------------
// K1, persistent multi-iteration variant (single shard, shards that do not take the TMA
// kernel): up to n_steps iterations of kcenters.py:217-226 in ONE cooperative launch.  On a
// small trajectory the data is L2-resident and an iteration is a few microseconds of work under
// ~20-30 us of per-launch cost (launch gap, prologue round trips, the single-CTA arg-max tail:
// 39.5 us per iteration at config 1, 20 000 frames x 264 atoms = 63 MB).  Here the grid stays
// resident: the shard arg-max is a grid barrier after which EVERY block reduces the per-block
// partials itself and stages the winner's frame straight from xyz.  Same arithmetic, summation
// order, tie rules and state protocol as k_kcenters_step_rmsd<true, kModeStep, 0, ROUNDS>, so
// results are bit-identical and single launches and multi launches mix freely.
// gbar: monotonically increasing arrival counter (a multiple of gridDim.x between launches).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int ROUNDS>
__global__ void __launch_bounds__(kStepThreads, 2)
k_kcenters_multi_rmsd(const float *__restrict__ xyz, const double *__restrict__ traces, long n,
                      int A, int A_pad, long frame_offset, unsigned char *cand, float *dist,
                      int *assign, int n_clusters_limit, double cutoff, eb_kc_state *state,
                      int64_t *center_list, Partial *partials, unsigned long long *gbar,
                      int n_steps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem *ss = reinterpret_cast<StepSmem *>(smem_raw);
    double *sums = reinterpret_cast<double *>(smem_raw + align16(sizeof(StepSmem)));
    double *center_base = sums + kStepWarps * 32 * kSumStride;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, l8 = lane & 7;
    const int A4 = A_pad >> 2;
    const unsigned G = gridDim.x;
    CenterSmem cs = center_smem_carve(center_base, A4);

    // nobody writes *state / the counter's base before the first barrier: these reads agree
    const int k0 = *reinterpret_cast<const volatile int32_t *>(&state->n_centers);
    const int done0 = *reinterpret_cast<const volatile int32_t *>(&state->done);
    const unsigned long long gbase = (ld_acquire_gpu(gbar) / G) * G;

    constexpr int FPC = 4 * ROUNDS;
    const long n_chunks = (n + FPC - 1) / FPC;
    const long warps_total = (long)G * kStepWarps;
    double *my_sums = sums + (size_t)warp * 32 * kSumStride;

    int it = 0;
    bool stopped = false;
    double win_d = 0.0;
    int64_t win_i = -1;
    for (; it < n_steps; ++it) {
        // ---- the centre of this iteration: candidate record (first) or the partials ---------
        const float *src;
        double Gb;
        bool have;
        if (it == 0) {
            const RecHeader *h = reinterpret_cast<const RecHeader *>(cand);
            win_d = __ldcg(&h->dist);
            win_i = __ldcg(reinterpret_cast<const long long *>(&h->index));
            Gb = __ldcg(&h->trace);
            have = win_i >= 0;
            src = reinterpret_cast<const float *>(cand + sizeof(RecHeader));
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
            src = xyz + (size_t)(have ? i : 0) * 3 * A_pad;
            Gb = have ? traces[i] : 0.0;
        }
        const int k = k0 + it;
        if (!(have && !done0 && k < n_clusters_limit && win_d > cutoff)) {   // kcenters.py:217
            stopped = true;
            break;
        }
        center_smem_fill(cs, src, A_pad);
        if (blockIdx.x == 0 && threadIdx.x == 0) center_list[k] = win_i;
        __syncthreads();

        // ---- body: identical to k_kcenters_step_rmsd<true, kModeStep, 0, ROUNDS> -------------
        double best_d = -2.0;
        int64_t best_i = INT64_MAX;
        for (long chunk = (long)blockIdx.x * kStepWarps + warp; chunk < n_chunks;
             chunk += warps_total) {
            const long base = chunk * FPC;
            const long f = (lane < FPC) ? base + lane : n;
            // requested before the inner products so that their latency hides behind them
            // (the same lane wrote dist[f] in the previous iteration: static chunk -> warp map)
            const double Ga_f = (f < n) ? traces[f] : 0.0;
            const float old = (f < n) ? dist[f] : 0.0f;
#pragma unroll 1
            for (int s = 0; s < ROUNDS; ++s) {
                const long fs = base + 4 * s + g;
                double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                if (fs < n)
                    frame_inner_products<true, 0>(m, xyz + (size_t)fs * 3 * A_pad, A4, l8, cs);
                group8_reduce(m);
                if (l8 == 0) {
                    double *dst = my_sums + (4 * s + g) * kSumStride;
#pragma unroll
                    for (int e = 0; e < 9; ++e) dst[e] = m[e];
                }
            }
            __syncwarp();
            if (f < n) {
                double m[9];
                const double *srcm = my_sums + lane * kSumStride;
#pragma unroll
                for (int e = 0; e < 9; ++e) m[e] = srcm[e];
                const float d = rmsd_from_msd(qcp_msd(m, Ga_f, Gb, A));
                if (d < old) {  // strict '<', kcenters.py:304
                    dist[f] = d;
                    assign[f] = k;
                }
                const double cur = (double)((d < old) ? d : old);
                if (cur > best_d) {  // frames arrive in increasing f per lane: first max kept
                    best_d = cur;
                    best_i = f;
                }
            }
            __syncwarp();
        }
        // ---- shard arg-max: per-block partial, grid barrier ------------------------------------
        block_argmax(best_d, best_i, ss->red);
        if (threadIdx.x == 0) {
            Partial *pp = partials + (size_t)(it & 1) * G;
            pp[blockIdx.x].dist = best_d;
            pp[blockIdx.x].index = best_i;
            __threadfence();
            atomicAdd(gbar, 1ull);
            const unsigned long long target = gbase + (unsigned long long)(it + 1) * G;
            while (ld_acquire_gpu(gbar) < target) __nanosleep(32);
        }
        __syncthreads();
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
            const float4 *srow = reinterpret_cast<const float4 *>(xyz + (size_t)i * 3 * A_pad);
            float4 *dst = reinterpret_cast<float4 *>(cand + sizeof(RecHeader));
            for (int t = threadIdx.x; t < 3 * A4; t += blockDim.x) dst[t] = __ldcg(srow + t);
        }
        if (threadIdx.x == 0) {
            out->dist = empty ? -1.0 : d;
            out->index = empty ? -1 : frame_offset + i;
            out->trace = empty ? 0.0 : traces[i];
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
            state->maxdist = win_d;
        }
        state->blocks_done = 0;
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------
// K1, TMA-staged variant of the step: same contract, same arithmetic and the SAME summation
// order as k_kcenters_step_rmsd<true, kModeStep, 0, 8> (8 lanes per frame, lane l8 adds atoms
// 4j..4j+3 for j = l8, l8+8, ... in increasing j, butterfly over the 8 lanes), so results are
// bit-identical.  The LDG version keeps its loads in flight in registers (6 x 16 B per lane at
// 128 registers, 16 warps/SM ~ 49 KB per SM), which is what bounds it (long_scoreboard, 74.7 %
// of DRAM peak).  Here every warp owns a two-stage ring of shared-memory buffers; a stage holds
// one HALF (in atoms) of the four frames of a round: twelve cp.async.bulk copies of 2*A_pad
// bytes (4 frames x 3 coordinate rows), issued by twelve lanes, one mbarrier.  While the warp
// (first version; 1.235 ms) -- now ONE 2-D tensor copy (cp.async.bulk.tensor.2d, box = 12 rows
// x half a row of the (3n x A_pad) matrix).  While the warp
// accumulates a half out of shared memory the next half is in flight: ~12 KB per warp, 96 KB
// per SM, and no register is held by a load.  One CTA of 8 warps per SM.
// ------------------------------------------------------------------------------------------
// P: parts a coordinate row is cut into (a stage holds A_pad / P atoms of four frames);
// S: ring depth (S - 1 stages are in flight while one is consumed).
// MODE: kModeStep (the k-centers iteration) or kModeDistOnly (md.rmsd(X, y): the centre comes
// from center_direct / center_trace_direct and out_only[f] receives the distance).
// FR: frames per 8-lane group and stage (1 or 2).  With FR = 2 a stage holds EIGHT frames (box of
// 24 rows) and a group accumulates two of them against ONE read of the float64 centre from
// shared memory: 12 instead of 18 LDS.128 per two frames (ncu had the L1TEX data pipe at 65 %,
// two thirds of it the centre).  Each frame's own summation order is unchanged, so results stay
// bit-identical.
template <int P, int S, int MODE = kModeStep, int FR = 1>
__global__ void __launch_bounds__(kStepThreads, 1)
k_kcenters_step_rmsd_tma(const float *__restrict__ xyz, const double *__restrict__ traces, long n,
                         int A, int A_pad, long frame_offset, const unsigned char *cand_in,
                         int n_cand, size_t rec_bytes, float *dist, int *assign,
                         int n_clusters_limit, double cutoff, eb_kc_state *state,
                         int64_t *center_list, Partial *partials, unsigned char *cand_out,
                         const __grid_constant__ CUtensorMap tmap, const float *center_direct,
                         const double *center_trace_direct, float *out_only, Exch exch)
{
    extern __shared__ __align__(128) unsigned char smem_tma[];
    constexpr int kFPR = 4 * FR;                                 // frames per round (= stage)
    constexpr int kStagesPerChunk = (32 / kFPR) * P;             // rounds x P parts
    const uint32_t part_row_bytes = 4u * (uint32_t)A_pad / P;    // one part of a coordinate row
    const uint32_t stage_bytes = 3u * kFPR * part_row_bytes;     // kFPR frames x 3 rows
    unsigned char *ring = smem_tma;                              // [warps][S][stage_bytes]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)kStepWarps * S * stage_bytes);
    StepSmem *ss = reinterpret_cast<StepSmem *>(bars + kStepWarps * S);
    double *sums = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ss) +
                                              align16(sizeof(StepSmem)));
    double *center_base = sums + kStepWarps * 32 * kSumStride;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, l8 = lane & 7;
    const int A4 = A_pad >> 2;
    const int A4p = A4 / P;                                      // float4 per part of a row
    unsigned char *my_ring = ring + (size_t)warp * S * stage_bytes;
    uint64_t *bar = bars + warp * S;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < S; ++b) mbar_init(&bar[b], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }

    // ---- prologue: winner among the gathered candidates, stop rule -------------------------
    int k = 0;
    CenterSmem cs = center_smem_carve(center_base, A4);
    double Gb;
    const size_t rec_len = rec_bytes;   // bytes of one record (rec_bytes becomes the stride)
    if (MODE == kModeStep) {
        if (exch.peers) {
            if (*reinterpret_cast<const volatile int32_t *>(&state->done)) {
                if (blockIdx.x == 0 && threadIdx.x == 0) state->n_noop += 1;
                return;
            }
            unsigned long long waited = 0;
            cand_in = exch_wait(exch, &waited);
            if (!cand_in) {   // a peer never published: fail the run instead of hanging
                if (threadIdx.x == 0) {
                    state->error = 1;
                    state->done = 1;
                }
                return;
            }
            if (blockIdx.x == 0 && threadIdx.x == 0) state->wait_ns += (long long)waited;
            n_cand = exch.size;
            rec_bytes = exch.rec_stride;
        }
        double cd;
        int64_t ci;
        const int r_win = pick_candidate(cand_in, n_cand, rec_bytes, cd, ci);
        const bool active = (r_win >= 0) && step_active(state, n_clusters_limit, cd, cutoff, k);
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
        const unsigned char *rec = cand_in + (size_t)r_win * rec_bytes;
        Gb = __ldcg(&reinterpret_cast<const RecHeader *>(rec)->trace);
        center_smem_fill(cs, reinterpret_cast<const float *>(rec + sizeof(RecHeader)), A_pad);
        if (threadIdx.x == 0) {
            ss->center_index = ci;
            ss->maxdist = cd;
        }
    } else {
        Gb = __ldg(center_trace_direct);
        center_smem_fill(cs, center_direct, A_pad);
    }
    __syncthreads();

    // ---- body ---------------------------------------------------------------------------------
    double best_d = -2.0;
    int64_t best_i = INT64_MAX;
    const long n_chunks = (n + 31) >> 5;
    const long warps_total = (long)gridDim.x * kStepWarps;
    double *my_sums = sums + (size_t)warp * 32 * kSumStride;

    // stage st of chunk c = part (st % P) of the four frames of round (st / P): ONE 2-D tensor
    // copy, box = 12 rows (4 frames x 3 coordinate rows of the (3n x A_pad) matrix) x one part
    // of a row; rows past the last frame are zero-filled by the TMA unit.  Nothing is requested
    // -- and nothing will be waited for -- when the whole round lies beyond frame n.  Stages
    // are numbered q = 0, 1, ... per warp: buffer q % S, mbarrier phase (q / S) & 1.
    long i_chunk = (long)blockIdx.x * kStepWarps + warp;   // producer cursor
    int i_st = 0;
    unsigned q_issue = 0;
    auto issue = [&]() {
        if (i_chunk < n_chunks) {
            const long f0 = (i_chunk << 5) + kFPR * (i_st / P);
            if (lane == 0 && f0 < n) {
                const int b = (int)(q_issue % S);
                mbar_expect_tx(&bar[b], stage_bytes);
                tma_load_2d(my_ring + (size_t)b * stage_bytes, &tmap, (i_st % P) * (A_pad / P),
                            (int)(f0 * 3), &bar[b]);
            }
            ++q_issue;
            if (++i_st == kStagesPerChunk) {
                i_st = 0;
                i_chunk += warps_total;
            }
        }
    };
#pragma unroll
    for (int pre = 0; pre < S - 1; ++pre) issue();

    unsigned q = 0;                                        // consumer stage number
    for (long chunk = (long)blockIdx.x * kStepWarps + warp; chunk < n_chunks;
         chunk += warps_total) {
        const long base = chunk << 5;
        double m[9], m2[9];
#pragma unroll 1
        for (int st = 0; st < kStagesPerChunk; ++st, ++q) {
            const int b = (int)(q % S);
            const int part = st % P;
            // the buffer consumed in the previous stage is free (all lanes are past its
            // __syncwarp): request the stage S - 1 ahead into it
            issue();
            const long f0 = base + kFPR * (st / P);
            if (f0 < n) mbar_wait(&bar[b], (q / S) & 1u);
            if (part == 0) {
#pragma unroll
                for (int e = 0; e < 9; ++e) m[e] = 0.0;
                if (FR == 2) {
#pragma unroll
                    for (int e = 0; e < 9; ++e) m2[e] = 0.0;
                }
            }
            const long fs = f0 + g;
            if (fs < n) {
                // rows of this group's frame(s) in the stage: pitch = one part of a row
                const float4 *px = reinterpret_cast<const float4 *>(
                    my_ring + (size_t)b * stage_bytes + (size_t)(3 * g) * part_row_bytes);
                const float4 *py = px + A4p;
                const float4 *pz = py + A4p;
                const int lo = part * A4p, hi = lo + A4p;
                // first j >= lo with j == l8 (mod 8): the LDG kernel's per-lane order
                int j = lo + ((l8 - lo) & 7);
                if (FR == 1) {
#pragma unroll 2
                    for (; j < hi; j += 8) {
                        const int jj = j - lo;
                        const float4 x = px[jj], y = py[jj], z = pz[jj];
                        const double2 cxl = cs.lo[0][j], cxh = cs.hi[0][j];
                        const double2 cyl = cs.lo[1][j], cyh = cs.hi[1][j];
                        const double2 czl = cs.lo[2][j], czh = cs.hi[2][j];
                        acc_atom(m, x.x, y.x, z.x, cxl.x, cyl.x, czl.x);
                        acc_atom(m, x.y, y.y, z.y, cxl.y, cyl.y, czl.y);
                        acc_atom(m, x.z, y.z, z.z, cxh.x, cyh.x, czh.x);
                        acc_atom(m, x.w, y.w, z.w, cxh.y, cyh.y, czh.y);
                    }
                } else {
                    // second frame of the group: four frames (12 rows) further down the stage;
                    // rows past the last frame were zero-filled by the TMA unit
                    const float4 *qx = px + 12 * A4p;
                    const float4 *qy = qx + A4p;
                    const float4 *qz = qy + A4p;
#pragma unroll 2
                    for (; j < hi; j += 8) {
                        const int jj = j - lo;
                        const float4 x = px[jj], y = py[jj], z = pz[jj];
                        const float4 u = qx[jj], v = qy[jj], w = qz[jj];
                        const double2 cxl = cs.lo[0][j], cxh = cs.hi[0][j];
                        const double2 cyl = cs.lo[1][j], cyh = cs.hi[1][j];
                        const double2 czl = cs.lo[2][j], czh = cs.hi[2][j];
                        acc_atom(m, x.x, y.x, z.x, cxl.x, cyl.x, czl.x);
                        acc_atom(m, x.y, y.y, z.y, cxl.y, cyl.y, czl.y);
                        acc_atom(m, x.z, y.z, z.z, cxh.x, cyh.x, czh.x);
                        acc_atom(m, x.w, y.w, z.w, cxh.y, cyh.y, czh.y);
                        acc_atom(m2, u.x, v.x, w.x, cxl.x, cyl.x, czl.x);
                        acc_atom(m2, u.y, v.y, w.y, cxl.y, cyl.y, czl.y);
                        acc_atom(m2, u.z, v.z, w.z, cxh.x, cyh.x, czh.x);
                        acc_atom(m2, u.w, v.w, w.w, cxh.y, cyh.y, czh.y);
                    }
                }
            }
            if (part == P - 1) {
                group8_reduce(m);
                if (FR == 2) group8_reduce(m2);
                if (l8 == 0) {
                    double *dst = my_sums + (kFPR * (st / P) + g) * kSumStride;
#pragma unroll
                    for (int e = 0; e < 9; ++e) dst[e] = m[e];
                    if (FR == 2) {
                        double *dst2 = dst + 4 * kSumStride;
#pragma unroll
                        for (int e = 0; e < 9; ++e) dst2[e] = m2[e];
                    }
                }
            }
            __syncwarp();
        }
        const long f = base + lane;
        if (f < n) {
            double mm[9];
            const double *src = my_sums + lane * kSumStride;
#pragma unroll
            for (int e = 0; e < 9; ++e) mm[e] = src[e];
            const float d = rmsd_from_msd(qcp_msd(mm, traces[f], Gb, A));
            if (MODE == kModeDistOnly) {
                out_only[f] = d;
            } else {
                const float old = dist[f];
                if (d < old) {  // strict '<', kcenters.py:304
                    dist[f] = d;
                    assign[f] = k;
                }
                const double cur = (double)((d < old) ? d : old);
                if (cur > best_d) {
                    best_d = cur;
                    best_i = f;
                }
            }
        }
        __syncwarp();
    }
    if (MODE == kModeDistOnly) return;

    // ---- epilogue: shard arg-max, candidate record, centre list ----------------------------
    if (!grid_argmax_last_block(best_d, best_i, partials, state, ss->red, &ss->flag)) return;

    RecHeader *out = reinterpret_cast<RecHeader *>(cand_out);
    const bool empty = (best_i == INT64_MAX);
    if (!empty) {
        const float4 *src = reinterpret_cast<const float4 *>(xyz + (size_t)best_i * 3 * A_pad);
        float4 *dst = reinterpret_cast<float4 *>(cand_out + sizeof(RecHeader));
        for (int t = threadIdx.x; t < 3 * A4; t += blockDim.x) dst[t] = __ldcg(src + t);
    }
    if (threadIdx.x == 0) {
        out->dist = empty ? -1.0 : best_d;
        out->index = empty ? -1 : frame_offset + best_i;
        out->trace = empty ? 0.0 : traces[best_i];
        out->reserved = 0;
        center_list[k] = ss->center_index;
        state->n_centers = k + 1;
        state->last_center = ss->center_index;
        state->maxdist = ss->maxdist;
        state->local_maxdist = empty ? -1.0 : best_d;
        state->blocks_done = 0;
        __threadfence();
    }
    if (exch.peers) exch_publish(exch, cand_out, rec_len);
}

static size_t step_tma_smem_bytes(int A_pad, int P, int S, int FR = 1)
{
    return (size_t)kStepWarps * S * 12 * FR * (4 * (size_t)A_pad / P) +
           sizeof(uint64_t) * kStepWarps * S + align16(sizeof(StepSmem)) +
           sizeof(double) * kStepWarps * 32 * kSumStride + sizeof(double) * 3 * (size_t)A_pad;
}

// Ring configuration of the TMA-staged kernel for rows of A_pad floats: P parts per row (whole
// 128-byte lines, <= 256 floats, P in 1..4) and S stages per warp.  Big copies win (measured at
// 500 atoms: (P,S) = (2,2) 1.11 ms, (4,4) 1.22, (8,8) 1.56), so the smallest P whose 2-stage
// ring fits shared memory is taken; when a stage is small (part rows under 1 KB) the ring is
// deepened to 4 where shared memory allows.  Parts under 640 bytes are not worth it (see
// rmsd_tma_parts in eb_common.cuh).  Returns false when the row cannot be cut.
static bool tma_config(int A_pad, int *P_out, int *S_out)
{
    if (A_pad <= 0 || (A_pad & 31)) return false;
    const int m = A_pad >> 5;
    for (int P = 1; P <= 4; ++P) {
        if (m % P != 0 || m / P > 8 || m / P < 5) continue;   // parts of 160..256 floats
        if (step_tma_smem_bytes(A_pad, P, 2) > 227 * 1024) continue;
        int S = 2;
        const size_t part_row_bytes = 4 * (size_t)A_pad / P;
        if (part_row_bytes < 1024 && step_tma_smem_bytes(A_pad, P, 4) <= 227 * 1024) S = 4;
        *P_out = P;
        *S_out = S;
        return true;
    }
    return false;
}
static int k1_tma_enabled();
template <int MODE>
static auto tma_kernel(int P, int S) -> decltype(&k_kcenters_step_rmsd_tma<1, 2, MODE>)
{
    if (S == 4) {
        switch (P) {
            case 1: return k_kcenters_step_rmsd_tma<1, 4, MODE>;
            case 2: return k_kcenters_step_rmsd_tma<2, 4, MODE>;
            case 3: return k_kcenters_step_rmsd_tma<3, 4, MODE>;
            default: return k_kcenters_step_rmsd_tma<4, 4, MODE>;
        }
    }
    switch (P) {
        case 1: return k_kcenters_step_rmsd_tma<1, 2, MODE>;
        case 2: return k_kcenters_step_rmsd_tma<2, 2, MODE>;
        case 3: return k_kcenters_step_rmsd_tma<3, 2, MODE>;
        default: return k_kcenters_step_rmsd_tma<4, 2, MODE>;
    }
}
static int pick_rounds(long n);
static int k1_variant();

// whether the exact step of a shard of n frames runs the TMA-staged kernel (see the dispatch in
// eb_kcenters_step_rmsd for the reasons behind each condition)
static bool step_uses_tma(long n, int n_atoms)
{
    const int A_pad = rmsd_apad(n_atoms);
    int P = 0, S = 0;
    return k1_tma_enabled() && k1_variant() == 0 && tma_config(A_pad, &P, &S) &&
           pick_rounds(n) == 8 && 3 * n < (int64_t(1) << 31);
}

// TMA-staged kernel for large shards (default); EB_K1_TMA=0 forces the LDG kernel (A/B switch)
static int k1_tma_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("EB_K1_TMA");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static size_t step_smem_bytes(int A_pad)
{
    return align16(sizeof(StepSmem)) + sizeof(double) * kStepWarps * 32 * kSumStride +
           sizeof(double) * 3 * (size_t)A_pad;
}

static int step_grid(long n, int frames_per_chunk = 32)
{
    const long chunks = (n + frames_per_chunk - 1) / frames_per_chunk;
    long blocks = (chunks + kStepWarps - 1) / kStepWarps;
    const long cap = 2L * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// 4-frame rounds per warp chunk: the largest of 8/4/2/1 that still gives every resident warp
// (2 CTAs x 8 warps per SM) a chunk
static int pick_rounds(long n)
{
    const long warps = 2L * sm_count() * kStepWarps;
    int r = 8;
    while (r > 1 && (n + 4 * r - 1) / (4 * r) < warps) r >>= 1;
    return r;
}

static int k1_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("EB_K1_VARIANT");  // developer A/B switch, see DESIGN.md
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 2) v = 0;
    }
    return v;
}

// optional extras of a launch: the centre store + centre-centre distances of the triangle mode
struct StepExtra {
    float *cstore = nullptr;
    double *cstore_traces = nullptr;
    float *cc = nullptr;
    long grid_frames = -1;   // size the grid for this many frames instead of n (kModeCC)
    Exch exch = {nullptr, 0, 0, 0};
};

template <bool EXACT, int MODE, int VAR, int ROUNDS>
static int launch_step_r(const float *xyz, const double *traces, long n, int A, long frame_offset,
                         const void *cand_in, int n_cand, float *dist, int *assign,
                         int n_clusters_limit, double cutoff, eb_kc_state *state,
                         int64_t *center_list, void *partials, void *cand_out,
                         const float *center_direct, const double *center_trace_direct,
                         float *out_only, cudaStream_t stream, StepExtra ex)
{
    const int A_pad = rmsd_apad(A);
    const size_t smem = step_smem_bytes(A_pad);
    if (smem > 227 * 1024)
        return fail(EB_ERR_LIMIT, "%s: n_atoms=%ld needs %ld bytes of shared memory (max 232448)",
                    "rmsd step", (long)A, (long)smem);
    auto kern = k_kcenters_step_rmsd<EXACT, MODE, VAR, ROUNDS>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        configured = smem;
    }
    const size_t rec_bytes = sizeof(RecHeader) + sizeof(float) * 3 * (size_t)A_pad;
    kern<<<step_grid(ex.grid_frames >= 0 ? ex.grid_frames : n, 4 * ROUNDS), kStepThreads, smem,
           stream>>>(
        xyz, traces, n, A, A_pad, frame_offset, (const unsigned char *)cand_in, n_cand, rec_bytes,
        dist, assign, n_clusters_limit, cutoff, state, center_list, (Partial *)partials,
        (unsigned char *)cand_out, center_direct, center_trace_direct, out_only, ex.cstore,
        ex.cstore_traces, ex.cc, ex.exch);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

template <bool EXACT, int MODE, int VAR>
static int launch_step_v(const float *xyz, const double *traces, long n, int A, long frame_offset,
                       const void *cand_in, int n_cand, float *dist, int *assign,
                       int n_clusters_limit, double cutoff, eb_kc_state *state,
                       int64_t *center_list, void *partials, void *cand_out,
                       const float *center_direct, const double *center_trace_direct,
                       float *out_only, cudaStream_t stream, StepExtra ex = StepExtra())
{
#define EB_STEP_ARGS                                                                              \
    xyz, traces, n, A, frame_offset, cand_in, n_cand, dist, assign, n_clusters_limit, cutoff,     \
        state, center_list, partials, cand_out, center_direct, center_trace_direct, out_only,     \
        stream, ex
    // short chunks only where they matter: the exact kernels without developer load variants
    if (EXACT && MODE != kModeSeed && (VAR == 0 || VAR == kVarTri)) {
        switch (pick_rounds(ex.grid_frames >= 0 ? ex.grid_frames : n)) {
            case 1: return launch_step_r<EXACT, MODE, (VAR == kVarTri ? kVarTri : 0), 1>(EB_STEP_ARGS);
            case 2: return launch_step_r<EXACT, MODE, (VAR == kVarTri ? kVarTri : 0), 2>(EB_STEP_ARGS);
            case 4: return launch_step_r<EXACT, MODE, (VAR == kVarTri ? kVarTri : 0), 4>(EB_STEP_ARGS);
        }
    }
    return launch_step_r<EXACT, MODE, VAR, 8>(EB_STEP_ARGS);
#undef EB_STEP_ARGS
}

// Persistent multi-iteration launch (single shard, exact arithmetic, non-TMA shapes).  *used = 0
// when not eligible; the caller then queues single steps.
template <int ROUNDS>
static int launch_multi_r(const float *xyz, const double *traces, long n, int A, long frame_offset,
                          void *cand, float *dist, int *assign, int limit, double cutoff,
                          eb_kc_state *state, int64_t *center_list, void *partials, int n_steps,
                          cudaStream_t stream, int *used)
{
    *used = 0;
    int A_pad = rmsd_apad(A);
    const size_t smem = step_smem_bytes(A_pad);
    if (smem > 227 * 1024) return EB_OK;
    static int coop = -1;
    if (coop < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
        coop = v;
    }
    if (!coop) return EB_OK;
    auto kern = k_kcenters_multi_rmsd<ROUNDS>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        configured = smem;
    }
    int per_sm = 0;
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStepThreads, smem));
    if (per_sm < 1) return EB_OK;
    int grid = step_grid(n, 4 * ROUNDS);
    const long cap = (long)per_sm * sm_count();
    if (grid > cap) grid = (int)cap;
    if (2 * grid > kMaxGrid) return EB_OK;
    unsigned char *candp = (unsigned char *)cand;
    Partial *pp = (Partial *)partials;
    unsigned long long *gbar = reinterpret_cast<unsigned long long *>(&state->reserved);
    void *args[] = {&xyz, &traces, &n, &A, &A_pad, &frame_offset, &candp, &dist, &assign, &limit,
                    &cutoff, &state, &center_list, &pp, &gbar, &n_steps};
    EB_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(kStepThreads), args,
                                        smem, stream));
    *used = 1;
    return EB_OK;
}

template <bool EXACT, int MODE, typename... Args> static int launch_step(Args... args)
{
    if (MODE == kModeSeed) return launch_step_v<EXACT, MODE, 0>(args...);
    switch (k1_variant()) {
        case 1: return launch_step_v<EXACT, MODE, 1>(args...);
        case 2: return launch_step_v<EXACT, MODE, 2>(args...);
    }
    return launch_step_v<EXACT, MODE, 0>(args...);
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_version(void) { return 100; }
const char *eb_last_error(void) { return eb::g_err; }
int eb_sm_count(void) { return eb::sm_count(); }

int eb_rmsd_apad(int n_atoms) { return rmsd_apad(n_atoms); }
int eb_kcenters_step_rmsd_uses_tma(int64_t n, int n_atoms) { return step_uses_tma(n, n_atoms); }
size_t eb_rmsd_record_bytes(int n_atoms)
{
    return sizeof(RecHeader) + sizeof(float) * 3 * (size_t)rmsd_apad(n_atoms);
}
size_t eb_kc_partials_bytes(void) { return sizeof(Partial) * kMaxGrid; }

int eb_center_and_trace(const float *xyz_aos, int64_t n, int n_atoms, int precentered,
                        float *xyz_soa, double *traces, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "center_and_trace: need n >= 0 and n_atoms > 0");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(xyz_aos && xyz_soa && traces, "center_and_trace: null pointer");
    const size_t smem = sizeof(float) * kCtrWarps * 3 * (size_t)n_atoms;
    if (smem > 227 * 1024)
        return fail(EB_ERR_LIMIT, "%s: n_atoms=%ld too large for the centring kernel",
                    "center_and_trace", (long)n_atoms);
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(k_center_and_trace,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    long blocks = (n + kCtrWarps - 1) / kCtrWarps;
    const long cap = 8L * sm_count();
    if (blocks > cap) blocks = cap;
    k_center_and_trace<<<(int)blocks, kCtrWarps * 32, smem, (cudaStream_t)stream>>>(
        xyz_aos, n, n_atoms, rmsd_apad(n_atoms), precentered, xyz_soa, traces);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_soa_to_aos(const float *xyz_soa, int64_t n, int n_atoms, float *xyz_aos, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "soa_to_aos: bad shape");
    if (n == 0) return EB_OK;
    long total = n * (long)n_atoms * 3;
    long blocks = (total + 255) / 256;
    if (blocks > 16L * sm_count()) blocks = 16L * sm_count();
    k_soa_to_aos<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(xyz_soa, n, n_atoms,
                                                                rmsd_apad(n_atoms), xyz_aos);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_gather_frames(const float *xyz_soa, const double *traces, int n_atoms, const int64_t *idx,
                     int64_t m, float *out_soa, double *out_traces, void *stream)
{
    EB_CHECK_ARG(m >= 0 && n_atoms > 0, "gather_frames: bad shape");
    if (m == 0) return EB_OK;
    long blocks = m < 4L * sm_count() ? m : 4L * sm_count();
    k_gather_frames<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(
        xyz_soa, traces, rmsd_apad(n_atoms), idx, m, out_soa, out_traces);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

// shared implementation of the step entry points: `exch.peers != nullptr` selects the fused
// peer-memory candidate exchange (cand_in / n_cand are then ignored)
static int step_rmsd_impl(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          int64_t frame_offset, const void *cand_in, int n_cand, float *dist,
                          int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                          eb_kc_state *state, int64_t *center_list, void *partials,
                          void *cand_out, int exact, int n_steps, Exch exch, void *stream)
{
    // TMA-staged kernel for large shards.  box = 12 rows x A_pad/P floats: the inner box
    // dimension is limited to 256 elements and must be a multiple of 16 bytes, the row index
    // (3 * frame) to int32; every part of a row must be whole 128-byte lines (aligned shared
    // memory stages, no partial-line fetches: at 264 atoms, 528-byte half rows, the TMA kernel
    // is slower than the LDG kernel).  Measured at 1.25M x 500 (ms per step): LDG kernel 1.26;
    // ring (P,S) = (2,2) 1.11; (4,4) 1.22; (8,8) 1.56 -- big copies win, so P is the smallest
    // split that respects the 256-element box limit and S = 2 fills shared memory.
    const int A_pad_ = rmsd_apad(n_atoms);
    int P = 1, S = 2;
    const bool use_tma = exact && step_uses_tma(n, n_atoms) && tma_config(A_pad_, &P, &S);
    // two frames per 8-lane group and centre read (FR = 2): stages of eight frames x 128 atoms
    // (24 rows x 512 bytes = 12 KB).  Measured with bench.py's data at 1.25M frames, FR 1 -> 2:
    // 256-float rows 0.604 -> 0.596 ms, 384 0.877 -> 0.854 ms, but 512-float rows 1.110 ->
    // 1.135 ms (four 512-byte parts instead of two 1 KB parts: big copies win), so 512 keeps
    // FR = 1.  EB_K1_FR=1 / 2 forces a form where it applies (developer A/B switch).
    static const int fr_env = [] {
        const char *e = getenv("EB_K1_FR");
        return e ? atoi(e) : 0;
    }();
    int FR = 1;
    const int parts128 = A_pad_ / 128;
    const bool fr2_ok = use_tma && A_pad_ % 128 == 0 && parts128 >= 2 && parts128 <= 4 &&
                        step_tma_smem_bytes(A_pad_, parts128, 2, 2) <= 227 * 1024;
    // ... in a BURST.  A long run is power-capped (the step draws > 1000 W at full clocks:
    // sw_power_cap, SM clock ~1.5 GHz) and then the form that moves fewer bytes through shared
    // memory wins also at 512 floats: 600 iterations at 1.25M x 500, 1.282 -> 1.249 ms per
    // iteration.  A batch of >= 64 queued iterations is taken as the sign of a long run.
    if (fr2_ok && (fr_env == 2 || (fr_env == 0 && (parts128 <= 3 || n_steps >= 64)))) {
        FR = 2;
        P = A_pad_ / 128;
        S = 2;
    }
    const size_t tma_smem = step_tma_smem_bytes(A_pad_, P, S, FR);
    if (use_tma) {
        CUtensorMap tmap;
        const int trc = make_tmap_2d(&tmap, xyz_soa, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                     sizeof(float), (uint64_t)(3 * n), (uint64_t)A_pad_,
                                     (uint64_t)A_pad_ * sizeof(float), 12 * FR,
                                     (uint32_t)(A_pad_ / P), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (trc != EB_OK) return trc;
        auto kern = tma_kernel<kModeStep>(P, S);
        if (FR == 2) {
            kern = k_kcenters_step_rmsd_tma<2, 2, kModeStep, 2>;
            if (P == 3) kern = k_kcenters_step_rmsd_tma<3, 2, kModeStep, 2>;
            if (P == 4) kern = k_kcenters_step_rmsd_tma<4, 2, kModeStep, 2>;
        }
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tma_smem));
        const size_t rec_bytes = sizeof(RecHeader) + sizeof(float) * 3 * (size_t)A_pad_;
        long chunks = (n + 31) / 32;
        long blocks = (chunks + kStepWarps - 1) / kStepWarps;
        if (blocks > sm_count()) blocks = sm_count();
        for (int it = 0; it < n_steps; ++it) {
            kern<<<(int)blocks, kStepThreads, tma_smem, (cudaStream_t)stream>>>(
                xyz_soa, traces, n, n_atoms, A_pad_, frame_offset, (const unsigned char *)cand_in,
                n_cand, rec_bytes, dist, assign, n_clusters_limit, dist_cutoff, state,
                center_list, (Partial *)partials, (unsigned char *)cand_out, tmap, nullptr,
                nullptr, nullptr, exch);
            EB_LAUNCH_CHECK();
        }
        return EB_OK;
    }
    // single shard, several iterations queued at once, not the TMA kernel: one persistent
    // cooperative launch for the whole batch (EB_K1_MULTI=0 forces one launch per iteration)
    static const int multi_on = [] {
        const char *e = getenv("EB_K1_MULTI");
        return e ? atoi(e) : 1;
    }();
    if (multi_on && n_steps > 1 && exact && k1_variant() == 0 && !exch.peers && n_cand == 1 &&
        cand_in == cand_out && n > 0) {
        int used = 0, rc = EB_OK;
        switch (pick_rounds(n)) {
#define EB_MULTI_ARGS                                                                             \
    xyz_soa, traces, (long)n, n_atoms, (long)frame_offset, cand_out, dist, assign,                \
        (int)n_clusters_limit, dist_cutoff, state, center_list, partials, n_steps,                \
        (cudaStream_t)stream, &used
            case 1: rc = launch_multi_r<1>(EB_MULTI_ARGS); break;
            case 2: rc = launch_multi_r<2>(EB_MULTI_ARGS); break;
            case 4: rc = launch_multi_r<4>(EB_MULTI_ARGS); break;
            default: rc = launch_multi_r<8>(EB_MULTI_ARGS); break;
#undef EB_MULTI_ARGS
        }
        if (rc != EB_OK) return rc;
        if (used) return EB_OK;
    }
    StepExtra ex;
    ex.exch = exch;
    for (int it = 0; it < n_steps; ++it) {
        int rc;
        if (exact && k1_variant() == 0)
            rc = launch_step_v<true, kModeStep, 0>(xyz_soa, traces, n, n_atoms, frame_offset,
                                                   cand_in, n_cand, dist, assign,
                                                   n_clusters_limit, dist_cutoff, state,
                                                   center_list, partials, cand_out, nullptr,
                                                   nullptr, nullptr, (cudaStream_t)stream, ex);
        else if (exact)
            rc = launch_step<true, kModeStep>(xyz_soa, traces, n, n_atoms, frame_offset, cand_in,
                                              n_cand, dist, assign, n_clusters_limit, dist_cutoff,
                                              state, center_list, partials, cand_out, nullptr,
                                              nullptr, nullptr, (cudaStream_t)stream);
        else
            rc = launch_step_v<false, kModeStep, 0>(xyz_soa, traces, n, n_atoms, frame_offset,
                                                    cand_in, n_cand, dist, assign,
                                                    n_clusters_limit, dist_cutoff, state,
                                                    center_list, partials, cand_out, nullptr,
                                                    nullptr, nullptr, (cudaStream_t)stream, ex);
        if (rc != EB_OK) return rc;
    }
    return EB_OK;
}

int eb_kcenters_step_rmsd(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          int64_t frame_offset, const void *cand_in, int n_cand, float *dist,
                          int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                          eb_kc_state *state, int64_t *center_list, void *partials,
                          void *cand_out, int exact, int n_steps, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "kcenters_step_rmsd: bad shape");
    EB_CHECK_ARG(n_steps >= 1, "kcenters_step_rmsd: n_steps < 1");
    EB_CHECK_ARG(n_steps == 1 || (n_cand == 1 && cand_in == cand_out),
                 "kcenters_step_rmsd: n_steps > 1 needs a single shard (cand_in == cand_out)");
    EB_CHECK_ARG(n_cand >= 1 && cand_in && cand_out && state && partials && center_list,
                 "kcenters_step_rmsd: null pointer / n_cand < 1");
    EB_CHECK_ARG(n < (int64_t(1) << 40), "kcenters_step_rmsd: shard too large");
    const Exch off = {nullptr, 0, 0, 0};
    return step_rmsd_impl(xyz_soa, traces, n, n_atoms, frame_offset, cand_in, n_cand, dist, assign,
                          n_clusters_limit, dist_cutoff, state, center_list, partials, cand_out,
                          exact, n_steps, off, stream);
}

// ---- fused peer-memory candidate exchange (see eb_common.cuh: struct Exch) ------------------
size_t eb_exch_bytes(int n_atoms, int n_ranks)
{
    const size_t rec = sizeof(RecHeader) + sizeof(float) * 3 * (size_t)rmsd_apad(n_atoms);
    const size_t stride = (rec + 127) & ~size_t(127);
    return kExchRecordsOff + 2 * (size_t)n_ranks * stride;
}

static int make_exch(Exch *e, const void *peers_dev, int n_ranks, int rank, int n_atoms)
{
    EB_CHECK_ARG(peers_dev && n_ranks >= 1 && n_ranks <= kExchMaxRanks && rank >= 0 &&
                     rank < n_ranks,
                 "p2p exchange: need 1..8 ranks and the device array of peer buffers");
    const size_t rec = sizeof(RecHeader) + sizeof(float) * 3 * (size_t)rmsd_apad(n_atoms);
    e->peers = (const long long *)peers_dev;
    e->size = n_ranks;
    e->rank = rank;
    e->rec_stride = (unsigned)((rec + 127) & ~size_t(127));
    // bounded spin (seconds, EB_EXCH_TIMEOUT_S; default 60): a step normally waits microseconds
    static const double timeout_s = [] {
        const char *v = getenv("EB_EXCH_TIMEOUT_S");
        const double t = v ? atof(v) : 60.0;
        return t > 0.0 ? t : 60.0;
    }();
    e->timeout_ns = (unsigned long long)(timeout_s * 1e9);
    return EB_OK;
}

int eb_kcenters_step_rmsd_p2p(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *peers_dev, int n_ranks, int rank,
                              float *dist, int32_t *assign, int32_t n_clusters_limit,
                              double dist_cutoff, eb_kc_state *state, int64_t *center_list,
                              void *partials, void *cand_out, int exact, int n_steps,
                              void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0 && n_steps >= 1, "kcenters_step_rmsd_p2p: bad shape");
    EB_CHECK_ARG(cand_out && state && partials && center_list && (n == 0 || (dist && assign)),
                 "kcenters_step_rmsd_p2p: null pointer");   // an empty shard has no dist / assign
    Exch e;
    const int rc = make_exch(&e, peers_dev, n_ranks, rank, n_atoms);
    if (rc != EB_OK) return rc;
    return step_rmsd_impl(xyz_soa, traces, n, n_atoms, frame_offset, nullptr, 0, dist, assign,
                          n_clusters_limit, dist_cutoff, state, center_list, partials, cand_out,
                          exact, n_steps, e, stream);
}

int eb_kcenters_seed_rmsd_p2p(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *peers_dev, int n_ranks, int rank,
                              const float *dist, int32_t first_center_id, eb_kc_state *state,
                              void *partials, void *cand_out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "kcenters_seed_rmsd_p2p: bad shape");
    EB_CHECK_ARG(state && partials && cand_out, "kcenters_seed_rmsd_p2p: null pointer");
    StepExtra ex;
    const int rc = make_exch(&ex.exch, peers_dev, n_ranks, rank, n_atoms);
    if (rc != EB_OK) return rc;
    EB_CUDA(cudaMemsetAsync(state, 0, sizeof(eb_kc_state), (cudaStream_t)stream));
    return launch_step_v<true, kModeSeed, 0>(xyz_soa, traces, n, n_atoms, frame_offset, nullptr,
                                             0, const_cast<float *>(dist), nullptr,
                                             first_center_id, 0.0, state, nullptr, partials,
                                             cand_out, nullptr, nullptr, nullptr,
                                             (cudaStream_t)stream, ex);
}

// Triangle-inequality variant (kcenters.py:287-296, `use_triangle_inequality=True`): per
// iteration one small launch computes the distances from the new centre to all stored centres
// (and stores the new centre), then the step skips every frame whose current distance is at
// most half the distance between its centre and the new one.  Same results as the plain step.
int eb_kcenters_step_rmsd_tri(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              int64_t frame_offset, const void *cand_in, int n_cand, float *dist,
                              int32_t *assign, int32_t n_clusters_limit, double dist_cutoff,
                              eb_kc_state *state, int64_t *center_list, void *partials,
                              void *cand_out, float *center_store, double *center_store_traces,
                              float *cc, int64_t store_capacity, int64_t k_upper, int n_steps,
                              void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "kcenters_step_rmsd_tri: bad shape");
    EB_CHECK_ARG(n_steps >= 1, "kcenters_step_rmsd_tri: n_steps < 1");
    EB_CHECK_ARG(n_steps == 1 || (n_cand == 1 && cand_in == cand_out),
                 "kcenters_step_rmsd_tri: n_steps > 1 needs a single shard");
    EB_CHECK_ARG(n_cand >= 1 && cand_in && cand_out && state && partials && center_list &&
                     center_store && center_store_traces && cc,
                 "kcenters_step_rmsd_tri: null pointer / n_cand < 1");
    EB_CHECK_ARG(k_upper >= 0 && k_upper + n_steps <= store_capacity,
                 "kcenters_step_rmsd_tri: centre store too small");
    for (int it = 0; it < n_steps; ++it) {
        StepExtra ex;
        ex.cstore = center_store;
        ex.cstore_traces = center_store_traces;
        ex.grid_frames = k_upper + it;          // host-side upper bound of the centre count
        int rc = launch_step_v<true, kModeCC, 0>(
            center_store, center_store_traces, 0, n_atoms, 0, cand_in, n_cand, nullptr, nullptr,
            n_clusters_limit, dist_cutoff, state, nullptr, nullptr, nullptr, nullptr, nullptr, cc,
            (cudaStream_t)stream, ex);
        if (rc != EB_OK) return rc;
        StepExtra ex2;
        ex2.cc = cc;
        rc = launch_step_v<true, kModeStep, kVarTri>(
            xyz_soa, traces, n, n_atoms, frame_offset, cand_in, n_cand, dist, assign,
            n_clusters_limit, dist_cutoff, state, center_list, partials, cand_out, nullptr,
            nullptr, nullptr, (cudaStream_t)stream, ex2);
        if (rc != EB_OK) return rc;
    }
    return EB_OK;
}

int eb_kcenters_seed_rmsd(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          int64_t frame_offset, const float *dist, int32_t first_center_id,
                          eb_kc_state *state, void *partials, void *cand_out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "kcenters_seed_rmsd: bad shape");
    EB_CHECK_ARG(state && partials && cand_out, "kcenters_seed_rmsd: null pointer");
    EB_CUDA(cudaMemsetAsync(state, 0, sizeof(eb_kc_state), (cudaStream_t)stream));
    return launch_step<true, kModeSeed>(xyz_soa, traces, n, n_atoms, frame_offset, nullptr, 0,
                                        const_cast<float *>(dist), nullptr, first_center_id, 0.0,
                                        state,
                                        nullptr, partials, cand_out, nullptr, nullptr, nullptr,
                                        (cudaStream_t)stream);
}

int eb_rmsd_one_to_all(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                       const float *center_soa, const double *center_trace, float *out,
                       int exact, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "rmsd_one_to_all: bad shape");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && center_soa && center_trace && out,
                 "rmsd_one_to_all: null pointer");
    if (exact && step_uses_tma(n, n_atoms)) {
        const int A_pad = rmsd_apad(n_atoms);
        int P = 1, S = 2;
        tma_config(A_pad, &P, &S);
        const size_t tma_smem = step_tma_smem_bytes(A_pad, P, S);
        CUtensorMap tmap;
        const int trc = make_tmap_2d(&tmap, xyz_soa, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                     sizeof(float), (uint64_t)(3 * n), (uint64_t)A_pad,
                                     (uint64_t)A_pad * sizeof(float), 12, (uint32_t)(A_pad / P),
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
        if (trc != EB_OK) return trc;
        auto kern = tma_kernel<kModeDistOnly>(P, S);
        EB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tma_smem));
        long blocks = ((n + 31) / 32 + kStepWarps - 1) / kStepWarps;
        if (blocks > sm_count()) blocks = sm_count();
        kern<<<(int)blocks, kStepThreads, tma_smem, (cudaStream_t)stream>>>(
            xyz_soa, traces, n, n_atoms, A_pad, 0, nullptr, 0, 0, nullptr, nullptr, 0, 0.0,
            nullptr, nullptr, nullptr, nullptr, tmap, center_soa, center_trace, out,
            Exch{nullptr, 0, 0, 0});
        EB_LAUNCH_CHECK();
        return EB_OK;
    }
    if (exact)
        return launch_step<true, kModeDistOnly>(xyz_soa, traces, n, n_atoms, 0, nullptr, 0,
                                                nullptr, nullptr, 0, 0.0, nullptr, nullptr,
                                                nullptr, nullptr, center_soa, center_trace, out,
                                                (cudaStream_t)stream);
    return launch_step<false, kModeDistOnly>(xyz_soa, traces, n, n_atoms, 0, nullptr, 0, nullptr,
                                             nullptr, 0, 0.0, nullptr, nullptr, nullptr, nullptr,
                                             center_soa, center_trace, out,
                                             (cudaStream_t)stream);
}

// One-vs-all distances for a PAM proposal with triangle-inequality pruning: out[f] = +inf (and
// frame f is not read) when assign[f] != cid and dist[f] <= (1 - 1e-5)/2 * cc[assign[f]], where
// cc[j] = d(proposal, medoid j).  For every other frame out[f] is exactly what
// eb_rmsd_one_to_all writes.  The three-way split of kmedoids.py:644-658 classifies a pruned
// frame as "unchanged", which is what its true distance would do.
int eb_rmsd_one_to_all_pruned(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              const float *center_soa, const double *center_trace,
                              const float *dist, const int32_t *assign, const float *cc,
                              int32_t cid, float *out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0, "rmsd_one_to_all_pruned: bad shape");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && center_soa && center_trace && out && dist && assign && cc,
                 "rmsd_one_to_all_pruned: null pointer");
    StepExtra ex;
    ex.cc = const_cast<float *>(cc);
    return launch_step_v<true, kModeDistOnly, kVarTri>(
        xyz_soa, traces, n, n_atoms, 0, nullptr, 0, const_cast<float *>(dist),
        const_cast<int32_t *>(assign), cid, 0.0, nullptr, nullptr, nullptr, nullptr, center_soa,
        center_trace, out, (cudaStream_t)stream, ex);
}

}  // extern "C"
