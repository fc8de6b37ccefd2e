// K4 / K6: the n-length bookkeeping of one PAM proposal, on device vectors.
//
// Reference behaviour reproduced (paths under /root/reference/enspara/):
//   cluster/kmedoids.py:639-658  three-way split after the full-pass distance to the proposal:
//        dn       = distances >  new   -> (cid, new)
//        up_other = distances <= new and assign != cid -> unchanged
//        up_this  = distances <= new and assign == cid -> needs the nearest of ALL medoids
//   cluster/kmedoids.py:478-479  cost = mean(d^2) (numerator here; deterministic order)
//   cluster/kmedoids.py:611,514  state_inds = where(assign == cid)[0]; choice(state_inds)
//        == state_inds[randint(len(state_inds))]  -> count + k-th member select
// All kernels are HBM bound over 4..8-byte-per-frame vectors; they exist so that distances and
// assignments never leave the device between proposals.
#include "eb_common.cuh"
#include "eb_rmsd.cuh"

namespace eb {

constexpr int kPamThreads = 256;
constexpr int kPamMaxBlocks = 2048;

template <typename D>
__global__ void __launch_bounds__(kPamThreads)
k_pam_classify(const D *__restrict__ new_ctr_dist, const D *__restrict__ dist,
               const int *__restrict__ assign, long n, int cid, D *__restrict__ new_dist,
               int *__restrict__ new_assign, int64_t *__restrict__ ambig_idx,
               unsigned long long *n_ambig)
{
    const int lane = threadIdx.x & 31;
    const long stride = (long)gridDim.x * blockDim.x;
    const long n_round = (n + 31) & ~31L;  // keep whole warps in the loop for the ballot
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool ambig = false;
        if (i < n) {
            const D dn = new_ctr_dist[i];
            const D d_old = dist[i];
            const int a = assign[i];
            if (d_old > dn) {
                new_assign[i] = cid;
                new_dist[i] = dn;
            } else if (a != cid) {
                new_assign[i] = a;
                new_dist[i] = d_old;
            } else {
                ambig = true;
                new_assign[i] = -1;
                new_dist[i] = (D)-1;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ambig);
        if (mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_ambig, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (ambig) ambig_idx[base + __popc(mask & ((1u << lane) - 1))] = i;
        }
    }
}

// Triangle-inequality pre-pass of a PAM proposal's full distance pass (kmedoids.py:637): a frame
// of ANOTHER cluster whose distance to its medoid is at most (1 - 1e-5)/2 of that medoid's
// distance to the proposal (cc[assign]) cannot get closer to the proposal than it is to its
// medoid, so the three-way split leaves it untouched: it gets +inf and is never read.  Every
// other frame goes to a compact index list; the exact kernel then evaluates the proposal
// against exactly those frames with all its lanes busy (the fused pruned pass of round 1
// skipped 4-frame rounds inside 32-frame chunks: with ~3 % of the frames scattered over the
// shard almost every round still had one live frame and three idle groups).
__global__ void __launch_bounds__(kPamThreads)
k_pam_need_list(const float *__restrict__ dist, const int *__restrict__ assign,
                const float *__restrict__ cc, long n, int cid, float *__restrict__ out,
                int64_t *__restrict__ need_idx, unsigned long long *n_need)
{
    const int lane = threadIdx.x & 31;
    const long stride = (long)gridDim.x * blockDim.x;
    const long n_round = (n + 31) & ~31L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool need = false;
        if (i < n) {
            const int a = assign[i];
            need = (a < 0) || (a == cid) || !(dist[i] <= 0.499995f * __ldg(cc + a));
            if (!need) out[i] = INFINITY;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_need, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) need_idx[base + __popc(mask & ((1u << lane) - 1))] = i;
        }
    }
}

// deterministic sum of squares: fixed block ranges, fixed tree, ordered final pass
// one warp, fixed order: lane l adds partials l, l+32, ... in sequence, then a fixed shuffle tree
__device__ __forceinline__ void sum_final_warp(const double *partials, int nb, double *out)
{
    const int lane = threadIdx.x;
    double s = 0.0;
    for (int b = lane; b < nb; b += 32) s += __ldcg(partials + b);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += shfl_xor_d(s, m);
    if (lane == 0) *out = s;
}

// The same partial sums, and the block that finishes last (ticket counter behind the partials,
// zero between uses) runs the ordered final pass: one launch, the same bits as two.
template <typename D>
__global__ void __launch_bounds__(kPamThreads)
k_sumsq(const D *__restrict__ x, long n, double *__restrict__ partials, unsigned int *ticket,
        double *out)
{
    __shared__ double sh[kPamThreads];
    __shared__ int last;
    const long per_block = (n + gridDim.x - 1) / gridDim.x;
    const long lo = (long)blockIdx.x * per_block;
    const long hi = min(n, lo + per_block);
    double s = 0.0;
    for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double v = (double)x[i];
        s = fma(v, v, s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = kPamThreads / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) sum_final_warp(partials, (int)gridDim.x, out);
    if (threadIdx.x == 0) *ticket = 0u;
}

constexpr int kHistMaxBins = 8192;  // 32 KB of shared counters per block

// members per cluster: per-block shared-memory histogram, then one global add per non-empty bin
__global__ void __launch_bounds__(kPamThreads)
k_count_members_smem(const int *__restrict__ assign, long n, int k, unsigned long long *counts)
{
    extern __shared__ unsigned int hist[];
    for (int b = threadIdx.x; b < k; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int a = assign[i];
        if (a >= 0 && a < k) atomicAdd(hist + a, 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < k; b += blockDim.x) {
        const unsigned int c = hist[b];
        if (c) atomicAdd(counts + b, (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(kPamThreads)
k_count_members(const int *__restrict__ assign, long n, int k, unsigned long long *counts)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int a = assign[i];
        if (a >= 0 && a < k) atomicAdd(counts + a, 1ULL);
    }
}

// k-th member select, pass 1: members per fixed block range
__global__ void __launch_bounds__(kPamThreads)
k_select_count(const int *__restrict__ assign, long n, int cid, long per_block,
               unsigned long long *block_counts)
{
    __shared__ unsigned int sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    const long lo = (long)blockIdx.x * per_block;
    const long hi = min(n, lo + per_block);
    unsigned int c = 0;
    for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) c += (assign[i] == cid);
    for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = sh;
}

// pass 2 (one block): find the counting block holding the kth member, then the member in it
__global__ void __launch_bounds__(1024)
k_select_pick(const int32_t *__restrict__ assign, long n, int32_t cid, long kth, long per_block,
              const unsigned long long *block_counts, int nb, int64_t *out, long offset,
              int64_t *global_out)
{
    // One block of 1024 threads; every load of a phase is independent of the others (the first
    // version walked 31 + 32 dependent 32-wide steps with one warp: 11 us per proposal).
    __shared__ long s_warp[32];
    __shared__ long s_block, s_rem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- phase 1: which counting block holds the kth member ---------------------------------
    const int cpt = (nb + 1023) / 1024;                  // block counts per thread
    long own = 0;
    for (int j = 0; j < cpt; ++j) {
        const int b = tid * cpt + j;
        if (b < nb) own += (long)block_counts[b];
    }
    long inc = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long up = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    if (tid == 0) {
        s_block = nb;
        s_rem = 0;
    }
    __syncthreads();
    long before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const long ex = before + inc - own;                  // members in blocks before this thread's
    if (kth >= ex && kth < ex + own) {                   // exactly one thread
        long rem = kth - ex;
        for (int j = 0; j < cpt; ++j) {
            const int b = tid * cpt + j;
            const long c = (long)block_counts[b];
            if (rem < c) {
                s_block = b;
                s_rem = rem;
                break;
            }
            rem -= c;
        }
    }
    __syncthreads();
    if (s_block >= nb) {
        if (tid == 0) {
            *out = -1;
            if (global_out) *global_out = -1;
        }
        return;
    }
    // ---- phase 2: the (rem)-th member inside that block's range -----------------------------
    const long lo = s_block * per_block;
    const long hi = min(n, lo + per_block);
    long rem = s_rem;
    for (long base = lo; base < hi; base += 1024) {
        const long i = base + tid;
        const bool hit = (i < hi) && (assign[i] == cid);
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        __syncthreads();                                  // s_warp free again
        if (lane == 0) s_warp[warp] = __popc(mask);
        __syncthreads();
        long pre = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const long c = s_warp[w];
            if (w < warp) pre += c;
            total += c;
        }
        if (rem < total) {
            if (hit && pre + __popc(mask & ((1u << lane) - 1u)) == rem) {
                *out = i;
                if (global_out) *global_out = i + offset;   // GLOBAL frame index of the member
            }
            return;
        }
        rem -= total;
    }
    if (tid == 0) {
        *out = -1;
        if (global_out) *global_out = -1;
    }
}

// The proposal takes medoid cid's slot, the displaced medoid is kept for a rejection
// (kmedoids.py:660-664); scal_i[0] = the proposal's global index; the screen's overflow counter
// is cleared.  One launch instead of five small copies and a memset.
__global__ void k_take_slot(float *slot, double *slot_trace, float *saved, double *saved_trace,
                            const float *prop, const double *prop_trace, long frame_floats,
                            int64_t *scal_i, const int64_t *prop_idx, int32_t *tc_ovf)
{
    for (long i = threadIdx.x; i < frame_floats; i += blockDim.x) {
        saved[i] = slot[i];
        slot[i] = prop[i];
    }
    if (threadIdx.x == 0) {
        *saved_trace = *slot_trace;
        *slot_trace = *prop_trace;
        scal_i[0] = prop_idx[0];
        if (tc_ovf) *tc_ovf = 0;
    }
}

__global__ void k_restore_slot(float *slot, double *slot_trace, const float *saved,
                               const double *saved_trace, long frame_floats)
{
    for (long i = threadIdx.x; i < frame_floats; i += blockDim.x) slot[i] = saved[i];
    if (threadIdx.x == 0) *slot_trace = *saved_trace;
}

// ------------------------------------------------------------------------------------------
// Re-assignment of a proposal's ambiguous frames (kmedoids.py:666-670) against the medoids the
// triangle inequality leaves.  RMSD is a metric: d(x, m_j) >= d(p, m_j) - d(x, p), so medoid j
// can only win (or tie) frame x if cc[j] = d(p, m_j) < 2 d(x, p); with the float margin of
// k_pam_need_list the test is !(d(x, p) <= 0.499995f * cc[j]).  d(x, p) is new_ctr_dist[x]
// (the full pass) and is exactly what slot cid -- now holding p -- would score.
// k_pam_medoid_list: ascending list of the medoids that pass the test for the LARGEST d(x, p)
// of the subset (one block; more than `cap` -> overflow, the caller takes the general path).
// k_pam_reassign_listed: one warp per ambiguous frame; the listed medoids that pass the
// frame's own test are scored exactly (the arithmetic of the exact kernel / re-score: same
// lanes, same atom order, same QCP) by the warp's four 8-lane groups; the nearest of those and
// (d(x, p), cid) wins, lowest index on ties.
// ------------------------------------------------------------------------------------------
struct SlotExchange {            // k_take_slot's arguments, for the kernels that fold it in
    float *slot;
    double *slot_trace;
    float *saved;
    double *saved_trace;
    const float *prop;
    const double *prop_trace;
    long frame_floats;
    int64_t *scal_i;
    const int64_t *prop_idx;
};

__global__ void __launch_bounds__(1024)
k_pam_medoid_list(const int64_t *__restrict__ ambig_idx, const int64_t *__restrict__ n_ambig,
                  const float *__restrict__ new_ctr_dist, const float *__restrict__ cc, int k,
                  int cid, int cap, int32_t *list, int32_t *list_n, int32_t *ovf,
                  SlotExchange ex)
{
    __shared__ float s_max[32];
    __shared__ int s_cnt[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ex.slot) {
        // the slot exchange of k_take_slot rides along (the listed re-assignment never reads
        // slot cid: the proposal's distances are new_ctr_dist)
        for (long i = tid; i < ex.frame_floats; i += 1024) {
            ex.saved[i] = ex.slot[i];
            ex.slot[i] = ex.prop[i];
        }
        if (tid == 0) {
            *ex.saved_trace = *ex.slot_trace;
            *ex.slot_trace = *ex.prop_trace;
            ex.scal_i[0] = ex.prop_idx[0];
        }
    }
    const long m = (long)*n_ambig;
    float dmax = 0.0f;
    for (long p = tid; p < m; p += 1024) dmax = fmaxf(dmax, new_ctr_dist[ambig_idx[p]]);
#pragma unroll
    for (int msk = 16; msk > 0; msk >>= 1)
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, msk));
    if (lane == 0) s_max[warp] = dmax;
    __syncthreads();
    dmax = s_max[0];
    for (int w = 1; w < 32; ++w) dmax = fmaxf(dmax, s_max[w]);
    int base = 0;
    for (int j0 = 0; j0 < k; j0 += 1024) {
        const int j = j0 + tid;
        const bool keep = j < k && j != cid && !(dmax <= 0.499995f * cc[j]);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[warp] = __popc(mask);
        __syncthreads();
        int pre = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = s_cnt[w];
            if (w < warp) pre += c;
            total += c;
        }
        const int pos = base + pre + __popc(mask & ((1u << lane) - 1u));
        if (keep && pos < cap) list[pos] = j;
        base += total;
        __syncthreads();
    }
    if (tid == 0) {
        if (base > cap) {
            *list_n = 0;
            atomicAdd(ovf, 1);
        } else {
            *list_n = base;
        }
    }
}

constexpr int kListWarps = 4;
constexpr int kListThreads = kListWarps * 32;

__global__ void __launch_bounds__(kListThreads, 4)
k_pam_reassign_listed(const float *__restrict__ xyz, const double *__restrict__ traces,
                      int n_atoms, int A_pad, const float *__restrict__ centers,
                      const double *__restrict__ ctraces,
                      const int64_t *__restrict__ ambig_idx, const int64_t *__restrict__ n_ambig,
                      const float *__restrict__ new_ctr_dist, const float *__restrict__ cc,
                      int cid, const int32_t *__restrict__ list,
                      const int32_t *__restrict__ list_n, int cap, float *new_dist,
                      int32_t *new_assign)
{
    // One WARP per ambiguous frame, warps independent of each other (no block barrier): a
    // frame costs a chain of dependent latencies (index, distance, list, coordinates) whatever
    // the number of survivors, so what matters is how many frames are in flight per SM
    // (a block per frame: 4-5; this form: 16).
    extern __shared__ int cl_all[];                      // [kListWarps][cap]
    const long m = (long)*n_ambig;
    const int ln = *list_n;
    const int lane = threadIdx.x & 31, l8 = lane & 7, g = lane >> 3, warp = threadIdx.x >> 5;
    int *cl = cl_all + (size_t)warp * cap;
    const int A4 = A_pad >> 2;
    const size_t stride = 3 * (size_t)A_pad;
    const long warps_total = (long)gridDim.x * kListWarps;
    for (long p = (long)blockIdx.x * kListWarps + warp; p < m; p += warps_total) {
        const long f = (long)ambig_idx[p];
        const float dn = new_ctr_dist[f];
        // the listed medoids that pass this frame's own test, in list (= index) order
        int cnt = 0;
        for (int i0 = 0; i0 < ln; i0 += 32) {
            const int i = i0 + lane;
            const int j = i < ln ? list[i] : 0;
            const bool keep = i < ln && !(dn <= 0.499995f * __ldg(cc + j));
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) cl[cnt + __popc(mask & ((1u << lane) - 1u))] = j;
            cnt += __popc(mask);
        }
        __syncwarp();
        const float4 *px = reinterpret_cast<const float4 *>(xyz + (size_t)f * stride);
        const double Ga = traces[f];
        float best_d = INFINITY;
        int best_c = 0x7fffffff;
        // two candidates per group share the frame's loads; with at most 4 survivors one
        // candidate per group keeps twice as many groups busy
        const int per = cnt > 4 ? 2 : 1;
        for (int base = 0; base < cnt; base += per * 4) {
            const int i0 = base + per * g;
            const bool act0 = i0 < cnt, act1 = per == 2 && i0 + 1 < cnt;
            const int c0 = act0 ? cl[i0] : 0, c1 = act1 ? cl[i0 + 1] : c0;
            const float4 *p0 = reinterpret_cast<const float4 *>(centers + (size_t)c0 * stride);
            const float4 *p1 = reinterpret_cast<const float4 *>(centers + (size_t)c1 * stride);
            double m0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, m1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (act1) {
#pragma unroll 2
                for (int j = l8; j < A4; j += 8) {
                    const float4 x = __ldg(px + j), y = __ldg(px + A4 + j),
                                 z = __ldg(px + 2 * A4 + j);
                    const float4 ax = __ldg(p0 + j), ay = __ldg(p0 + A4 + j),
                                 az = __ldg(p0 + 2 * A4 + j);
                    const float4 bx = __ldg(p1 + j), by = __ldg(p1 + A4 + j),
                                 bz = __ldg(p1 + 2 * A4 + j);
                    acc_atom(m0, x.x, y.x, z.x, (double)ax.x, (double)ay.x, (double)az.x);
                    acc_atom(m0, x.y, y.y, z.y, (double)ax.y, (double)ay.y, (double)az.y);
                    acc_atom(m0, x.z, y.z, z.z, (double)ax.z, (double)ay.z, (double)az.z);
                    acc_atom(m0, x.w, y.w, z.w, (double)ax.w, (double)ay.w, (double)az.w);
                    acc_atom(m1, x.x, y.x, z.x, (double)bx.x, (double)by.x, (double)bz.x);
                    acc_atom(m1, x.y, y.y, z.y, (double)bx.y, (double)by.y, (double)bz.y);
                    acc_atom(m1, x.z, y.z, z.z, (double)bx.z, (double)by.z, (double)bz.z);
                    acc_atom(m1, x.w, y.w, z.w, (double)bx.w, (double)by.w, (double)bz.w);
                }
            } else if (act0) {
#pragma unroll 4
                for (int j = l8; j < A4; j += 8) {
                    const float4 x = __ldg(px + j), y = __ldg(px + A4 + j),
                                 z = __ldg(px + 2 * A4 + j);
                    const float4 ax = __ldg(p0 + j), ay = __ldg(p0 + A4 + j),
                                 az = __ldg(p0 + 2 * A4 + j);
                    acc_atom(m0, x.x, y.x, z.x, (double)ax.x, (double)ay.x, (double)az.x);
                    acc_atom(m0, x.y, y.y, z.y, (double)ax.y, (double)ay.y, (double)az.y);
                    acc_atom(m0, x.z, y.z, z.z, (double)ax.z, (double)ay.z, (double)az.z);
                    acc_atom(m0, x.w, y.w, z.w, (double)ax.w, (double)ay.w, (double)az.w);
                }
            }
            group8_reduce(m0);
            if (per == 2) group8_reduce(m1);
            if (act0) {
                const float d = rmsd_from_msd(qcp_msd(m0, Ga, ctraces[c0], n_atoms));
                if (d < best_d || (d == best_d && c0 < best_c)) {
                    best_d = d;
                    best_c = c0;
                }
            }
            if (act1) {
                const float d = rmsd_from_msd(qcp_msd(m1, Ga, ctraces[c1], n_atoms));
                if (d < best_d || (d == best_d && c1 < best_c)) {
                    best_d = d;
                    best_c = c1;
                }
            }
        }
        // nearest over the four groups, then against (d(x, p), cid): slot cid holds the proposal
#pragma unroll
        for (int msk = 8; msk < 32; msk <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, best_d, msk);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, msk);
            if (od < best_d || (od == best_d && oc < best_c)) {
                best_d = od;
                best_c = oc;
            }
        }
        if (lane == 0) {
            float bd = dn;
            int bc = cid;
            if (best_d < bd || (best_d == bd && best_c < bc)) {
                bd = best_d;
                bc = best_c;
            }
            new_dist[f] = bd;
            new_assign[f] = bc;
        }
        __syncwarp();
    }
}

static int pam_blocks(long n)
{
    long b = (n + kPamThreads * 4 - 1) / (kPamThreads * 4);
    const long cap = 8L * sm_count();
    if (b > cap) b = cap;
    if (b > kPamMaxBlocks) b = kPamMaxBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace eb

using namespace eb;

extern "C" {

size_t eb_pam_scratch_bytes(int64_t n)
{
    (void)n;
    return sizeof(double) * kPamMaxBlocks + 16;   // partials / block counts + the ticket
}

int eb_pam_classify(const void *new_ctr_dist, const void *dist, const int32_t *assign, int64_t n,
                    int dist_is_f64, int32_t cid, void *new_dist, int32_t *new_assign,
                    int64_t *ambig_idx, int64_t *n_ambig, void *stream)
{
    EB_CHECK_ARG(n >= 0, "pam_classify: n < 0");
    EB_CHECK_ARG(n_ambig, "pam_classify: null counter");
    cudaStream_t s = (cudaStream_t)stream;
    EB_CUDA(cudaMemsetAsync(n_ambig, 0, sizeof(int64_t), s));
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(new_ctr_dist && dist && assign && new_dist && new_assign && ambig_idx,
                 "pam_classify: null pointer");
    if (dist_is_f64)
        k_pam_classify<double><<<pam_blocks(n), kPamThreads, 0, s>>>(
            (const double *)new_ctr_dist, (const double *)dist, assign, n, cid,
            (double *)new_dist, new_assign, ambig_idx, (unsigned long long *)n_ambig);
    else
        k_pam_classify<float><<<pam_blocks(n), kPamThreads, 0, s>>>(
            (const float *)new_ctr_dist, (const float *)dist, assign, n, cid, (float *)new_dist,
            new_assign, ambig_idx, (unsigned long long *)n_ambig);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_pam_need_list(const float *dist, const int32_t *assign, const float *cc, int64_t n,
                     int32_t cid, float *out, int64_t *need_idx, int64_t *n_need, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_need, "pam_need_list: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    EB_CUDA(cudaMemsetAsync(n_need, 0, sizeof(int64_t), s));
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(dist && assign && cc && out && need_idx, "pam_need_list: null pointer");
    k_pam_need_list<<<pam_blocks(n), kPamThreads, 0, s>>>(dist, assign, cc, n, cid, out, need_idx,
                                                         (unsigned long long *)n_need);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_sum_squares(const void *dist, int64_t n, int dist_is_f64, double *out, void *scratch,
                   void *stream)
{
    EB_CHECK_ARG(n >= 0 && out && scratch, "sum_squares: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        EB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
        return EB_OK;
    }
    const int nb = pam_blocks(n);
    unsigned int *ticket = reinterpret_cast<unsigned int *>((double *)scratch + kPamMaxBlocks);
    if (dist_is_f64)
        k_sumsq<double><<<nb, kPamThreads, 0, s>>>((const double *)dist, n, (double *)scratch,
                                                   ticket, out);
    else
        k_sumsq<float><<<nb, kPamThreads, 0, s>>>((const float *)dist, n, (double *)scratch,
                                                  ticket, out);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_count_members(const int32_t *assign, int64_t n, int32_t k, int64_t *counts, void *stream)
{
    EB_CHECK_ARG(n >= 0 && k >= 0 && counts, "count_members: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    EB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)k, s));
    if (n == 0 || k == 0) return EB_OK;
    if (k <= kHistMaxBins) {
        int nb = pam_blocks(n);
        if (nb > 2 * sm_count()) nb = 2 * sm_count();   // few blocks: fewer flushes
        k_count_members_smem<<<nb, kPamThreads, sizeof(unsigned int) * (size_t)k, s>>>(
            assign, n, k, (unsigned long long *)counts);
    } else {
        k_count_members<<<pam_blocks(n), kPamThreads, 0, s>>>(assign, n, k,
                                                              (unsigned long long *)counts);
    }
    EB_LAUNCH_CHECK();
    return EB_OK;
}

static int select_member(const int32_t *assign, int64_t n, int32_t cid, int64_t kth, int64_t *out,
                         void *scratch, int64_t offset, int64_t *global_out, void *stream)
{
    EB_CHECK_ARG(n >= 0 && kth >= 0 && out && scratch, "select_member: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = pam_blocks(n);
    const long per_block = (n + nb - 1) / nb;
    k_select_count<<<nb, kPamThreads, 0, s>>>(assign, n, cid, per_block,
                                              (unsigned long long *)scratch);
    EB_LAUNCH_CHECK();
    k_select_pick<<<1, 1024, 0, s>>>(assign, n, cid, kth, per_block,
                                     (const unsigned long long *)scratch, nb, out, (long)offset,
                                     global_out);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_select_member(const int32_t *assign, int64_t n, int32_t cid, int64_t kth, int64_t *out,
                     void *scratch, void *stream)
{
    return select_member(assign, n, cid, kth, out, scratch, 0, nullptr, stream);
}

int eb_pam_propose_rmsd(const eb_pam_ctx *c, int32_t cid, int64_t kth, int64_t m_max, int stages,
                        void *stream)
{
    EB_CHECK_ARG(c && c->xyz && c->traces && c->medoid_xyz && c->prop_xyz && c->scal_i &&
                     c->scal_d && c->scratch,
                 "pam_propose: missing buffers");
    EB_CHECK_ARG(cid >= 0 && cid < c->k && c->k >= 2, "pam_propose: bad cluster id");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t frame_floats = 3 * (size_t)rmsd_apad(c->n_atoms);
    int rc;
    if (stages & EB_PAM_SELECT) {
        // k-th member of the cluster -> scal_i[0] (local) and prop_idx (global), then its frame
        rc = select_member(c->assign, c->n, cid, kth, c->scal_i, c->scratch, c->frame_offset,
                           c->prop_idx, stream);
        if (rc != EB_OK) return rc;
        rc = eb_gather_frames(c->xyz, c->traces, c->n_atoms, c->scal_i, 1, c->prop_xyz,
                              c->prop_traces, stream);
        if (rc != EB_OK) return rc;
    }
    if (stages & EB_PAM_TRIAL) {
        EB_CHECK_ARG(c->dist && c->assign && c->new_dist && c->new_assign && c->new_ctr_dist &&
                         c->cc && c->need_idx && c->need_n && c->need_assign && c->ambig_idx &&
                         c->saved_xyz,
                     "pam_propose: missing buffers");
        // the three device counters of a proposal {n_ambig, overflow, n_need} are cleared with
        // one memset when the caller laid them out next to each other
        const bool counters_adjacent =
            c->tc_ovf && (const char *)c->tc_ovf == (const char *)(c->scal_i + 2) &&
            (const char *)c->need_n == (const char *)(c->scal_i + 3);
        if (counters_adjacent) {
            EB_CUDA(cudaMemsetAsync(c->scal_i + 1, 0, 3 * sizeof(int64_t), s));
        } else {
            EB_CUDA(cudaMemsetAsync(c->scal_i + 1, 0, sizeof(int64_t), s));
            EB_CUDA(cudaMemsetAsync(c->need_n, 0, sizeof(int64_t), s));
            if (c->tc_ovf) EB_CUDA(cudaMemsetAsync(c->tc_ovf, 0, sizeof(int32_t), s));
        }
        // distances proposal -> every medoid, then the pruned full pass (kmedoids.py:637)
        rc = eb_rmsd_one_to_all(c->medoid_xyz, c->medoid_traces, c->k, c->n_atoms, c->prop_xyz,
                                c->prop_traces, c->cc, 1, stream);
        if (rc != EB_OK) return rc;
        if (c->n > 0) {
            k_pam_need_list<<<pam_blocks(c->n), kPamThreads, 0, s>>>(
                c->dist, c->assign, c->cc, c->n, cid, c->new_ctr_dist, c->need_idx,
                (unsigned long long *)c->need_n);
            EB_LAUNCH_CHECK();
            rc = eb_rmsd_assign_dev(c->xyz, c->traces, c->n, c->n_atoms, c->prop_xyz,
                                    c->prop_traces, 1, c->need_idx, c->n, c->new_ctr_dist,
                                    c->need_assign, 0, 1, (const int32_t *)c->need_n, stream);
            if (rc != EB_OK) return rc;
            // three-way split (kmedoids.py:644-658)
            k_pam_classify<float><<<pam_blocks(c->n), kPamThreads, 0, s>>>(
                c->new_ctr_dist, c->dist, c->assign, c->n, cid, c->new_dist, c->new_assign,
                c->ambig_idx, (unsigned long long *)(c->scal_i + 1));
            EB_LAUNCH_CHECK();
        }
        // the proposal takes the medoid's slot (kmedoids.py:660-664); the old one is kept
        SlotExchange ex;
        ex.slot = c->medoid_xyz + (size_t)cid * frame_floats;
        ex.slot_trace = c->medoid_traces + cid;
        ex.saved = c->saved_xyz;
        ex.saved_trace = c->saved_traces;
        ex.prop = c->prop_xyz;
        ex.prop_trace = c->prop_traces;
        ex.frame_floats = (long)frame_floats;
        ex.scal_i = c->scal_i;
        ex.prop_idx = c->prop_idx;
        const bool listed = m_max > 0 && c->use_list;
        if (!listed) {
            k_take_slot<<<1, 512, 0, s>>>(ex.slot, ex.slot_trace, ex.saved, ex.saved_trace, ex.prop,
                                          ex.prop_trace, ex.frame_floats, ex.scal_i, ex.prop_idx,
                                          nullptr);
            EB_LAUNCH_CHECK();
        }
        // ambiguous frames against all medoids (kmedoids.py:666-670)
        if (listed) {
            EB_CHECK_ARG(c->med_list && c->med_list_n && c->med_list_cap > 0 && c->tc_ovf &&
                             c->med_list_cap <= 2048,
                         "pam_propose: missing medoid list");
            k_pam_medoid_list<<<1, 1024, 0, s>>>(c->ambig_idx, c->scal_i + 1, c->new_ctr_dist,
                                                 c->cc, c->k, cid, c->med_list_cap, c->med_list,
                                                 c->med_list_n, c->tc_ovf, ex);
            EB_LAUNCH_CHECK();
            long blocks = (m_max + kListWarps - 1) / kListWarps;
            if (blocks > 4L * sm_count()) blocks = 4L * sm_count();
            k_pam_reassign_listed<<<(int)blocks, kListThreads,
                                    sizeof(int) * (size_t)kListWarps * c->med_list_cap, s>>>(
                c->xyz, c->traces, c->n_atoms, rmsd_apad(c->n_atoms), c->medoid_xyz,
                c->medoid_traces, c->ambig_idx, c->scal_i + 1, c->new_ctr_dist, c->cc, cid,
                c->med_list, c->med_list_n, c->med_list_cap, c->new_dist, c->new_assign);
            EB_LAUNCH_CHECK();
        } else if (m_max > 0) {
            if (c->use_tc) {
                EB_CHECK_ARG(c->tc_cand && c->tc_scratch && c->tc_ovf,
                             "pam_propose: missing screen workspace");
                rc = eb_rmsd_assign_tc_dev(c->xyz, c->traces, m_max, c->n_atoms, c->medoid_xyz,
                                           c->medoid_traces, c->k, c->kappa, c->ambig_idx, 1,
                                           c->new_dist, c->new_assign, c->tc_cand, c->tc_scratch,
                                           nullptr, 1, (const int32_t *)(c->scal_i + 1),
                                           c->tc_ovf, stream);
            } else {
                rc = eb_rmsd_assign_dev(c->xyz, c->traces, c->n, c->n_atoms, c->medoid_xyz,
                                        c->medoid_traces, c->k, c->ambig_idx, m_max, c->new_dist,
                                        c->new_assign, 0, 1, (const int32_t *)(c->scal_i + 1),
                                        stream);
            }
            if (rc != EB_OK) return rc;
        }
        rc = eb_sum_squares(c->new_dist, c->n, 0, c->scal_d, c->scratch, stream);
        if (rc != EB_OK) return rc;
    }
    if (stages & EB_PAM_READBACK) {
        EB_CHECK_ARG(c->pin_d && c->pin_i, "pam_propose: missing pinned buffers");
        const bool packed = c->pin_o && c->tc_ovf &&
                            (const char *)c->scal_i == (const char *)c->scal_d + 8 &&
                            (const char *)c->tc_ovf == (const char *)c->scal_d + 24 &&
                            (const char *)c->pin_i == (const char *)c->pin_d + 8 &&
                            (const char *)c->pin_o == (const char *)c->pin_d + 24;
        if (packed) {
            // {cost, proposal, n_ambig, overflow} are adjacent on both sides: one copy
            EB_CUDA(cudaMemcpyAsync(c->pin_d, c->scal_d, 32, cudaMemcpyDeviceToHost, s));
        } else {
            EB_CUDA(cudaMemcpyAsync(c->pin_d, c->scal_d, sizeof(double), cudaMemcpyDeviceToHost,
                                    s));
            EB_CUDA(cudaMemcpyAsync(c->pin_i, c->scal_i, 2 * sizeof(int64_t),
                                    cudaMemcpyDeviceToHost, s));
            if (c->pin_o && c->tc_ovf)
                EB_CUDA(cudaMemcpyAsync(c->pin_o, c->tc_ovf, sizeof(int32_t),
                                        cudaMemcpyDeviceToHost, s));
        }
    }
    return EB_OK;
}

int eb_pam_restore_medoid(const eb_pam_ctx *c, int32_t cid, void *stream)
{
    EB_CHECK_ARG(c && c->medoid_xyz && c->saved_xyz && cid >= 0 && cid < c->k,
                 "pam_restore: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t frame_floats = 3 * (size_t)rmsd_apad(c->n_atoms);
    k_restore_slot<<<1, 512, 0, s>>>(c->medoid_xyz + (size_t)cid * frame_floats,
                                     c->medoid_traces + cid, c->saved_xyz, c->saved_traces,
                                     (long)frame_floats);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

size_t eb_struct_bytes(int which)
{
    return which == 0 ? sizeof(eb_kc_state) : which == 1 ? sizeof(eb_pam_ctx) : 0;
}

}  // extern "C"
