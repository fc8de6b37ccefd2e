// K3 (exact path): nearest of k centres for every frame, RMSD, float64 inner products.
//
// Reference behaviour reproduced (paths under /root/reference/enspara/):
//   cluster/util.py:159-205   assign_to_nearest_center: centres visited in order, strict '<',
//                             so the lowest centre index wins ties; init assignment 0, dist inf
//   cluster/kmedoids.py:666-667  the same pass restricted to X[dst_up_assig_this]
//
// Work decomposition: a group of 8 lanes owns one frame for the whole pass; centres are staged
// in shared memory as float64 tiles; each pass over the frame's atoms accumulates the 3x3
// matrices against C centres at once (the frame's float->double conversions are shared), the
// 9*C sums are butterflied over the 8 lanes and parked; after 8 centres are parked the 32 lanes
// of the warp each solve one (frame, centre) QCP quartic, the group takes the (distance,
// centre index) minimum and folds it into the frame's running best with a strict '<'.
// This is the exact-arithmetic path (parity); the tensor-core screening path feeds it the few
// centres per frame that survive the screen.
#include "eb_rmsd.cuh"

namespace eb {

constexpr int kAsgThreads = 256;
constexpr int kAsgWarps = kAsgThreads / 32;
constexpr int kAsgC = 2;  // centres per accumulation pass (the code below assumes 2)

template <int C>
__device__ __forceinline__ void frame_inner_products_multi(double m[C][9], const float *frame,
                                                           int A4, int l8,
                                                           const double *center_base,
                                                           size_t center_stride)
{
    const float4 *px = reinterpret_cast<const float4 *>(frame);
    const float4 *py = px + A4;
    const float4 *pz = py + A4;
    for (int j = l8; j < A4; j += 8) {
        const float4 x = __ldg(px + j);
        const float4 y = __ldg(py + j);
        const float4 z = __ldg(pz + j);
        const double fx[4] = {(double)x.x, (double)x.y, (double)x.z, (double)x.w};
        const double fy[4] = {(double)y.x, (double)y.y, (double)y.z, (double)y.w};
        const double fz[4] = {(double)z.x, (double)z.y, (double)z.z, (double)z.w};
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const CenterSmem cs =
                center_smem_carve(const_cast<double *>(center_base) + c * center_stride, A4);
            const double2 cxl = cs.lo[0][j], cxh = cs.hi[0][j];
            const double2 cyl = cs.lo[1][j], cyh = cs.hi[1][j];
            const double2 czl = cs.lo[2][j], czh = cs.hi[2][j];
            const double cx[4] = {cxl.x, cxl.y, cxh.x, cxh.y};
            const double cy[4] = {cyl.x, cyl.y, cyh.x, cyh.y};
            const double cz[4] = {czl.x, czl.y, czh.x, czh.y};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                m[c][0] = fma(fx[a], cx[a], m[c][0]);
                m[c][1] = fma(fx[a], cy[a], m[c][1]);
                m[c][2] = fma(fx[a], cz[a], m[c][2]);
                m[c][3] = fma(fy[a], cx[a], m[c][3]);
                m[c][4] = fma(fy[a], cy[a], m[c][4]);
                m[c][5] = fma(fy[a], cz[a], m[c][5]);
                m[c][6] = fma(fz[a], cx[a], m[c][6]);
                m[c][7] = fma(fz[a], cy[a], m[c][7]);
                m[c][8] = fma(fz[a], cz[a], m[c][8]);
            }
        }
    }
}

__global__ void __launch_bounds__(kAsgThreads)
k_rmsd_assign(const float *__restrict__ xyz, const double *__restrict__ traces, long n, int A,
              int A_pad, const float *__restrict__ centers, const double *__restrict__ ctraces,
              int k, const int64_t *__restrict__ frame_idx, long m, float *out_dist,
              int *out_assign, int accumulate, int scatter, int TC,
              const int *__restrict__ m_dev)
{
    if (m_dev) m = min(m, (long)__ldg(m_dev));   // device-side subset size (<= the host's m)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sums = reinterpret_cast<double *>(smem_raw);                 // [warps][32][9]
    double *ctr_trace = sums + kAsgWarps * 32 * 9;                       // [TC]
    double *ctr = ctr_trace + ((TC + 1) & ~1);                           // [TC][3*A_pad]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, l8 = lane & 7;
    const int A4 = A_pad >> 2;
    const size_t cstride = 3 * (size_t)A_pad;
    double *my_sums = sums + (size_t)warp * 32 * 9;
    const long groups_total = (long)gridDim.x * kAsgWarps * 4;

    for (int tile0 = 0; tile0 < k; tile0 += TC) {
        const int tc = min(TC, k - tile0);
        __syncthreads();
        for (int c = 0; c < tc; ++c) {
            const CenterSmem cs = center_smem_carve(ctr + c * cstride, A4);
            center_smem_fill(cs, centers + (size_t)(tile0 + c) * cstride, A_pad);
        }
        for (int c = threadIdx.x; c < tc; c += blockDim.x) ctr_trace[c] = ctraces[tile0 + c];
        __syncthreads();

        for (long wbase = ((long)blockIdx.x * kAsgWarps + warp) * 4; wbase < m;
             wbase += groups_total) {
            const long item = wbase + g;
            const bool valid = item < m;
            const long f = valid ? (frame_idx ? frame_idx[item] : item) : 0;
            float best_d = INFINITY;
            int best_c = 0;
            const long opos = scatter ? f : item;  // where this frame's result lives
            if (valid && (tile0 > 0 || accumulate)) {
                best_d = out_dist[opos];
                best_c = out_assign[opos];
            }
            const double Ga = valid ? traces[f] : 0.0;
            const float *frame = xyz + (size_t)f * cstride;

            for (int c8 = 0; c8 < tc; c8 += 8) {
#pragma unroll 1
                for (int cc = 0; cc < 8 && c8 + cc < tc; cc += kAsgC) {
                    const int c0 = c8 + cc;
                    const bool two = (c0 + 1 < tc);  // block-uniform
                    double acc[kAsgC][9];
#pragma unroll
                    for (int c = 0; c < kAsgC; ++c)
#pragma unroll
                        for (int e = 0; e < 9; ++e) acc[c][e] = 0.0;
                    if (valid) {
                        if (two)
                            frame_inner_products_multi<2>(acc, frame, A4, l8,
                                                          ctr + c0 * cstride, cstride);
                        else
                            frame_inner_products_multi<1>(
                                reinterpret_cast<double(*)[9]>(acc), frame, A4, l8,
                                ctr + c0 * cstride, cstride);
                    }
                    group8_reduce(acc[0]);
                    if (two) group8_reduce(acc[1]);
                    if (l8 == 0) {
                        double *dst = my_sums + (g * 8 + cc) * 9;
#pragma unroll
                        for (int e = 0; e < 9; ++e) dst[e] = acc[0][e];
                        if (two) {
#pragma unroll
                            for (int e = 0; e < 9; ++e) dst[9 + e] = acc[1][e];
                        }
                    }
                }
                __syncwarp();
                const int ci = c8 + l8;
                float d = INFINITY;
                int cidx = tile0 + ci;
                if (valid && ci < tc) {
                    double M[9];
                    const double *src = my_sums + (g * 8 + l8) * 9;
#pragma unroll
                    for (int e = 0; e < 9; ++e) M[e] = src[e];
                    d = rmsd_from_msd(qcp_msd(M, Ga, ctr_trace[ci], A));
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, d, off);
                    const int oc = __shfl_xor_sync(0xffffffffu, cidx, off);
                    if (od < d || (od == d && oc < cidx)) {
                        d = od;
                        cidx = oc;
                    }
                }
                if (d < best_d) {  // strict '<': an earlier centre keeps the frame on ties
                    best_d = d;
                    best_c = cidx;
                }
                __syncwarp();
            }
            if (valid && l8 == 0) {
                out_dist[opos] = best_d;
                out_assign[opos] = best_c;
            }
        }
    }
}

}  // namespace eb

using namespace eb;

extern "C" int eb_rmsd_assign(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                              const float *centers_soa, const double *center_traces, int32_t k,
                              const int64_t *frame_idx, int64_t n_idx, float *out_dist,
                              int32_t *out_assign, int accumulate, int scatter, void *stream)
{
    return eb_rmsd_assign_dev(xyz_soa, traces, n, n_atoms, centers_soa, center_traces, k,
                              frame_idx, n_idx, out_dist, out_assign, accumulate, scatter,
                              nullptr, stream);
}

// n_idx is the host's upper bound of the subset size, *n_idx_dev (optional, <= n_idx) the real
// one, known only on the device (PAM's ambiguous frames).
extern "C" int eb_rmsd_assign_dev(const float *xyz_soa, const double *traces, int64_t n,
                                  int n_atoms, const float *centers_soa,
                                  const double *center_traces, int32_t k,
                                  const int64_t *frame_idx, int64_t n_idx, float *out_dist,
                                  int32_t *out_assign, int accumulate, int scatter,
                                  const int32_t *n_idx_dev, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0 && k >= 0 && n_idx >= 0, "rmsd_assign: bad shape");
    const long m = frame_idx ? n_idx : n;
    if (m == 0 || k == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && centers_soa && center_traces && out_dist && out_assign,
                 "rmsd_assign: null pointer");
    const int A_pad = rmsd_apad(n_atoms);
    const size_t per_center = sizeof(double) * 3 * (size_t)A_pad;
    const size_t fixed = sizeof(double) * kAsgWarps * 32 * 9;
    const size_t budget = 200 * 1024;
    if (fixed + per_center + 16 > 227 * 1024)
        return fail(EB_ERR_LIMIT, "%s: n_atoms=%ld does not fit shared memory", "rmsd_assign",
                    (long)n_atoms);
    long TC = (long)((budget - fixed) / (per_center + 8));
    if (TC < 1) TC = 1;
    if (TC >= 8) TC &= ~7L;
    const long k8 = (k + 7) & ~7L;
    if (TC > k8) TC = k8;
    const size_t smem = fixed + sizeof(double) * ((TC + 1) & ~1L) + per_center * TC;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EB_CUDA(cudaFuncSetAttribute(k_rmsd_assign, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        configured = 227 * 1024;
    }
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;
    long blocks = (m + 31) / 32;
    if (blocks > (long)per_sm * sm_count()) blocks = (long)per_sm * sm_count();
    k_rmsd_assign<<<(int)blocks, kAsgThreads, smem, (cudaStream_t)stream>>>(
        xyz_soa, traces, n, n_atoms, A_pad, centers_soa, center_traces, k, frame_idx, m, out_dist,
        out_assign, accumulate, scatter, (int)TC, n_idx_dev);
    EB_LAUNCH_CHECK();
    return EB_OK;
}
