// K3 on the tensor cores: many-centres RMSD as 3xTF32 tcgen05 GEMMs with a fused QCP epilogue
// ("screen"), followed by exact float64 re-scoring of the few centres per frame that survive.
//
// Reference behaviour reproduced: cluster/util.py:159-205 (assign_to_nearest_center) -- the
// result is defined by the exact path (eb_rmsd_assign); this file only removes centres that
// provably (within the error model below) cannot be the nearest one.
//
// Formulation.  With the SoA layout, coordinate row i of frame f is a contiguous K-major vector
// of A_pad floats, and so is row j of centre c.  For each coordinate i the matrix
//     D_i[f, 3c+j] = sum_a X[f,i,a] * Y[c,j,a]
// is a plain (frames x atoms) . (atoms x 3*centres) GEMM; the three D_i share the B operand and
// put all nine entries of M(f,c) into ONE thread of the epilogue (TMEM lane f holds row f of
// D_0, D_1, D_2), so the 3x3 matrix is never exchanged between threads or written to memory.
// TF32 has 10 mantissa bits, so every operand is split x = hi + lo (hi = x truncated to TF32,
// lo = x - hi, exact) and D = Ahi.Bhi + Ahi.Blo + Alo.Bhi is accumulated in FP32 in TMEM
// (3xTF32).  Operands are pre-split in global memory by k_split_tf32 and arrive by TMA
// (64-byte rows, SWIZZLE_64B) in a 3-stage mbarrier ring; one elected thread issues
// tcgen05.mma.kind::tf32 (M=128 frames, N=96 = 32 centres x 3), 9 MMAs per 8 atoms.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5
// epilogue (TMEM lane quarter = warp_id % 4).
#include "eb_rmsd.cuh"
#include "eb_tma.cuh"

namespace eb {
namespace tc {

constexpr int BM = 128;                  // frames per tile  (= TMEM lanes)
constexpr int NC = 32;                   // centres per tile
constexpr int BN = 3 * NC;               // B rows per tile: (centre, coordinate)
constexpr int BK = 16;                   // atoms per stage: 64-byte rows
constexpr int STAGES = 3;
constexpr int A_TILE = BM * BK * 4;      // 8192
constexpr int B_TILE = BN * BK * 4;      // 6144
constexpr int STAGE_BYTES = 6 * A_TILE + 2 * B_TILE;  // {hi,lo} x 3 coords of A, {hi,lo} of B
constexpr int TMEM_COLS = 512;           // 3 accumulators x 96 columns = 288 -> power of two
constexpr int THREADS = 192;
constexpr int MAX_CAND = 128;            // survivors kept per frame before falling back

// ---- operand split -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_split_tf32(const float4 *__restrict__ x, long n4, float4 *__restrict__ hi,
             float4 *__restrict__ lo)
{
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n4;
         t += (long)gridDim.x * blockDim.x) {
        const float4 v = x[t];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        hi[t] = h;
        lo[t] = l;
    }
}

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int c0, int c1,
                                            int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float v[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                   "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_64B, 64-byte rows: 8-row groups are 512 B apart
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);   // start address
    d |= (uint64_t)1 << 16;                   // leading byte offset (16-B units; adjacent)
    d |= (uint64_t)(512 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                   // layout type SWIZZLE_64B
    return d;
}
constexpr uint32_t kIdesc = (1u << 4)          // D format F32
                            | (2u << 7)        // A format TF32
                            | (2u << 10)       // B format TF32
                            | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Smem {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full;
    uint64_t tmem_empty;
    uint32_t tmem_base;
};

// mode 0: debug, write the approximate M (9 floats per pair) to `dbg` (n x k x 9)
// mode 1: screen, maintain per-frame candidate lists
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
k_tc_screen(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
            const double *__restrict__ traces, const double *__restrict__ ctraces, long n, int k,
            int n_atoms, int A_pad, double kappa, float *dbg, int *cand_count, int *cand_list,
            float *cand_bound)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    Smem *sm = reinterpret_cast<Smem *>(tiles + (size_t)STAGES * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = A_pad / BK;
    const int n_ct = (k + NC - 1) / NC;
    const long n_ft = (n + BM - 1) / BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm->full[s], 1);
            mbar_init(&sm->empty[s], 1);
        }
        mbar_init(&sm->tmem_full, 1);
        mbar_init(&sm->tmem_empty, 4);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&sm->tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer ===================================================
        if (lane == 0) {
            tma_prefetch_desc(&tm_a_hi);
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_b_hi);
            tma_prefetch_desc(&tm_b_lo);
            int stage = 0;
            uint32_t phase = 0;
            for (long ft = blockIdx.x; ft < n_ft; ft += gridDim.x) {
                for (int ct = 0; ct < n_ct; ++ct) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&sm->empty[stage], phase ^ 1);
                        unsigned char *st = tiles + (size_t)stage * STAGE_BYTES;
                        mbar_expect_tx(&sm->full[stage], STAGE_BYTES);
                        for (int i = 0; i < 3; ++i) {
                            tma_load_3d(st + (2 * i) * A_TILE, &tm_a_hi, kb * BK, i,
                                        (int)(ft * BM), &sm->full[stage]);
                            tma_load_3d(st + (2 * i + 1) * A_TILE, &tm_a_lo, kb * BK, i,
                                        (int)(ft * BM), &sm->full[stage]);
                        }
                        tma_load_2d(st + 6 * A_TILE, &tm_b_hi, kb * BK, ct * BN,
                                    &sm->full[stage]);
                        tma_load_2d(st + 6 * A_TILE + B_TILE, &tm_b_lo, kb * BK, ct * BN,
                                    &sm->full[stage]);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (long ft = blockIdx.x; ft < n_ft; ft += gridDim.x) {
                for (int ct = 0; ct < n_ct; ++ct) {
                    mbar_wait(&sm->tmem_empty, tphase ^ 1);  // epilogue drained the accumulators
                    tc_fence_after();
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&sm->full[stage], phase);
                        tc_fence_after();
                        const uint32_t st = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
                        const uint64_t bhi = smem_desc_sw64(st + 6 * A_TILE);
                        const uint64_t blo = smem_desc_sw64(st + 6 * A_TILE + B_TILE);
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint64_t koff = (uint64_t)((ks * 32) >> 4);  // 8 tf32 = 32 B
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                const uint64_t ahi = smem_desc_sw64(st + (2 * i) * A_TILE);
                                const uint64_t alo = smem_desc_sw64(st + (2 * i + 1) * A_TILE);
                                const uint32_t d = tmem_base + (uint32_t)(i * BN);
                                const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                                tc_mma_tf32(d, ahi + koff, bhi + koff, kIdesc, first);
                                tc_mma_tf32(d, ahi + koff, blo + koff, kIdesc, 1u);
                                tc_mma_tf32(d, alo + koff, bhi + koff, kIdesc, 1u);
                            }
                        }
                        tc_commit(&sm->empty[stage]);  // smem slot reusable once these MMAs retire
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc_commit(&sm->tmem_full);
                    tphase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> QCP ============================
        const int q = warp & 3;                    // TMEM lane quarter of this warp
        const int row = q * 32 + lane;             // frame within the tile
        uint32_t tphase = 0;
        for (long ft = blockIdx.x; ft < n_ft; ft += gridDim.x) {
            const long f = ft * BM + row;
            const bool fvalid = f < n;
            const double Ga = fvalid ? traces[f] : 0.0;
            // screen state of this frame (thread-private across all centre tiles)
            double umin = 1e300;
            int ncand = 0;
            for (int ct = 0; ct < n_ct; ++ct) {
                mbar_wait(&sm->tmem_full, tphase);
                tphase ^= 1;
                tc_fence_after();
                const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                for (int g = 0; g < NC / 8; ++g) {  // 8 centres = 24 columns per accumulator
                    float m[3][24];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint32_t col = (uint32_t)(i * BN + g * 24);
                        tc_ld8(lane_addr + col, &m[i][0]);
                        tc_ld8(lane_addr + col + 8, &m[i][8]);
                        tc_ld8(lane_addr + col + 16, &m[i][16]);
                    }
                    tc_wait_ld();
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        const int c = ct * NC + g * 8 + cc;
                        if (!fvalid || c >= k) continue;
                        if (MODE == 0) {
                            float *o = dbg + ((size_t)f * k + c) * 9;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) o[3 * i + j] = m[i][3 * cc + j];
                        } else {
                            double M[9];
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) M[3 * i + j] = (double)m[i][3 * cc + j];
                            const double Gb = ctraces[c];
                            const double msd = qcp_msd(M, Ga, Gb, n_atoms);
                            const double e = kappa * sqrt(Ga * Gb) / (double)n_atoms;
                            const double lo = msd - e, up = msd + e;
                            if (lo <= umin) {
                                if (ncand < MAX_CAND) {
                                    cand_list[(size_t)f * MAX_CAND + ncand] = c;
                                    cand_bound[(size_t)f * MAX_CAND + ncand] = __double2float_rd(lo);
                                }
                                ++ncand;
                            }
                            if (up < umin) umin = up;
                            if (ncand == MAX_CAND + 1) {
                                // list overflowed: compact it against the current bound once
                                int w = 0;
                                for (int s = 0; s < MAX_CAND; ++s) {
                                    const float b = cand_bound[(size_t)f * MAX_CAND + s];
                                    const int cs = cand_list[(size_t)f * MAX_CAND + s];
                                    if ((double)b <= umin) {
                                        cand_list[(size_t)f * MAX_CAND + w] = cs;
                                        cand_bound[(size_t)f * MAX_CAND + w] = b;
                                        ++w;
                                    }
                                }
                                if (w < MAX_CAND) {
                                    cand_list[(size_t)f * MAX_CAND + w] = c;
                                    cand_bound[(size_t)f * MAX_CAND + w] = __double2float_rd(lo);
                                    ncand = w + 1;
                                } else {
                                    ncand = MAX_CAND + 2;  // sticky overflow: exact fallback
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm->tmem_empty);
            }
            if (MODE == 1 && fvalid) {
                if (ncand > MAX_CAND) {
                    cand_count[f] = -1;  // overflow -> the host routes this frame to the exact path
                } else {
                    // final prune against the final bound (order preserved: ascending centre id)
                    int w = 0;
                    for (int s = 0; s < ncand; ++s) {
                        const float b = cand_bound[(size_t)f * MAX_CAND + s];
                        const int cs = cand_list[(size_t)f * MAX_CAND + s];
                        if ((double)b <= umin) {
                            cand_list[(size_t)f * MAX_CAND + w] = cs;
                            ++w;
                        }
                    }
                    cand_count[f] = w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(TMEM_COLS));
    }
}

// ---- exact re-scoring of the surviving (frame, centre) pairs ---------------------------------
// A group of 8 lanes owns one frame and walks its candidate list in ascending centre order with
// the strict '<' of cluster/util.py:201; centres are read from global memory (few pairs).
__global__ void __launch_bounds__(256)
k_rescore(const float *__restrict__ xyz, const double *__restrict__ traces, long n, int n_atoms,
          int A_pad, const float *__restrict__ centers, const double *__restrict__ ctraces,
          const int *__restrict__ cand_count, const int *__restrict__ cand_list, float *out_dist,
          int *out_assign)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, l8 = lane & 7;
    const int A4 = A_pad >> 2;
    const size_t stride = 3 * (size_t)A_pad;
    const long groups = (long)gridDim.x * (blockDim.x >> 3);
    for (long fb = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; fb < n;
         fb += groups) {
        const long f = fb + g;
        const bool valid = f < n;
        const int cnt = valid ? cand_count[f] : 0;
        int cnt_max = cnt;
        cnt_max = max(cnt_max, __shfl_xor_sync(0xffffffffu, cnt_max, 8));
        cnt_max = max(cnt_max, __shfl_xor_sync(0xffffffffu, cnt_max, 16));
        float best_d = INFINITY;
        int best_c = 0;
        const float4 *px = reinterpret_cast<const float4 *>(xyz + (size_t)(valid ? f : 0) * stride);
        for (int s = 0; s < cnt_max; ++s) {
            const bool act = valid && s < cnt;
            const int c = act ? cand_list[(size_t)f * MAX_CAND + s] : 0;
            const float4 *pc = reinterpret_cast<const float4 *>(centers + (size_t)c * stride);
            double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (act) {
                for (int j = l8; j < A4; j += 8) {
                    const float4 x = __ldg(px + j), y = __ldg(px + A4 + j), z = __ldg(px + 2 * A4 + j);
                    const float4 cx = __ldg(pc + j), cy = __ldg(pc + A4 + j), cz = __ldg(pc + 2 * A4 + j);
                    acc_atom(m, x.x, y.x, z.x, (double)cx.x, (double)cy.x, (double)cz.x);
                    acc_atom(m, x.y, y.y, z.y, (double)cx.y, (double)cy.y, (double)cz.y);
                    acc_atom(m, x.z, y.z, z.z, (double)cx.z, (double)cy.z, (double)cz.z);
                    acc_atom(m, x.w, y.w, z.w, (double)cx.w, (double)cy.w, (double)cz.w);
                }
            }
            group8_reduce(m);
            if (act) {
                const float d = rmsd_from_msd(qcp_msd(m, traces[f], ctraces[c], n_atoms));
                if (d < best_d) {
                    best_d = d;
                    best_c = c;
                }
            }
        }
        if (valid && l8 == 0 && cnt > 0) {
            out_dist[f] = best_d;
            out_assign[f] = best_c;
        }
    }
}

static int make_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1,
                        uint32_t b2, CUtensorMapSwizzle swz)
{
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(EB_ERR_CUDA, "%s", "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstride[2] = {s1_bytes, s2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), gdim,
                           gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(EB_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%s %ld)", "code", (long)r);
    return EB_OK;
}

}  // namespace tc
}  // namespace eb

using namespace eb;

extern "C" {

size_t eb_tc_scratch_bytes(int64_t n, int n_atoms, int32_t k)
{
    const size_t row = sizeof(float) * 3 * (size_t)rmsd_apad(n_atoms);
    // split copies of frames and centres + candidate lists
    return 2 * row * (size_t)n + 2 * row * (size_t)(k + tc::NC) +
           (size_t)n * (sizeof(int) + tc::MAX_CAND * (sizeof(int) + sizeof(float))) + 4096;
}

// mode 0 (debug): dbg receives the approximate inner-product matrices, n x k x 9 floats.
// mode 1: screen + exact re-score.  out_dist/out_assign get the exact result for every frame whose
//         candidate list did not overflow; n_overflow (device int) counts the others, whose
//         cand_count is -1 and which the caller must send through eb_rmsd_assign.
int eb_rmsd_assign_tc(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                      const float *centers_soa, const double *center_traces, int32_t k,
                      double kappa, float *out_dist, int32_t *out_assign, int32_t *cand_count,
                      void *scratch, float *dbg, int mode, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0 && k >= 0, "rmsd_assign_tc: bad shape");
    if (n == 0 || k == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && centers_soa && center_traces && scratch,
                 "rmsd_assign_tc: null pointer");
    EB_CHECK_ARG(n < (1L << 31) / 3, "rmsd_assign_tc: too many frames for one pass");
    cudaStream_t s = (cudaStream_t)stream;
    const int A_pad = rmsd_apad(n_atoms);
    EB_CHECK_ARG(A_pad % tc::BK == 0, "rmsd_assign_tc: padded atom count must be a multiple of 16");
    const size_t row = 3 * (size_t)A_pad;  // floats per frame
    float *a_hi = (float *)scratch;
    float *a_lo = a_hi + row * n;
    float *b_hi = a_lo + row * n;
    const size_t kb_rows = (size_t)k + tc::NC;  // slack so TMA boxes never start out of bounds
    float *b_lo = b_hi + row * kb_rows;
    int *cand_list = (int *)(b_lo + row * kb_rows);
    float *cand_bound = (float *)(cand_list + (size_t)n * tc::MAX_CAND);

    {
        const long n4 = (long)(row * n / 4);
        long blocks = (n4 + 255) / 256;
        if (blocks > 16L * sm_count()) blocks = 16L * sm_count();
        tc::k_split_tf32<<<(int)blocks, 256, 0, s>>>((const float4 *)xyz_soa, n4, (float4 *)a_hi,
                                                     (float4 *)a_lo);
        EB_LAUNCH_CHECK();
        const long k4 = (long)(row * k / 4);
        blocks = (k4 + 255) / 256;
        tc::k_split_tf32<<<(int)blocks, 256, 0, s>>>((const float4 *)centers_soa, k4,
                                                     (float4 *)b_hi, (float4 *)b_lo);
        EB_LAUNCH_CHECK();
    }
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const uint64_t pitch = (uint64_t)A_pad * 4;
    int rc;
    rc = tc::make_tmap_3d(&ta_hi, a_hi, A_pad, 3, (uint64_t)n, pitch, 3 * pitch, tc::BK, 1, tc::BM,
                          CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = tc::make_tmap_3d(&ta_lo, a_lo, A_pad, 3, (uint64_t)n, pitch, 3 * pitch, tc::BK, 1, tc::BM,
                          CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_tmap_2d(&tb_hi, b_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 3 * (uint64_t)k,
                      (uint64_t)A_pad, pitch, tc::BN, tc::BK, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_tmap_2d(&tb_lo, b_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 3 * (uint64_t)k,
                      (uint64_t)A_pad, pitch, tc::BN, tc::BK, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;

    const size_t smem = (size_t)tc::STAGES * tc::STAGE_BYTES + sizeof(tc::Smem) + 1024;
    const long n_ft = (n + tc::BM - 1) / tc::BM;
    const int grid = (int)(n_ft < sm_count() ? n_ft : sm_count());
    if (mode == 0) {
        EB_CHECK_ARG(dbg, "rmsd_assign_tc: debug buffer missing");
        EB_CUDA(cudaFuncSetAttribute(tc::k_tc_screen<0>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::k_tc_screen<0><<<grid, tc::THREADS, smem, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, traces,
                                                           center_traces, n, k, n_atoms, A_pad,
                                                           kappa, dbg, nullptr, nullptr, nullptr);
        EB_LAUNCH_CHECK();
        return EB_OK;
    }
    EB_CHECK_ARG(out_dist && out_assign && cand_count, "rmsd_assign_tc: null output");
    EB_CUDA(cudaFuncSetAttribute(tc::k_tc_screen<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    tc::k_tc_screen<1><<<grid, tc::THREADS, smem, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, traces,
                                                       center_traces, n, k, n_atoms, A_pad, kappa,
                                                       nullptr, cand_count, cand_list, cand_bound);
    EB_LAUNCH_CHECK();
    long blocks = (n + 31) / 32;
    if (blocks > 8L * sm_count()) blocks = 8L * sm_count();
    tc::k_rescore<<<(int)blocks, 256, 0, s>>>(xyz_soa, traces, n, n_atoms, A_pad, centers_soa,
                                              center_traces, cand_count, cand_list, out_dist,
                                              out_assign);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

}  // extern "C"
