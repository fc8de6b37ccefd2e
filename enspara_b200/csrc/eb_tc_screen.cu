// K3 on the tensor cores: many-centres RMSD as split-FP16 tcgen05 GEMMs with a fused QCP epilogue
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
// Operand precision: every coordinate is scaled by 2^8 and split into TWO FP16 numbers,
// x*256 = h1 + h2 with h1 = rn_fp16(x*256), h2 = rn_fp16(x*256 - h1).  Round-to-nearest keeps
// 11 + 11 significant bits plus the sign of h2, i.e. |x*256 - h1 - h2| <= 2^-24 |x*256| (FP32
// input precision) as long as h2 is a normal FP16 number and an absolute 2^-25 otherwise, and
// D = A1.B1 + A1.B2 + A2.B1 is accumulated in FP32 in TMEM (the dropped A2.B2 term is 2^-24
// relative).  Compared with the 3xTF32 split of round 1 (hi/lo FP32 words: 8 bytes per element,
// 8 atoms per MMA) this is 4 bytes per element and 16 atoms per MMA at the same cycles per
// dispatch: half the operand traffic and half the tensor time for the same accuracy class --
// the error that matters is the FP32 accumulation inside the tensor core (see TC_KAPPA in
// cluster/_ops.py).  Coordinates beyond +-255 nm would overflow FP16: the pack kernel raises
// a flag and the whole pass falls back to the exact kernel.  Operands are pre-split AND
// pre-packed by k_pack_f16x2 into the exact swizzled shared-memory image of every tile
// (64-byte rows = 32 atoms, SWIZZLE_64B), so one pipeline stage arrives with two contiguous
// bulk copies (cp.async.bulk) in a 3-stage mbarrier ring; one elected thread issues
// tcgen05.mma.kind::f16 (M=128 frames, N=144 = 48 centres x 3), 9 MMAs per 16 atoms.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5
// epilogue (TMEM lane quarter = warp_id % 4).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "eb_rmsd.cuh"
#include "eb_tma.cuh"

namespace eb {
namespace tc {

constexpr int BM = 128;                  // frames per tile  (= TMEM lanes)
constexpr int NC = 48;                   // centres per tile
constexpr int BN_USED = 3 * NC;          // B rows per tile that carry (centre, coordinate) ...
constexpr int BN = 144;                  // ... padded with zero rows to the UMMA N step of 16
constexpr int BK = 32;                   // atoms per stage: 64-byte rows of FP16
constexpr int STAGES = 3;
constexpr int A_TILE = BM * BK * 2;      // 8192
constexpr int B_TILE = BN * BK * 2;      // 9216
constexpr int STAGE_BYTES = 6 * A_TILE + 2 * B_TILE;  // {h1,h2} x 3 coords of A, {h1,h2} of B
constexpr float kOperandScale = 256.0f;               // x*2^8 before the FP16 split
constexpr float kAccumUnscale = 1.0f / 65536.0f;      // accumulators hold 2^16 * M
constexpr float kF16Limit = 65000.0f;                 // |x*2^8| beyond this cannot be split
// Tile shape.  tcgen05 operands are read from shared memory at 128 B/clk/SM: an M=128 x N MMA
// of 16 FP16 atoms takes max(N/2, (4096 + 32 N) / 128) clocks, so N = 144 (72 clk for 8.6 KB)
// is tensor-bound while N = 80 -- the widest tile that leaves TMEM room for a SECOND accumulator
// set (2 x 3 x 80 = 480 columns) -- is operand-read-bound: measured, the MMA stream of the
// 262144 x 1008 x 500 pass takes 4.97 ms at N = 144 and 7.05 ms at N = 80.  So: one accumulator
// set of 3 x 144 columns, and the overlap comes from draining it early instead -- twelve
// epilogue warps (three per TMEM lane quarter) each own 16 centres = two groups of 8; a warp
// pulls a group's 72 values into registers, and after the LAST group's load it hands TMEM back,
// so the next tile's MMAs run under the second half of the QCP arithmetic.
constexpr int ACC_SETS = (2 * 3 * BN <= 512) ? 2 : 1;
constexpr int ACC_COLS = 3 * BN;         // columns of one accumulator set
constexpr int TMEM_COLS = 512;
constexpr int EPI_PARTS = 3;   // epilogue warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * EPI_PARTS;
constexpr int GROUPS_PER_WARP = NC / 8 / EPI_PARTS;   // groups of 8 centres (24 columns) per warp
constexpr int THREADS = 64 + 32 * EPI_WARPS;  // producer, MMA, 12 epilogue warps
constexpr int CAND_BUDGET = 512;  // candidate entries per frame, shared out over its lists
constexpr int MIN_CAND = 8;       // ... but never fewer than this per list
constexpr int DEF_SEG = 8;     // centre segments: CTAs sharing a frame tile hit it in L2 (measured: 4 -> 8 takes dram reads from 2.9x to 1.6x the operand image, screen 8.15 -> 7.4 ms)
constexpr int MAX_SEG = 32;    // small frame subsets (PAM) are spread over more segments
static_assert(NC == 8 * EPI_PARTS * GROUPS_PER_WARP, "groups of 8 centres per epilogue warp");
static_assert(ACC_SETS * ACC_COLS <= TMEM_COLS, "two accumulator sets must fit TMEM");
static_assert(BN % 16 == 0 && BN <= 256 && BN >= BN_USED, "UMMA N constraint for M = 128");
static_assert(B_TILE % 512 == 0 && STAGE_BYTES % 1024 == 0, "SWIZZLE_64B atoms stay aligned");

// ---- operand split + packing -----------------------------------------------------------------
// Writes, for every (tile T, k-block kb, sub-row-set s, h1|h2), the exact shared-memory image of
// the MMA operand tile (RT rows x 64 bytes = 32 FP16 atoms, SWIZZLE_64B: 16-byte chunk c of row
// r lives at chunk c ^ ((r >> 1) & 3)), contiguously in global memory, so that a whole pipeline
// stage arrives with ONE bulk copy instead of hundreds of 64-byte TMA row requests.
//   image index = (((T * KB + kb) * S + s) * 2 + hl) * (RT * 64 bytes)
//   source row  = (T * RT + r) * S + s   of the (rows_total x A_pad) float matrix `x`
// Frames: RT = RU = 128, S = 3 (the three coordinate rows of a frame); centres: S = 1, RT = BN
// image rows per tile of which the first RU = BN_USED carry source rows (T*RU + r), the rest
// zeros (BN == BN_USED = 144 today).
// `row_idx` (optional, frames only): tile row (T*RT + r) is frame row_idx[T*RT + r] of `x`.
// One thread = one 16-byte chunk = 8 atoms.  *overflow is set when a coordinate cannot be
// represented (|x| * 2^8 beyond the FP16 range): the caller then uses the exact kernel.
__device__ __forceinline__ void split_f16x2(float x, __half &h1, __half &h2, bool &bad)
{
    const float v = x * kOperandScale;             // exact (power of two)
    bad |= !(fabsf(v) <= kF16Limit);               // also catches NaN
    h1 = __float2half_rn(v);
    h2 = __float2half_rn(v - __half2float(h1));    // the difference is exact in FP32
}

__global__ void __launch_bounds__(256)
k_pack_f16x2(const float *__restrict__ x, long rows_total, int A_pad, int RT, int RU, int S,
             long n_tiles, const int64_t *__restrict__ row_idx, unsigned char *__restrict__ img,
             int *overflow, const int *__restrict__ n_items_dev)
{
    // optional device-side item count (PAM: the number of ambiguous frames is only known on
    // the device): rows of items beyond it are zero-filled and their row_idx is never read
    if (n_items_dev) rows_total = min(rows_total, (long)S * (long)__ldg(n_items_dev));
    const int Qs = A_pad >> 3;               // 8-atom chunks per source row (A_pad % 8 == 0)
    const int KB = (A_pad + BK - 1) / BK;    // k-blocks of the image (zero-padded to 32 atoms)
    const int Q = KB * 4;                    // 16-byte chunks per image row
    const long total = n_tiles * RT * S * (long)Q;
    const size_t tile_bytes = (size_t)RT * 64;
    bool bad = false;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long)gridDim.x * blockDim.x) {
        const int q = (int)(t % Q);
        const long ri = t / Q;           // image row (T, r, s) flattened as ((T*RT + r)*S + s)
        const int sidx = (int)(ri % S);
        const long tri = ri / S;         // T*RT + r
        const long T = tri / RT;
        const int r = (int)(tri - T * RT);
        const long tr = T * RU + r;      // source item (frame / centre row) of this image row
        const long rr = tr * S + sidx;   // source row
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (r < RU && rr < rows_total && q < Qs) {
            const long src = row_idx ? (long)__ldg(row_idx + tr) * S + sidx : rr;
            const float4 *p = reinterpret_cast<const float4 *>(x + (size_t)src * A_pad) + 2 * q;
            v0 = __ldg(p);
            v1 = __ldg(p + 1);
        }
        const float in[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        __align__(16) __half h1[8], h2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_f16x2(in[e], h1[e], h2[e], bad);
        const int kb = q >> 2, c = q & 3;
        const size_t base = ((((size_t)T * KB + kb) * S + sidx) * 2) * tile_bytes +
                            (size_t)r * 64 + (size_t)((c ^ ((r >> 1) & 3)) << 4);
        *reinterpret_cast<uint4 *>(img + base) = *reinterpret_cast<const uint4 *>(h1);
        *reinterpret_cast<uint4 *>(img + base + tile_bytes) = *reinterpret_cast<const uint4 *>(h2);
    }
    if (bad) atomicOr(overflow, 1);
}

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
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
                            | (0u << 7)        // A format F16
                            | (0u << 10)       // B format F16
                            | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);


// FP32 QCP for the screen, four independent pairs at a time (instruction-level parallelism: an
// epilogue warp has its scheduler almost to itself).  Inputs are normalised by s = sqrt(Ga*Gb)
// so that |M'| <= 1 and the start g = (Ga+Gb)/(2s) >= 1 whatever the size of the molecule;
// lam' = lambda_max / s.  Eight Newton steps from above, no early exit; `dl` returns the size
// of the last step so that the caller can refuse to trust a solve that has not settled.
#ifndef EB_TC_NEWTON
#define EB_TC_NEWTON 8
#endif
constexpr int kNewtonSteps = EB_TC_NEWTON;
__device__ __forceinline__ void qcp4_f32(const float M[4][9], const float g[4], float lam[4],
                                         float dl[4])
{
    float c2[4], c1[4], c0[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float Sxx = M[p][0], Sxy = M[p][1], Sxz = M[p][2];
        const float Syx = M[p][3], Syy = M[p][4], Syz = M[p][5];
        const float Szx = M[p][6], Szy = M[p][7], Szz = M[p][8];
        const float k00 = Sxx + Syy + Szz, k01 = Syz - Szy, k02 = Szx - Sxz, k03 = Sxy - Syx;
        const float k11 = Sxx - Syy - Szz, k12 = Sxy + Syx, k13 = Szx + Sxz;
        const float k22 = -Sxx + Syy - Szz, k23 = Syz + Szy;
        const float k33 = -Sxx - Syy + Szz;
        c2[p] = -2.0f * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz +
                         Szx * Szx + Szy * Szy + Szz * Szz);
        c1[p] = -8.0f * (Sxx * (Syy * Szz - Syz * Szy) - Sxy * (Syx * Szz - Syz * Szx) +
                         Sxz * (Syx * Szy - Syy * Szx));
        const float a01 = k00 * k11 - k01 * k01, a02 = k00 * k12 - k02 * k01;
        const float a03 = k00 * k13 - k03 * k01, a12 = k01 * k12 - k02 * k11;
        const float a13 = k01 * k13 - k03 * k11, a23 = k02 * k13 - k03 * k12;
        const float b01 = k02 * k13 - k12 * k03, b02 = k02 * k23 - k22 * k03;
        const float b03 = k02 * k33 - k23 * k03, b12 = k12 * k23 - k22 * k13;
        const float b13 = k12 * k33 - k23 * k13, b23 = k22 * k33 - k23 * k23;
        c0[p] = a01 * b23 - a02 * b13 + a03 * b12 + a12 * b03 - a13 * b02 + a23 * b01;
        lam[p] = g[p];
        dl[p] = 0.f;
    }
#pragma unroll
    for (int it = 0; it < kNewtonSteps; ++it) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float l = lam[p], l2 = l * l;
            const float b = (l2 + c2[p]) * l;
            const float a = b + c1[p];
            const float den = 2.0f * l2 * l + b + a;
            const float delta = (den != 0.0f) ? __fdividef(a * l + c0[p], den) : 0.0f;
            lam[p] = l - delta;
            dl[p] = delta;
        }
    }
}
// slack for the FP32 solve, relative to (Ga+Gb)/2, in units of N*msd (measured: < 8e-6)
constexpr float kQcpSlackF = 6.8e-5f;   // 6.4e-5 + 4e-6 for the FP32 rounding of the bound arithmetic

struct Smem {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[ACC_SETS];
    uint64_t tmem_empty[ACC_SETS];
    uint32_t tmem_base;
};

// mode 0: debug, write the approximate M (9 floats per pair) to `dbg` (n x k x 9)
// mode 1: screen, maintain per-frame candidate lists
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
k_tc_screen(const unsigned char *__restrict__ a_img, const unsigned char *__restrict__ b_img,
            const double *__restrict__ traces, const double *__restrict__ ctraces, long n, int k,
            int n_atoms, int A_pad, double kappa, float *dbg, int *cand_count, int *cand_list,
            float *cand_bound, float *cand_umin, int n_seg, const int64_t *__restrict__ frame_idx,
            int MAX_CAND, const int *__restrict__ n_dev)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw;
    if (n_dev) n = min(n, (long)__ldg(n_dev));   // device-side frame count (<= the host's n)
    Smem *sm = reinterpret_cast<Smem *>(tiles + (size_t)STAGES * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = (A_pad + BK - 1) / BK;
    const int n_ct = (k + NC - 1) / NC;
    const long n_ft = (n + BM - 1) / BM;
    // work item = (frame tile, centre segment); consecutive CTAs take the segments of the SAME
    // frame tile, so its A operand is fetched from DRAM once and then served by L2
    const long n_items = n_ft * n_seg;
    const int ct_per_seg = (n_ct + n_seg - 1) / n_seg;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm->full[s], 1);
            mbar_init(&sm->empty[s], 1);
        }
        for (int a = 0; a < ACC_SETS; ++a) {
            mbar_init(&sm->tmem_full[a], 1);
            mbar_init(&sm->tmem_empty[a], EPI_WARPS);
        }
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
            int stage = 0;
            uint32_t phase = 0;
            for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
                const long ft = item / n_seg;
                const int seg = (int)(item - ft * n_seg);
                const int ct_hi = min(n_ct, (seg + 1) * ct_per_seg);
                for (int ct = seg * ct_per_seg; ct < ct_hi; ++ct) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&sm->empty[stage], phase ^ 1);
                        unsigned char *st = tiles + (size_t)stage * STAGE_BYTES;
                        mbar_expect_tx(&sm->full[stage], STAGE_BYTES);
                        // stage image = [coord i][hi|lo] A tiles (48 KB) then [hi|lo] B tiles
                        bulk_g2s(st, a_img + ((size_t)ft * KB + kb) * (6 * A_TILE), 6 * A_TILE,
                                 &sm->full[stage]);
                        bulk_g2s(st + 6 * A_TILE, b_img + ((size_t)ct * KB + kb) * (2 * B_TILE),
                                 2 * B_TILE, &sm->full[stage]);
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
            uint32_t phase = 0;
            uint32_t tile_no = 0;          // centre tiles issued by this CTA: set = tile_no & 1
            for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
                const long ft = item / n_seg;
                const int seg = (int)(item - ft * n_seg);
                const int ct_hi = min(n_ct, (seg + 1) * ct_per_seg);
                (void)ft;
                for (int ct = seg * ct_per_seg; ct < ct_hi; ++ct, ++tile_no) {
                    const uint32_t as = tile_no % ACC_SETS;
                    // the epilogue drained this accumulator set (its previous use)
                    mbar_wait(&sm->tmem_empty[as], ((tile_no / ACC_SETS) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t acc = tmem_base + as * (uint32_t)ACC_COLS;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&sm->full[stage], phase);
                        tc_fence_after();
                        const uint32_t st = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
                        const uint64_t bhi = smem_desc_sw64(st + 6 * A_TILE);
                        const uint64_t blo = smem_desc_sw64(st + 6 * A_TILE + B_TILE);
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            const uint64_t koff = (uint64_t)((ks * 32) >> 4);  // 16 f16 = 32 B
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                const uint64_t ahi = smem_desc_sw64(st + (2 * i) * A_TILE);
                                const uint64_t alo = smem_desc_sw64(st + (2 * i + 1) * A_TILE);
                                const uint32_t d = acc + (uint32_t)(i * BN);
                                const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                                tc_mma_f16(d, ahi + koff, bhi + koff, kIdesc, first);
                                tc_mma_f16(d, ahi + koff, blo + koff, kIdesc, 1u);
                                tc_mma_f16(d, alo + koff, bhi + koff, kIdesc, 1u);
                            }
                        }
                        tc_commit(&sm->empty[stage]);  // smem slot reusable once these MMAs retire
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc_commit(&sm->tmem_full[as]);
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> QCP ============================
        // warp w (2..13): TMEM lane quarter w & 3 (hardware rule), centres 8*part .. 8*part+7 of
        // the tile with part = (w - 2) >> 2
        const int q = warp & 3;
        const int part = (warp - 2) >> 2;
        const int row = q * 32 + lane;             // frame within the tile
        const int n_lists = n_seg * EPI_PARTS;
        uint32_t tile_no = 0;
        for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
            const long ft = item / n_seg;
            const int seg = (int)(item - ft * n_seg);
            const int ct_hi = min(n_ct, (seg + 1) * ct_per_seg);
            const long f = ft * BM + row;
            const bool fvalid = f < n;
            const double Ga = fvalid ? traces[frame_idx ? frame_idx[f] : f] : 0.0;
            const float gaf = (float)Ga;
            const float sa = sqrtf(gaf);
            const float kappaf = (float)kappa * 1.000001f;
            // screen state of this (frame, list): thread-private across the centre tiles; all
            // bounds are kept in units of N * msd.  FP32 throughout: the rounding of
            // v = 2 (half - sc * lam) is <= 3e-7 * half, two orders below the slack that is
            // added anyway (kQcpSlackF carries an extra 4e-6 for it); the float64 version of
            // this bookkeeping (DADD / DMUL / DSETP / F2F) held 36 % of the kernel's stall
            // samples for 5 % of its instructions.
            float umin = 3.0e38f;
            int ncand = 0;
            const int lid = seg * EPI_PARTS + part;
            const size_t slot = ((size_t)(fvalid ? f : 0) * n_lists + lid) * MAX_CAND;
            for (int ct = seg * ct_per_seg; ct < ct_hi; ++ct, ++tile_no) {
                const uint32_t as = tile_no % ACC_SETS;
                // the traces of this warp's 16 centres (and their square roots), one per lane,
                // fetched while the tile's MMAs still run; the pair loop broadcasts them with
                // shuffles instead of every lane loading and converting the same value
                float my_gb = 0.f, my_sq = 0.f;
                if (lane < 16) {
                    my_gb = (float)__ldg(ctraces + min(ct * NC + part * 16 + lane, k - 1));
                    my_sq = sqrtf(my_gb);
                }
                mbar_wait(&sm->tmem_full[as], (tile_no / ACC_SETS) & 1u);
                tc_fence_after();
                const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                           as * (uint32_t)ACC_COLS;
#pragma unroll 1
                for (int gg = 0; gg < GROUPS_PER_WARP; ++gg) {   // 8 centres = 24 columns
                    const int g = part * GROUPS_PER_WARP + gg;
                    float m[3][24];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint32_t col = (uint32_t)(i * BN + g * 24);
                        tc_ld8(lane_addr + col, &m[i][0]);
                        tc_ld8(lane_addr + col + 8, &m[i][8]);
                        tc_ld8(lane_addr + col + 16, &m[i][16]);
                    }
                    tc_wait_ld();
                    if (gg == GROUPS_PER_WARP - 1) {
                        // the warp's last accumulator values are in registers: hand TMEM back
                        // BEFORE this group's QCP arithmetic, the next tile's MMAs run under it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm->tmem_empty[as]);
                    }
                    if (MODE == 2) {
                        // timing probe (mode 2): consume the accumulators without the QCP
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 24; ++j) acc += m[i][j];
                        if (acc == 1.2345e-30f && dbg) dbg[0] = acc;
                    }
                    if (MODE == 0) {
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            const int c = ct * NC + g * 8 + cc;
                            if (!fvalid || c >= k) continue;
                            float *o = dbg + ((size_t)f * k + c) * 9;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j)
                                    o[3 * i + j] = m[i][3 * cc + j] * kAccumUnscale;
                        }
                    }
#pragma unroll
                    for (int bq = 0; bq < (MODE == 1 ? 2 : 0); ++bq) {   // two batches of four
                        float Mn[4][9], gq[4], scf[4], lam[4], dl[4], halff[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int cc = bq * 4 + p;
                            const float gbf = __shfl_sync(0xffffffffu, my_gb, gg * 8 + cc);
                            const float sqg = __shfl_sync(0xffffffffu, my_sq, gg * 8 + cc);
                            scf[p] = fmaxf(sa * sqg, 1e-30f);
                            const float inv = __frcp_rn(scf[p]);
                            const float inv_acc = inv * kAccumUnscale;   // accumulators: 2^16 M
                            halff[p] = 0.5f * (gaf + gbf);
                            gq[p] = halff[p] * inv;
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j)
                                    Mn[p][3 * i + j] = m[i][3 * cc + j] * inv_acc;
                        }
                        qcp4_f32(Mn, gq, lam, dl);
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int c = ct * NC + g * 8 + bq * 4 + p;
                            if (!fvalid || c >= k) continue;
                            const float half = halff[p];
                            const float sc = scf[p];
                            const float v = 2.0f * __fmaf_rn(-sc, lam[p], half);  // N * msd
                            const float e = __fmaf_rn(kappaf, sc, kQcpSlackF * half);
                            // Newton from above decreases monotonically towards lambda_max, so
                            // an unsettled solve still OVER-estimates lambda: v - e stays a valid
                            // lower bound, only the upper bound needs a settled solve.  A solve
                            // that left the feasible range proves nothing: keep the pair.
                            const bool finite = (lam[p] == lam[p]) && lam[p] <= gq[p] * 1.001f &&
                                                lam[p] >= -gq[p];
                            const bool settled = fabsf(dl[p]) <= 4e-6f * fabsf(gq[p]);
                            const float lo = finite ? v - e : -3.0e38f;
                            const float up = (finite && settled) ? v + e : 3.0e38f;
                            if (lo <= umin) {
                                if (ncand < MAX_CAND) {
                                    cand_list[slot + ncand] = c;
                                    cand_bound[slot + ncand] = lo;
                                }
                                ++ncand;
                            }
                            if (up < umin) umin = up;
                            if (ncand == MAX_CAND + 1) {
                                // list overflowed: compact it against the current bound once
                                int w = 0;
                                for (int s2 = 0; s2 < MAX_CAND; ++s2) {
                                    const float b = cand_bound[slot + s2];
                                    const int cs = cand_list[slot + s2];
                                    if (b <= umin) {
                                        cand_list[slot + w] = cs;
                                        cand_bound[slot + w] = b;
                                        ++w;
                                    }
                                }
                                if (w < MAX_CAND) {
                                    cand_list[slot + w] = c;
                                    cand_bound[slot + w] = lo;
                                    ncand = w + 1;
                                } else {
                                    ncand = MAX_CAND + 2;  // sticky overflow: exact fallback
                                }
                            }
                        }
                    }
                }
            }
            if (MODE == 1 && fvalid) {
                cand_umin[(size_t)f * n_lists + lid] = umin;
                if (ncand > MAX_CAND) {
                    cand_count[(size_t)f * n_lists + lid] = -1;  // overflow -> exact path (host)
                } else {
                    int w = 0;  // final prune against the final bound of this list
                    for (int s2 = 0; s2 < ncand; ++s2) {
                        const float b = cand_bound[slot + s2];
                        const int cs = cand_list[slot + s2];
                        if (b <= umin) {
                            cand_list[slot + w] = cs;
                            cand_bound[slot + w] = b;
                            ++w;
                        }
                    }
                    cand_count[(size_t)f * n_lists + lid] = w;
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

// ---- load balance of the re-score: frames bucketed by their candidate count --------------------
// A warp of the re-score kernel owns four frames and runs as many passes (two centres each) as
// the frame with the most survivors needs.  On the synthetic ensemble neighbouring frames belong
// to different conformers and their counts differ a lot: measured at 262144 x 1008, 3.7 passes
// per frame on average but 8.3 per warp (x2.25).  A counting sort on min(63, passes) -- three
// tiny launches -- puts frames with equal counts next to each other, largest first.
constexpr int RS_BINS = 64;

__global__ void __launch_bounds__(256)
k_rs_key(const int *__restrict__ cand_count, const float *__restrict__ cand_bound,
         const float *__restrict__ cand_umin, long n, int n_lists, int MAX_CAND,
         unsigned char *key, int *hist)
{
    __shared__ int sh[RS_BINS];
    if (threadIdx.x < RS_BINS) sh[threadIdx.x] = 0;
    __syncthreads();
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < n;
         f += (long)gridDim.x * blockDim.x) {
        // exactly what the re-score will keep: entries whose lower bound does not exceed the
        // frame's best upper bound over all lists
        float U = INFINITY;
        bool ovf = false;
        for (int l = 0; l < n_lists; ++l) {
            U = fminf(U, __ldg(cand_umin + (size_t)f * n_lists + l));
            ovf |= __ldg(cand_count + (size_t)f * n_lists + l) < 0;
        }
        int tot = 0;
        if (!ovf) {
            for (int l = 0; l < n_lists; ++l) {
                const int c = __ldg(cand_count + (size_t)f * n_lists + l);
                const float *b = cand_bound + ((size_t)f * n_lists + l) * MAX_CAND;
                for (int e = 0; e < c; ++e) tot += __ldg(b + e) <= U;
            }
        }
        const int k = min(RS_BINS - 1, (tot + 1) >> 1);
        key[f] = (unsigned char)k;
        atomicAdd(&sh[k], 1);
    }
    __syncthreads();
    if (threadIdx.x < RS_BINS && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// one warp: offsets of the bins in DESCENDING key order; resets the cursors
__global__ void k_rs_scan(const int *hist, int *offsets, int *cursor)
{
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = RS_BINS - 1; k >= 0; --k) {
            offsets[k] = run;
            run += hist[k];
            cursor[k] = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
k_rs_scatter(const unsigned char *__restrict__ key, long n, const int *__restrict__ offsets,
             int *cursor, int *order)
{
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < n;
         f += (long)gridDim.x * blockDim.x) {
        const int k = key[f];
        // warp-aggregated cursor bump per bin
        const unsigned peers = __match_any_sync(__activemask(), k);
        const int leader = __ffs(peers) - 1;
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cursor[k], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        order[offsets[k] + base + __popc(peers & ((1u << lane) - 1u))] = (int)f;
    }
}

// ---- exact re-scoring of the surviving (frame, centre) pairs ---------------------------------
// A group of 8 lanes owns one frame: it merges the frame's candidate lists (global upper bound
// U = min over lists; survivors = entries whose lower bound does not exceed U) into one compact
// list in shared memory and then scores two centres per pass (the frame's loads and conversions
// are shared) with the exact float64 arithmetic of the reference path (float x float products
// are exact in double; QCP in double; float32 result).  Ties: lowest centre index, which is what
// strict '<' in centre order gives (cluster/util.py:201).  Frame and centres come from L1/L2.
// Measured alternatives (262144 frames x 1008 centres, 7.8 survivors per frame): this kernel
// 6.1 ms; frame staged in shared memory (2 blocks/SM) 7.5 ms; centres prefetched with
// cp.async.bulk into per-group double buffers (4 warps/SM, frame from L2) 14.1 ms.
constexpr int RS_GROUPS = 16;            // frames per block
constexpr int RS_THREADS = RS_GROUPS * 8;
constexpr long RS_TEAM_MAX_FRAMES = 4096;  // at most this many frames: one block per frame (k_rescore_team)

__global__ void __launch_bounds__(RS_THREADS, 5)
k_rescore(const float *__restrict__ xyz, const double *__restrict__ traces, long n, int n_atoms,
          int A_pad, const float *__restrict__ centers, const double *__restrict__ ctraces,
          const int *__restrict__ cand_count, const int *__restrict__ cand_list,
          const float *__restrict__ cand_bound, const float *__restrict__ cand_umin, int n_seg,
          float *out_dist, int *out_assign, int *frame_flag,
          const int64_t *__restrict__ frame_idx, int scatter, int MAX_CAND,
          const int *__restrict__ f16_overflow, const int *__restrict__ n_dev, int *ovf_count,
          const int *__restrict__ order)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    if (n_dev) n = min(n, (long)__ldg(n_dev));
    const int list_cap = n_seg * MAX_CAND;
    int *cl_all = reinterpret_cast<int *>(rs_smem);                       // [RS_GROUPS][list_cap]
    const int lane = threadIdx.x & 31, g = lane >> 3, l8 = lane & 7;
    const int grp = threadIdx.x >> 3;                                    // group within the block
    const int A4 = A_pad >> 2;
    const size_t stride = 3 * (size_t)A_pad;
    int *cl = cl_all + (size_t)grp * list_cap;
    const long groups = (long)gridDim.x * RS_GROUPS;
    const long n_round = (n + groups - 1) / groups * groups;   // every warp runs the same trips
    for (long fb = (long)blockIdx.x * RS_GROUPS + grp; fb < n_round; fb += groups) {
        // `order` (optional): frames sorted by their number of candidates, so that the four
        // frames of a warp need the same number of passes (the warp runs the maximum)
        const bool valid = fb < n;
        const long f = valid ? (order ? (long)__ldg(order + fb) : fb) : fb;
        // merge the lists: global upper bound, overflow if any list overflowed
        float U = INFINITY;
        int ovf = 0;
        if (valid) {
            for (int sgm = l8; sgm < n_seg; sgm += 8) {
                U = fminf(U, cand_umin[(size_t)f * n_seg + sgm]);
                ovf |= cand_count[(size_t)f * n_seg + sgm] < 0;
            }
        }
#pragma unroll
        for (int msk = 1; msk < 8; msk <<= 1) {
            U = fminf(U, __shfl_xor_sync(0xffffffffu, U, msk));
            ovf |= __shfl_xor_sync(0xffffffffu, ovf, msk);
        }
        // a coordinate outside the FP16 split's range voids the whole screen: exact fallback
        const bool overflow = ovf != 0 || __ldg(f16_overflow) != 0;
        const long src = valid ? (frame_idx ? (long)frame_idx[f] : f) : 0;  // row of xyz / traces
        const float4 *px = reinterpret_cast<const float4 *>(xyz + (size_t)src * stride);
        // compact the survivors of all lists into cl[0 .. cnt_g)
        int cnt_g = 0;
        for (int sgm = 0; sgm < n_seg; ++sgm) {
            const int cnt = (valid && !overflow) ? cand_count[(size_t)f * n_seg + sgm] : 0;
            int cmax = cnt;
            cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 8));
            cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
            const size_t slot = ((size_t)(valid ? f : 0) * n_seg + sgm) * MAX_CAND;
            for (int s0 = 0; s0 < cmax; s0 += 8) {
                const int sidx = s0 + l8;
                const bool act = sidx < cnt && cand_bound[slot + sidx] <= U;
                const int c = act ? cand_list[slot + sidx] : 0;
                const unsigned b = __ballot_sync(0xffffffffu, act);
                const unsigned gb = (b >> (8 * g)) & 0xffu;
                if (act) cl[cnt_g + __popc(gb & ((1u << l8) - 1u))] = c;
                cnt_g += __popc(gb);
            }
        }
        __syncwarp();
        int tmax = cnt_g;
        tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, 8));
        tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, 16));
        float best_d = INFINITY;
        int best_c = 0;
        const double Ga = valid ? traces[src] : 0.0;
        for (int i0 = 0; i0 < tmax; i0 += 2) {
            const bool act0 = i0 < cnt_g, act1 = i0 + 1 < cnt_g;
            const int c0 = act0 ? cl[i0] : 0, c1 = act1 ? cl[i0 + 1] : c0;
            const float4 *p0 = reinterpret_cast<const float4 *>(centers + (size_t)c0 * stride);
            const float4 *p1 = reinterpret_cast<const float4 *>(centers + (size_t)c1 * stride);
            double m0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, m1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (act0) {
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
            }
            group8_reduce(m0);
            group8_reduce(m1);
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
        if (valid && l8 == 0) {
            if (overflow || cnt_g == 0) {
                frame_flag[f] = -1;  // the host sends this frame through the exact kernel
                if (ovf_count) atomicAdd(ovf_count, 1);
            } else {
                frame_flag[f] = cnt_g;
                const long o = scatter ? src : f;
                out_dist[o] = best_d;
                out_assign[o] = best_c;
            }
        }
        __syncwarp();
    }
}


// Re-score of a SMALL frame set (PAM's ambiguous subset: ~10^3 frames whose count lives on the
// device): k_rescore gives every frame to one 8-lane group that walks the frame's survivors two
// at a time, so its duration is the longest list (58 us per proposal at config 3, measured).
// Here one BLOCK owns a frame: warp 0 merges the lists, the 16 groups take the survivors two at
// a time in parallel, and the nearest (lowest index on ties) is reduced through shared memory.
// Per pair the arithmetic is k_rescore's (same lanes, same atom order, same QCP), and the
// minimum over (distance, index) does not depend on the order of evaluation, so the results are
// identical.
__global__ void __launch_bounds__(RS_THREADS, 5)
k_rescore_team(const float *__restrict__ xyz, const double *__restrict__ traces, long n,
               int n_atoms, int A_pad, const float *__restrict__ centers,
               const double *__restrict__ ctraces, const int *__restrict__ cand_count,
               const int *__restrict__ cand_list, const float *__restrict__ cand_bound,
               const float *__restrict__ cand_umin, int n_seg, float *out_dist, int *out_assign,
               int *frame_flag, const int64_t *__restrict__ frame_idx, int scatter, int MAX_CAND,
               const int *__restrict__ f16_overflow, const int *__restrict__ n_dev,
               int *ovf_count)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    __shared__ int sh_cnt, sh_ovf;
    __shared__ float red_d[RS_GROUPS];
    __shared__ int red_c[RS_GROUPS];
    if (n_dev) n = min(n, (long)__ldg(n_dev));
    int *cl = reinterpret_cast<int *>(rs_smem);                           // [n_seg * MAX_CAND]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane & 7;
    const int grp = threadIdx.x >> 3;
    const int A4 = A_pad >> 2;
    const size_t stride = 3 * (size_t)A_pad;
    for (long f = blockIdx.x; f < n; f += gridDim.x) {
        // merge the frame's lists with every load independent of the others (a serial walk of
        // 63 lists costs 63 dependent L2 round trips): global upper bound and overflow flag
        // first, then every (list, slot) is tested by its own thread and the survivors are
        // appended through a shared counter -- their order does not matter for the minimum
        float U = INFINITY;
        int ovf = 0;
        for (int sgm = threadIdx.x; sgm < n_seg; sgm += RS_THREADS) {
            U = fminf(U, cand_umin[(size_t)f * n_seg + sgm]);
            ovf |= cand_count[(size_t)f * n_seg + sgm] < 0;
        }
#pragma unroll
        for (int msk = 1; msk < 32; msk <<= 1) {
            U = fminf(U, __shfl_xor_sync(0xffffffffu, U, msk));
            ovf |= __shfl_xor_sync(0xffffffffu, ovf, msk);
        }
        if (lane == 0) {
            red_d[warp] = U;
            red_c[warp] = ovf;
        }
        if (threadIdx.x == 0) sh_cnt = 0;
        __syncthreads();
        U = fminf(fminf(red_d[0], red_d[1]), fminf(red_d[2], red_d[3]));
        ovf = red_c[0] | red_c[1] | red_c[2] | red_c[3];
        const bool overflow_all = ovf != 0 || __ldg(f16_overflow) != 0;
        if (!overflow_all) {
            const int slots = n_seg * MAX_CAND;
            for (int sl = threadIdx.x; sl < slots; sl += RS_THREADS) {
                const int sgm = sl / MAX_CAND, sidx = sl - sgm * MAX_CAND;
                const int cnt = cand_count[(size_t)f * n_seg + sgm];
                const size_t slot = ((size_t)f * n_seg + sgm) * MAX_CAND + sidx;
                if (sidx < cnt && cand_bound[slot] <= U) cl[atomicAdd(&sh_cnt, 1)] = cand_list[slot];
            }
        }
        if (threadIdx.x == 0) sh_ovf = overflow_all ? 1 : 0;
        __syncthreads();
        const int cnt_g = sh_cnt;
        const bool overflow = sh_ovf != 0;
        const long src = frame_idx ? (long)frame_idx[f] : f;
        const float4 *px = reinterpret_cast<const float4 *>(xyz + (size_t)src * stride);
        const double Ga = traces[src];
        float best_d = INFINITY;
        int best_c = 0x7fffffff;
        // every group runs every trip (the shuffles below are warp-wide); idle ones skip the walk
        for (int base = 0; base < cnt_g; base += 2 * RS_GROUPS) {
            const int i0 = base + 2 * grp;
            const bool act0 = i0 < cnt_g, act1 = i0 + 1 < cnt_g;
            const int c0 = act0 ? cl[i0] : 0, c1 = act1 ? cl[i0 + 1] : c0;
            const float4 *p0 = reinterpret_cast<const float4 *>(centers + (size_t)c0 * stride);
            const float4 *p1 = reinterpret_cast<const float4 *>(centers + (size_t)c1 * stride);
            double m0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, m1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (act0) {
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
            }
            group8_reduce(m0);
            group8_reduce(m1);
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
        if (l8 == 0) {
            red_d[grp] = best_d;
            red_c[grp] = best_c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (overflow || cnt_g == 0) {
                frame_flag[f] = -1;  // the host sends this frame through the exact kernel
                if (ovf_count) atomicAdd(ovf_count, 1);
            } else {
                float bd = red_d[0];
                int bc = red_c[0];
                for (int g2 = 1; g2 < RS_GROUPS; ++g2) {
                    const float d = red_d[g2];
                    const int c = red_c[g2];
                    if (d < bd || (d == bd && c < bc)) {
                        bd = d;
                        bc = c;
                    }
                }
                frame_flag[f] = cnt_g;
                const long o = scatter ? src : f;
                out_dist[o] = bd;
                out_assign[o] = bc;
            }
        }
        __syncthreads();
    }
}

}  // namespace tc
}  // namespace eb

using namespace eb;

// Centre segments per frame tile: the default keeps consecutive CTAs on one frame tile (its A
// operand is then served by L2); a small frame subset (PAM's ambiguous frames) is cut into more
// segments so that every SM gets a work item.
static int tc_def_seg()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("EB_TC_SEG");   // developer A/B switch
        v = e ? atoi(e) : tc::DEF_SEG;
        if (v < 1 || v > tc::MAX_SEG) v = tc::DEF_SEG;
    }
    return v;
}
static int tc_max_cand(int n_lists)
{
    const int c = tc::CAND_BUDGET / n_lists;
    return c < tc::MIN_CAND ? tc::MIN_CAND : c;
}
static int tc_pick_nseg(long n_ft, int n_ct)
{
    const int def_seg = tc_def_seg();
    int n_seg = n_ct < def_seg ? n_ct : def_seg;
    const long want = 2L * sm_count();
    if (n_ft * n_seg < want) {
        long s2 = (want + n_ft - 1) / n_ft;
        if (s2 > tc::MAX_SEG) s2 = tc::MAX_SEG;
        if (s2 > n_ct) s2 = n_ct;
        if (s2 > n_seg) n_seg = (int)s2;
    }
    if (n_seg < 1) n_seg = 1;
    // drop segments that would be empty: ceil(n_ct / ceil(n_ct / n_seg))
    const int per = (n_ct + n_seg - 1) / n_seg;
    return (n_ct + per - 1) / per;
}

extern "C" {

// bytes of the packed operand images per tile row (one frame: 3 coordinate rows; one centre:
// 3 rows too): {h1, h2} x 2 bytes x 3 x A_img
static size_t tc_img_row_bytes(int n_atoms)
{
    const int A_img = (rmsd_apad(n_atoms) + tc::BK - 1) / tc::BK * tc::BK;
    return 2 * sizeof(__half) * 3 * (size_t)A_img;
}

// bytes of the packed centre image: per centre tile and k-block {h1, h2} x (BN rows x 64 B)
static size_t tc_b_img_bytes(int n_atoms, int32_t k)
{
    const int KB = (rmsd_apad(n_atoms) + tc::BK - 1) / tc::BK;
    const size_t n_ct = (size_t)((k + tc::NC - 1) / tc::NC) + 1;
    return n_ct * KB * 2 * (size_t)tc::B_TILE;
}

size_t eb_tc_scratch_bytes(int64_t n, int n_atoms, int32_t k)
{
    const size_t row = tc_img_row_bytes(n_atoms);
    const long n_ft = (n + tc::BM - 1) / tc::BM;
    const int n_ct = (k + tc::NC - 1) / tc::NC;
    const size_t n_lists = (size_t)tc_pick_nseg(n_ft < 1 ? 1 : n_ft, n_ct < 1 ? 1 : n_ct) *
                           tc::EPI_PARTS;
    // split copies of frames and centres + candidate lists + flags
    return row * (size_t)(n + tc::BM) + tc_b_img_bytes(n_atoms, k) +
           (size_t)n * n_lists *
               (sizeof(int) + sizeof(float) +
                tc_max_cand((int)n_lists) * (sizeof(int) + sizeof(float))) +
           (size_t)n * (sizeof(int) + 1) + 4096;   // + re-score order / keys, bins, flags
}

// mode 0 (debug): dbg receives the approximate inner-product matrices, n x k x 9 floats.
// mode 1: screen + exact re-score.  out_dist/out_assign get the exact result for every frame whose
//         candidate list did not overflow; the others have cand_count[i] == -1 and the caller
//         must send them through eb_rmsd_assign.
// frame_idx (optional, int64[n]): the pass covers frames frame_idx[0..n) of xyz_soa / traces;
//         cand_count is indexed by position i; results go to position i, or to frame_idx[i]
//         when `scatter` is set.
int eb_rmsd_assign_tc(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                      const float *centers_soa, const double *center_traces, int32_t k,
                      double kappa, const int64_t *frame_idx, int scatter, float *out_dist,
                      int32_t *out_assign, int32_t *cand_count, void *scratch, float *dbg,
                      int mode, void *stream)
{
    return eb_rmsd_assign_tc_dev(xyz_soa, traces, n, n_atoms, centers_soa, center_traces, k,
                                 kappa, frame_idx, scatter, out_dist, out_assign, cand_count,
                                 scratch, dbg, mode, nullptr, nullptr, stream);
}

// The same pass with the number of frames known only on the DEVICE: n is the host's upper bound
// (grids, scratch and list layout are sized by it), *n_dev (<= n) the real count; positions
// beyond it are neither read nor written.  overflow_count (optional) is incremented once per
// frame whose candidate lists overflowed (cand_count = -1), so that a caller can defer the
// exact fallback to its next read-back instead of synchronising here.
int eb_rmsd_assign_tc_dev(const float *xyz_soa, const double *traces, int64_t n, int n_atoms,
                          const float *centers_soa, const double *center_traces, int32_t k,
                          double kappa, const int64_t *frame_idx, int scatter, float *out_dist,
                          int32_t *out_assign, int32_t *cand_count, void *scratch, float *dbg,
                          int mode, const int32_t *n_dev, int32_t *overflow_count, void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0 && k >= 0, "rmsd_assign_tc: bad shape");
    if (n == 0 || k == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && centers_soa && center_traces && scratch,
                 "rmsd_assign_tc: null pointer");
    EB_CHECK_ARG(n < (1L << 31) / 3, "rmsd_assign_tc: too many frames for one pass");
    cudaStream_t s = (cudaStream_t)stream;
    const int A_pad = rmsd_apad(n_atoms);
    const int A_img = (A_pad + tc::BK - 1) / tc::BK * tc::BK;   // image rows padded to 32 atoms
    const size_t row = tc_img_row_bytes(n_atoms);  // image bytes per frame / per centre
    // packed operand images (tile-padded): frames then centres
    const size_t a_rows = (size_t)((n + tc::BM - 1) / tc::BM) * tc::BM;   // frames, padded
    unsigned char *a_img = (unsigned char *)scratch;
    unsigned char *b_img = a_img + row * a_rows;
    unsigned char *lists = b_img + tc_b_img_bytes(n_atoms, k);
    const int n_ct_total = (k + tc::NC - 1) / tc::NC;
    const long n_ft = (n + tc::BM - 1) / tc::BM;
    const int n_seg = tc_pick_nseg(n_ft, n_ct_total);
    const size_t n_lists = (size_t)n_seg * tc::EPI_PARTS;
    int *cand_list = (int *)lists;
    const int max_cand = tc_max_cand((int)n_lists);
    float *cand_bound = (float *)(cand_list + (size_t)n * n_lists * max_cand);
    float *cand_umin = cand_bound + (size_t)n * n_lists * max_cand;
    int *seg_count = (int *)(cand_umin + (size_t)n * n_lists);
    // set by the pack kernels when a coordinate does not fit the FP16 split (|x| > 253 nm)
    int *f16_overflow = (int *)(((uintptr_t)(seg_count + (size_t)n * n_lists) + 15) & ~(uintptr_t)15);
    EB_CUDA(cudaMemsetAsync(f16_overflow, 0, sizeof(int), s));

    {
        long total = n_ft * tc::BM * 3 * (long)(A_img / 8);
        long blocks = (total + 255) / 256;
        if (blocks > 32L * sm_count()) blocks = 32L * sm_count();
        tc::k_pack_f16x2<<<(int)blocks, 256, 0, s>>>(xyz_soa, 3L * n, A_pad, tc::BM, tc::BM, 3,
                                                     n_ft, frame_idx, a_img, f16_overflow,
                                                     n_dev);
        EB_LAUNCH_CHECK();
        const long b_tiles = n_ct_total;
        total = b_tiles * tc::BN * (long)(A_img / 8);
        blocks = (total + 255) / 256;
        if (blocks > 32L * sm_count()) blocks = 32L * sm_count();
        tc::k_pack_f16x2<<<(int)blocks, 256, 0, s>>>(centers_soa, 3L * k, A_pad, tc::BN,
                                                     tc::BN_USED, 1, b_tiles, nullptr, b_img,
                                                     f16_overflow, nullptr);
        EB_LAUNCH_CHECK();
    }
    const size_t smem = (size_t)tc::STAGES * tc::STAGE_BYTES + sizeof(tc::Smem) + 1024;
    const long n_items = n_ft * n_seg;
    const int grid = (int)(n_items < sm_count() ? n_items : sm_count());
    if (mode == 0) {
        EB_CHECK_ARG(dbg, "rmsd_assign_tc: debug buffer missing");
        EB_CUDA(cudaFuncSetAttribute(tc::k_tc_screen<0>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::k_tc_screen<0><<<grid, tc::THREADS, smem, s>>>(a_img, b_img, traces,
                                                           center_traces, n, k, n_atoms, A_pad,
                                                           kappa, dbg, nullptr, nullptr, nullptr,
                                                           nullptr, n_seg, frame_idx, 8, n_dev);
        EB_LAUNCH_CHECK();
        return EB_OK;
    }
    if (mode == 2) {   // timing probe: the screen without its QCP epilogue
        EB_CUDA(cudaFuncSetAttribute(tc::k_tc_screen<2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::k_tc_screen<2><<<grid, tc::THREADS, smem, s>>>(a_img, b_img, traces,
                                                           center_traces, n, k, n_atoms, A_pad,
                                                           kappa, dbg, nullptr, nullptr, nullptr,
                                                           nullptr, n_seg, frame_idx, 8, n_dev);
        EB_LAUNCH_CHECK();
        return EB_OK;
    }
    EB_CHECK_ARG(out_dist && out_assign && cand_count, "rmsd_assign_tc: null output");
    EB_CUDA(cudaFuncSetAttribute(tc::k_tc_screen<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    tc::k_tc_screen<1><<<grid, tc::THREADS, smem, s>>>(a_img, b_img, traces,
                                                       center_traces, n, k, n_atoms, A_pad, kappa,
                                                       nullptr, seg_count, cand_list, cand_bound,
                                                       cand_umin, n_seg, frame_idx, max_cand,
                                                       n_dev);
    EB_LAUNCH_CHECK();
    // frames bucketed by candidate count (dense passes only: a few microseconds that a small
    // PAM subset would not earn back, and its size is only known on the device)
    int *order = nullptr;
    if (!n_dev && n >= 16384 && n < (1L << 31)) {
        int *bins = (int *)(((uintptr_t)(f16_overflow + 4) + 15) & ~(uintptr_t)15);   // 3 x 64
        order = bins + 3 * tc::RS_BINS;
        unsigned char *key = (unsigned char *)(order + n);
        EB_CUDA(cudaMemsetAsync(bins, 0, sizeof(int) * tc::RS_BINS, s));
        long kb = (n + 255) / 256;
        if (kb > 8L * sm_count()) kb = 8L * sm_count();
        tc::k_rs_key<<<(int)kb, 256, 0, s>>>(seg_count, cand_bound, cand_umin, n, (int)n_lists,
                                             max_cand, key, bins);
        tc::k_rs_scan<<<1, 32, 0, s>>>(bins, bins + tc::RS_BINS, bins + 2 * tc::RS_BINS);
        tc::k_rs_scatter<<<(int)kb, 256, 0, s>>>(key, n, bins + tc::RS_BINS,
                                                 bins + 2 * tc::RS_BINS, order);
        EB_LAUNCH_CHECK();
    }
    if (n <= tc::RS_TEAM_MAX_FRAMES) {
        // small frame sets (PAM's ambiguous subset): one block per frame
        const size_t rs_smem = align16(sizeof(int) * n_lists * max_cand);
        long blocks = n < 5L * sm_count() ? n : 5L * sm_count();
        if (blocks < 1) blocks = 1;
        EB_CHECK_ARG(rs_smem <= 200 * 1024, "rmsd_assign_tc: candidate lists too large");
        EB_CUDA(cudaFuncSetAttribute(tc::k_rescore_team,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(rs_smem > 48 * 1024 ? rs_smem : 48 * 1024)));
        tc::k_rescore_team<<<(int)blocks, tc::RS_THREADS, rs_smem, s>>>(
            xyz_soa, traces, n, n_atoms, A_pad, centers_soa, center_traces, seg_count, cand_list,
            cand_bound, cand_umin, (int)n_lists, out_dist, out_assign, cand_count, frame_idx,
            scatter, max_cand, f16_overflow, n_dev, overflow_count);
        EB_LAUNCH_CHECK();
    } else {
        const size_t rs_smem = align16(sizeof(int) * tc::RS_GROUPS * n_lists * max_cand);
        long blocks = (n + tc::RS_GROUPS - 1) / tc::RS_GROUPS;
        if (blocks > 10L * sm_count()) blocks = 10L * sm_count();
        EB_CHECK_ARG(rs_smem <= 227 * 1024, "rmsd_assign_tc: candidate lists too large");
        EB_CUDA(cudaFuncSetAttribute(tc::k_rescore, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rs_smem));
        tc::k_rescore<<<(int)blocks, tc::RS_THREADS, rs_smem, s>>>(
            xyz_soa, traces, n, n_atoms, A_pad, centers_soa, center_traces, seg_count, cand_list,
            cand_bound, cand_umin, (int)n_lists, out_dist, out_assign, cand_count, frame_idx,
            scatter, max_cand, f16_overflow, n_dev, overflow_count, order);
        EB_LAUNCH_CHECK();
    }
    return EB_OK;
}

// Exact scoring of explicit per-position centre lists (the re-score kernel on its own): position
// p covers frame frame_idx[p] (or p) and the list_len-strided list cand_list[p * list_len ..
// + cand_count[p]); out_dist/out_assign[p] receive the nearest listed centre (exact float64
// path, lowest index on ties).  Used by the host's audit of the tensor-core screen: a sample of
// frames is scored against ALL centres, cut into lists of list_len, and compared with what the
// screen + re-score produced.  bound_lo / bound_up: list_len * n_pos floats of -inf and n_pos
// floats of +inf (every listed centre is scored); zero_flag: one int holding 0.
int eb_rmsd_score_lists(const float *xyz_soa, const double *traces, int64_t n_pos, int n_atoms,
                        const float *centers_soa, const double *center_traces,
                        const int32_t *cand_count, const int32_t *cand_list, int list_len,
                        const float *bound_lo, const float *bound_up, const int32_t *zero_flag,
                        const int64_t *frame_idx, float *out_dist, int32_t *out_assign,
                        int32_t *frame_flag, void *stream)
{
    EB_CHECK_ARG(n_pos >= 0 && n_atoms > 0 && list_len > 0, "rmsd_score_lists: bad shape");
    if (n_pos == 0) return EB_OK;
    EB_CHECK_ARG(xyz_soa && traces && centers_soa && center_traces && cand_count && cand_list &&
                     bound_lo && bound_up && zero_flag && out_dist && out_assign && frame_flag,
                 "rmsd_score_lists: null pointer");
    const size_t rs_smem = align16(sizeof(int) * tc::RS_GROUPS * (size_t)list_len);
    EB_CHECK_ARG(rs_smem <= 227 * 1024, "rmsd_score_lists: lists too long");
    long blocks = (n_pos + tc::RS_GROUPS - 1) / tc::RS_GROUPS;
    if (blocks > 10L * sm_count()) blocks = 10L * sm_count();
    EB_CUDA(cudaFuncSetAttribute(tc::k_rescore, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(rs_smem > 48 * 1024 ? rs_smem : 48 * 1024)));
    tc::k_rescore<<<(int)blocks, tc::RS_THREADS, rs_smem, (cudaStream_t)stream>>>(
        xyz_soa, traces, n_pos, n_atoms, rmsd_apad(n_atoms), centers_soa, center_traces,
        cand_count, cand_list, bound_lo, bound_up, 1, out_dist, out_assign, frame_flag, frame_idx,
        0, list_len, zero_flag, nullptr, nullptr, nullptr);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

}  // extern "C"
