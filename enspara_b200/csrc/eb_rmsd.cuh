// Device-side RMSD building blocks shared by the k-centers step (K1) and the many-centres
// assignment (K3): the QCP solve and the 8-lane inner-product accumulation.
#pragma once
#include "eb_common.cuh"

namespace eb {

// ------------------------------------------------------------------------------------------
// Theobald QCP: msd from the 3x3 inner-product matrix M (M[3i+j] = sum_a x_i y_j), the two
// traces and the atom count.  Same formula sequence as oracle/enspara_oracle.c:qcp_msd
// (restating mdtraj's msdFromMandG), all in float64.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double det2(double a, double b, double c, double d)
{
    return a * d - b * c;
}

__device__ __forceinline__ double qcp_msd(const double M[9], double Ga, double Gb, int n_atoms)
{
    const double Sxx = M[0], Sxy = M[1], Sxz = M[2];
    const double Syx = M[3], Syy = M[4], Syz = M[5];
    const double Szx = M[6], Szy = M[7], Szz = M[8];

    const double k00 = Sxx + Syy + Szz, k01 = Syz - Szy, k02 = Szx - Sxz, k03 = Sxy - Syx;
    const double k11 = Sxx - Syy - Szz, k12 = Sxy + Syx, k13 = Szx + Sxz;
    const double k22 = -Sxx + Syy - Szz, k23 = Syz + Szy;
    const double k33 = -Sxx - Syy + Szz;

    const double c2 = -2.0 * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy +
                              Syz * Syz + Szx * Szx + Szy * Szy + Szz * Szz);
    const double detM = Sxx * (Syy * Szz - Syz * Szy) - Sxy * (Syx * Szz - Syz * Szx) +
                        Sxz * (Syx * Szy - Syy * Szx);
    const double c1 = -8.0 * detM;

    const double r01_01 = det2(k00, k01, k01, k11), r01_02 = det2(k00, k02, k01, k12);
    const double r01_03 = det2(k00, k03, k01, k13), r01_12 = det2(k01, k02, k11, k12);
    const double r01_13 = det2(k01, k03, k11, k13), r01_23 = det2(k02, k03, k12, k13);
    const double r23_01 = det2(k02, k12, k03, k13), r23_02 = det2(k02, k22, k03, k23);
    const double r23_03 = det2(k02, k23, k03, k33), r23_12 = det2(k12, k22, k13, k23);
    const double r23_13 = det2(k12, k23, k13, k33), r23_23 = det2(k22, k23, k23, k33);
    const double c0 = r01_01 * r23_23 - r01_02 * r23_13 + r01_03 * r23_12 + r01_12 * r23_03 -
                      r01_13 * r23_02 + r01_23 * r23_01;

    double lambda = 0.5 * (Ga + Gb);
#pragma unroll 1
    for (int it = 0; it < 50; ++it) {
        const double l2 = lambda * lambda;
        const double b = (l2 + c2) * lambda;
        const double a = b + c1;
        const double denom = 2.0 * l2 * lambda + b + a;
        if (denom == 0.0) break;
        const double delta = (a * lambda + c0) / denom;
        lambda -= delta;
        if (fabs(delta) < fabs(1e-11 * lambda)) break;
    }
    double msd = (Ga + Gb - 2.0 * lambda) / (double)n_atoms;
    if (!(msd > 0.0)) msd = 0.0;
    return msd;
}

__device__ __forceinline__ float rmsd_from_msd(double msd) { return sqrtf((float)msd); }

// ------------------------------------------------------------------------------------------
// centre staged in shared memory as float64, split so that the 8 lanes of a frame group read
// 8 consecutive 16-byte entries (conflict free): lo[c][j] = atoms (4j, 4j+1), hi[c][j] =
// atoms (4j+2, 4j+3) of coordinate row c.  Size: 3 * A_pad doubles.
// ------------------------------------------------------------------------------------------
struct CenterSmem {
    double2 *lo[3];
    double2 *hi[3];
};

__device__ __forceinline__ CenterSmem center_smem_carve(double *base, int A4)
{
    CenterSmem c;
    double2 *p = reinterpret_cast<double2 *>(base);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        c.lo[r] = p + (2 * r) * A4;
        c.hi[r] = p + (2 * r + 1) * A4;
    }
    return c;
}

// all threads of the block cooperatively convert one SoA centre (3*A_pad floats) to the layout
__device__ __forceinline__ void center_smem_fill(const CenterSmem &c, const float *center_soa,
                                                 int A_pad)
{
    const int A4 = A_pad >> 2;
    const float4 *src = reinterpret_cast<const float4 *>(center_soa);
    for (int t = threadIdx.x; t < 3 * A4; t += blockDim.x) {
        const int r = t / A4, j = t - r * A4;
        const float4 v = __ldcg(src + t);
        c.lo[r][j] = make_double2((double)v.x, (double)v.y);
        c.hi[r][j] = make_double2((double)v.z, (double)v.w);
    }
}

// 9 running sums for one atom, exact products (float x float fits a double exactly)
__device__ __forceinline__ void acc_atom(double m[9], float fx, float fy, float fz, double cx,
                                         double cy, double cz)
{
    const double x = (double)fx, y = (double)fy, z = (double)fz;
    m[0] = fma(x, cx, m[0]); m[1] = fma(x, cy, m[1]); m[2] = fma(x, cz, m[2]);
    m[3] = fma(y, cx, m[3]); m[4] = fma(y, cy, m[4]); m[5] = fma(y, cz, m[5]);
    m[6] = fma(z, cx, m[6]); m[7] = fma(z, cy, m[7]); m[8] = fma(z, cz, m[8]);
}

// One frame against the staged centre, computed by the 8 lanes l8 = 0..7 of a group.
// EXACT: every product goes through the float64 pipe.
// !EXACT: 4-atom float32 FMA blocks per lane, block sums added in float64 (fast mode).
template <bool EXACT, int LDVAR = 0>
__device__ __forceinline__ void frame_inner_products(double m[9], const float *frame, int A4,
                                                     int l8, const CenterSmem &c)
{
    const float4 *px = reinterpret_cast<const float4 *>(frame);
    const float4 *py = px + A4;
    const float4 *pz = py + A4;
#pragma unroll 2
    for (int j = l8; j < A4; j += 8) {
        const float4 x = LDVAR == 1 ? ldg_stream_256(px + j) : ldg_stream(px + j);
        const float4 y = LDVAR == 1 ? ldg_stream_256(py + j) : ldg_stream(py + j);
        const float4 z = LDVAR == 1 ? ldg_stream_256(pz + j) : ldg_stream(pz + j);
        const double2 cxl = c.lo[0][j], cxh = c.hi[0][j];
        const double2 cyl = c.lo[1][j], cyh = c.hi[1][j];
        const double2 czl = c.lo[2][j], czh = c.hi[2][j];
        if (EXACT) {
            acc_atom(m, x.x, y.x, z.x, cxl.x, cyl.x, czl.x);
            acc_atom(m, x.y, y.y, z.y, cxl.y, cyl.y, czl.y);
            acc_atom(m, x.z, y.z, z.z, cxh.x, cyh.x, czh.x);
            acc_atom(m, x.w, y.w, z.w, cxh.y, cyh.y, czh.y);
        } else {
            const float cx[4] = {(float)cxl.x, (float)cxl.y, (float)cxh.x, (float)cxh.y};
            const float cy[4] = {(float)cyl.x, (float)cyl.y, (float)cyh.x, (float)cyh.y};
            const float cz[4] = {(float)czl.x, (float)czl.y, (float)czh.x, (float)czh.y};
            const float fx[4] = {x.x, x.y, x.z, x.w};
            const float fy[4] = {y.x, y.y, y.z, y.w};
            const float fz[4] = {z.x, z.y, z.z, z.w};
            float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                s[0] = fmaf(fx[a], cx[a], s[0]); s[1] = fmaf(fx[a], cy[a], s[1]);
                s[2] = fmaf(fx[a], cz[a], s[2]); s[3] = fmaf(fy[a], cx[a], s[3]);
                s[4] = fmaf(fy[a], cy[a], s[4]); s[5] = fmaf(fy[a], cz[a], s[5]);
                s[6] = fmaf(fz[a], cx[a], s[6]); s[7] = fmaf(fz[a], cy[a], s[7]);
                s[8] = fmaf(fz[a], cz[a], s[8]);
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) m[e] += (double)s[e];
        }
    }
}

// butterfly over the 8 lanes of a group; afterwards every lane of the group holds the total
__device__ __forceinline__ void group8_reduce(double m[9])
{
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        m[e] += shfl_xor_d(m[e], 1);
        m[e] += shfl_xor_d(m[e], 2);
        m[e] += shfl_xor_d(m[e], 4);
    }
}

}  // namespace eb
