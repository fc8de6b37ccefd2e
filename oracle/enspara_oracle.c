/*
 * enspara_oracle.c -- CPU ORACLE for the clustering hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (enspara_b200/) never does.
 *
 * What is restated here, and from where:
 *
 *  (1) mdtraj.rmsd -- THIRD-PARTY, NOT under /root/reference and NOT installed in this image
 *      (dependency `mdtraj>=1.7`, /root/reference/pyproject.toml:41; no lock file).  Reached
 *      from the reference at enspara/cluster/util.py:290-291 ('rmsd' -> md.rmsd),
 *      enspara/apps/cluster.py:210 and enspara/cluster/util.py:625-629 (precentered=True).
 *      The published algorithm (Theobald 2005 QCP; Liu, Agrafiotis & Theobald 2010; as wired
 *      up in mdtraj's _rmsd.pyx / theobald_rmsd / center code) is restated:
 *        - per-frame centring in place on float32 coordinates, trace G = sum |x|^2,
 *        - 3x3 inner-product matrix M = sum_a x_a y_a^T,
 *        - quartic characteristic polynomial of the 4x4 key matrix, Newton-Raphson from
 *          (Ga+Gb)/2 in double, msd = max(0,(Ga+Gb-2*lambda)/N), result sqrtf((float)msd).
 *      Two arithmetic variants:
 *        orc_rmsd_f64  : M and G accumulated in double ("truth"; this is what the CUDA path
 *                        is held to, bit-for-bit after the final float32 rounding in the
 *                        overwhelming majority of frames and to 1e-5 relative always)
 *        orc_rmsd_f32  : mdtraj-like arithmetic -- four float32 SSE-lane accumulators per
 *                        matrix entry (atoms a = l mod 4 go to lane l), mul and add rounded
 *                        separately, float32 traces -- used to QUANTIFY mdtraj's own float32
 *                        noise and as the timed CPU baseline ("restated mdtraj CPU path").
 *      PINNING: the restatement is checked against the reference's golden statistics on
 *      frame0.xtc (enspara/test/test_cluster.py:200-238) in tests/test_oracle_golden.py.
 *
 *  (2) enspara.geometry.libdist euclidean / manhattan
 *      (/root/reference/enspara/geometry/libdist.pyx:100-145): one point vs many rows,
 *      typed difference, float64 accumulation IN j ORDER, sqrt in double.  Verified here
 *      against the reference's own Cython build (oracle/_ref) by tests/test_oracle_ref.py.
 *
 *  (3) the O(n) bookkeeping of one k-centers step (kcenters.py:282,304-306) for the timed
 *      CPU baseline: strict-< min update and first-occurrence argmax.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* QCP: largest eigenvalue of the key matrix via Newton on the quartic.  M is row-major with  */
/* M[3*i+j] = sum_a x_i(a) * y_j(a), x = frame ("target"), y = centre ("reference").          */
/* SURVEY.md App. B step 4.                                                                   */
/* ------------------------------------------------------------------------------------------ */
static double det2(double a, double b, double c, double d) { return a * d - b * c; }

static double qcp_msd(const double M[9], double Ga, double Gb, int n_atoms)
{
    const double Sxx = M[0], Sxy = M[1], Sxz = M[2];
    const double Syx = M[3], Syy = M[4], Syz = M[5];
    const double Szx = M[6], Szy = M[7], Szz = M[8];

    /* symmetric traceless 4x4 key matrix K */
    const double k00 = Sxx + Syy + Szz, k01 = Syz - Szy, k02 = Szx - Sxz, k03 = Sxy - Syx;
    const double k11 = Sxx - Syy - Szz, k12 = Sxy + Syx, k13 = Szx + Sxz;
    const double k22 = -Sxx + Syy - Szz, k23 = Syz + Szy;
    const double k33 = -Sxx - Syy + Szz;

    /* P(l) = l^4 + c2 l^2 + c1 l + c0 */
    const double c2 = -2.0 * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy +
                              Syz * Syz + Szx * Szx + Szy * Szy + Szz * Szz);
    const double detM = Sxx * (Syy * Szz - Syz * Szy) - Sxy * (Syx * Szz - Syz * Szx) +
                        Sxz * (Syx * Szy - Syy * Szx);
    const double c1 = -8.0 * detM;

    /* c0 = det K by Laplace expansion over rows (0,1) x rows (2,3) */
    const double r01_01 = det2(k00, k01, k01, k11), r01_02 = det2(k00, k02, k01, k12);
    const double r01_03 = det2(k00, k03, k01, k13), r01_12 = det2(k01, k02, k11, k12);
    const double r01_13 = det2(k01, k03, k11, k13), r01_23 = det2(k02, k03, k12, k13);
    const double r23_01 = det2(k02, k12, k03, k13), r23_02 = det2(k02, k22, k03, k23);
    const double r23_03 = det2(k02, k23, k03, k33), r23_12 = det2(k12, k22, k13, k23);
    const double r23_13 = det2(k12, k23, k13, k33), r23_23 = det2(k22, k23, k23, k33);
    const double c0 = r01_01 * r23_23 - r01_02 * r23_13 + r01_03 * r23_12 + r01_12 * r23_03 -
                      r01_13 * r23_02 + r01_23 * r23_01;

    double lambda = 0.5 * (Ga + Gb);
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

/* exported for the unit test that checks the quartic against numpy's eigvalsh */
double orc_qcp_msd(const double *M, double Ga, double Gb, int n_atoms)
{
    return qcp_msd(M, Ga, Gb, n_atoms);
}

/* ------------------------------------------------------------------------------------------ */
/* Centring + trace.  xyz is (n, A, 3) float32 C-order, modified in */
/* place: centroid accumulated in double, subtracted, result rounded to float32; the trace is */
/* taken from the ROUNDED coordinates so Ga is consistent with what QCP later reads.          */
/* ------------------------------------------------------------------------------------------ */
void orc_center_and_trace(float *xyz, long n, int A, double *traces64, float *traces32)
{
#pragma omp parallel for schedule(static)
    for (long f = 0; f < n; ++f) {
        float *p = xyz + (size_t)f * A * 3;
        double sx = 0, sy = 0, sz = 0;
        for (int a = 0; a < A; ++a) {
            sx += p[3 * a];
            sy += p[3 * a + 1];
            sz += p[3 * a + 2];
        }
        const double mx = sx / A, my = sy / A, mz = sz / A;
        double g = 0;
        for (int a = 0; a < A; ++a) {
            const float x = (float)((double)p[3 * a] - mx);
            const float y = (float)((double)p[3 * a + 1] - my);
            const float z = (float)((double)p[3 * a + 2] - mz);
            p[3 * a] = x;
            p[3 * a + 1] = y;
            p[3 * a + 2] = z;
            g += (double)x * x + (double)y * y + (double)z * z;
        }
        if (traces64) traces64[f] = g;
        if (traces32) traces32[f] = (float)g;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* "truth": every frame of a PRE-CENTRED (n,A,3) block against one pre-centred (A,3) centre.  */
/* float32 x float32 products are exact in double; only the 9 running sums round.             */
/* ------------------------------------------------------------------------------------------ */
void orc_rmsd_f64(const float *xyz, const double *traces, long n, int A, const float *ref,
                  double ref_trace, float *out)
{
#pragma omp parallel for schedule(static)
    for (long f = 0; f < n; ++f) {
        const float *p = xyz + (size_t)f * A * 3;
        double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < A; ++a) {
            const double x0 = p[3 * a], x1 = p[3 * a + 1], x2 = p[3 * a + 2];
            const double y0 = ref[3 * a], y1 = ref[3 * a + 1], y2 = ref[3 * a + 2];
            M[0] += x0 * y0; M[1] += x0 * y1; M[2] += x0 * y2;
            M[3] += x1 * y0; M[4] += x1 * y1; M[5] += x1 * y2;
            M[6] += x2 * y0; M[7] += x2 * y1; M[8] += x2 * y2;
        }
        const double msd = qcp_msd(M, traces[f], ref_trace, A);
        out[f] = sqrtf((float)msd);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* mdtraj-like float32 arithmetic: 4 atoms per SSE step, one float  */
/* lane accumulator per (entry, lane), separate mul/add roundings, horizontal add at the end, */
/* remainder atoms added afterwards; float32 traces; Newton in double.                        */
/* ------------------------------------------------------------------------------------------ */
void orc_rmsd_f32(const float *xyz, const float *traces32, long n, int A, const float *ref,
                  float ref_trace, float *out)
{
#pragma omp parallel for schedule(static)
    for (long f = 0; f < n; ++f) {
        const float *p = xyz + (size_t)f * A * 3;
        float acc[9][4];
        memset(acc, 0, sizeof acc);
        const int A4 = A & ~3;
        for (int a = 0; a < A4; a += 4) {
            for (int l = 0; l < 4; ++l) {
                const float *x = p + 3 * (a + l), *y = ref + 3 * (a + l);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        const float prod = x[i] * y[j]; /* built with -ffp-contract=off */
                        acc[3 * i + j][l] = acc[3 * i + j][l] + prod;
                    }
            }
        }
        float Mf[9];
        for (int e = 0; e < 9; ++e) Mf[e] = (acc[e][0] + acc[e][1]) + (acc[e][2] + acc[e][3]);
        for (int a = A4; a < A; ++a) {
            const float *x = p + 3 * a, *y = ref + 3 * a;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    const float prod = x[i] * y[j];
                    Mf[3 * i + j] = Mf[3 * i + j] + prod;
                }
        }
        double M[9];
        for (int e = 0; e < 9; ++e) M[e] = Mf[e];
        const double msd = qcp_msd(M, (double)traces32[f], (double)ref_trace, A);
        out[f] = sqrtf((float)msd);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* The same arithmetic with SSE intrinsics, the way mdtraj's msd_atom_major runs it (SURVEY.md */
/* App. B step 3): 12 floats = 4 atoms are loaded as three vectors, shuffled AoS -> SoA, nine  */
/* mulps + nine addps per step, one vector accumulator per matrix entry, horizontal add at the */
/* end.  Lane l of accumulator e is exactly acc[e][l] of orc_rmsd_f32 above (mul and add round */
/* separately, -ffp-contract=off), so the two are BIT-IDENTICAL (tests/test_oracle_golden.py); */
/* this one is what the CPU baselines time.                                                   */
/* ------------------------------------------------------------------------------------------ */
#include <immintrin.h>
static inline void aos4_to_soa(const float *p, __m128 *x, __m128 *y, __m128 *z)
{
    const __m128 a = _mm_loadu_ps(p);      /* x0 y0 z0 x1 */
    const __m128 b = _mm_loadu_ps(p + 4);  /* y1 z1 x2 y2 */
    const __m128 c = _mm_loadu_ps(p + 8);  /* z2 x3 y3 z3 */
    const __m128 t0 = _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 3, 2)); /* x2 y2 z2 x3 */
    const __m128 t1 = _mm_shuffle_ps(a, b, _MM_SHUFFLE(1, 0, 3, 2)); /* z0 x1 y1 z1 */
    *x = _mm_shuffle_ps(a, t0, _MM_SHUFFLE(3, 0, 3, 0));             /* x0 x1 x2 x3 */
    *y = _mm_shuffle_ps(t1, c, _MM_SHUFFLE(2, 1, 2, 1));             /* y1.. fixed below */
    /* y: y0 = a[1], y1 = b[0], y2 = b[3], y3 = c[2] */
    {
        const __m128 ay = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 1, 1));   /* y0 y0 y1 y1 */
        const __m128 cy = _mm_shuffle_ps(b, c, _MM_SHUFFLE(2, 2, 3, 3));   /* y2 y2 y3 y3 */
        *y = _mm_shuffle_ps(ay, cy, _MM_SHUFFLE(2, 0, 2, 0));              /* y0 y1 y2 y3 */
    }
    /* z: z0 = a[2], z1 = b[1], z2 = c[0], z3 = c[3] */
    {
        const __m128 az = _mm_shuffle_ps(a, b, _MM_SHUFFLE(1, 1, 2, 2));   /* z0 z0 z1 z1 */
        const __m128 cz = _mm_shuffle_ps(c, c, _MM_SHUFFLE(3, 3, 0, 0));   /* z2 z2 z3 z3 */
        *z = _mm_shuffle_ps(az, cz, _MM_SHUFFLE(2, 0, 2, 0));              /* z0 z1 z2 z3 */
    }
}

static inline float hsum4(__m128 v)
{
    float l[4];
    _mm_storeu_ps(l, v);
    return (l[0] + l[1]) + (l[2] + l[3]);
}

void orc_rmsd_f32_sse(const float *xyz, const float *traces32, long n, int A, const float *ref,
                      float ref_trace, float *out)
{
    const int A4 = A & ~3;
    /* the centre's SoA form is shared by every frame */
    float *rsoa = (float *)malloc((size_t)(A4 + 4) * 3 * sizeof(float));
    for (int a = 0; a < A4; a += 4) {
        __m128 x, y, z;
        aos4_to_soa(ref + 3 * a, &x, &y, &z);
        _mm_storeu_ps(rsoa + 3 * a, x);
        _mm_storeu_ps(rsoa + 3 * a + 4, y);
        _mm_storeu_ps(rsoa + 3 * a + 8, z);
    }
#pragma omp parallel for schedule(static)
    for (long f = 0; f < n; ++f) {
        const float *p = xyz + (size_t)f * A * 3;
        __m128 acc[9];
        for (int e = 0; e < 9; ++e) acc[e] = _mm_setzero_ps();
        for (int a = 0; a < A4; a += 4) {
            __m128 x[3], y[3];
            aos4_to_soa(p + 3 * a, &x[0], &x[1], &x[2]);
            y[0] = _mm_loadu_ps(rsoa + 3 * a);
            y[1] = _mm_loadu_ps(rsoa + 3 * a + 4);
            y[2] = _mm_loadu_ps(rsoa + 3 * a + 8);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    acc[3 * i + j] = _mm_add_ps(acc[3 * i + j], _mm_mul_ps(x[i], y[j]));
        }
        float Mf[9];
        for (int e = 0; e < 9; ++e) Mf[e] = hsum4(acc[e]);
        for (int a = A4; a < A; ++a) {
            const float *x = p + 3 * a, *y = ref + 3 * a;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    const float prod = x[i] * y[j];
                    Mf[3 * i + j] = Mf[3 * i + j] + prod;
                }
        }
        double M[9];
        for (int e = 0; e < 9; ++e) M[e] = Mf[e];
        const double msd = qcp_msd(M, (double)traces32[f], (double)ref_trace, A);
        out[f] = sqrtf((float)msd);
    }
    free(rsoa);
}

/* ------------------------------------------------------------------------------------------ */
/* md.rmsd(traj, ref) exactly as the reference calls it (NOT precentered): copy + centre both */
/* every call, then QCP.  This is what the reference pays per k-centers iteration (SURVEY.md  */
/* 3.2), so it is what the CPU baseline times.  mode 0 = f64 truth, 1 = mdtraj-like float32 (scalar lanes), 2 = the same with SSE. */
/* ------------------------------------------------------------------------------------------ */
int orc_md_rmsd(const float *xyz, long n, int A, const float *ref, int mode, float *out)
{
    const size_t nf = (size_t)n * A * 3;
    float *copy = (float *)malloc(nf * sizeof(float));
    float *rc = (float *)malloc((size_t)A * 3 * sizeof(float));
    double *t64 = (double *)malloc((size_t)n * sizeof(double));
    float *t32 = (float *)malloc((size_t)n * sizeof(float));
    if (!copy || !rc || !t64 || !t32) {
        free(copy); free(rc); free(t64); free(t32);
        return 1;
    }
#pragma omp parallel for schedule(static)
    for (long f = 0; f < n; ++f)
        memcpy(copy + (size_t)f * A * 3, xyz + (size_t)f * A * 3, (size_t)A * 3 * sizeof(float));
    memcpy(rc, ref, (size_t)A * 3 * sizeof(float));
    orc_center_and_trace(copy, n, A, t64, t32);
    double rt64;
    float rt32;
    orc_center_and_trace(rc, 1, A, &rt64, &rt32);
    if (mode == 0)
        orc_rmsd_f64(copy, t64, n, A, rc, rt64, out);
    else if (mode == 1)
        orc_rmsd_f32(copy, t32, n, A, rc, rt32, out);
    else
        orc_rmsd_f32_sse(copy, t32, n, A, rc, rt32, out);
    free(copy); free(rc); free(t64); free(t32);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* libdist restatement (libdist.pyx:100-145).  The generated C for float32 input computes     */
/* d = X[i,j] - y[j] in float, squares it in float (powf(d, 2.0f), which gcc folds to d*d),   */
/* and adds it to the float64 out[i] in j order; integer inputs subtract and square in the    */
/* promoted C integer type.  sqrt is the C double sqrt.                                       */
/* ------------------------------------------------------------------------------------------ */
#define ORC_DEFINE_FLOAT_DIST(SUFFIX, T)                                                       \
    void orc_euclidean_##SUFFIX(const T *X, long n, long F, const T *y, double *out)           \
    {                                                                                          \
        _Pragma("omp parallel for schedule(static)") for (long i = 0; i < n; ++i)              \
        {                                                                                      \
            const T *r = X + (size_t)i * F;                                                    \
            double acc = 0;                                                                    \
            for (long j = 0; j < F; ++j) {                                                     \
                const T d = r[j] - y[j];                                                    \
                const T s = d * d;                                                          \
                acc += (double)s;                                                              \
            }                                                                                  \
            out[i] = sqrt(acc);                                                                \
        }                                                                                      \
    }                                                                                          \
    void orc_manhattan_##SUFFIX(const T *X, long n, long F, const T *y, double *out)           \
    {                                                                                          \
        _Pragma("omp parallel for schedule(static)") for (long i = 0; i < n; ++i)              \
        {                                                                                      \
            const T *r = X + (size_t)i * F;                                                    \
            double acc = 0;                                                                    \
            for (long j = 0; j < F; ++j) {                                                     \
                const T d = r[j] - y[j];                                                    \
                acc += fabs((double)d);                                                        \
            }                                                                                  \
            out[i] = acc;                                                                      \
        }                                                                                      \
    }

ORC_DEFINE_FLOAT_DIST(f32, float)
ORC_DEFINE_FLOAT_DIST(f64, double)

#define ORC_DEFINE_INT_DIST(SUFFIX, T, W)                                                      \
    void orc_euclidean_##SUFFIX(const T *X, long n, long F, const T *y, double *out)           \
    {                                                                                          \
        for (long i = 0; i < n; ++i) {                                                         \
            const T *r = X + (size_t)i * F;                                                    \
            double acc = 0;                                                                    \
            for (long j = 0; j < F; ++j) {                                                     \
                const W d = (W)r[j] - (W)y[j];                                                 \
                acc += (double)(W)(d * d);                                                     \
            }                                                                                  \
            out[i] = sqrt(acc);                                                                \
        }                                                                                      \
    }                                                                                          \
    void orc_manhattan_##SUFFIX(const T *X, long n, long F, const T *y, double *out)           \
    {                                                                                          \
        for (long i = 0; i < n; ++i) {                                                         \
            const T *r = X + (size_t)i * F;                                                    \
            double acc = 0;                                                                    \
            for (long j = 0; j < F; ++j) {                                                     \
                const W d = (W)r[j] - (W)y[j];                                                 \
                acc += fabs((double)d);                                                        \
            }                                                                                  \
            out[i] = acc;                                                                      \
        }                                                                                      \
    }

ORC_DEFINE_INT_DIST(i8, int8_t, int)
ORC_DEFINE_INT_DIST(i16, int16_t, int)
ORC_DEFINE_INT_DIST(i32, int32_t, int)
ORC_DEFINE_INT_DIST(i64, int64_t, long)

/* ------------------------------------------------------------------------------------------ */
/* one k-centers bookkeeping pass (kcenters.py:304-306 and the argmax of :282 for the next    */
/* iteration): strict '<' update, first-occurrence argmax.  Returns the argmax index.         */
/* ------------------------------------------------------------------------------------------ */
long orc_kcenters_update_f32(const float *dist, long n, long center_id, double *distances,
                             long *assignments)
{
    long best = 0;
    double bestv = -1.0;
    for (long i = 0; i < n; ++i) {
        const double d = (double)dist[i];
        if (d < distances[i]) {
            distances[i] = d;
            assignments[i] = center_id;
        }
        if (distances[i] > bestv) {
            bestv = distances[i];
            best = i;
        }
    }
    return best;
}

/* ------------------------------------------------------------------------------------------ */
/* The synthetic-trajectory generator of enspara_b200/synth.py, restated in C   */
/* so that the CPU arms of bench.py can make their host sample without numpy's minutes and     */
/* without loading the product's CUDA library.  Integer hashing + separately rounded float32   */
/* operations only (-ffp-contract=off): bit-identical to synth.trajectory (tested).            */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t syn_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline float syn_u01(uint64_t key, uint64_t idx, uint64_t counter)
{
    const uint64_t z = syn_mix64(key ^ syn_mix64(idx * 0xD1B54A32D192ED03ULL + counter));
    return (float)(z >> 40) * 5.9604644775390625e-08f;
}
static inline float syn_gauss4(uint64_t key, uint64_t idx, uint64_t c)
{
    const float u0 = syn_u01(key, idx, c), u1 = syn_u01(key, idx, c + 1);
    const float u2 = syn_u01(key, idx, c + 2), u3 = syn_u01(key, idx, c + 3);
    const float s = (u0 + u1) + (u2 + u3);
    return (s - 2.0f) * 1.7320508f;
}

void orc_synth_trajectory(float *out, long n, int A, long first_frame, uint64_t seed,
                          const float *base, int n_base)
{
    const uint64_t key = syn_mix64(seed + 0x9E3779B97F4A7C15ULL * 2ULL); /* stream 1: frames */
#pragma omp parallel for schedule(static)
    for (long fl = 0; fl < n; ++fl) {
        const uint64_t f = (uint64_t)(first_frame + fl);
        const float sigma = 0.02f + 0.13f * syn_u01(key, f, 0);
        const float q0 = syn_gauss4(key, f, 1), q1 = syn_gauss4(key, f, 5);
        const float q2 = syn_gauss4(key, f, 9), q3 = syn_gauss4(key, f, 13);
        float qn = sqrtf((q0 * q0 + q1 * q1) + (q2 * q2 + q3 * q3));
        if (!(qn > 1e-6f)) qn = 1e-6f;
        const float w = q0 / qn, x = q1 / qn, y = q2 / qn, z = q3 / qn;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
        const float xw = x * w, yw = y * w, zw = z * w;
        const float r00 = 1.0f - 2.0f * (yy + zz), r01 = 2.0f * (xy - zw), r02 = 2.0f * (xz + yw);
        const float r10 = 2.0f * (xy + zw), r11 = 1.0f - 2.0f * (xx + zz), r12 = 2.0f * (yz - xw);
        const float r20 = 2.0f * (xz - yw), r21 = 2.0f * (yz + xw), r22 = 1.0f - 2.0f * (xx + yy);
        const float t0 = 2.0f * syn_u01(key, f, 17) - 1.0f;
        const float t1 = 2.0f * syn_u01(key, f, 18) - 1.0f;
        const float t2 = 2.0f * syn_u01(key, f, 19) - 1.0f;
        const float *b = base + (size_t)(f % (uint64_t)n_base) * A * 3;
        float *o = out + (size_t)fl * A * 3;
        for (int a = 0; a < A; ++a) {
            const uint64_t c0 = 32 + (uint64_t)(3 * a) * 4;
            const float px = b[3 * a] + sigma * syn_gauss4(key, f, c0);
            const float py = b[3 * a + 1] + sigma * syn_gauss4(key, f, c0 + 4);
            const float pz = b[3 * a + 2] + sigma * syn_gauss4(key, f, c0 + 8);
            o[3 * a] = ((r00 * px + r01 * py) + r02 * pz) + t0;
            o[3 * a + 1] = ((r10 * px + r11 * py) + r12 * pz) + t1;
            o[3 * a + 2] = ((r20 * px + r21 * py) + r22 * pz) + t2;
        }
    }
}

void orc_synth_features(float *X, long n, long F, long first_row, uint64_t seed)
{
    const uint64_t key = syn_mix64(seed + 0x9E3779B97F4A7C15ULL * 3ULL); /* stream 2: features */
#pragma omp parallel for schedule(static)
    for (long r = 0; r < n; ++r)
        for (long j = 0; j < F; ++j)
            X[r * F + j] = syn_u01(key, (uint64_t)(first_row + r), (uint64_t)j);
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
