// Counter-based synthetic data generator, bit-identical to enspara_b200/synth.py (numpy):
// integer hashing (splitmix64 finaliser) and correctly rounded float32 + - * / sqrt only, every
// operation rounded separately (no FMA contraction), so a shard generated in HBM equals the
// same frames generated on any host.  Lets a 10M-frame x 500-atom trajectory (60 GB) exist
// without ever touching host memory.
#include "eb_common.cuh"

namespace eb {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, int stream)
{
    return mix64(seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(stream + 1));
}

__device__ __forceinline__ float u01(uint64_t key, uint64_t idx, uint64_t counter)
{
    const uint64_t z = mix64(key ^ mix64(idx * 0xD1B54A32D192ED03ULL + counter));
    return __fmul_rn((float)(z >> 40), 5.9604644775390625e-08f);  // 2^-24, exact
}

__device__ __forceinline__ float gauss4(uint64_t key, uint64_t idx, uint64_t c)
{
    const float u0 = u01(key, idx, c), u1 = u01(key, idx, c + 1);
    const float u2 = u01(key, idx, c + 2), u3 = u01(key, idx, c + 3);
    const float s = __fadd_rn(__fadd_rn(u0, u1), __fadd_rn(u2, u3));
    return __fmul_rn(__fsub_rn(s, 2.0f), 1.7320508f);
}

constexpr int kStreamFrame = 1, kStreamFeat = 2;

__global__ void __launch_bounds__(256)
k_synth_trajectory(float *__restrict__ out, long n, int A, long first_frame, uint64_t key,
                   const float *__restrict__ base, int n_base)
{
    const long total = n * (long)A;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long)gridDim.x * blockDim.x) {
        const long fl = t / A;
        const int a = (int)(t - fl * A);
        const uint64_t f = (uint64_t)(first_frame + fl);
        // per-frame scalars (recomputed per atom: ~25 hashes, cheap next to the 12 below)
        const float sigma = __fadd_rn(0.02f, __fmul_rn(0.13f, u01(key, f, 0)));
        const float q0 = gauss4(key, f, 1), q1 = gauss4(key, f, 5);
        const float q2 = gauss4(key, f, 9), q3 = gauss4(key, f, 13);
        float qn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)),
                                        __fadd_rn(__fmul_rn(q2, q2), __fmul_rn(q3, q3))));
        qn = fmaxf(qn, 1e-6f);
        const float w = __fdiv_rn(q0, qn), x = __fdiv_rn(q1, qn);
        const float y = __fdiv_rn(q2, qn), z = __fdiv_rn(q3, qn);
        const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
        const float xy = __fmul_rn(x, y), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z);
        const float xw = __fmul_rn(x, w), yw = __fmul_rn(y, w), zw = __fmul_rn(z, w);
        const float r00 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(yy, zz)));
        const float r01 = __fmul_rn(2.0f, __fsub_rn(xy, zw));
        const float r02 = __fmul_rn(2.0f, __fadd_rn(xz, yw));
        const float r10 = __fmul_rn(2.0f, __fadd_rn(xy, zw));
        const float r11 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(xx, zz)));
        const float r12 = __fmul_rn(2.0f, __fsub_rn(yz, xw));
        const float r20 = __fmul_rn(2.0f, __fsub_rn(xz, yw));
        const float r21 = __fmul_rn(2.0f, __fadd_rn(yz, xw));
        const float r22 = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(xx, yy)));
        const float t0 = __fsub_rn(__fmul_rn(2.0f, u01(key, f, 17)), 1.0f);
        const float t1 = __fsub_rn(__fmul_rn(2.0f, u01(key, f, 18)), 1.0f);
        const float t2 = __fsub_rn(__fmul_rn(2.0f, u01(key, f, 19)), 1.0f);

        const float *b = base + ((size_t)(f % (uint64_t)n_base) * A + a) * 3;
        const uint64_t c0 = 32 + (uint64_t)(3 * a) * 4;
        const float px = __fadd_rn(b[0], __fmul_rn(sigma, gauss4(key, f, c0)));
        const float py = __fadd_rn(b[1], __fmul_rn(sigma, gauss4(key, f, c0 + 4)));
        const float pz = __fadd_rn(b[2], __fmul_rn(sigma, gauss4(key, f, c0 + 8)));
        float *o = out + (size_t)t * 3;
        o[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r00, px), __fmul_rn(r01, py)),
                                   __fmul_rn(r02, pz)), t0);
        o[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r10, px), __fmul_rn(r11, py)),
                                   __fmul_rn(r12, pz)), t1);
        o[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r20, px), __fmul_rn(r21, py)),
                                   __fmul_rn(r22, pz)), t2);
    }
}

__global__ void __launch_bounds__(256)
k_synth_features(float *__restrict__ X, long n, long F, long first_row, uint64_t key)
{
    const long total = n * F;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long)gridDim.x * blockDim.x) {
        const long r = t / F;
        const long j = t - r * F;
        X[t] = u01(key, (uint64_t)(first_row + r), (uint64_t)j);
    }
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_synth_trajectory_aos(float *xyz_aos, int64_t n, int n_atoms, int64_t first_frame,
                            uint64_t seed, const float *base_conformers, int n_base,
                            void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_atoms > 0 && n_base > 0, "synth_trajectory: bad shape");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(xyz_aos && base_conformers, "synth_trajectory: null pointer");
    long blocks = (n * (long)n_atoms + 255) / 256;
    if (blocks > 32L * sm_count()) blocks = 32L * sm_count();
    k_synth_trajectory<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        xyz_aos, n, n_atoms, first_frame, stream_key(seed, kStreamFrame), base_conformers,
        n_base);
    EB_LAUNCH_CHECK();
    return EB_OK;
}

int eb_synth_features(float *X, int64_t n, int64_t n_features, int64_t first_row, uint64_t seed,
                      void *stream)
{
    EB_CHECK_ARG(n >= 0 && n_features > 0, "synth_features: bad shape");
    if (n == 0) return EB_OK;
    EB_CHECK_ARG(X, "synth_features: null pointer");
    long blocks = (n * (long)n_features + 255) / 256;
    if (blocks > 32L * sm_count()) blocks = 32L * sm_count();
    k_synth_features<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        X, n, n_features, first_row, stream_key(seed, kStreamFeat));
    EB_LAUNCH_CHECK();
    return EB_OK;
}

}  // extern "C"
