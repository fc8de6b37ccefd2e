// TMA / mbarrier helpers (sm_100a): tensor-map creation on the host through the driver entry
// point (no link-time dependency on libcuda) and the device-side PTX wrappers.
#pragma once
#include <cuda.h>

#include "eb_common.cuh"

namespace eb {

// ------------------------------------------------------------------------------------------
// host: 2-D row-major tensor map  [rows][cols] of `elem_bytes`-sized elements
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                        const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encode_fn()
{
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// rows x cols matrix with row pitch `pitch_bytes`; box = box_rows x box_cols; swizzle 128 B
// requires box_cols * elem_bytes == 128.  Returns 0 on success.
inline int make_tmap_2d(CUtensorMap *out, const void *base, CUtensorMapDataType dt,
                        size_t elem_bytes, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz)
{
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(EB_ERR_CUDA, "%s", "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(out, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(EB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%s %ld)", "code", (long)r);
    (void)elem_bytes;
    return EB_OK;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// plain bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 2-D tiled tensor copy global -> shared; c0 = innermost (column) coordinate, c1 = row
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// the same with an L2 eviction policy (createpolicy): rows a persistent kernel re-reads every
// iteration are loaded evict_last, the streamed remainder evict_first, so the re-read part stays
// L2-resident across iterations
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *tmap, int c0,
                                                 int c1, uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
#endif

}  // namespace eb
