// Early-ray-termination cast: instantiations and launch (see ert.h).
#include "ert.h"

#define YTB_CAST_KERNEL_ONLY
#include "cast_kernels.cuh"

namespace ytb {

template <int TFM, bool PARTIAL, bool MANY>
static cudaError_t launch_one_ert(const CastParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    void (*k)(const CastParams) = cast_kernel<KIND_VR, false, TFM, PARTIAL, true, float, MANY, true>;
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, CTA_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

template <int TFM>
static cudaError_t launch_tfm(const CastParams &p, unsigned grid, size_t smem, bool partial, bool many, cudaStream_t st)
{
    if (partial) return many ? launch_one_ert<TFM, true, true>(p, grid, smem, st) : launch_one_ert<TFM, true, false>(p, grid, smem, st);
    return many ? launch_one_ert<TFM, false, true>(p, grid, smem, st) : launch_one_ert<TFM, false, false>(p, grid, smem, st);
}

cudaError_t launch_ert(const CastParams &p, unsigned grid, size_t smem, int tfm_bits, bool partial, bool many, cudaStream_t st)
{
    switch (tfm_bits & (TFM_ZERO_INSIDE | TFM_EDGE)) {
    case 0: return launch_tfm<TFM_COLOUR + TFM_GREY>(p, grid, smem, partial, many, st);
    case TFM_ZERO_INSIDE: return launch_tfm<TFM_COLOUR + TFM_GREY + TFM_ZERO_INSIDE>(p, grid, smem, partial, many, st);
    case TFM_EDGE: return launch_tfm<TFM_COLOUR + TFM_GREY + TFM_EDGE>(p, grid, smem, partial, many, st);
    default: return launch_tfm<TFM_COLOUR + TFM_GREY + TFM_ZERO_INSIDE + TFM_EDGE>(p, grid, smem, partial, many, st);
    }
}

}  // namespace ytb
