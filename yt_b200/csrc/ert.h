// Host entry point of the early-ray-termination cast (ert.cu): the grey-opacity ColorTransferFunction volume render
// walked from the far end of the composition order (see cast_kernel<..., ERT> in cast_kernels.cuh).  Its own translation
// unit: sixteen more instantiations of the heavy kernel, compiled beside ytb200.cu.
#pragma once
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace ytb {

// `tfm_bits`: TFM_ZERO_INSIDE | TFM_EDGE of the table (TFM_GREY is implied); `many`: the long-brick-list instantiation.
cudaError_t launch_ert(const CastParams &p, unsigned grid, size_t smem, int tfm_bits, bool partial, bool many, cudaStream_t st);

}  // namespace ytb
