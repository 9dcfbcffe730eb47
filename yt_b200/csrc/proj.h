// Host entry points of the slab-walk projection cast (proj.cu / proj_kernels.cuh); its own translation unit so that
// the kernel rebuilds in seconds.
#pragma once
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace ytb {

// lane -> pixel map of the projection cast, chosen per launch on the host (make_proj_map)
struct ProjMap {
    int mode;        // 0: 16x16 tiles of 4x8 patches (as cast_kernel); 1: sheared rows (32 pixels along f, 8 rows along s)
    int f_is_j;      // mode 1: the fast pixel axis f is the image's j axis (else i)
    int nf, ns;      // mode 1: image size along f and s
    float slope;     // mode 1: shift along s per pixel along f, |slope| <= 1
    int s_lo;        // mode 1: s of row 0 of tile row 0 (<= 0: the shear needs rows above the image)
    int n_ftiles, n_stiles, band;   // tile grid and the number of s-tiles issued together (both modes)
    int s_first;     // mode 0: the launch covers s-tiles [s_first, s_first + n_stiles) only (one row band of a staged frame)
};

// `first`: the first brick of the order list (its cell size fixes the dominant axis), or nullptr.
// `host_image`: the cast reads / writes a page-locked host image in place over PCIe.
ProjMap make_proj_map(const DevCamera &cam, const DevBrick *first, bool host_image);
cudaError_t launch_proj(const CastParams &p, const ProjMap &m, int n_fields, bool partial, cudaStream_t st);

}  // namespace ytb
