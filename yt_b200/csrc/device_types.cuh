// Device-side data layout shared by the kernels and the C ABI (libytb200.so).
//
// Naming follows the reference's domain: a *brick* is one PartitionedGrid /
// VolumeContainer (yt/utilities/lib/volume_container.pxd:13-33), a *frame* is the
// image + z-buffer + ray set of one ImageSampler (image_samplers.pyx:59-130), a
// *table* is one FieldInterpolationTable (field_interpolation_tables.pxd:27-39).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define YTB_MAX_FIELDS 6
#define YTB_MAX_TABLES 6

namespace ytb {

// One brick as the kernels see it.  Geometry stays fp64 (the DDA is fp64); `data`
// points at fp32 or fp64 arrays depending on the owning set's storage.
struct DevBrick {
    double le[3], re[3], dds[3], idds[3];
    int dims[3];
    int n_fields;
    // element strides of the data arrays: plane (x) and row (y) pitch.  Contiguous C order unless the owner padded
    // the rows -- a z-row whose byte length is a multiple of 256 B puts every row of a ray's neighbourhood into the
    // same L1 sets (r01z: 1024^3 fp32 cells cast 24 % slower than 1000^3 or 1032^3), so library-owned arrays pad it.
    int pitch_x, pitch_y;
    const void *data[YTB_MAX_FIELDS];
    const unsigned char *mask;  // nullptr == all ones
};

// Ray description.  Plane-parallel rays are generated in the kernel from vp_pos/vp_dir
// exactly as lenses.pyx:179-192 does; every other lens reads per-pixel fp64 arrays.
struct DevCamera {
    int lens, nx, ny;
    double vp_pos[12];
    double vp_dir[3];
    double width[3];
    const double *pos_px;  // device (nx,ny,3) or nullptr
    const double *dir_px;  // device (nx,ny,3) or nullptr
    // conservative pixel-footprint helpers (ours, not the reference's bbox: SURVEY.md H5/H6/H12)
    double ext_rows[6];    // plane-parallel: rows of inv([east north dir]) giving (px,py) of a point
    double ext_origin[3];  // plane-parallel: back_center; perspective: camera position
    double persp[15];      // perspective: camera_data rows (position,width,east,north,normal)
    int ext_mode;          // 0 full image, 1 plane-parallel projection, 2 pinhole projection
};

// Transfer-function description.  Table payloads live in global memory and are staged
// into shared memory by each CTA.
struct DevTF {
    int n_fits, grey, colour_fast;  // colour_fast: ColorTransferFunction layout (see tf_create)
    int nbins[YTB_MAX_TABLES];
    int offset[YTB_MAX_TABLES];     // element offset of table i inside the payload arrays
    int field_id[YTB_MAX_TABLES], weight_field_id[YTB_MAX_TABLES], weight_table_id[YTB_MAX_TABLES];
    int field_table_id[6];
    int total_bins;
    // fp64 (EXACT): values / d0 / dy exactly as FIT_initialize_table builds them
    double b0[YTB_MAX_TABLES], b1[YTB_MAX_TABLES], idbin[YTB_MAX_TABLES];
    const double *values, *d0, *dy;
    // fp32 (FAST): (value, next-value - value) pairs; bounds split hi/lo so x-b0 keeps fp64-ish accuracy
    float b0_hi[YTB_MAX_TABLES], b0_lo[YTB_MAX_TABLES], b0_f[YTB_MAX_TABLES], b1_f[YTB_MAX_TABLES], idbin_f[YTB_MAX_TABLES];
    const float2 *pairs;
    // colour_fast payload: (r,g,b,a) rows [0, 0, y_0 .. y_{n-1}, y_{n-1}]; replicated 2^colour_copy_shift times in smem
    const float4 *colour;
    int colour_copy_shift;
    // the same rows as a cast CTA holds them in shared memory (every row 2^colour_copy_shift times): one bulk async copy
    const float4 *colour_image;
    // row-coordinate constants of the colour fast path (see ColourTable in raycast.cuh); they stay in the
    // kernel-parameter constant bank instead of in registers
    float colour_c0, colour_clamp_bound, colour_lo, colour_hi, colour_mid, colour_half, colour_band, colour_zrow;
    float colour_clamp_hi;   // delta layout: row coordinates are clamped to [colour_clamp_lo, colour_clamp_hi] = [0.5, nbins + 1.5] - origin
    // LINE form of the delta layout (colour_line, decided per table at ytb200_tf_create): a row holds (c, d) with the bin's
    // line written in the row coordinate itself, y = d * fs + c, so the lerp needs no fractional part.  Every row coordinate
    // of the fast path is then measured from `origin` (an integer, folded into colour_c0): clamps and the round-down magic
    // constant move with it (colour_magic = 2^23 + origin, so the mantissa bits still hold the true row index).
    float colour_clamp_lo, colour_magic;
    int colour_line;
    int colour_zero_inside, colour_edge, colour_delta, colour_rows, colour_nonneg;
    // LightSourceRenderSampler (image_samplers.pyx:496-600): one directional light, gradient shading
    int light_on;
    double light_dir[3], light_rgba[4];
};

// Formats of a partial image: the per-pixel affine map C -> A + B*C a rank's bricks apply to whatever lies behind them
// (SURVEY.md A.4).  What crosses NVLink is the smallest exact description of that map:
//   AB64 / AB32   (A[4], B[4]) in fp64 / fp32: any transfer function (64 / 32 B per pixel)
//   A32 / A32G    the ColorTransferFunction render (non-grey / grey opacity), fp32 A[4] only (16 B): while no opacity
//                 clamp has fired B = 1 - A per channel (grey: 1 - A_alpha for all four), and a clamp zeroes B for good
//                 (ColourAcc), so B is "1 - A, or 0 when flagged"; flags: non-grey in the unused 4th channel (bit c =
//                 channel c), grey in the sign of A_alpha (alpha tables are non-negative: checked at tf_create)
//   SUM1/2/4      the projection samplers: B == 1, A = 1 / 2 / 4 fp64 sums (8 / 16 / 32 B)
enum PartialFormat { PF_AB64 = 0, PF_AB32 = 1, PF_A32 = 2, PF_A32G = 3, PF_SUM1 = 4, PF_SUM2 = 5, PF_SUM4 = 6, PF_COUNT = 7 };
__host__ __device__ constexpr int partial_bytes(int fmt)
{
    return fmt == PF_AB64 ? 64 : fmt == PF_AB32 ? 32 : (fmt == PF_A32 || fmt == PF_A32G) ? 16 : fmt == PF_SUM1 ? 8 : fmt == PF_SUM2 ? 16 : 32;
}

struct CastParams {
    DevCamera cam;
    const DevBrick *bricks;
    const int *order;       // brick indices in the reference's traversal order
    const int4 *extent;     // per order entry: [i0,i1) x [j0,j1) pixel footprint
    const int4 *chunk_extent;   // per CTA_THREADS consecutive order entries: the union of their footprints
    int n_order;
    int n_samples;
    double inv_ns;          // 1 / n_samples
    float iproj_s0, iproj_s1, iproj_s2, iproj_s3;   // sum over s = 0 .. n_samples-1 of s^0, s^1, s^2, s^3 (interpolated projection)
    double *image;          // (nx,ny,4) fp64, in/out
    const double *zbuffer;  // (nx,ny) fp64 or nullptr (== ones)
    double *partial;        // affine-map output in format `partial_fmt` (see PartialFormat), or nullptr
    int partial_fmt;
    // PARTIAL casts feeding a peer swap: when the whole grid has stored its pixels, the last CTA writes `tail_seq` into
    // `tail_flag` (the round-0 partner's flag word, over NVLink); `tail_counter` counts the CTAs (this rank's memory)
    unsigned *tail_counter;
    unsigned long long *tail_flag;
    unsigned long long tail_seq;
    DevTF tf;
    unsigned long long *counters;  // [0] cells handed to sample(); [1] cells the early-termination cast sampled
    // Early ray termination (ert.cu; grey-opacity ColorTransferFunction only): > 0 selects the REVERSED cast -- the order
    // list arrives reversed and every ray is walked from its far end, compositing front to back, until the transmittance
    // of everything processed so far drops below ert_eps (see cast_kernel<..., ERT>)
    float ert_eps;
};

}  // namespace ytb
