// The cast kernels: one CTA per 16x16-pixel image tile, one thread per ray, walking the
// view-ordered brick list (the reference's `for brick in traverse(): sampler(brick)` loop,
// render_source.py:594-597 + image_samplers.pyx:146-197, fused into one launch).
//
// Because a pixel is owned by exactly one thread for the whole brick list, the reference's
// compositing order (bricks in kd order, cells in increasing t inside a brick -- SURVEY.md
// finding 4 / H14) is preserved by construction and there are no races.
#pragma once
#include "raycast.cuh"

namespace ytb {

constexpr int TILE_I = 16, TILE_J = 16, CTA_THREADS = 256;
#ifndef YTB_MIN_BLOCKS
#define YTB_MIN_BLOCKS 2
#endif
enum { KIND_PROJ = 0, KIND_IPROJ = 1, KIND_VR = 2 };
enum { TFM_NONE = 0, TFM_GENERAL = 1, TFM_COLOUR = 2 };

// ---------------------------------------------------------------------------------------------
// Per-thread sampler state.  KIND/EXACT/TFM/PARTIAL/NF1 are compile-time; T is the storage type.
//   NF1 : the brick set has exactly one field (corners cached in registers per cell).
// ---------------------------------------------------------------------------------------------
template <int KIND, bool EXACT, int TFM, bool PARTIAL, bool NF1, typename T>
struct CellSampler {
    // ray
    double vp[3], vd[3];
    // per-brick
    const DevBrick *br;
    double dds[3], idds[3], le[3];
    double o[3], dc[3];  // FAST: ray origin / direction in cell units of this brick
    int ny, nz;          // cell dims (y,z)
    int n_fields;
    const unsigned char *mask;
    // sampler constants
    int n_samples;
    double inv_ns;
    const DevTF *tf;
    TFShared64 t64;
    const float2 *pairs;
    ColourTable ct;
    ColourAcc ca;
    // accumulators
    double acc[4];   // PROJ / IPROJ / EXACT VR
    double B64[4];
    float rgba[4];   // FAST VR
    float B[4];
    unsigned long long n_cells;

    __device__ __forceinline__ void set_brick(const DevBrick *b)
    {
        br = b;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dds[i] = b->dds[i]; idds[i] = b->idds[i]; le[i] = b->le[i];
            if (!EXACT) { o[i] = (vp[i] - le[i]) * idds[i]; dc[i] = vd[i] * idds[i]; }
        }
        ny = b->dims[1]; nz = b->dims[2];
        n_fields = b->n_fields;
        mask = b->mask;
    }

    __device__ __forceinline__ void cell(double enter_t, double exit_t, int ix, int iy, int iz)
    {
        if (KIND == KIND_PROJ) {
            const double dl = dsub(exit_t, enter_t);
            const size_t di = ((size_t)ix * ny + iy) * nz + iz;
            const int nf = n_fields < 4 ? n_fields : 4;
#pragma unroll
            for (int f = 0; f < 4; f++) {
                if (f < nf) {
                    const double v = (double)__ldg((const T *)br->data[f] + di);
                    acc[f] = EXACT ? dadd(acc[f], dmul(v, dl)) : fma(v, dl, acc[f]);
                }
            }
            return;
        } else {
            if (KIND == KIND_VR && mask != nullptr) {
                if (mask[((size_t)ix * ny + iy) * nz + iz] != 1) return;
            }
            const int sy = nz + 1, sx = (ny + 1) * (nz + 1);
            const size_t offset = (size_t)ix * sx + (size_t)iy * sy + iz;
            if (EXACT) {
                cell_exact(enter_t, exit_t, ix, iy, iz, sx, sy, offset);
            } else {
                cell_fast(enter_t, exit_t, ix, iy, iz, sx, sy, offset);
            }
        }
    }

    // Reference evaluation order, all fp64 (image_samplers.pyx:372-402 / 453-487).
    __device__ __forceinline__ void cell_exact(double enter_t, double exit_t, int ix, int iy, int iz, int sx, int sy,
                                               size_t offset)
    {
        const int idx[3] = {ix, iy, iz};
        double dp[3], ds[3], dvs[6];
        const double dt = ddiv(dsub(exit_t, enter_t), (double)n_samples);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dp[i] = dadd(dmul(dadd(enter_t, dmul(0.5, dt)), vd[i]), vp[i]);
            dp[i] = dsub(dp[i], dadd(dmul((double)idx[i], dds[i]), le[i]));
            dp[i] = dmul(dp[i], idds[i]);
            ds[i] = dmul(dmul(vd[i], idds[i]), dt);
        }
        for (int s = 0; s < n_samples; s++) {
#pragma unroll
            for (int j = 0; j < 6; j++)
                dvs[j] = (j < n_fields) ? offset_interpolate64((const double *)br->data[j] + offset, sx, sy, dp) : 0.0;
            if (KIND == KIND_VR) {
                eval_transfer64<PARTIAL>(*tf, t64, dt, dvs, acc, B64);
            } else {
                const int nf = n_fields < 3 ? n_fields : 3;
#pragma unroll
                for (int j = 0; j < 3; j++)
                    if (j < nf) acc[j] = dadd(acc[j], dmul(dvs[j], dt));
            }
#pragma unroll
            for (int i = 0; i < 3; i++) dp[i] = dadd(dp[i], ds[i]);
        }
    }

    // fp64 cell entry, fp32 samples.
    __device__ __forceinline__ void cell_fast(double enter_t, double exit_t, int ix, int iy, int iz, int sx, int sy,
                                              size_t offset)
    {
        const double dt64 = (exit_t - enter_t) * inv_ns;
        const double tmid = fma(0.5, dt64, enter_t);
        float x = (float)(fma(tmid, dc[0], o[0]) - (double)ix);
        float y = (float)(fma(tmid, dc[1], o[1]) - (double)iy);
        float z = (float)(fma(tmid, dc[2], o[2]) - (double)iz);
        const float sxs = (float)(dc[0] * dt64), sys = (float)(dc[1] * dt64), szs = (float)(dc[2] * dt64);
        const float dt = (float)dt64;
        if (NF1) {
            Corners c;
            load_corners(c, (const T *)br->data[0] + offset, sx, sy);
            if (KIND == KIND_IPROJ) {
                // the projection family carries the 1e-6 bar: sample in fp64 (contraction allowed)
                const T *d = (const T *)br->data[0] + offset;
                const double a0 = (double)__ldg(d), a1 = (double)__ldg(d + 1), b0 = (double)__ldg(d + sy), b1 = (double)__ldg(d + sy + 1);
                const double e0 = (double)__ldg(d + sx), e1 = (double)__ldg(d + sx + 1), f0 = (double)__ldg(d + sx + sy), f1 = (double)__ldg(d + sx + sy + 1);
                double xd = fma(tmid, dc[0], o[0]) - (double)ix, yd = fma(tmid, dc[1], o[1]) - (double)iy, zd = fma(tmid, dc[2], o[2]) - (double)iz;
                const double sxd = dc[0] * dt64, syd = dc[1] * dt64, szd = dc[2] * dt64;
                double sum = 0.0;
                for (int s = 0; s < n_samples; s++) {
                    const double v00 = fma(zd, a1 - a0, a0), v01 = fma(zd, b1 - b0, b0);
                    const double v10 = fma(zd, e1 - e0, e0), v11 = fma(zd, f1 - f0, f0);
                    const double v0 = fma(yd, v01 - v00, v00), v1 = fma(yd, v11 - v10, v10);
                    sum += fma(xd, v1 - v0, v0);
                    xd += sxd; yd += syd; zd += szd;
                }
                acc[0] = fma(sum, dt64, acc[0]);
            } else {
                if (TFM == TFM_COLOUR) {
                    CornersP cp;
                    load_corners_p(cp, (const T *)br->data[0] + offset, sx, sy);
                    if (tf->grey) colour_cell_run<PARTIAL, true>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dt, ca);
                    else colour_cell_run<PARTIAL, false>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dt, ca);
                } else {
                    float dvs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int s = 0; s < n_samples; s++) {
                        dvs[0] = trilinear(c, x, y, z);
                        eval_transfer32<PARTIAL>(*tf, pairs, dt, dvs, rgba, B);
                        x += sxs; y += sys; z += szs;
                    }
                }
            }
        } else {
            // any number of fields: corners re-read (L1) per sample; general but not the tuned path
            if (KIND == KIND_IPROJ) {
                double xd = fma(tmid, dc[0], o[0]) - (double)ix, yd = fma(tmid, dc[1], o[1]) - (double)iy, zd = fma(tmid, dc[2], o[2]) - (double)iz;
                const double sxd = dc[0] * dt64, syd = dc[1] * dt64, szd = dc[2] * dt64;
                const int nf = n_fields < 3 ? n_fields : 3;
                double sum[3] = {0.0, 0.0, 0.0};
                for (int s = 0; s < n_samples; s++) {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        if (j < nf) {
                            const T *d = (const T *)br->data[j] + offset;
                            const double a0 = (double)__ldg(d), a1 = (double)__ldg(d + 1), b0 = (double)__ldg(d + sy), b1 = (double)__ldg(d + sy + 1);
                            const double e0 = (double)__ldg(d + sx), e1 = (double)__ldg(d + sx + 1), f0 = (double)__ldg(d + sx + sy), f1 = (double)__ldg(d + sx + sy + 1);
                            const double v00 = fma(zd, a1 - a0, a0), v01 = fma(zd, b1 - b0, b0);
                            const double v10 = fma(zd, e1 - e0, e0), v11 = fma(zd, f1 - f0, f0);
                            const double v0 = fma(yd, v01 - v00, v00), v1 = fma(yd, v11 - v10, v10);
                            sum[j] += fma(xd, v1 - v0, v0);
                        }
                    }
                    xd += sxd; yd += syd; zd += szd;
                }
#pragma unroll
                for (int j = 0; j < 3; j++)
                    if (j < nf) acc[j] = fma(sum[j], dt64, acc[j]);
            } else {
                for (int s = 0; s < n_samples; s++) {
                    float dvs[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        dvs[j] = 0.f;
                        if (j < n_fields) {
                            Corners c;
                            load_corners(c, (const T *)br->data[j] + offset, sx, sy);
                            dvs[j] = trilinear(c, x, y, z);
                        }
                    }
                    eval_transfer32<PARTIAL>(*tf, pairs, dt, dvs, rgba, B);
                    x += sxs; y += sys; z += szs;
                }
            }
        }
    }
};

// lenses.pyx:179-192 (plane-parallel, same operation order) / 196-207 (per-pixel arrays)
__device__ __forceinline__ void make_ray(const DevCamera &cam, int vi, int vj, double (&vp)[3], double (&vd)[3])
{
    if (cam.lens == 0) {
        const double px = dsub(ddiv(dmul(cam.width[0], (double)vi), dsub((double)cam.nx, 1.0)), ddiv(cam.width[0], 2.0));
        const double py = dsub(ddiv(dmul(cam.width[1], (double)vj), dsub((double)cam.ny, 1.0)), ddiv(cam.width[1], 2.0));
#pragma unroll
        for (int i = 0; i < 3; i++) {
            vp[i] = dadd(dadd(dmul(cam.vp_pos[i], px), dmul(cam.vp_pos[3 + i], py)), cam.vp_pos[9 + i]);
            vd[i] = cam.vp_dir[i];
        }
    } else {
        const size_t o = ((size_t)vi * cam.ny + vj) * 3;
#pragma unroll
        for (int i = 0; i < 3; i++) { vp[i] = cam.pos_px[o + i]; vd[i] = cam.dir_px[o + i]; }
    }
}

template <int KIND, bool EXACT, int TFM, bool PARTIAL, bool NF1, typename T>
__global__ void __launch_bounds__(CTA_THREADS, YTB_MIN_BLOCKS)
cast_kernel(const __grid_constant__ CastParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_list[CTA_THREADS];
    __shared__ int s_warp_cnt[CTA_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    CellSampler<KIND, EXACT, TFM, PARTIAL, NF1, T> smp;
    smp.tf = &p.tf;
    smp.pairs = nullptr;
    smp.t64.values = smp.t64.d0 = smp.t64.dy = nullptr;
    // ---- stage the transfer-function tables into shared memory -------------------------------
    if (KIND == KIND_VR) {
        if (EXACT) {
            double *sv = reinterpret_cast<double *>(smem_raw);
            const int n = p.tf.total_bins;
            for (int i = tid; i < n; i += CTA_THREADS) {
                sv[i] = p.tf.values[i]; sv[n + i] = p.tf.d0[i]; sv[2 * n + i] = p.tf.dy[i];
            }
            smp.t64.values = sv; smp.t64.d0 = sv + n; smp.t64.dy = sv + 2 * n;
        } else if (TFM == TFM_COLOUR) {
            // replicate each bin's (r,g,b,a) row COLOUR_COPIES times: copy k sits in 16-byte bank group k
            float4 *sc = reinterpret_cast<float4 *>(smem_raw);
            const int nb = p.tf.nbins[0];
            for (int i = tid; i < nb * COLOUR_COPIES; i += CTA_THREADS) sc[i] = p.tf.colour[i / COLOUR_COPIES];
            smp.ct.base = (unsigned)__cvta_generic_to_shared(sc) + (unsigned)(lane & (COLOUR_COPIES - 1)) * 16u;
            smp.ct.b0_hi = p.tf.b0_hi[0]; smp.ct.b0_lo = p.tf.b0_lo[0]; smp.ct.b1 = p.tf.b1_f[0];
            smp.ct.idbin = p.tf.idbin_f[0]; smp.ct.nbins_m2 = nb - 2;
        } else {
            float2 *sp = reinterpret_cast<float2 *>(smem_raw);
            const int n = p.tf.total_bins;
            for (int i = tid; i < n; i += CTA_THREADS) sp[i] = p.tf.pairs[i];
            smp.pairs = sp;
        }
    }

    // ---- this thread's pixel: warps are 4(i) x 8(j) pixel patches, j fastest (image is C-order) ----
    const int tiles_j = (p.cam.ny + TILE_J - 1) / TILE_J;
    const int ti0 = (blockIdx.x / tiles_j) * TILE_I, tj0 = (blockIdx.x % tiles_j) * TILE_J;
    const int vi = ti0 + (warp >> 1) * 4 + (lane >> 3);
    const int vj = tj0 + (warp & 1) * 8 + (lane & 7);
    const bool active = vi < p.cam.nx && vj < p.cam.ny;
    const size_t pix = (size_t)(active ? vi : 0) * p.cam.ny + (active ? vj : 0);

    double max_t = 1.0;
    double init[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) {
        make_ray(p.cam, vi, vj, smp.vp, smp.vd);
        if (p.zbuffer) { const double z = p.zbuffer[pix]; max_t = fmin(fmax(z, 0.0), 1.0); }
        if (!PARTIAL) {
            const double4 px = *reinterpret_cast<const double4 *>(p.image + pix * 4);
            init[0] = px.x; init[1] = px.y; init[2] = px.z; init[3] = px.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) { smp.vp[i] = 0.0; smp.vd[i] = 0.0; }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        smp.acc[c] = init[c]; smp.B64[c] = 1.0;
        smp.rgba[c] = (float)init[c]; smp.B[c] = 1.f;
    }
    smp.ca.c01 = make_float2((float)init[0], (float)init[1]); smp.ca.c2 = (float)init[2]; smp.ca.c3 = (float)init[3];
    smp.ca.b01 = make_float2(1.f, 1.f); smp.ca.b2 = 1.f; smp.ca.b3 = 1.f;
    smp.n_samples = p.n_samples;
    smp.inv_ns = 1.0 / (double)p.n_samples;
    smp.n_cells = 0;
    __syncthreads();

    // ---- walk the ordered brick list; cull bricks whose pixel footprint misses this tile ----
    const int ti1 = ti0 + TILE_I, tj1 = tj0 + TILE_J;
    for (int base = 0; base < p.n_order; base += CTA_THREADS) {
        const int k = base + tid;
        bool hit = false;
        int b = -1;
        if (k < p.n_order) {
            const int4 e = p.extent[k];
            hit = e.x < ti1 && e.y > ti0 && e.z < tj1 && e.w > tj0;
            b = p.order[k];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CTA_THREADS / 32; w++) {
            const int cnt = s_warp_cnt[w];
            if (w < warp) off += cnt;
            total += cnt;
        }
        if (hit) s_list[off + __popc(m & ((1u << lane) - 1u))] = b;
        __syncthreads();
        if (active) {
            for (int q = 0; q < total; q++) {
                const DevBrick *br = p.bricks + s_list[q];
                BrickGeom g;
                load_geom(g, br);
                smp.set_brick(br);
                walk_brick(g, smp.vp, smp.vd, max_t, smp);
            }
        }
        __syncthreads();
    }

    // ---- write back ---------------------------------------------------------------------------------
    if (!active) return;
    double out[4], Bout[4];
    if (KIND == KIND_VR && !EXACT) {
        if (TFM == TFM_COLOUR) {
            smp.rgba[0] = smp.ca.c01.x; smp.rgba[1] = smp.ca.c01.y; smp.rgba[2] = smp.ca.c2; smp.rgba[3] = smp.ca.c3;
            smp.B[0] = smp.ca.b01.x; smp.B[1] = smp.ca.b01.y; smp.B[2] = smp.ca.b2; smp.B[3] = smp.ca.b3;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) { out[c] = (double)smp.rgba[c]; Bout[c] = (double)smp.B[c]; }
        if (!PARTIAL && !p.tf.grey) out[3] = init[3];  // non-grey never touches alpha (H9)
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) { out[c] = smp.acc[c]; Bout[c] = smp.B64[c]; }
    }
    if (PARTIAL) {
        if (KIND == KIND_VR && !p.tf.grey) { out[3] = 0.0; Bout[3] = 1.0; }
        double4 *dst = reinterpret_cast<double4 *>(p.partial + pix * 8);
        dst[0] = make_double4(out[0], out[1], out[2], out[3]);
        dst[1] = make_double4(Bout[0], Bout[1], Bout[2], Bout[3]);
    } else {
        *reinterpret_cast<double4 *>(p.image + pix * 4) = make_double4(out[0], out[1], out[2], out[3]);
    }
}

// Counting pass: the same traversal, `hit` summed over rays (VR excludes masked cells).
struct CountSampler {
    const unsigned char *mask;
    int ny, nz;
    bool use_mask;
    unsigned long long n;
    __device__ __forceinline__ void cell(double, double, int ix, int iy, int iz)
    {
        if (use_mask && mask != nullptr) {
            if (mask[((size_t)ix * ny + iy) * nz + iz] != 1) return;
        }
        n++;
    }
};

__global__ void __launch_bounds__(CTA_THREADS) count_kernel(const __grid_constant__ CastParams p, int kind)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_j = (p.cam.ny + TILE_J - 1) / TILE_J;
    const int ti0 = (blockIdx.x / tiles_j) * TILE_I, tj0 = (blockIdx.x % tiles_j) * TILE_J;
    const int vi = ti0 + (warp >> 1) * 4 + (lane >> 3);
    const int vj = tj0 + (warp & 1) * 8 + (lane & 7);
    const bool active = vi < p.cam.nx && vj < p.cam.ny;
    CountSampler smp;
    smp.n = 0;
    smp.use_mask = (kind == KIND_VR);
    if (active) {
        double vp[3], vd[3], max_t = 1.0;
        make_ray(p.cam, vi, vj, vp, vd);
        const size_t pix = (size_t)vi * p.cam.ny + vj;
        if (p.zbuffer) max_t = fmin(fmax(p.zbuffer[pix], 0.0), 1.0);
        const int ti1 = ti0 + TILE_I, tj1 = tj0 + TILE_J;
        for (int k = 0; k < p.n_order; k++) {
            const int4 e = p.extent[k];
            if (!(e.x < ti1 && e.y > ti0 && e.z < tj1 && e.w > tj0)) continue;
            const DevBrick *br = p.bricks + p.order[k];
            BrickGeom g;
            load_geom(g, br);
            smp.mask = br->mask; smp.ny = br->dims[1]; smp.nz = br->dims[2];
            walk_brick(g, vp, vd, max_t, smp);
        }
    }
    unsigned long long n = smp.n;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
    if (lane == 0 && n) atomicAdd(p.counters, n);
}

// Conservative pixel footprint of every brick in the order list (replaces calculate_extent_*,
// lenses.pyx:22-175; any conservative footprint is parity-safe because each ray still does its
// own exact slab test -- SURVEY.md H5/H6/H12).
__global__ void extent_kernel(DevCamera cam, const DevBrick *bricks, const int *order, int n_order, int4 *extent)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_order) return;
    int4 e = make_int4(0, cam.nx, 0, cam.ny);
    const DevBrick *b = bricks + order[k];
    if (cam.ext_mode == 1 || cam.ext_mode == 2) {
        double lo_i = 1e300, hi_i = -1e300, lo_j = 1e300, hi_j = -1e300;
        bool ok = true;
        for (int c = 0; c < 8; c++) {
            const double X[3] = {(c & 1) ? b->re[0] : b->le[0], (c & 2) ? b->re[1] : b->le[1], (c & 4) ? b->re[2] : b->le[2]};
            const double r[3] = {X[0] - cam.ext_origin[0], X[1] - cam.ext_origin[1], X[2] - cam.ext_origin[2]};
            double fi, fj;
            if (cam.ext_mode == 1) {
                const double px = cam.ext_rows[0] * r[0] + cam.ext_rows[1] * r[1] + cam.ext_rows[2] * r[2];
                const double py = cam.ext_rows[3] * r[0] + cam.ext_rows[4] * r[1] + cam.ext_rows[5] * r[2];
                fi = (px / cam.width[0] + 0.5) * (cam.nx - 1);
                fj = (py / cam.width[1] + 0.5) * (cam.ny - 1);
            } else {
                const double *w = cam.persp + 3, *ea = cam.persp + 6, *no = cam.persp + 9, *nm = cam.persp + 12;
                const double depth = r[0] * nm[0] + r[1] * nm[1] + r[2] * nm[2];
                const double scale = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                if (!(depth > 1e-6 * scale) || !(depth > 0.0)) { ok = false; break; }
                const double w1 = w[0] * (double)cam.ny / (double)cam.nx;
                const double ex = (r[0] * ea[0] + r[1] * ea[1] + r[2] * ea[2]) / depth * w[2];
                const double ey = (r[0] * no[0] + r[1] * no[1] + r[2] * no[2]) / depth * w[2];
                fi = (ex / w[0] + 0.5) * (cam.nx - 1);
                fj = (ey / w1 + 0.5) * (cam.ny - 1);
            }
            lo_i = fmin(lo_i, fi); hi_i = fmax(hi_i, fi);
            lo_j = fmin(lo_j, fj); hi_j = fmax(hi_j, fj);
        }
        if (ok && hi_i - lo_i < 1e9 && hi_j - lo_j < 1e9) {
            const double big = 2.0e9;
            lo_i = fmax(lo_i, -big); lo_j = fmax(lo_j, -big); hi_i = fmin(hi_i, big); hi_j = fmin(hi_j, big);
            e.x = max(0, (int)floor(lo_i) - 1);
            e.y = min(cam.nx, (int)ceil(hi_i) + 2);
            e.z = max(0, (int)floor(lo_j) - 1);
            e.w = min(cam.ny, (int)ceil(hi_j) + 2);
        }
    }
    extent[k] = e;
}

// ---- affine partial images -----------------------------------------------------
__global__ void compose_kernel(const double4 *earlier, const double4 *later, double4 *out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 A1 = earlier[2 * i], B1 = earlier[2 * i + 1];
    const double4 A2 = later[2 * i], B2 = later[2 * i + 1];
    out[2 * i] = make_double4(fma(B2.x, A1.x, A2.x), fma(B2.y, A1.y, A2.y), fma(B2.z, A1.z, A2.z), fma(B2.w, A1.w, A2.w));
    out[2 * i + 1] = make_double4(B2.x * B1.x, B2.y * B1.y, B2.z * B1.z, B2.w * B1.w);
}

__global__ void apply_kernel(const double4 *partial, double4 *image, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 A = partial[2 * i], B = partial[2 * i + 1];
    const double4 c = image[i];
    image[i] = make_double4(fma(B.x, c.x, A.x), fma(B.y, c.y, A.y), fma(B.z, c.z, A.z), fma(B.w, c.w, A.w));
}

__global__ void f64_to_f32_kernel(const double *src, float *dst, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

__global__ void fill_f64_kernel(double *dst, double v, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

}  // namespace ytb
