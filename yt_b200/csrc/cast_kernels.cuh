// The cast kernels: one CTA per 16x16-pixel image tile, one thread per ray, walking the
// view-ordered brick list (the reference's `for brick in traverse(): sampler(brick)` loop,
// render_source.py:594-597 + image_samplers.pyx:146-197, fused into one launch).
//
// Because a pixel is owned by exactly one thread for the whole brick list, the reference's
// compositing order (bricks in kd order, cells in increasing t inside a brick -- SURVEY.md
// finding 4 / H14) is preserved by construction and there are no races.
#pragma once
#include <type_traits>

#include "raycast.cuh"

namespace ytb {

#ifndef YTB_MIN_BLOCKS
#define YTB_MIN_BLOCKS 2
#endif
#ifndef YTB_MIN_BLOCKS_PROJ
#define YTB_MIN_BLOCKS_PROJ 3   // the projection cast is load-latency bound: occupancy pays (ncu r01c)
#endif
#ifndef YTB_MIN_BLOCKS_IPROJ
#define YTB_MIN_BLOCKS_IPROJ 3   // (single field, fp32 storage: the corner-reuse walk) the interpolated projection's cell body is too short to hide its corner loads (ncu r02y: 40 % of
                                 // the stall samples wait on the prefetched face): a third resident CTA at 80 registers, 17.8 -> 15.8 ms
#endif
#ifndef YTB_PROJ_DEPTH
#define YTB_PROJ_DEPTH 8
#endif
constexpr int PROJ_DEPTH = YTB_PROJ_DEPTH;   // cells of one ray in flight in the fused projection walk
enum { KIND_PROJ = 0, KIND_IPROJ = 1, KIND_VR = 2 };
// transfer-function mode of the FAST path.  The ColorTransferFunction fast path is specialised on
// grey_opacity and on "0 lies strictly inside the table bounds" (both uniform per launch).
//   TFM_COLOUR | GREY         grey_opacity
//   TFM_COLOUR | ZERO_INSIDE  b0 < 0 < b1: cells whose 8 corners are exactly 0 are rejected (FIT_get_value, isnormal)
//   TFM_COLOUR | EDGE         the table does not vanish at its bounds: samples within fp32 rounding of a bound
//                             are flagged and the brick is re-walked carefully (see CarefulSampler)
enum { TFM_NONE = 0, TFM_GENERAL = 1, TFM_COLOUR = 8, TFM_GREY = 1, TFM_ZERO_INSIDE = 2, TFM_EDGE = 4 };
__host__ __device__ constexpr bool tfm_is_colour(int tfm) { return tfm >= TFM_COLOUR; }
__host__ __device__ constexpr bool tfm_grey(int tfm) { return tfm >= TFM_COLOUR && (tfm & TFM_GREY); }
__host__ __device__ constexpr bool tfm_zero_inside(int tfm) { return tfm >= TFM_COLOUR && (tfm & TFM_ZERO_INSIDE); }
__host__ __device__ constexpr bool tfm_edge(int tfm) { return tfm >= TFM_COLOUR && (tfm & TFM_EDGE); }

// ---------------------------------------------------------------------------------------------
// EXACT sampler: the reference's evaluation order, all fp64, no contraction
// (image_samplers.pyx:327-343 / 372-402 / 453-487).  Validation mode, not the tuned path.
// ---------------------------------------------------------------------------------------------
template <int KIND, bool PARTIAL>
struct ExactSampler {
    const double *ray;   // shared memory: this CTA's rays, component-major [6][CTA_THREADS] (vp then vd)
    double vp[3], vd[3];
    const DevBrick *br;
    double dds[3], idds[3], le[3];
    int ny, nz, n_fields;
    const unsigned char *mask;
    int n_samples;
    const DevTF *tf;
    TFShared64 t64;
    int cur[3], stp[3];
    double acc[4], B64[4];

    __device__ __forceinline__ void set_brick(const DevBrick *b)
    {
        br = b;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dds[i] = b->dds[i]; idds[i] = b->idds[i]; le[i] = b->le[i];
            vp[i] = ray[i * CTA_THREADS + threadIdx.x]; vd[i] = ray[(3 + i) * CTA_THREADS + threadIdx.x];
        }
        ny = b->dims[1]; nz = b->dims[2];
        n_fields = b->n_fields;
        mask = b->mask;
    }
    template <bool MASKED>
    __device__ __forceinline__ void begin(const int (&c)[3], const int (&s)[3])
    {
#pragma unroll
        for (int i = 0; i < 3; i++) { cur[i] = c[i]; stp[i] = s[i]; }
    }
    __device__ __forceinline__ void skipped_cell() {}
    template <bool MASKED>
    __device__ __forceinline__ void advance(bool is0, bool is1, bool is2)
    {
        if (is0) cur[0] += stp[0];
        if (is1) cur[1] += stp[1];
        if (is2) cur[2] += stp[2];
    }
    template <bool MASKED>
    __device__ __forceinline__ void cell(double enter_t, double exit_t)
    {
        const int ix = cur[0], iy = cur[1], iz = cur[2];
        if (KIND == KIND_PROJ) {
            const double dl = dsub(exit_t, enter_t);
            const size_t di = (size_t)ix * br->pitch_x + (size_t)iy * br->pitch_y + iz;
            const int nf = n_fields < 4 ? n_fields : 4;
#pragma unroll
            for (int f = 0; f < 4; f++)
                if (f < nf) acc[f] = dadd(acc[f], dmul(((const double *)br->data[f])[di], dl));
            return;
        }
        if (KIND == KIND_VR && MASKED && !tf->light_on) {   // LightSourceRenderSampler.sample has no mask test (image_samplers.pyx:551-589)
            if (mask[((size_t)ix * ny + iy) * nz + iz] != 1) return;
        }
        const int sy = br->pitch_y, sx = br->pitch_x;
        const size_t offset = (size_t)ix * sx + (size_t)iy * sy + iz;
        double dp[3], ds[3], dvs[6];
        const double dt = ddiv(dsub(exit_t, enter_t), (double)n_samples);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dp[i] = dadd(dmul(dadd(enter_t, dmul(0.5, dt)), vd[i]), vp[i]);
            dp[i] = dsub(dp[i], dadd(dmul((double)cur[i], dds[i]), le[i]));
            dp[i] = dmul(dp[i], idds[i]);
            ds[i] = dmul(dmul(vd[i], idds[i]), dt);
        }
        for (int s = 0; s < n_samples; s++) {
#pragma unroll
            for (int j = 0; j < 6; j++)
                dvs[j] = (j < n_fields) ? offset_interpolate64((const double *)br->data[j] + offset, sx, sy, dp) : 0.0;
            if (KIND == KIND_VR && tf->light_on) {
                double grad[3];
                eval_gradient_t((const double *)br->data[0] + offset, sx, sy, dp, grad);
                eval_transfer64_light<PARTIAL>(*tf, t64, dt, dvs, grad, acc, B64);
            } else if (KIND == KIND_VR) {
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
};

// Rare path of the ColorTransferFunction cast.  FIT_get_value rejects x <= b0, x >= b1 and x == 0
// (field_interpolation_tables.pxd:70-71) and the table need not vanish there, so a sample that lands
// within fp32 rounding of such a threshold must be classified the way the reference classifies it, or
// a whole sample's emission appears / disappears.  The fast sample loop flags such samples ("doubt");
// the kernel then re-walks that brick for that ray with this sampler: sample values and validity tests
// in fp64, in the reference's evaluation order (image_samplers.pyx:472-487, fixed_interpolator.cpp:34-52),
// followed by the fp32 table lookup / compositing of the fast path.  The re-walk sits outside the hot
// loop, so its registers cost the fast path nothing.
template <bool PARTIAL, bool GREY, typename T, bool ERT = false>
struct CarefulSampler {
    const double *ray;
    const DevBrick *br;
    const DevTF *tf;
    unsigned base;
    int n_samples;
    int cur[3], stp[3];
    ColourAcc ca;
    float dt_scale;   // ERT: the reversed ray is parametrised over [0, max_t] of the caller's ray (see cast_kernel)

    template <bool MASKED>
    __device__ __forceinline__ void begin(const int (&c)[3], const int (&s)[3])
    {
#pragma unroll
        for (int i = 0; i < 3; i++) { cur[i] = c[i]; stp[i] = s[i]; }
    }
    __device__ __forceinline__ void skipped_cell() {}
    template <bool MASKED>
    __device__ __forceinline__ void advance(bool is0, bool is1, bool is2)
    {
        if (is0) cur[0] += stp[0];
        if (is1) cur[1] += stp[1];
        if (is2) cur[2] += stp[2];
    }
    template <bool MASKED>
    __device__ __noinline__ void cell(double enter_t, double exit_t)
    {
        const int ny = br->dims[1], nz = br->dims[2];
        if (MASKED) {
            if (br->mask[((size_t)cur[0] * ny + cur[1]) * nz + cur[2]] != 1) return;
        }
        const int sy = br->pitch_y, sx = br->pitch_x;
        const T *d = (const T *)br->data[0] + ((size_t)cur[0] * sx + (size_t)cur[1] * sy + cur[2]);
        const double c000 = (double)__ldg(d), c001 = (double)__ldg(d + 1), c010 = (double)__ldg(d + sy), c011 = (double)__ldg(d + sy + 1);
        const double c100 = (double)__ldg(d + sx), c101 = (double)__ldg(d + sx + 1), c110 = (double)__ldg(d + sx + sy), c111 = (double)__ldg(d + sx + sy + 1);
        const double dt = ddiv(dsub(exit_t, enter_t), (double)n_samples);
        float dtf = (float)((exit_t - enter_t) * (1.0 / (double)n_samples));
        if (ERT) dtf *= dt_scale;
        const double b0 = tf->b0[0], b1 = tf->b1[0], idbin = tf->idbin[0];
        // a valid sample must read a data row: clamp its row coordinate into [2, hi) against fp32 rounding
        const float hi = __int_as_float(__float_as_int(tf->colour_hi) - 1);
        ColourTable ct;
        ct.base = base;
        ct.tf = tf;
        double dp[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double vp = ray[i * CTA_THREADS + threadIdx.x], vd = ray[(3 + i) * CTA_THREADS + threadIdx.x];
            dp[i] = dadd(dmul(dadd(enter_t, dmul(0.5, dt)), vd), vp);
            dp[i] = dsub(dp[i], dadd(dmul((double)cur[i], br->dds[i]), br->le[i]));
            dp[i] = dmul(dp[i], br->idds[i]);
            ds[i] = dmul(dmul(vd, br->idds[i]), dt);
        }
        for (int s = 0; s < n_samples; s++) {
            double dv = dsub(1.0, dp[2]);
            double v0 = dadd(dmul(dv, c000), dmul(dp[2], c001)), v1 = dadd(dmul(dv, c010), dmul(dp[2], c011));
            double v2 = dadd(dmul(dv, c100), dmul(dp[2], c101)), v3 = dadd(dmul(dv, c110), dmul(dp[2], c111));
            dv = dsub(1.0, dp[1]);
            v0 = dadd(dmul(dv, v0), dmul(dp[1], v1));
            v1 = dadd(dmul(dv, v2), dmul(dp[1], v3));
            dv = dsub(1.0, dp[0]);
            const double xv = dadd(dmul(dv, v0), dmul(dp[0], v1));
            const bool valid = !(xv >= b1 || xv <= b0) && is_normal64(xv);
            float fs = 0.5f;
            if (valid) fs = fminf(fmaxf((float)fma(xv - b0, idbin, 2.0), 2.0f), hi);
            colour_sample<PARTIAL, GREY, ERT>(ct, fs, dtf, ca);
#pragma unroll
            for (int i = 0; i < 3; i++) dp[i] = dadd(dp[i], ds[i]);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// FAST sampler (the product path): fp64 cell entry, fp32 samples, fp64 sums for the
// projection family.  It follows the walk with running offsets instead of index arithmetic:
//   off    element offset of the current cell (cell-centred data: PROJ) or of its low vertex
//   moff   cell offset into the mask (VR with a non-trivial mask only)
//   ob[i]  ray origin in cell units minus the current cell index, fp64: the cell-local coordinate of
//          the point at parameter t is fma(t, dc[i], ob[i])
// NF1: the brick set has exactly one field (corners cached in registers per cell).
// ---------------------------------------------------------------------------------------------
template <int KIND, int TFM, bool PARTIAL, bool NF1, typename T, bool ERT = false>
struct FastSampler {
    const double *ray;   // shared memory: this CTA's rays, component-major [6][CTA_THREADS] (vp then vd)
    const DevBrick *br;
    float dt_scale;      // ERT: max_t of the caller's ray (the reversed ray runs over [0, 1] of [0, max_t])
    unsigned cells;      // ERT: cells this ray sampled (the rest of the algorithmic count was skipped)
    double ob[3], dc[3];
    float dcf[3];
    unsigned off, moff;
    int s0, s1, s2;      // signed element strides of one step along each axis
    int sx, sy;          // vertex strides (VR / IPROJ)
    int n_fields;
    const unsigned char *mask;
    int n_samples;
    const CastParams *prm;   // kernel parameters (constant bank)
    const DevTF *tf;
    const float2 *pairs;
    ColourTable ct;
    ColourAcc ca;
    bool doubt;      // colour path: a sample of this brick needs the careful re-walk
    double acc[4];   // PROJ / IPROJ sums
    float rgba[4];   // general-TF VR
    float B[4];

    __device__ __forceinline__ void set_brick(const DevBrick *b)
    {
        br = b;
        n_fields = b->n_fields;
        mask = b->mask;
    }
    template <bool MASKED>
    __device__ __forceinline__ void begin(const int (&c)[3], const int (&s)[3])
    {
        const int ny = br->dims[1], nz = br->dims[2];
        sx = br->pitch_x; sy = br->pitch_y;
        if (KIND != KIND_PROJ) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double idds = br->idds[i];
                dc[i] = ray[(3 + i) * CTA_THREADS + threadIdx.x] * idds;
                ob[i] = (ray[i * CTA_THREADS + threadIdx.x] - br->le[i]) * idds - (double)c[i];
                dcf[i] = (float)dc[i];
            }
        }
        off = (unsigned)c[0] * (unsigned)sx + (unsigned)c[1] * (unsigned)sy + (unsigned)c[2];
        s0 = s[0] * sx; s1 = s[1] * sy; s2 = s[2];
        moff = 0;
        if (KIND == KIND_VR && MASKED) moff = ((unsigned)c[0] * (unsigned)ny + (unsigned)c[1]) * (unsigned)nz + (unsigned)c[2];
    }
    __device__ __forceinline__ void skipped_cell() {}
    template <bool MASKED>
    __device__ __forceinline__ void advance(bool is0, bool is1, bool is2)
    {
        off += (unsigned)(is0 ? s0 : (is1 ? s1 : s2));
        if (KIND != KIND_PROJ) {
            // ob[i] -= step[i]; the step has the sign of the ray direction on that axis
            if (is0) ob[0] -= copysign(1.0, dc[0]);
            if (is1) ob[1] -= copysign(1.0, dc[1]);
            if (is2) ob[2] -= copysign(1.0, dc[2]);
        }
        if (KIND == KIND_VR && MASKED) {
            const int ny = br->dims[1], nz = br->dims[2];
            const int m0 = s0 > 0 ? ny * nz : -ny * nz, m1 = s1 > 0 ? nz : -nz;
            moff += (unsigned)(is0 ? m0 : (is1 ? m1 : s2));
        }
    }

    // ------------------------------------------------------------------------------------------------
    // The fused walk of the ColorTransferFunction path (the north-star hot loop).  Same cells, same
    // (enter_t, exit_t) as walk_brick; what differs is the bookkeeping around them (ncu r01c: 133 of the
    // 445 instructions per cell were not samples, and 20 % of the stall cycles waited for corner loads):
    //   * one PTX block per cell does the axis choice, exit_t, the sample spacing and the step with
    //     PREDICATED adds (nvcc turns `if (is0) t0 += d0` into an add plus two selects);
    //   * CORNER REUSE: consecutive cells of a ray share a face, so a step needs 4 new corner loads, not 8 (the
    //     cast sits on the L1 LSU data pipe and the 8 corner LDG were a third of its wavefronts).  The cell is kept
    //     in RAY-ORIENTED corner order -- corner (i,j,k) is the vertex at origin + i*S0 + j*S1 + k*S2 with
    //     S_a = step_a * stride_a (entry side 0, exit side 1 on every axis) -- so a step along axis a is always
    //     "exit face becomes entry face, load the new exit face" whatever the sign of the direction: 12 selects;
    //   * the cell coordinates are mirrored with the corners (x' = 1 - x on an axis the ray descends), and need no
    //     running origin: the ray leaves the cell through the exit face of a stepping axis at t = tmax_axis, so
    //     the first sample sits at 1 + (t_mid - tmax_axis) * |dc|; the cancellation t_mid - tmax is done in fp64
    //     (H13), the rest in fp32;
    //   * the 4 corners of the NEXT cell's exit face are requested before the current cell is sampled, so their
    //     latency hides behind ~300 sample instructions.
    // r01zi: 44.6 -> 43.8 ms on the north-star frame against the walk that loaded all 8 corners of every cell.
    // ------------------------------------------------------------------------------------------------
    template <bool MASKED>
    __device__ __forceinline__ void walk_colour(const WalkState &w)
    {
        const int ny = br->dims[1], nz = br->dims[2];
        const int vsy = br->pitch_y, vsx = (ny + 1) * vsy;
        float dcx[3], bse[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double idds = br->idds[i];
            const double vdi = ray[(3 + i) * CTA_THREADS + threadIdx.x];
            dcx[i] = fabsf((float)(vdi * idds));
            if (w.step[i] != 0) bse[i] = 1.f;
            else bse[i] = (float)((ray[i * CTA_THREADS + threadIdx.x] - br->le[i]) * idds - (double)w.cur[i]);
        }
        // ray-oriented origin of the start cell and signed strides (an axis the ray does not move along keeps +stride)
        const int S0 = w.step[0] < 0 ? -vsx : vsx, S1 = w.step[1] < 0 ? -vsy : vsy, S2 = w.step[2] < 0 ? -1 : 1;
        unsigned o = (unsigned)(w.cur[0] + (w.step[0] < 0)) * (unsigned)vsx + (unsigned)(w.cur[1] + (w.step[1] < 0)) * (unsigned)vsy +
                     (unsigned)(w.cur[2] + (w.step[2] < 0));
        unsigned mo = 0;
        int ms0 = 0, ms1 = 0;
        if (MASKED) {
            mo = ((unsigned)w.cur[0] * (unsigned)ny + (unsigned)w.cur[1]) * (unsigned)nz + (unsigned)w.cur[2];
            ms0 = w.step[0] * ny * nz; ms1 = w.step[1] * nz;
        }
        double t0 = w.step[0] == 0 ? 1e30 : w.t0, t1 = w.step[1] == 0 ? 1e30 : w.t1, t2 = w.step[2] == 0 ? 1e30 : w.t2;
        const double d0 = w.d0, d1 = w.d1, d2 = w.d2, max_t = w.max_t, inv_ns = prm->inv_ns;
        int rem0 = w.rem0, rem1 = w.rem1, rem2 = w.rem2;
        double enter_t = w.enter_t;
        const T *data = (const T *)br->data[0];
        // c[i][j][k] in ray-oriented order, named as RawCorners names them: a=c00k, b=c01k, e=c10k, f=c11k
        float a0, a1, b0, b1, e0, e1, f0, f1;
        {
            const T *d = data + o;
            a0 = (float)__ldg(d); a1 = (float)__ldg(d + S2);
            b0 = (float)__ldg(d + S1); b1 = (float)__ldg(d + S1 + S2);
            e0 = (float)__ldg(d + S0); e1 = (float)__ldg(d + S0 + S2);
            f0 = (float)__ldg(d + S0 + S1); f1 = (float)__ldg(d + S0 + S1 + S2);
        }
        unsigned cont;
        do {
            const unsigned mo_cur = mo;
            float dt, q0, q1, q2;
            unsigned axis, fo;
            int U, V;
            asm volatile(
                "{\n\t"
                ".reg .pred a01, a02, a12, n01, p0, p1, p2, pc, pm;\n\t"
                ".reg .f64 tm, ex, dl, dt, tmid, q;\n\t"
                ".reg .s32 ds, rr;\n\t"
                "setp.lt.f64 a01, %0, %1;\n\t"
                "setp.lt.f64 a02, %0, %2;\n\t"
                "setp.lt.f64 a12, %1, %2;\n\t"
                "and.pred p0, a01, a02;\n\t"
                "not.pred n01, a01;\n\t"
                "and.pred p1, n01, a12;\n\t"
                "or.pred p2, p0, p1;\n\t"            // p2 == NOT (axis 2)
                "selp.f64 tm, %1, %2, p1;\n\t"
                "selp.f64 tm, %0, tm, p0;\n\t"
                "setp.lt.f64 pm, tm, %20;\n\t"
                "selp.f64 ex, tm, %20, pm;\n\t"
                "sub.rn.f64 dl, ex, %7;\n\t"
                "mul.rn.f64 dt, dl, %21;\n\t"
                "fma.rn.f64 tmid, dt, 0d3FE0000000000000, %7;\n\t"
                "cvt.rn.f32.f64 %8, dt;\n\t"
                "sub.rn.f64 q, tmid, %0;\n\t"
                "cvt.rn.f32.f64 %9, q;\n\t"
                "sub.rn.f64 q, tmid, %1;\n\t"
                "cvt.rn.f32.f64 %10, q;\n\t"
                "sub.rn.f64 q, tmid, %2;\n\t"
                "cvt.rn.f32.f64 %11, q;\n\t"
                "@p0 add.rn.f64 %0, %0, %17;\n\t"
                "@p1 add.rn.f64 %1, %1, %18;\n\t"
                "@!p2 add.rn.f64 %2, %2, %19;\n\t"
                "@p0 add.s32 %3, %3, -1;\n\t"
                "@p1 add.s32 %4, %4, -1;\n\t"
                "@!p2 add.s32 %5, %5, -1;\n\t"
                "selp.b32 ds, %23, %24, p1;\n\t"
                "selp.b32 ds, %22, ds, p0;\n\t"
                "add.s32 %6, %6, ds;\n\t"
                "add.s32 %14, %6, ds;\n\t"          // origin of the new cell's exit face
                "selp.b32 %15, %24, %23, p2;\n\t"   // in-face strides: (S2,S1) axis 0, (S2,S0) axis 1, (S1,S0) axis 2
                "selp.b32 %16, %23, %22, p0;\n\t"
                "selp.u32 %13, 1, 2, p1;\n\t"
                "selp.u32 %13, 0, %13, p0;\n\t"
                "or.b32 rr, %3, %4;\n\t"
                "or.b32 rr, rr, %5;\n\t"
                "setp.ge.s32 pc, rr, 0;\n\t"
                "setp.ge.f64 pm, tm, %20;\n\t"
                "not.pred pm, pm;\n\t"
                "and.pred pc, pc, pm;\n\t"
                "selp.u32 %12, 1, 0, pc;\n\t"
                "mov.f64 %7, tm;\n\t"
                "}"
                : "+d"(t0), "+d"(t1), "+d"(t2), "+r"(rem0), "+r"(rem1), "+r"(rem2), "+r"(o), "+d"(enter_t),
                  "=f"(dt), "=f"(q0), "=f"(q1), "=f"(q2), "=r"(cont), "=r"(axis), "=r"(fo), "=r"(U), "=r"(V)
                : "d"(d0), "d"(d1), "d"(d2), "d"(max_t), "d"(inv_ns), "r"(S0), "r"(S1), "r"(S2));
            if (MASKED) mo += (unsigned)(axis == 0 ? ms0 : (axis == 1 ? ms1 : w.step[2]));
            // request the 4 corners of the next cell's exit face now; they are consumed one iteration later
            float L0, L1, L2, L3;
            if (cont) {
                L0 = (float)__ldg(data + fo); L1 = (float)__ldg(data + (fo + (unsigned)U));
                L2 = (float)__ldg(data + (fo + (unsigned)V)); L3 = (float)__ldg(data + (fo + (unsigned)U + (unsigned)V));
            }
            if constexpr (KIND == KIND_IPROJ) {
                // InterpolatedProjectionSampler (image_samplers.pyx:372-402), fp32 storage: along the ray the trilinear
                // interpolant is a cubic in the sample index, so the n_samples midpoint samples of the cell are summed in
                // closed form (sum of s^k, k = 0..3) from its four coefficients -- in fp32, like every other sample of the
                // FAST precision; the cell sums are accumulated in fp64 (rtol 1e-6 bar; measured 1e-7, test_parity_gpu)
                const float x = fmaf(q0, dcx[0], bse[0]), y = fmaf(q1, dcx[1], bse[1]), z = fmaf(q2, dcx[2], bse[2]);
                const float sxs = dcx[0] * dt, sys = dcx[1] * dt, szs = dcx[2] * dt;
                const float2 p0 = make_float2(a0, e0), dz0 = make_float2(a1 - a0, e1 - e0);   // z-edges at y0: (x0 | x1)
                const float2 p1 = make_float2(b0, f0), dz1 = make_float2(b1 - b0, f1 - f0);   // z-edges at y1
                const float2 neg1 = make_float2(-1.f, -1.f);
                const float2 za = fma2(bcast2(z), dz0, p0), zs = mul2(bcast2(szs), dz0);
                const float2 zb = fma2(bcast2(z), dz1, p1), zt = mul2(bcast2(szs), dz1);
                const float2 g0 = fma2(za, neg1, zb), g1 = fma2(zs, neg1, zt);
                const float2 v0 = fma2(bcast2(y), g0, za);
                const float2 v1 = fma2(bcast2(sys), g0, fma2(bcast2(y), g1, zs));
                const float2 v2 = mul2(bcast2(sys), g1);
                const float h0 = v0.y - v0.x, h1 = v1.y - v1.x, h2 = v2.y - v2.x;
                const float k0 = fmaf(x, h0, v0.x);
                const float k1 = fmaf(sxs, h0, fmaf(x, h1, v1.x));
                const float k2 = fmaf(sxs, h1, fmaf(x, h2, v2.x));
                const float k3 = sxs * h2;
                const float sum = fmaf(k3, prm->iproj_s3, fmaf(k2, prm->iproj_s2, fmaf(k1, prm->iproj_s1, k0 * prm->iproj_s0)));
                acc[0] += (double)(sum * dt);
            }
            bool sample = KIND == KIND_VR;
            if (MASKED) sample = mask[mo_cur] == 1;
            if (sample) {
                RawCorners cur_c;
                cur_c.a0 = a0; cur_c.a1 = a1; cur_c.b0 = b0; cur_c.b1 = b1; cur_c.e0 = e0; cur_c.e1 = e1; cur_c.f0 = f0; cur_c.f1 = f1;
                CornersP cp;
                const bool all_zero = rows_from_raw(cp, cur_c, ct);
                if (!(tfm_zero_inside(TFM) && all_zero)) {
                    const float x = fmaf(q0, dcx[0], bse[0]), y = fmaf(q1, dcx[1], bse[1]), z = fmaf(q2, dcx[2], bse[2]);
                    const float sxs = dcx[0] * dt, sys = dcx[1] * dt, szs = dcx[2] * dt;
                    float m;
                    const float dte = ERT ? dt * dt_scale : dt;   // dt of the caller's ray: what the emission is weighted with
                    const bool clamp = !(dte * tf->colour_clamp_bound <= 1.f);
                    if (tf->colour_line) {
                        if (!clamp) m = colour_cell_run<PARTIAL, tfm_grey(TFM), false, tfm_edge(TFM), true, ERT, true>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                        else m = colour_cell_run<PARTIAL, tfm_grey(TFM), true, tfm_edge(TFM), true, ERT, true>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                    } else if (tf->colour_delta) {
                        if (!clamp) m = colour_cell_run<PARTIAL, tfm_grey(TFM), false, tfm_edge(TFM), true, ERT>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                        else m = colour_cell_run<PARTIAL, tfm_grey(TFM), true, tfm_edge(TFM), true, ERT>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                    } else {
                        if (!clamp) m = colour_cell_run<PARTIAL, tfm_grey(TFM), false, tfm_edge(TFM), false, ERT>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                        else m = colour_cell_run<PARTIAL, tfm_grey(TFM), true, tfm_edge(TFM), false, ERT>(cp, ct, n_samples, x, y, z, sxs, sys, szs, dte, ca);
                    }
                    if (tfm_edge(TFM)) doubt = doubt || (m < tf->colour_band);
                    if (ERT) {
                        // everything still to come is seen through `tb`: below the threshold the ray is finished
                        cells++;
                        if (ca.tb < prm->ert_eps) cont = 0u;
                    }
                }
            }
            // the exit face of the step axis becomes the entry face; the loaded face is the new exit face
            const bool x0 = axis == 0, x1 = axis == 1, x2 = axis == 2;
            const float n_a0 = x0 ? e0 : (x1 ? b0 : a1);
            const float n_e0 = x0 ? L0 : (x1 ? f0 : e1);
            const float n_b0 = x0 ? f0 : (x1 ? L0 : b1);
            const float n_a1 = x0 ? e1 : (x1 ? b1 : L0);
            const float n_f0 = x2 ? f1 : L2;
            const float n_e1 = x0 ? L1 : (x1 ? f1 : L2);
            const float n_b1 = x0 ? f1 : L1;
            a0 = n_a0; e0 = n_e0; b0 = n_b0; a1 = n_a1; f0 = n_f0; e1 = n_e1; b1 = n_b1; f1 = L3;
        } while (cont);
    }

    // ------------------------------------------------------------------------------------------------
    // The fused walk of the single-field InterpolatedProjectionSampler (image_samplers.pyx:372-402 inside
    // grid_traversal.pyx:119-138): the step block of walk_colour with fp64 outputs, the next cell's 8 corners
    // requested before the current cell is integrated, and the cell integrated in closed form (see cell()).
    // The first sample's cell-local coordinate along a stepping axis is base + (t_mid - tmax_axis) * dc, base = 1
    // (stepping up) or 0 (stepping down): the ray leaves the cell through that face at t = tmax_axis.
    // ------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void walk_iproj(const WalkState &w)
    {
        const int ny = br->dims[1];
        const int vsy = br->pitch_y, vsx = (ny + 1) * vsy;
        double dcx[3], bse[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double idds = br->idds[i];
            dcx[i] = ray[(3 + i) * CTA_THREADS + threadIdx.x] * idds;
            if (w.step[i] > 0) bse[i] = 1.0;
            else if (w.step[i] < 0) bse[i] = 0.0;
            else bse[i] = (ray[i * CTA_THREADS + threadIdx.x] - br->le[i]) * idds - (double)w.cur[i];
        }
        unsigned o = (unsigned)w.cur[0] * (unsigned)vsx + (unsigned)w.cur[1] * (unsigned)vsy + (unsigned)w.cur[2];
        const int st0 = w.step[0] * vsx, st1 = w.step[1] * vsy, st2 = w.step[2];
        double t0 = w.step[0] == 0 ? 1e30 : w.t0, t1 = w.step[1] == 0 ? 1e30 : w.t1, t2 = w.step[2] == 0 ? 1e30 : w.t2;
        const double d0 = w.d0, d1 = w.d1, d2 = w.d2, max_t = w.max_t, inv_ns = prm->inv_ns;
        int rem0 = w.rem0, rem1 = w.rem1, rem2 = w.rem2;
        double enter_t = w.enter_t;
        const T *data = (const T *)br->data[0];
        const double nn = (double)n_samples;
        const double S1 = 0.5 * nn * (nn - 1.0), S2 = nn * (nn - 1.0) * (2.0 * nn - 1.0) * (1.0 / 6.0), S3 = S1 * S1;
        double a = acc[0];
        T c[8], nx[8];
        auto load8 = [&](T (&v)[8], const T *d) {
            v[0] = __ldg(d); v[1] = __ldg(d + 1); v[2] = __ldg(d + vsy); v[3] = __ldg(d + vsy + 1);
            v[4] = __ldg(d + vsx); v[5] = __ldg(d + vsx + 1); v[6] = __ldg(d + vsx + vsy); v[7] = __ldg(d + vsx + vsy + 1);
        };
        load8(c, data + o);
        unsigned cont;
        do {
            double dt, q0, q1, q2;
            asm volatile(
                "{\n\t"
                ".reg .pred a01, a02, a12, n01, p0, p1, p2, pc, pm;\n\t"
                ".reg .f64 tm, ex, dl, tmid;\n\t"
                ".reg .s32 ds, rr;\n\t"
                "setp.lt.f64 a01, %0, %1;\n\t"
                "setp.lt.f64 a02, %0, %2;\n\t"
                "setp.lt.f64 a12, %1, %2;\n\t"
                "and.pred p0, a01, a02;\n\t"
                "not.pred n01, a01;\n\t"
                "and.pred p1, n01, a12;\n\t"
                "or.pred p2, p0, p1;\n\t"
                "selp.f64 tm, %1, %2, p1;\n\t"
                "selp.f64 tm, %0, tm, p0;\n\t"
                "setp.lt.f64 pm, tm, %16;\n\t"
                "selp.f64 ex, tm, %16, pm;\n\t"
                "sub.rn.f64 dl, ex, %7;\n\t"
                "mul.rn.f64 %8, dl, %17;\n\t"
                "fma.rn.f64 tmid, %8, 0d3FE0000000000000, %7;\n\t"
                "sub.rn.f64 %9, tmid, %0;\n\t"
                "sub.rn.f64 %10, tmid, %1;\n\t"
                "sub.rn.f64 %11, tmid, %2;\n\t"
                "@p0 add.rn.f64 %0, %0, %13;\n\t"
                "@p1 add.rn.f64 %1, %1, %14;\n\t"
                "@!p2 add.rn.f64 %2, %2, %15;\n\t"
                "@p0 add.s32 %3, %3, -1;\n\t"
                "@p1 add.s32 %4, %4, -1;\n\t"
                "@!p2 add.s32 %5, %5, -1;\n\t"
                "selp.b32 ds, %19, %20, p1;\n\t"
                "selp.b32 ds, %18, ds, p0;\n\t"
                "add.s32 %6, %6, ds;\n\t"
                "or.b32 rr, %3, %4;\n\t"
                "or.b32 rr, rr, %5;\n\t"
                "setp.ge.s32 pc, rr, 0;\n\t"
                "setp.ge.f64 pm, tm, %16;\n\t"
                "not.pred pm, pm;\n\t"
                "and.pred pc, pc, pm;\n\t"
                "selp.u32 %12, 1, 0, pc;\n\t"
                "mov.f64 %7, tm;\n\t"
                "}"
                : "+d"(t0), "+d"(t1), "+d"(t2), "+r"(rem0), "+r"(rem1), "+r"(rem2), "+r"(o), "+d"(enter_t),
                  "=&d"(dt), "=&d"(q0), "=&d"(q1), "=&d"(q2), "=&r"(cont)
                : "d"(d0), "d"(d1), "d"(d2), "d"(max_t), "d"(inv_ns), "r"(st0), "r"(st1), "r"(st2));
            if (cont) load8(nx, data + o);
            {
                const double xd = fma(q0, dcx[0], bse[0]), yd = fma(q1, dcx[1], bse[1]), zd = fma(q2, dcx[2], bse[2]);
                const double sxd = dcx[0] * dt, syd = dcx[1] * dt, szd = dcx[2] * dt;
                const double a0 = (double)c[0], a1 = (double)c[1], b0 = (double)c[2], b1 = (double)c[3];
                const double e0 = (double)c[4], e1 = (double)c[5], f0 = (double)c[6], f1 = (double)c[7];
                const double pa = fma(zd, a1 - a0, a0), qa = szd * (a1 - a0);
                const double pb = fma(zd, b1 - b0, b0), qb = szd * (b1 - b0);
                const double pe = fma(zd, e1 - e0, e0), qe = szd * (e1 - e0);
                const double pf = fma(zd, f1 - f0, f0), qf = szd * (f1 - f0);
                const double d0p = pb - pa, d0q = qb - qa, d1p = pf - pe, d1q = qf - qe;
                const double u00 = fma(yd, d0p, pa), u01 = fma(syd, d0p, fma(yd, d0q, qa)), u02 = syd * d0q;
                const double u10 = fma(yd, d1p, pe), u11 = fma(syd, d1p, fma(yd, d1q, qe)), u12 = syd * d1q;
                const double g0 = u10 - u00, g1 = u11 - u01, g2 = u12 - u02;
                const double k0 = fma(xd, g0, u00);
                const double k1 = fma(sxd, g0, fma(xd, g1, u01));
                const double k2 = fma(sxd, g1, fma(xd, g2, u02));
                const double k3 = sxd * g2;
                const double sum = fma(k3, S3, fma(k2, S2, fma(k1, S1, k0 * nn)));
                a = fma(sum, dt, a);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = nx[k];
        } while (cont);
        acc[0] = a;
    }

    // ------------------------------------------------------------------------------------------------
    // The fused walk of the single-field ProjectionSampler (image_samplers.pyx:327-343 inside
    // grid_traversal.pyx:119-138).  Same cells and the same (enter_t, exit_t) as walk_brick.  ncu r01e on
    // the generic loop: 75 instructions per cell, the data pointer re-read from the brick table for every
    // cell and the value consumed right after its load (long_scoreboard 6.2 of 9.7 stall cycles per issue).
    // Here the pointer lives in a register, the step is one PTX block with predicated adds, and the NEXT
    // cell's value is requested before the current one is accumulated.
    // ------------------------------------------------------------------------------------------------
    template <int NF, int DEPTH>
    __device__ __forceinline__ void walk_proj(const WalkState &w)
    {
        const int px = br->pitch_x, py = br->pitch_y;
        unsigned o = (unsigned)w.cur[0] * (unsigned)px + (unsigned)w.cur[1] * (unsigned)py + (unsigned)w.cur[2];
        const int st0 = w.step[0] * px, st1 = w.step[1] * py, st2 = w.step[2];
        double t0 = w.t0, t1 = w.t1, t2 = w.t2;
        const double d0 = w.d0, d1 = w.d1, d2 = w.d2, max_t = w.max_t;
        int rem0 = w.rem0, rem1 = w.rem1, rem2 = w.rem2;
        double enter_t = w.enter_t;
        const T *data[NF];
        double a[NF];
#pragma unroll
        for (int f = 0; f < NF; f++) { data[f] = (const T *)br->data[f]; a[f] = acc[f]; }
        // The walk runs DEPTH cells AHEAD of the accumulation: a ring of (value, path length) slots keeps
        // that many loads of one ray in flight (r01g: with one load in flight a warp paid a full L2/DRAM
        // latency, ~620 cycles, per cell).  `live` goes false when the reference's loop would have ended
        // (it never comes back: the exit tests are monotone along the walk); dead slots hold 0 * 0.
        T v[DEPTH][NF];
        double dl[DEPTH];
        unsigned live = 1u;   // walk_setup returned true: the start cell is sampled
        auto produce = [&](T (&vj)[NF], double &dlj) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                vj[f] = T(0);
                if (live) vj[f] = __ldg(data[f] + o);
            }
            double len;
            unsigned cont;
            asm volatile(
                "{\n\t"
                ".reg .pred a01, a02, a12, n01, p0, p1, p2, pc, pm;\n\t"
                ".reg .f64 tm, ex;\n\t"
                ".reg .s32 ds, rr;\n\t"
                // axis of the smallest tmax, reference tie rule (grid_traversal.pyx:122-131)
                "setp.lt.f64 a01, %0, %1;\n\t"
                "setp.lt.f64 a02, %0, %2;\n\t"
                "setp.lt.f64 a12, %1, %2;\n\t"
                "and.pred p0, a01, a02;\n\t"
                "not.pred n01, a01;\n\t"
                "and.pred p1, n01, a12;\n\t"
                "or.pred p2, p0, p1;\n\t"            // p2 == NOT (axis 2)
                "selp.f64 tm, %1, %2, p1;\n\t"
                "selp.f64 tm, %0, tm, p0;\n\t"
                "setp.lt.f64 pm, tm, %13;\n\t"
                "selp.f64 ex, tm, %13, pm;\n\t"       // exit_t = tm < max_t ? tm : max_t
                "sub.rn.f64 %8, ex, %7;\n\t"          // path length in t through this cell
                // the step: cur[i] += step[i]; enter_t = tmax[i]; tmax[i] += tdelta[i]
                "@p0 add.rn.f64 %0, %0, %10;\n\t"
                "@p1 add.rn.f64 %1, %1, %11;\n\t"
                "@!p2 add.rn.f64 %2, %2, %12;\n\t"
                "@p0 add.s32 %3, %3, -1;\n\t"
                "@p1 add.s32 %4, %4, -1;\n\t"
                "@!p2 add.s32 %5, %5, -1;\n\t"
                "selp.b32 ds, %15, %16, p1;\n\t"
                "selp.b32 ds, %14, ds, p0;\n\t"
                "add.s32 %6, %6, ds;\n\t"
                // go on while the stepped index is inside the brick and enter_t < max_t (NaN goes on, like the reference)
                "or.b32 rr, %3, %4;\n\t"
                "or.b32 rr, rr, %5;\n\t"
                "setp.ge.s32 pc, rr, 0;\n\t"
                "setp.ge.f64 pm, tm, %13;\n\t"
                "not.pred pm, pm;\n\t"
                "and.pred pc, pc, pm;\n\t"
                "selp.u32 %9, 1, 0, pc;\n\t"
                "mov.f64 %7, tm;\n\t"
                "}"
                : "+d"(t0), "+d"(t1), "+d"(t2), "+r"(rem0), "+r"(rem1), "+r"(rem2), "+r"(o), "+d"(enter_t),
                  "=d"(len), "=r"(cont)
                : "d"(d0), "d"(d1), "d"(d2), "d"(max_t), "r"(st0), "r"(st1), "r"(st2));
            dlj = live ? len : 0.0;
            live &= cont;
        };
        auto consume = [&](const T (&vj)[NF], double dlj) {
#pragma unroll
            for (int f = 0; f < NF; f++) a[f] = fma((double)vj[f], dlj, a[f]);
        };
#pragma unroll
        for (int j = 0; j < DEPTH; j++) produce(v[j], dl[j]);
        while (live) {
#pragma unroll
            for (int j = 0; j < DEPTH; j++) {
                consume(v[j], dl[j]);
                produce(v[j], dl[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < DEPTH; j++) consume(v[j], dl[j]);
#pragma unroll
        for (int f = 0; f < NF; f++) acc[f] = a[f];
    }

    template <bool MASKED>
    __device__ __forceinline__ void cell(double enter_t, double exit_t)
    {
        if (KIND == KIND_PROJ) {
            const double dl = exit_t - enter_t;
            const int nf = n_fields < 4 ? n_fields : 4;
#pragma unroll
            for (int f = 0; f < 4; f++)
                if (f < nf) acc[f] = fma((double)__ldg((const T *)br->data[f] + off), dl, acc[f]);
            return;
        }
        if (KIND == KIND_VR && MASKED && !(!tfm_is_colour(TFM) && tf->light_on)) {
            if (mask[moff] != 1) return;
        }
        const double dt64 = (exit_t - enter_t) * prm->inv_ns;
        const double tmid = fma(0.5, dt64, enter_t);
        const double xd = fma(tmid, dc[0], ob[0]), yd = fma(tmid, dc[1], ob[1]), zd = fma(tmid, dc[2], ob[2]);
        if (KIND == KIND_IPROJ) {
            // The projection family carries the 1e-6 bar: fp64.  Along the ray the trilinear interpolant
            // (fixed_interpolator.cpp:34-52) is a cubic in the sample index s, so the n_samples midpoint samples
            // of image_samplers.pyx:393-401 sum in closed form: sum_s (c0 + c1 s + c2 s^2 + c3 s^3) =
            // n c0 + c1 S1 + c2 S2 + c3 S3 with the power sums of 0..n-1 -- O(1) per cell instead of O(n_samples),
            // equal to the sample loop up to fp64 rounding.
            const double sxd = dc[0] * dt64, syd = dc[1] * dt64, szd = dc[2] * dt64;
            const double nn = (double)n_samples;
            const double S1 = 0.5 * nn * (nn - 1.0), S2 = nn * (nn - 1.0) * (2.0 * nn - 1.0) * (1.0 / 6.0), S3 = S1 * S1;
            const int nf = NF1 ? 1 : (n_fields < 3 ? n_fields : 3);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (j < nf) {
                    const T *d = (const T *)br->data[j] + off;
                    const double a0 = (double)__ldg(d), a1 = (double)__ldg(d + 1), b0 = (double)__ldg(d + sy), b1 = (double)__ldg(d + sy + 1);
                    const double e0 = (double)__ldg(d + sx), e1 = (double)__ldg(d + sx + 1), f0 = (double)__ldg(d + sx + sy), f1 = (double)__ldg(d + sx + sy + 1);
                    // z-edges, linear in s: v = p + q s
                    const double pa = fma(zd, a1 - a0, a0), qa = szd * (a1 - a0);
                    const double pb = fma(zd, b1 - b0, b0), qb = szd * (b1 - b0);
                    const double pe = fma(zd, e1 - e0, e0), qe = szd * (e1 - e0);
                    const double pf = fma(zd, f1 - f0, f0), qf = szd * (f1 - f0);
                    // faces x0 | x1, quadratic in s: v = u0 + u1 s + u2 s^2  (y-lerp of two z-edges)
                    const double d0p = pb - pa, d0q = qb - qa, d1p = pf - pe, d1q = qf - qe;
                    const double u00 = fma(yd, d0p, pa), u01 = fma(syd, d0p, fma(yd, d0q, qa)), u02 = syd * d0q;
                    const double u10 = fma(yd, d1p, pe), u11 = fma(syd, d1p, fma(yd, d1q, qe)), u12 = syd * d1q;
                    // x-lerp: cubic
                    const double g0 = u10 - u00, g1 = u11 - u01, g2 = u12 - u02;
                    const double c0 = fma(xd, g0, u00);
                    const double c1 = fma(sxd, g0, fma(xd, g1, u01));
                    const double c2 = fma(sxd, g1, fma(xd, g2, u02));
                    const double c3 = sxd * g2;
                    const double sum = fma(c3, S3, fma(c2, S2, fma(c1, S1, c0 * nn)));
                    acc[j] = fma(sum, dt64, acc[j]);
                }
            }
            return;
        }
        float x = (float)xd, y = (float)yd, z = (float)zd;
        const float dt = (float)dt64;
        const float sxs = dcf[0] * dt, sys = dcf[1] * dt, szs = dcf[2] * dt;
        if (tfm_is_colour(TFM)) {
            // the ColorTransferFunction path has its own fused walk (walk_colour)
        } else if (KIND == KIND_VR && tf->light_on) {
            // LightSourceRenderSampler (image_samplers.pyx:551-589): positions and the gradient of field 0 in fp64, table
            // lookups and the lit composite in fp32.  General, not tuned (the reference reaches it from old_camera only).
            double dp[3] = {xd, yd, zd};
            const double st[3] = {dc[0] * dt64, dc[1] * dt64, dc[2] * dt64};
            // field 0's cell in registers: its value and the six gradient probes of every sample read these 8 corners
            double c0[8];
            {
                const T *d = (const T *)br->data[0] + off;
                c0[0] = (double)__ldg(d); c0[1] = (double)__ldg(d + 1); c0[2] = (double)__ldg(d + sy); c0[3] = (double)__ldg(d + sy + 1);
                c0[4] = (double)__ldg(d + sx); c0[5] = (double)__ldg(d + sx + 1); c0[6] = (double)__ldg(d + sx + sy); c0[7] = (double)__ldg(d + sx + sy + 1);
            }
            for (int s = 0; s < n_samples; s++) {
                float dvs[6];
                dvs[0] = (float)cell_interpolate64(c0, dp);
#pragma unroll
                for (int j = 1; j < 6; j++)
                    dvs[j] = j < n_fields ? (float)offset_interpolate_t((const T *)br->data[j] + off, sx, sy, dp) : 0.f;
                double grad[3];
                cell_gradient64(c0, dp, grad);
                eval_transfer32_light<PARTIAL>(*tf, pairs, dt, dvs, grad, rgba, B);
#pragma unroll
                for (int i = 0; i < 3; i++) dp[i] += st[i];
            }
        } else if (NF1) {
            Corners c;
            load_corners(c, (const T *)br->data[0] + off, sx, sy);
            float dvs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < n_samples; s++) {
                dvs[0] = trilinear(c, x, y, z);
                eval_transfer32<PARTIAL>(*tf, pairs, dt, dvs, rgba, B);
                x += sxs; y += sys; z += szs;
            }
        } else {
            // any number of fields: corners re-read (L1) per sample; general but not the tuned path
            for (int s = 0; s < n_samples; s++) {
                float dvs[6];
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    dvs[j] = 0.f;
                    if (j < n_fields) {
                        Corners c;
                        load_corners(c, (const T *)br->data[j] + off, sx, sy);
                        dvs[j] = trilinear(c, x, y, z);
                    }
                }
                eval_transfer32<PARTIAL>(*tf, pairs, dt, dvs, rgba, B);
                x += sxs; y += sys; z += szs;
            }
        }
    }
};

// walk_volume for the ColorTransferFunction fast path: reference entry logic, then the fused walk
template <bool HAVE_IV = false, class S>
__device__ __forceinline__ void cast_colour(const BrickGeom &g, const double (&vp)[3], const double (&vd)[3], double max_t, S &smp, bool masked,
                                            const double *ivp = nullptr)
{
    WalkState w;
    if (!walk_setup<HAVE_IV>(g, vp, vd, max_t, smp, w, ivp)) return;
    if (masked) smp.template walk_colour<true>(w);
    else smp.template walk_colour<false>(w);
}

template <bool EXACT, class A, class B> struct pick_sampler { using type = A; };
template <class A, class B> struct pick_sampler<false, A, B> { using type = B; };

// MANY: instantiation for long brick lists (AMR: thousands of small bricks per frame; ncu r01z6: 1800 thread-instructions
// of set-up per (ray, brick) against ~4900 of sampling).  It skips, warp-wide, the bricks whose pixel footprint misses the
// warp's 4x8 patch, and computes 1 / v_dir once per ray.  A separate instantiation so that the single-brick kernel of the
// north-star frame keeps its register allocation (an in-place attempt cost it 13 %).
// ERT (ert.cu; grey-opacity ColorTransferFunction casts only): EARLY RAY TERMINATION.  The reference composites a ray's
// samples in walk order with C <- e + a C (field_interpolation_tables.pxd:103-110), so the samples processed LAST weigh
// most and everything before them is seen through the product of their a's.  The ERT cast therefore runs the composition
// from the other end: the host hands the order list REVERSED, every ray is reversed in place (origin vp + max_t vd,
// direction -max_t vd: the same cells and the same sample points, met in the opposite order), the affine maps are
// applied front to back (A <- A + B e, B <- B a) and a ray stops as soon as B < ert_eps -- what is left
// can change the pixel by at most B times the brightest colour the rest could reach.  A warp whose rays are all
// finished skips the remaining bricks (one vote per brick), a CTA whose warps are all finished leaves the list.
template <int KIND, bool EXACT, int TFM, bool PARTIAL, bool NF1, typename T, bool MANY = false, bool ERT = false>
__global__ void __launch_bounds__(CTA_THREADS, (KIND == KIND_PROJ && !EXACT) ? YTB_MIN_BLOCKS_PROJ
                                            : ((KIND == KIND_IPROJ && !EXACT && NF1 && std::is_same<T, float>::value) ? YTB_MIN_BLOCKS_IPROJ : YTB_MIN_BLOCKS))
cast_kernel(const __grid_constant__ CastParams p)
{
    static_assert(!ERT || (KIND == KIND_VR && !EXACT && tfm_grey(TFM) && NF1), "early termination: grey-opacity colour casts only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // MANY: the tile's WHOLE culled list is gathered before any brick is walked (LIST_CAP entries at a time), so the warps
    // of a CTA meet at a barrier once per ~1000 list entries instead of once per 256 scanned bricks (ncu r02m, cfg4: 58
    // barrier pairs per CTA with 3-4 bricks of uneven work between them; `barrier` was 9 % of the stall samples)
    constexpr int LIST_CAP = MANY ? 4 * CTA_THREADS : CTA_THREADS;
    __shared__ int s_list[LIST_CAP];
    __shared__ int s_warp_cnt[2][CTA_THREADS / 32];
    // the rays of this tile, component-major: only needed when a ray enters a brick, so they live here
    // instead of in 12 registers per thread
    __shared__ double s_ray[(MANY ? 9 : 6) * CTA_THREADS];
    __shared__ int4 s_ext[MANY ? LIST_CAP : 1];
    __shared__ __align__(8) unsigned long long s_table_bar;   // mbarrier of the transfer-table bulk copy

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- this thread's pixel: warps are 4(i) x 8(j) pixel patches, j fastest (image is C-order) ----
    const int tiles_j = (p.cam.ny + TILE_J - 1) / TILE_J;
    const int ti0 = (blockIdx.x / tiles_j) * TILE_I, tj0 = (blockIdx.x % tiles_j) * TILE_J;
    const int vi = ti0 + (warp >> 1) * 4 + (lane >> 3);
    const int vj = tj0 + (warp & 1) * 8 + (lane & 7);
    const bool active = vi < p.cam.nx && vj < p.cam.ny;
    const size_t pix = (size_t)(active ? vi : 0) * p.cam.ny + (active ? vj : 0);

    if (!tile_has_bricks(p.extent, p.chunk_extent, p.n_order, ti0, ti0 + TILE_I, tj0, tj0 + TILE_J)) {
        // nothing of the brick list reaches this tile: the image keeps its pixels; a partial gets the identity map
        if (PARTIAL) {
            if (active) {
                const double zero[4] = {0.0, 0.0, 0.0, 0.0}, one[4] = {1.0, 1.0, 1.0, 1.0};
                store_partial(p.partial, p.partial_fmt, pix, zero, one);
            }
            if (p.tail_flag != nullptr) {
                SwapSignal sg;
                sg.n = 1; sg.flag[0] = p.tail_flag;
                grid_tail_signal(p.tail_counter, sg, p.tail_seq);
            }
        }
        return;
    }


    typename pick_sampler<EXACT, ExactSampler<KIND, PARTIAL>, FastSampler<KIND, TFM, PARTIAL, NF1, T, ERT>>::type smp;
    smp.tf = &p.tf;
    smp.ray = s_ray;
    // ---- stage the transfer-function tables into shared memory -------------------------------
    if constexpr (EXACT) {
        smp.t64.values = smp.t64.d0 = smp.t64.dy = nullptr;
        if (KIND == KIND_VR) {
            double *sv = reinterpret_cast<double *>(smem_raw);
            const int n = p.tf.total_bins;
            for (int i = tid; i < n; i += CTA_THREADS) {
                sv[i] = p.tf.values[i]; sv[n + i] = p.tf.d0[i]; sv[2 * n + i] = p.tf.dy[i];
            }
            smp.t64.values = sv; smp.t64.d0 = sv + n; smp.t64.dy = sv + 2 * n;
        }
    } else {
        smp.pairs = nullptr;
        if (KIND == KIND_VR && tfm_is_colour(TFM)) {
            // replicate each row COLOUR_COPIES times: copy k sits in 16-byte bank group k (rows are already
            // padded with the zero / duplicate rows by ytb200_tf_create)
            // (delta layout: row = [8 copies of the values | 8 copies of the increments].)  The replicated image is
            // kept in global memory (ytb200_tf_create) and arrives by ONE bulk async copy issued by thread 0 -- the TMA
            // engine moves the 64 KiB while the CTA generates its rays and reads its pixels; every thread waits on the
            // mbarrier right before the walk.
            float4 *sc = reinterpret_cast<float4 *>(smem_raw);
            if (tid == 0) {
                const unsigned bytes = ((unsigned)p.tf.colour_rows * COLOUR_COPIES * (unsigned)sizeof(float4)) * (p.tf.colour_delta ? 2u : 1u);
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_table_bar);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"((unsigned)__cvta_generic_to_shared(sc)), "l"(p.tf.colour_image), "r"(bytes), "r"(bar) : "memory");
            }
            // colour_cell_run forms the address as float_as_int(2^23 + row) * ROW_BYTES + base: fold the
            // (wrapping) bias of the exponent bits into base
            smp.ct.base = (unsigned)__cvta_generic_to_shared(sc) + (unsigned)(lane & (COLOUR_COPIES - 1)) * 16u
                          - 0x4B000000u * (p.tf.colour_delta ? COLOUR_DELTA_ROW_BYTES : COLOUR_ROW_BYTES);
            smp.ct.tf = &p.tf;
        } else if (KIND == KIND_VR) {
            float2 *sp = reinterpret_cast<float2 *>(smem_raw);
            const int n = p.tf.total_bins;
            for (int i = tid; i < n; i += CTA_THREADS) sp[i] = p.tf.pairs[i];
            smp.pairs = sp;
        }
    }

    double max_t = 1.0;
    {
        double vp[3] = {0.0, 0.0, 0.0}, vd[3] = {0.0, 0.0, 0.0};
        double init[4] = {0.0, 0.0, 0.0, 0.0};   // initial colours (opaque sources); not kept live through the walk
        if (active) {
            make_ray(p.cam, vi, vj, vp, vd);
            if (p.zbuffer) { const double z = p.zbuffer[pix]; max_t = fmin(fmax(z, 0.0), 1.0); }
            if (!PARTIAL && !ERT) {
                const double4 px = *reinterpret_cast<const double4 *>(p.image + pix * 4);
                init[0] = px.x; init[1] = px.y; init[2] = px.z; init[3] = px.w;
            }
        }
        if constexpr (ERT) {
            // the reversed ray: t' in [0, 1] is t = max_t (1 - t') of the caller's ray; a cell's dt' times max_t is its dt
#pragma unroll
            for (int i = 0; i < 3; i++) { vp[i] = fma(max_t, vd[i], vp[i]); vd[i] = -(max_t * vd[i]); }
            smp.dt_scale = (float)max_t;
            smp.cells = 0u;
            max_t = 1.0;
        }
#pragma unroll
        for (int i = 0; i < 3; i++) { s_ray[i * CTA_THREADS + tid] = vp[i]; s_ray[(3 + i) * CTA_THREADS + tid] = vd[i]; }
        if constexpr (MANY) {
#pragma unroll
            for (int i = 0; i < 3; i++) s_ray[(6 + i) * CTA_THREADS + tid] = vd[i] == 0.0 ? 0.0 : ddiv(1.0, vd[i]);   // grid_traversal.pyx:60
        }
        if constexpr (EXACT) {
#pragma unroll
            for (int c = 0; c < 4; c++) { smp.acc[c] = init[c]; smp.B64[c] = 1.0; }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) { smp.acc[c] = init[c]; smp.rgba[c] = (float)init[c]; smp.B[c] = 1.f; }
            smp.ca.c01 = make_float2((float)init[0], (float)init[1]); smp.ca.c2 = (float)init[2]; smp.ca.c3 = (float)init[3];
            smp.ca.fired = 0u;
            smp.ca.tb = 1.f;   // ERT: nothing composited yet (its accumulators start from zero: init[] stays 0)
            smp.prm = &p;
        }
    }
    smp.n_samples = p.n_samples;
    __syncthreads();
    if constexpr (!EXACT && KIND == KIND_VR && tfm_is_colour(TFM)) {
        // the table image has landed (phase 0 of the mbarrier completes when the copy's bytes have)
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_table_bar);
        asm volatile("{\n\t.reg .pred p;\n\tTABLE_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra TABLE_DONE;\n\t"
                     "bra TABLE_WAIT;\n\tTABLE_DONE:\n\t}" :: "r"(bar) : "memory");
    }

    // ---- walk the ordered brick list; cull bricks whose pixel footprint misses this tile ----
    const int ti1 = ti0 + TILE_I, tj1 = tj0 + TILE_J;
    const int wi0 = ti0 + (warp >> 1) * 4, wj0 = tj0 + (warp & 1) * 8;   // this warp's 4x8 pixel patch
    const unsigned wmask = __ballot_sync(0xffffffffu, active);             // the lanes of this warp that own a pixel
    for (int base = 0, chunk = 0; base < p.n_order;) {
        // ordered compaction of the next chunk(s) of the brick list: one chunk unless MANY, else until the list is full.
        // s_warp_cnt is double-buffered so that a chunk costs one barrier (a warp rewrites a buffer two chunks later,
        // after every warp has passed the barrier of the chunk in between).
        int total = 0;
        do {
            // the union footprint of these CTA_THREADS consecutive bricks (kd order: a compact subtree) misses the tile:
            // nothing to scan (CTA-uniform, so the barrier below stays collective)
            const int4 ce = p.chunk_extent[base / CTA_THREADS];
            if (ce.x < ti1 && ce.y > ti0 && ce.z < tj1 && ce.w > tj0) {
                const int k = base + tid;
                bool hit = false;
                int b = -1;
                int4 e = make_int4(0, 0, 0, 0);
                if (k < p.n_order) {
                    e = p.extent[k];
                    hit = e.x < ti1 && e.y > ti0 && e.z < tj1 && e.w > tj0;
                    b = p.order[k];
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_warp_cnt[chunk & 1][warp] = __popc(m);
                __syncthreads();
                int off = 0, ctotal = 0;
#pragma unroll
                for (int w = 0; w < CTA_THREADS / 32; w++) {
                    const int cnt = s_warp_cnt[chunk & 1][w];
                    if (w < warp) off += cnt;
                    ctotal += cnt;
                }
                if (hit) {
                    const int slot = total + off + __popc(m & ((1u << lane) - 1u));
                    s_list[slot] = b;
                    if constexpr (MANY) s_ext[slot] = e;
                }
                total += ctotal;
                chunk++;
            }
            base += CTA_THREADS;
        } while (MANY && base < p.n_order && total + CTA_THREADS <= LIST_CAP);
        __syncthreads();
        if (active) {
            for (int q = 0; q < total; q++) {
                if constexpr (ERT) {
                    // the vote: bricks are skipped by the warp once every ray of it is finished (lanes of a warp that are
                    // off the image never get here, hence the active mask)
                    const bool live = smp.ca.tb >= p.ert_eps;
                    if (!__any_sync(wmask, live)) break;
                    if (!live) continue;
                }
                if constexpr (MANY) {
                    const int4 we = s_ext[q];   // warp-uniform: the footprint misses every pixel of this warp
                    if (!(we.x < wi0 + 4 && we.y > wi0 && we.z < wj0 + 8 && we.w > wj0)) continue;
                }
                const DevBrick *br = p.bricks + s_list[q];
                BrickGeom g;
                load_geom(g, br);
                smp.set_brick(br);
                double vp[3], vd[3];
#pragma unroll
                for (int i = 0; i < 3; i++) { vp[i] = s_ray[i * CTA_THREADS + tid]; vd[i] = s_ray[(3 + i) * CTA_THREADS + tid]; }
                const bool masked = KIND == KIND_VR && br->mask != nullptr;
                if constexpr (!EXACT && KIND == KIND_VR && tfm_edge(TFM)) {
                    // accumulators at brick entry, parked in local memory for the (rare) careful re-walk
                    volatile float entry[8];
                    entry[0] = smp.ca.c01.x; entry[1] = smp.ca.c01.y; entry[2] = smp.ca.c2; entry[3] = smp.ca.c3;
                    if (PARTIAL) entry[4] = __uint_as_float(smp.ca.fired);
                    if (ERT) entry[5] = smp.ca.tb;
                    smp.doubt = false;
                    cast_colour(g, vp, vd, max_t, smp, masked);
                    if (smp.doubt) {
                        CarefulSampler<PARTIAL, tfm_grey(TFM), T, ERT> cs;
                        cs.ray = s_ray; cs.br = br; cs.tf = &p.tf; cs.base = smp.ct.base; cs.n_samples = p.n_samples;
                        cs.ca.c01 = make_float2(entry[0], entry[1]); cs.ca.c2 = entry[2]; cs.ca.c3 = entry[3];
                        cs.ca.fired = 0u;
                        cs.ca.tb = 1.f;
                        if (PARTIAL) cs.ca.fired = __float_as_uint(entry[4]);
                        if (ERT) { cs.ca.tb = entry[5]; cs.dt_scale = smp.dt_scale; }
                        // reload geometry and ray so that nothing of them stays live across the fast walk
                        BrickGeom g2;
                        load_geom(g2, br);
                        double vp2[3], vd2[3];
#pragma unroll
                        for (int i = 0; i < 3; i++) { vp2[i] = s_ray[i * CTA_THREADS + tid]; vd2[i] = s_ray[(3 + i) * CTA_THREADS + tid]; }
                        if (masked) walk_brick<true>(g2, vp2, vd2, max_t, cs);
                        else walk_brick<false>(g2, vp2, vd2, max_t, cs);
                        smp.ca = cs.ca;
                    }
                } else if constexpr (!EXACT && KIND == KIND_VR && tfm_is_colour(TFM)) {
                    if constexpr (MANY) {
                        const double iv[3] = {s_ray[6 * CTA_THREADS + tid], s_ray[7 * CTA_THREADS + tid], s_ray[8 * CTA_THREADS + tid]};
                        cast_colour<true>(g, vp, vd, max_t, smp, masked, iv);
                    } else {
                        cast_colour(g, vp, vd, max_t, smp, masked);
                    }
                } else if constexpr (!EXACT && KIND == KIND_IPROJ && NF1) {
                    WalkState w;
                    if (walk_setup(g, vp, vd, max_t, smp, w)) {
                        if constexpr (std::is_same<T, float>::value) smp.template walk_colour<false>(w);   // corner-reuse walk, fp32 cell sums
                        else smp.walk_iproj(w);                                                          // fp64 storage: fp64 closed form
                    }
                } else if constexpr (!EXACT && KIND == KIND_PROJ) {
                    // image_samplers.pyx:337-343 sums the first min(4, n_fields) fields
                    WalkState w;
                    if (walk_setup(g, vp, vd, max_t, smp, w)) {
                        if constexpr (NF1) smp.template walk_proj<1, PROJ_DEPTH>(w);
                        else {
                            const int nf = br->n_fields;
                            if (nf == 2) smp.template walk_proj<2, 4>(w);
                            else if (nf == 3) smp.template walk_proj<3, 3>(w);
                            else smp.template walk_proj<4, 3>(w);
                        }
                    }
                } else {
                    if (masked) walk_brick<true>(g, vp, vd, max_t, smp);
                    else walk_brick<false>(g, vp, vd, max_t, smp);
                }
            }
        }
        if constexpr (ERT) {
            if (__syncthreads_and(!active || smp.ca.tb < p.ert_eps)) break;   // every ray of the tile is finished
        } else {
            __syncthreads();
        }
    }

    // ---- write back ---------------------------------------------------------------------------------
    if (active) {
    double out[4], Bout[4];
    if constexpr (EXACT) {
#pragma unroll
        for (int c = 0; c < 4; c++) { out[c] = smp.acc[c]; Bout[c] = smp.B64[c]; }
    } else if (ERT) {
        // (A, B) of everything composited, front to back; B counts as 0 once the ray was cut (what it hides stays hidden)
        const float b = smp.ca.tb;
        out[0] = (double)smp.ca.c01.x; out[1] = (double)smp.ca.c01.y; out[2] = (double)smp.ca.c2; out[3] = (double)smp.ca.c3;
        Bout[0] = Bout[1] = Bout[2] = Bout[3] = (double)b;
        if (!PARTIAL) {
            const double4 c0 = *reinterpret_cast<const double4 *>(p.image + pix * 4);
            if (b == 1.f && smp.ca.c3 == 0.f && smp.ca.c01.x == 0.f && smp.ca.c01.y == 0.f && smp.ca.c2 == 0.f) {
                out[0] = c0.x; out[1] = c0.y; out[2] = c0.z; out[3] = c0.w;   // untouched pixel: bit-identical
            } else {
                out[0] = fma((double)b, c0.x, out[0]); out[1] = fma((double)b, c0.y, out[1]);
                out[2] = fma((double)b, c0.z, out[2]); out[3] = fma((double)b, c0.w, out[3]);
            }
        }
        // cells sampled by this warp's rays
        unsigned n = smp.cells;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) n += __shfl_xor_sync(wmask, n, sft);
        if (p.counters != nullptr && lane == (__ffs(wmask) - 1) && n) atomicAdd(p.counters + 1, (unsigned long long)n);
    } else if (KIND == KIND_VR) {
        if (tfm_is_colour(TFM)) {
            smp.rgba[0] = smp.ca.c01.x; smp.rgba[1] = smp.ca.c01.y; smp.rgba[2] = smp.ca.c2; smp.rgba[3] = smp.ca.c3;
            // transmittance of the partial: B = 1 - A unless a clamp fired for the channel (see ColourAcc)
            const unsigned fz = smp.ca.fired;
            if (tfm_grey(TFM)) {
                const float b = (fz & 1u) ? 0.f : 1.f - smp.ca.c3;
                smp.B[0] = smp.B[1] = smp.B[2] = smp.B[3] = b;
            } else {
                smp.B[0] = (fz & 1u) ? 0.f : 1.f - smp.ca.c01.x;
                smp.B[1] = (fz & 2u) ? 0.f : 1.f - smp.ca.c01.y;
                smp.B[2] = (fz & 4u) ? 0.f : 1.f - smp.ca.c2;
                smp.B[3] = 1.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) { out[c] = (double)smp.rgba[c]; Bout[c] = (double)smp.B[c]; }
        if (!PARTIAL) {
            // A channel whose fp32 accumulator still holds the (rounded) initial colour was not changed by any sample:
            // give it back its fp64 value, so that pixels no ray touches -- and the alpha of a non-grey render (H9) --
            // come back bit-identical, as in the reference where they are never written with anything else.
            const double4 c0 = *reinterpret_cast<const double4 *>(p.image + pix * 4);
            if (smp.rgba[0] == (float)c0.x) out[0] = c0.x;
            if (smp.rgba[1] == (float)c0.y) out[1] = c0.y;
            if (smp.rgba[2] == (float)c0.z) out[2] = c0.z;
            if (smp.rgba[3] == (float)c0.w || !p.tf.grey) out[3] = c0.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) { out[c] = smp.acc[c]; Bout[c] = 1.0; }
    }
    if (PARTIAL) {
        if (KIND == KIND_VR && !p.tf.grey) { out[3] = 0.0; Bout[3] = 1.0; }
        store_partial(p.partial, p.partial_fmt, pix, out, Bout);
    } else {
        *reinterpret_cast<double4 *>(p.image + pix * 4) = make_double4(out[0], out[1], out[2], out[3]);
    }
    }
    if (PARTIAL && p.tail_flag != nullptr) {
        SwapSignal sg;
        sg.n = 1; sg.flag[0] = p.tail_flag;
        grid_tail_signal(p.tail_counter, sg, p.tail_seq);
    }
}

#ifndef YTB_CAST_KERNEL_ONLY   // (ert.cu instantiates cast_kernel only: the kernels below have one home, ytb200.cu)
// Counting pass: the same traversal, `hit` summed over rays (VR excludes masked cells).
struct CountSampler {
    const unsigned char *mask;
    int ny, nz;
    bool use_mask;
    unsigned long long n;
    unsigned moff;
    int m0, m1, m2;
    template <bool MASKED>
    __device__ __forceinline__ void begin(const int (&c)[3], const int (&s)[3])
    {
        moff = ((unsigned)c[0] * (unsigned)ny + (unsigned)c[1]) * (unsigned)nz + (unsigned)c[2];
        m0 = s[0] * ny * nz; m1 = s[1] * nz; m2 = s[2];
    }
    __device__ __forceinline__ void skipped_cell() { n++; }
    template <bool MASKED>
    __device__ __forceinline__ void advance(bool is0, bool is1, bool is2) { moff += (unsigned)(is0 ? m0 : (is1 ? m1 : m2)); }
    template <bool MASKED>
    __device__ __forceinline__ void cell(double, double)
    {
        if (use_mask && mask != nullptr) {
            if (mask[moff] != 1) return;
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
            walk_brick<false>(g, vp, vd, max_t, smp);
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

// Union footprint of every CTA_THREADS consecutive entries of the order list (the chunk a cast CTA scans at a time): in
// kd order such a chunk is a compact subtree, so most chunks of a long list miss a given 16x16 tile and are skipped
// without reading their entries.  An entry with an empty footprint does not widen the union.
__global__ void __launch_bounds__(CTA_THREADS) chunk_extent_kernel(const int4 *extent, int n_order, int4 *chunk_extent)
{
    const int k = blockIdx.x * CTA_THREADS + threadIdx.x;
    int x = 0x7fffffff, y = -0x7fffffff, z = 0x7fffffff, w = -0x7fffffff;
    if (k < n_order) {
        const int4 e = extent[k];
        if (e.x < e.y && e.z < e.w) { x = e.x; y = e.y; z = e.z; w = e.w; }
    }
    x = __reduce_min_sync(0xffffffffu, x); y = __reduce_max_sync(0xffffffffu, y);
    z = __reduce_min_sync(0xffffffffu, z); w = __reduce_max_sync(0xffffffffu, w);
    __shared__ int4 s[CTA_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = make_int4(x, y, z, w);
    __syncthreads();
    if (threadIdx.x == 0) {
        int4 r = s[0];
        for (int i = 1; i < CTA_THREADS / 32; i++) { r.x = min(r.x, s[i].x); r.y = max(r.y, s[i].y); r.z = min(r.z, s[i].z); r.w = max(r.w, s[i].w); }
        chunk_extent[blockIdx.x] = r;
    }
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

// ---- vertex centring on the device ---------------------------------------------
// No-ghost vertex-centred data of a grid patch: grid_patch.py:284-296 calls ghost_zone_interpolate(1, cells,
// [0.5]*3, out, [0]*3) (interpolators.pyx:177-227): with those arguments ids = ods = iids = 1 exactly, the
// vertex g of an axis with n cells reads cells ii = clip(int(g - 0.5), 0, n - 2), ii + 1 with weight
// xp = (g - 0.5) - ii (-0.5 at the low face, 1.5 at the high face: linear extrapolation, 0.5 inside) and the
// output is the 8-term sum of interpolators.pyx:216-224, reproduced here term for term with non-contracting
// fp64 so that fp64 storage is bit-identical to the reference.  Fused behind it: the log10 of
// amr_kdtree.py:346-352 (values <= 0 -> NaN first), the [li:ri+1] slice of :362-365 and the conversion to the
// brick set's storage type -- one pass over HBM instead of three host passes and a 3x host-memory blow-up.
// One vertex by the reference's general formula (any vertex, including the extrapolated faces).
template <typename TIN>
__device__ __forceinline__ double vertex_generic(const TIN *__restrict__ cells, int nx, int ny, int nz, int gx, int gy, int gz)
{
    const int ii = gx == 0 ? 0 : min(gx - 1, nx - 2), ij = gy == 0 ? 0 : min(gy - 1, ny - 2), ik = gz == 0 ? 0 : min(gz - 1, nz - 2);
    const double xp = dsub(dsub((double)gx, 0.5), (double)ii), yp = dsub(dsub((double)gy, 0.5), (double)ij),
                 zp = dsub(dsub((double)gz, 0.5), (double)ik);
    const double xm = dsub(1.0, xp), ym = dsub(1.0, yp), zm = dsub(1.0, zp);
    const size_t sy = (size_t)nz, sx = (size_t)ny * nz;
    const TIN *c = cells + ((size_t)ii * sx + (size_t)ij * sy + (size_t)ik);
    const double c000 = (double)__ldg(c), c001 = (double)__ldg(c + 1), c010 = (double)__ldg(c + sy), c011 = (double)__ldg(c + sy + 1);
    const double c100 = (double)__ldg(c + sx), c101 = (double)__ldg(c + sx + 1), c110 = (double)__ldg(c + sx + sy),
                 c111 = (double)__ldg(c + sx + sy + 1);
    double v = dmul(c000, dmul(dmul(xm, ym), zm));
    v = dadd(v, dmul(c100, dmul(dmul(xp, ym), zm)));
    v = dadd(v, dmul(c010, dmul(dmul(xm, yp), zm)));
    v = dadd(v, dmul(c001, dmul(dmul(xm, ym), zp)));
    v = dadd(v, dmul(c101, dmul(dmul(xp, ym), zp)));
    v = dadd(v, dmul(c011, dmul(dmul(xm, yp), zp)));
    v = dadd(v, dmul(c110, dmul(dmul(xp, yp), zm)));
    v = dadd(v, dmul(c111, dmul(dmul(xp, yp), zp)));
    return v;
}

// The marching kernel.  A thread owns one z-vertex of one (y) row and marches along x, carrying the four cells
// (x-1, y-1 | y, z-1 | z) of the previous plane in registers: 4 coalesced loads per vertex instead of 8, no
// index division, no weight products (the first version of this kernel: 8 LDG, 24 weight products and 64-bit
// div/mod per vertex, 0.77 TB/s).
// Interior vertices have all eight weights equal to 0.125, a power of two, so every product of the reference's
// sum is exact and scaling commutes with rounding: 0.125 * (((c000 + c100) + c010) + ...) in the reference's
// term order IS the reference's result bit for bit (absent subnormal / overflowing intermediates, which are
// sent to the general formula).  Face vertices (extrapolation weights -0.5 / 1.5) take the general formula.
constexpr int VC_THREADS = 128, VC_ZSPAN = VC_THREADS, VC_XCHUNK = 32;

template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(VC_THREADS) vertex_centre_kernel(const TIN *__restrict__ cells, TOUT *__restrict__ out,
                                                                    int nx, int ny, int nz, int lo0, int lo1, int lo2,
                                                                    int m0, int m1, int m2, int take_log, int out_px, int out_py)
{
    const int oz = blockIdx.x * VC_THREADS + threadIdx.x;
    const int oy = blockIdx.y;
    if (oz >= m2) return;
    const int gz = lo2 + oz, gy = lo1 + oy;
    const int ox0 = blockIdx.z * VC_XCHUNK, ox1 = min(m0, ox0 + VC_XCHUNK);
    const bool col_ok = gy >= 1 && gy <= ny - 1 && gz >= 1 && gz <= nz - 1;   // cell rows gy-1, gy and cells gz-1, gz exist
    const size_t sx = (size_t)ny * nz;
    // cell (x = 0, gy - 1, gz - 1) of this column; only dereferenced when col_ok
    const TIN *col = cells + (col_ok ? (size_t)(gy - 1) * nz + (size_t)(gz - 1) : 0);
    TOUT *dst = out + (size_t)ox0 * out_px + (size_t)oy * out_py + oz;
    const size_t dstep = (size_t)out_px;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double a0m = 0.0, a0 = 0.0, b0m = 0.0, b0 = 0.0;
    int gx = lo0 + ox0;
    // the carry starts as the plane below the first interior x of this chunk (gx = 0 itself is a face)
    if (col_ok && max(gx, 1) <= nx - 1) {
        const TIN *p = col + (size_t)(max(gx, 1) - 1) * sx;
        a0m = (double)__ldg(p); a0 = (double)__ldg(p + 1); b0m = (double)__ldg(p + nz); b0 = (double)__ldg(p + nz + 1);
    }
    for (int ox = ox0; ox < ox1; ox++, gx++, dst += dstep) {
        const bool interior = col_ok && gx >= 1 && gx <= nx - 1;
        double v;
        bool fast = false;
        if (interior) {
            const TIN *p = col + (size_t)gx * sx;
            const double a1m = (double)__ldg(p), a1 = (double)__ldg(p + 1), b1m = (double)__ldg(p + nz), b1 = (double)__ldg(p + nz + 1);
            // c000 + c100 + c010 + c001 + c101 + c011 + c110 + c111 (interpolators.pyx:216-224), weights 1/8
            double sum = dadd(a0m, a1m);
            sum = dadd(sum, b0m);
            sum = dadd(sum, a0);
            sum = dadd(sum, a1);
            sum = dadd(sum, b0);
            sum = dadd(sum, b1m);
            sum = dadd(sum, b1);
            const double mag = fabs(sum);
            fast = (mag > 1e-280 && mag < 1e300) || sum == 0.0;
            v = dmul(sum, 0.125);
            a0m = a1m; a0 = a1; b0m = b1m; b0 = b1;
        }
        if (!fast) v = vertex_generic(cells, nx, ny, nz, gx, gy, gz);
        if (take_log) v = (v <= 0.0) ? qnan : log10(v);
        *dst = (TOUT)v;
    }
}

// Ghost-zone vertex centring (grid_patch.py:298-310): `cells` is the grid padded with one layer of ghost zones --
// (nx+2, ny+2, nz+2), built on the host by the smoothed covering grid (yt_b200/loaders.py) -- and every vertex is
// 0.125 * the sum of its 8 cells, added in the reference's order c111, c011, c101, c110, c010, c100, c001, c000
// (first index x) with non-contracting fp64 adds, so that fp64 storage is bit-identical.  Same marching scheme,
// fused log10 / slice / storage conversion as vertex_centre_kernel.
template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(VC_THREADS) vertex_centre_ghost_kernel(const TIN *__restrict__ cells, TOUT *__restrict__ out,
                                                                          int nx, int ny, int nz, int lo0, int lo1, int lo2,
                                                                          int m0, int m1, int m2, int take_log, int out_px, int out_py)
{
    const int oz = blockIdx.x * VC_THREADS + threadIdx.x;
    const int oy = blockIdx.y;
    if (oz >= m2) return;
    const int gz = lo2 + oz, gy = lo1 + oy;
    const int ox0 = blockIdx.z * VC_XCHUNK, ox1 = min(m0, ox0 + VC_XCHUNK);
    const size_t sy = (size_t)nz + 2, sx = ((size_t)ny + 2) * sy;
    (void)nx;
    const TIN *col = cells + (size_t)gy * sy + (size_t)gz;     // padded cell (x = 0, gy, gz): the vertex's low corner
    TOUT *dst = out + (size_t)ox0 * out_px + (size_t)oy * out_py + oz;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    int gx = lo0 + ox0;
    const TIN *p = col + (size_t)gx * sx;
    double c000 = (double)__ldg(p), c001 = (double)__ldg(p + 1), c010 = (double)__ldg(p + sy), c011 = (double)__ldg(p + sy + 1);
    for (int ox = ox0; ox < ox1; ox++, gx++, dst += (size_t)out_px) {
        const TIN *q = col + (size_t)(gx + 1) * sx;
        const double c100 = (double)__ldg(q), c101 = (double)__ldg(q + 1), c110 = (double)__ldg(q + sy), c111 = (double)__ldg(q + sy + 1);
        double v = dadd(0.0, c111);
        v = dadd(v, c011);
        v = dadd(v, c101);
        v = dadd(v, c110);
        v = dadd(v, c010);
        v = dadd(v, c100);
        v = dadd(v, c001);
        v = dadd(v, c000);
        v = dmul(v, 0.125);
        if (take_log) v = (v <= 0.0) ? qnan : log10(v);
        *dst = (TOUT)v;
        c000 = c100; c001 = c101; c010 = c110; c011 = c111;
    }
}

// ---- binary-swap compositing over NVLink peer memory -----------------------------------
// One swap round = ONE kernel: it waits (every CTA spins on a flag word in this rank's own memory) until the partner has
// finished the buffer it is about to read, composes this rank's kept half with the partner's copy of the same pixels --
// read straight out of the partner's HBM through its NVLink peer mapping: no send/recv, no receive buffer -- stores the
// result (locally; in the LAST round into the destination ranks' slabs, or as finished fp64 pixels into the caller's
// page-locked host image) and, when the whole grid has stored, pushes the frame's sequence number into the flag word
// of whoever reads that result next.  A frame is k + 2 launches on the destination rank (cast, k rounds, apply) and
// k + 1 elsewhere, whatever the image size.

// the map of pixel i in compute precision CT (float for the fp32 formats, double otherwise)
template <int FMT> struct pf_compute { using type = double; };
template <> struct pf_compute<PF_AB32> { using type = float; };
template <> struct pf_compute<PF_A32> { using type = float; };
template <> struct pf_compute<PF_A32G> { using type = float; };

template <int FMT, typename CT>
__device__ __forceinline__ void pf_load(const void *__restrict__ buf, long long i, CT (&A)[4], CT (&B)[4])
{
    if constexpr (FMT == PF_AB64) {
        const double4 a = reinterpret_cast<const double4 *>(buf)[2 * i], b = reinterpret_cast<const double4 *>(buf)[2 * i + 1];
        A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = a.w; B[0] = b.x; B[1] = b.y; B[2] = b.z; B[3] = b.w;
    } else if constexpr (FMT == PF_AB32) {
        const float4 a = reinterpret_cast<const float4 *>(buf)[2 * i], b = reinterpret_cast<const float4 *>(buf)[2 * i + 1];
        A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = a.w; B[0] = b.x; B[1] = b.y; B[2] = b.z; B[3] = b.w;
    } else if constexpr (FMT == PF_A32) {
        const float4 a = reinterpret_cast<const float4 *>(buf)[i];
        const unsigned bits = (unsigned)a.w;
        A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = 0.f;
        B[0] = (bits & 1u) ? 0.f : 1.f - a.x; B[1] = (bits & 2u) ? 0.f : 1.f - a.y; B[2] = (bits & 4u) ? 0.f : 1.f - a.z; B[3] = 1.f;
    } else if constexpr (FMT == PF_A32G) {
        const float4 a = reinterpret_cast<const float4 *>(buf)[i];
        const float al = fabsf(a.w);
        const float b = signbit(a.w) ? 0.f : 1.f - al;
        A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = al;
        B[0] = B[1] = B[2] = B[3] = b;
    } else if constexpr (FMT == PF_SUM1) {
        A[0] = reinterpret_cast<const double *>(buf)[i]; A[1] = A[2] = A[3] = 0.0;
        B[0] = B[1] = B[2] = B[3] = 1.0;
    } else if constexpr (FMT == PF_SUM2) {
        const double2 a = reinterpret_cast<const double2 *>(buf)[i];
        A[0] = a.x; A[1] = a.y; A[2] = A[3] = 0.0;
        B[0] = B[1] = B[2] = B[3] = 1.0;
    } else {
        const double4 a = reinterpret_cast<const double4 *>(buf)[i];
        A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = a.w;
        B[0] = B[1] = B[2] = B[3] = 1.0;
    }
}

template <int FMT, typename CT>
__device__ __forceinline__ void pf_store(void *__restrict__ buf, long long i, const CT (&A)[4], const CT (&B)[4])
{
    if constexpr (FMT == PF_AB64) {
        reinterpret_cast<double4 *>(buf)[2 * i] = make_double4(A[0], A[1], A[2], A[3]);
        reinterpret_cast<double4 *>(buf)[2 * i + 1] = make_double4(B[0], B[1], B[2], B[3]);
    } else if constexpr (FMT == PF_AB32) {
        reinterpret_cast<float4 *>(buf)[2 * i] = make_float4(A[0], A[1], A[2], A[3]);
        reinterpret_cast<float4 *>(buf)[2 * i + 1] = make_float4(B[0], B[1], B[2], B[3]);
    } else if constexpr (FMT == PF_A32) {
        const unsigned bits = (B[0] == CT(0) ? 1u : 0u) | (B[1] == CT(0) ? 2u : 0u) | (B[2] == CT(0) ? 4u : 0u);
        reinterpret_cast<float4 *>(buf)[i] = make_float4(A[0], A[1], A[2], (float)bits);
    } else if constexpr (FMT == PF_A32G) {
        const float al = fabsf((float)A[3]);
        reinterpret_cast<float4 *>(buf)[i] = make_float4(A[0], A[1], A[2], B[3] == CT(0) ? -al : al);
    } else if constexpr (FMT == PF_SUM1) {
        reinterpret_cast<double *>(buf)[i] = A[0];
    } else if constexpr (FMT == PF_SUM2) {
        reinterpret_cast<double2 *>(buf)[i] = make_double2(A[0], A[1]);
    } else {
        reinterpret_cast<double4 *>(buf)[i] = make_double4(A[0], A[1], A[2], A[3]);
    }
}

// Every CTA's first `n` threads spin (in this rank's memory) until flag k has reached `seq`; gives up after ~10 s and
// records the failure.  The fence after the barrier orders the reads that follow behind the flag observation.
__device__ __forceinline__ void grid_wait(const unsigned long long *flags, int n, unsigned long long seq, int *error)
{
    if ((int)threadIdx.x < n) {
        const volatile unsigned long long *f = flags + threadIdx.x;
        unsigned long long t0 = 0;
        bool timing = false;
        while (*f < seq) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (!timing) { t0 = t; timing = true; }
            if (t - t0 > 10000000000ULL) { *error = 1; break; }
            __nanosleep(100);
        }
        __threadfence_system();   // the pollers' fence, then the barrier: every thread's reads come after the flag
    }
    __syncthreads();
}

struct SwapOut { void *buf[SWAP_MAX_DST]; int n; };   // destination buffers (own memory and / or peers')

struct SwapRoundArgs {
    const void *mine, *theirs;      // both start at the first kept pixel
    long long n;                    // kept pixels
    int mine_is_earlier;            // whose map is applied first
    const unsigned long long *wait_flag; unsigned long long seq; int *error;   // partner's buffer complete?
    SwapOut out;                    // partial-format destinations (STORE_PARTIAL)
    const double4 *c0; double4 *image;   // APPLY: initial colours and fp64 destination of the strip (both offset to it)
    unsigned *counter; SwapSignal done;  // whom to tell when this grid has stored
};

// APPLY == false: out[k][i] = compose(mine[i], theirs[i]);  APPLY == true: image[i] = compose(...)(c0[i])
template <int FMT, bool APPLY>
__global__ void __launch_bounds__(256) swap_round_kernel(const __grid_constant__ SwapRoundArgs a)
{
    using CT = typename pf_compute<FMT>::type;
    grid_wait(a.wait_flag, 1, a.seq, a.error);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        CT Am[4], Bm[4], At[4], Bt[4], A[4], B[4];
        pf_load<FMT, CT>(a.mine, i, Am, Bm);
        pf_load<FMT, CT>(a.theirs, i, At, Bt);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const CT A1 = a.mine_is_earlier ? Am[c] : At[c], B1 = a.mine_is_earlier ? Bm[c] : Bt[c];   // applied first
            const CT A2 = a.mine_is_earlier ? At[c] : Am[c], B2 = a.mine_is_earlier ? Bt[c] : Bm[c];   // applied second
            A[c] = B2 * A1 + A2;
            B[c] = B2 * B1;
        }
        if constexpr (APPLY) {
            const double4 c = a.c0[i];
            a.image[i] = make_double4(fma((double)B[0], c.x, (double)A[0]), fma((double)B[1], c.y, (double)A[1]),
                                      fma((double)B[2], c.z, (double)A[2]), fma((double)B[3], c.w, (double)A[3]));
        } else {
            for (int k = 0; k < a.out.n; k++) pf_store<FMT, CT>(a.out.buf[k], i, A, B);
        }
    }
    grid_tail_signal(a.counter, a.done, a.seq);
}

struct SwapApplyArgs {
    const void *partial;            // n pixels in format FMT (the composed strips of all ranks, or this rank's own cast)
    const double4 *c0; double4 *image; long long n;
    const unsigned long long *wait_flags; int n_wait; unsigned long long seq; int *error;
};

// image[i] = A[i] + B[i] * c0[i] once every strip of the frame has landed (n_wait flags)
template <int FMT>
__global__ void __launch_bounds__(256) swap_apply_kernel(const __grid_constant__ SwapApplyArgs a)
{
    using CT = typename pf_compute<FMT>::type;
    if (a.n_wait > 0) grid_wait(a.wait_flags, a.n_wait, a.seq, a.error);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        CT A[4], B[4];
        pf_load<FMT, CT>(a.partial, i, A, B);
        const double4 c = a.c0[i];
        a.image[i] = make_double4(fma((double)B[0], c.x, (double)A[0]), fma((double)B[1], c.y, (double)A[1]),
                                  fma((double)B[2], c.z, (double)A[2]), fma((double)B[3], c.w, (double)A[3]));
    }
}

// Spin (locally) until every flag has reached `seq`: the destination rank of a host-image frame, whose stream must not
// complete before every rank's strip has been stored.
__global__ void swap_wait_kernel(const unsigned long long *flags, int n, unsigned long long seq, int *error)
{
    grid_wait(flags, n, seq, error);
}

// One flag word (in a peer's slab) set from a kernel: stream-ordered behind whatever the stream did before (a copy-engine
// transfer of a finished strip, for one).
__global__ void swap_signal_kernel(unsigned long long *flag, unsigned long long seq)
{
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(flag) = seq;
    __threadfence_system();
}

__global__ void f64_to_f32_kernel(const double *src, float *dst, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

// host fp64 rows (contiguous, `row` elements each) -> device rows with pitch `pitch`, converted to TOUT
template <typename TOUT>
__global__ void repitch_kernel(const double *src, TOUT *dst, long long n, int row, int pitch, long long first_row)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = i / row;
    dst[(first_row + r) * pitch + (i - r * row)] = (TOUT)src[i];
}

__global__ void fill_f64_kernel(double *dst, double v, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

#endif  // YTB_CAST_KERNEL_ONLY

}  // namespace ytb
