// Ray/brick traversal and the three cell samplers, as device functions.
//
// What is reproduced (bit-for-bit in the fp64 parts) and where it comes from:
//   * ray/box entry, start cell, Amanatides-Woo stepping, tie rule, max_t clamp
//       -- walk_volume, yt/utilities/lib/grid_traversal.pyx:26-140
//   * ProjectionSampler / InterpolatedProjectionSampler / VolumeRenderSampler cell bodies
//       -- yt/utilities/lib/image_samplers.pyx:327-343, 372-402, 453-487
//   * trilinear vertex interpolation -- fixed_interpolator.cpp:34-52
//   * transfer-function lookup + compositing -- field_interpolation_tables.pxd:66-110
//
// How it is done here is ours: one thread per ray, fp64 DDA written with explicit
// round-to-nearest intrinsics (so nvcc never contracts a*b+c into an FMA and every
// discrete decision -- entry face, start cell, tie -- matches the reference on the same
// inputs), the 8 cell corners cached in registers for the n_samples samples of a cell,
// fp32 sampling/compositing in FAST mode, tables in shared memory.
#pragma once
#include "device_types.cuh"

namespace ytb {

// ---- non-contracting fp64 helpers ----------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct BrickGeom {
    double le[3], re[3], dds[3], idds[3];
    int dims[3];
};

__device__ __forceinline__ void load_geom(BrickGeom &g, const DevBrick *b)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g.le[i] = b->le[i]; g.re[i] = b->re[i]; g.dds[i] = b->dds[i]; g.idds[i] = b->idds[i];
        g.dims[i] = b->dims[i];
    }
}

// Snap a face coordinate onto the brick edge when within 1e-10 of a cell (grid_traversal.pyx:67-74).
__device__ __forceinline__ double snap_edge(double v, double le, double re, double tol)
{
    if (fabs(dsub(v, le)) < tol) return le;
    if (fabs(dsub(v, re)) < tol) return re;
    return v;
}

// walk_volume.  The sampler S is the reference's sampler_function plus the bookkeeping that lets it
// follow the walk without index arithmetic:
//   S::begin<MASKED>(cur, step)        start cell chosen (called once per brick the ray enters)
//   S::cell<MASKED>(enter_t, exit_t)   == sample(vc, v_pos, v_dir, enter_t, exit_t, cur_ind, data)
//   S::skipped_cell()                  the reference would sample an out-of-range cell here (see below)
//   S::advance<MASKED>(is0, is1, is2)  the walk steps along exactly one axis (cur[i] += step[i])
// MASKED: the brick has a non-trivial cell mask (uniform per brick; keeps mask bookkeeping out of the
// common instantiation).
// The stepping loop is branch-free: lanes of a warp step along different axes, so `if (i == 0) ...`
// would serialise three copies of the step (ncu r01a: BSSY/BSYNC around every step).  The axis choice
// keeps the reference's tie rule: `tmax0<tmax1 ? (tmax0<tmax2 ? 0 : 2) : (tmax1<tmax2 ? 1 : 2)`.
// Returns the number of cells visited ("hit").
struct WalkState {
    double t0, t1, t2;      // tmax per axis
    double d0, d1, d2;      // tdelta per axis
    double enter_t, max_t;
    int rem0, rem1, rem2;   // cells left before the walk leaves the brick along each axis
    int cur[3], step[3];    // start cell and step signs
    int hit;
};

// Everything of walk_volume up to the first sampled cell (grid_traversal.pyx:40-117).  Returns false when
// the ray does not sample the brick at all; w.hit then holds the return value of walk_volume.
// HAVE_IV: 1 / v_dir was computed once per ray by the caller (`ivp`, same expression) instead of once per (ray, brick).
template <bool HAVE_IV = false, class S>
__device__ __forceinline__ bool walk_setup(const BrickGeom &g, const double (&vp)[3], const double (&vd)[3],
                                           double max_t, S &smp, WalkState &w, const double *ivp = nullptr)
{
    w.hit = 0;
    if (max_t > 1.0) max_t = 1.0;
    double intersect_t = 1.1;
    int direction = -1;
    if (g.le[0] <= vp[0] && vp[0] < g.re[0] && g.le[1] <= vp[1] && vp[1] < g.re[1] &&
        g.le[2] <= vp[2] && vp[2] < g.re[2]) {
        intersect_t = 0.0;
        direction = 3;
    }
    int step[3];
    double iv[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        iv[i] = 0.0;
        if (vd[i] < 0.0) step[i] = -1;
        else if (vd[i] == 0.0) { step[i] = 0; continue; }
        else step[i] = 1;
        iv[i] = HAVE_IV ? ivp[i] : ddiv(1.0, vd[i]);
        if (direction == 3) continue;
        const int x = (i + 1) % 3, y = (i + 2) % 3;
        const double face = step[i] > 0 ? g.le[i] : g.re[i];
        const double tl = dmul(dsub(face, vp[i]), iv[i]);
        double tx = dadd(vp[x], dmul(tl, vd[x]));
        double ty = dadd(vp[y], dmul(tl, vd[y]));
        tx = snap_edge(tx, g.le[x], g.re[x], dmul(1e-10, g.dds[x]));
        ty = snap_edge(ty, g.le[y], g.re[y], dmul(1e-10, g.dds[y]));
        if (g.le[x] <= tx && tx <= g.re[x] && g.le[y] <= ty && ty <= g.re[y] && 0.0 <= tl && tl < intersect_t) {
            direction = i;
            intersect_t = tl;
        }
    }
    if (!((0.0 <= intersect_t) && (intersect_t < max_t))) return false;
    int cur[3];
    double tmax[3], tdelta[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        tdelta[i] = dmul(dmul((double)step[i], iv[i]), g.dds[i]);
        if (i == direction && step[i] > 0) cur[i] = 0;
        else if (i == direction && step[i] < 0) cur[i] = g.dims[i] - 1;
        else {
            const double pos = dadd(dmul(intersect_t, vd[i]), vp[i]);
            double c = dmul(dsub(pos, g.le[i]), g.idds[i]);
            if (-1.0 < c && c < 0.0 && step[i] > 0) c = 0.0;
            else if ((double)(g.dims[i] - 1) < c && c < (double)g.dims[i] && step[i] < 0) c = (double)(g.dims[i] - 1);
            cur[i] = (int)floor(c);
        }
        double edge = 0.0;
        if (step[i] > 0) edge = dadd(dmul((double)(cur[i] + 1), g.dds[i]), g.le[i]);
        else if (step[i] < 0) edge = dadd(dmul((double)cur[i], g.dds[i]), g.le[i]);
        tmax[i] = dmul(dsub(edge, vp[i]), iv[i]);
        if (step[i] == 0) tmax[i] = 1e60;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (cur[i] == g.dims[i] && step[i] >= 0) return false;
        if (cur[i] == -1 && step[i] <= -1) return false;
    }
    double enter_t = intersect_t;
    int hit = 0;
    // Rare: a start index the reference leaves unguarded (e.g. cur == dims while stepping down, a ray
    // entering exactly through an edge).  The reference samples that out-of-range cell over a zero-length
    // segment and walks on; we walk the same steps without touching memory.
    while (cur[0] < 0 || cur[0] >= g.dims[0] || cur[1] < 0 || cur[1] >= g.dims[1] || cur[2] < 0 || cur[2] >= g.dims[2]) {
        hit++;
        smp.skipped_cell();
        int i;
        if (tmax[0] < tmax[1]) i = (tmax[0] < tmax[2]) ? 0 : 2;
        else                   i = (tmax[1] < tmax[2]) ? 1 : 2;
        bool out;
        if (i == 0)      { enter_t = tmax[0]; cur[0] += step[0]; tmax[0] = dadd(tmax[0], tdelta[0]); out = cur[0] < 0 || cur[0] >= g.dims[0]; }
        else if (i == 1) { enter_t = tmax[1]; cur[1] += step[1]; tmax[1] = dadd(tmax[1], tdelta[1]); out = cur[1] < 0 || cur[1] >= g.dims[1]; }
        else             { enter_t = tmax[2]; cur[2] += step[2]; tmax[2] = dadd(tmax[2], tdelta[2]); out = cur[2] < 0 || cur[2] >= g.dims[2]; }
        if (out || enter_t >= max_t) { w.hit = hit; return false; }
    }
    // cells left before the walk leaves the brick along each axis (an axis with step 0 never runs out)
    w.rem0 = step[0] > 0 ? g.dims[0] - 1 - cur[0] : (step[0] < 0 ? cur[0] : 0x7fffffff);
    w.rem1 = step[1] > 0 ? g.dims[1] - 1 - cur[1] : (step[1] < 0 ? cur[1] : 0x7fffffff);
    w.rem2 = step[2] > 0 ? g.dims[2] - 1 - cur[2] : (step[2] < 0 ? cur[2] : 0x7fffffff);
    w.t0 = tmax[0]; w.t1 = tmax[1]; w.t2 = tmax[2];
    w.d0 = tdelta[0]; w.d1 = tdelta[1]; w.d2 = tdelta[2];
    w.enter_t = enter_t; w.max_t = max_t; w.hit = hit;
#pragma unroll
    for (int i = 0; i < 3; i++) { w.cur[i] = cur[i]; w.step[i] = step[i]; }
    return true;
}

template <bool MASKED, class S>
__device__ __forceinline__ int walk_brick(const BrickGeom &g, const double (&vp)[3], const double (&vd)[3],
                                          double max_t_in, S &smp)
{
    WalkState w;
    if (!walk_setup(g, vp, vd, max_t_in, smp, w)) return w.hit;
    int rem0 = w.rem0, rem1 = w.rem1, rem2 = w.rem2;
    double t0 = w.t0, t1 = w.t1, t2 = w.t2;
    const double d0 = w.d0, d1 = w.d1, d2 = w.d2;
    double enter_t = w.enter_t;
    const double max_t = w.max_t;
    int hit = w.hit;
    smp.template begin<MASKED>(w.cur, w.step);
    bool go;
    do {
        hit++;
        const bool a01 = t0 < t1, a02 = t0 < t2, a12 = t1 < t2;
        const bool is0 = a01 && a02, is1 = !a01 && a12;
        const bool is2 = !(is0 || is1);
        const double tm = is0 ? t0 : (is1 ? t1 : t2);
        const double exit_t = tm < max_t ? tm : max_t;
        smp.template cell<MASKED>(enter_t, exit_t);
        enter_t = tm;
        if (is0) { t0 = dadd(t0, d0); rem0--; }
        if (is1) { t1 = dadd(t1, d1); rem1--; }
        if (is2) { t2 = dadd(t2, d2); rem2--; }
        smp.template advance<MASKED>(is0, is1, is2);
        go = (rem0 | rem1 | rem2) >= 0 && !(enter_t >= max_t);
    } while (go);
    return hit;
}

// ---- small helpers ----------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R pick6(const R (&a)[6], int idx)
{
    R r = a[0];
    if (idx == 1) r = a[1];
    if (idx == 2) r = a[2];
    if (idx == 3) r = a[3];
    if (idx == 4) r = a[4];
    if (idx == 5) r = a[5];
    return r;
}

// std::isnormal for fp64 (field_interpolation_tables.pxd:71).
__device__ __forceinline__ bool is_normal64(double x)
{
    const double a = fabs(x);
    return a >= 2.2250738585072014e-308 && a <= 1.7976931348623157e308;  // false for NaN/0/subnormal/inf
}

// ---- EXACT (fp64) pieces: reference evaluation order, no contraction -----------------------------
__device__ __forceinline__ double offset_interpolate64(const double *d, int sx, int sy, const double (&dp)[3])
{
    double dv = dsub(1.0, dp[2]);
    double v0 = dadd(dmul(dv, d[0]),       dmul(dp[2], d[1]));
    double v1 = dadd(dmul(dv, d[sy]),      dmul(dp[2], d[sy + 1]));
    double v2 = dadd(dmul(dv, d[sx]),      dmul(dp[2], d[sx + 1]));
    double v3 = dadd(dmul(dv, d[sx + sy]), dmul(dp[2], d[sx + sy + 1]));
    dv = dsub(1.0, dp[1]);
    v0 = dadd(dmul(dv, v0), dmul(dp[1], v1));
    v1 = dadd(dmul(dv, v2), dmul(dp[1], v3));
    dv = dsub(1.0, dp[0]);
    return dadd(dmul(dv, v0), dmul(dp[0], v1));
}

struct TFShared64 {  // shared-memory views of the fp64 tables
    const double *values, *d0, *dy;
};

__device__ __forceinline__ double fit_get_value64(const DevTF &tf, const TFShared64 &t, int i, const double (&dvs)[6])
{
    const double x = pick6(dvs, tf.field_id[i]);
    if (x >= tf.b1[i] || x <= tf.b0[i]) return 0.0;
    if (!is_normal64(x)) return 0.0;
    int bin = (int)dmul(dsub(x, tf.b0[i]), tf.idbin[i]);
    bin = max(0, min(bin, tf.nbins[i] - 2));
    const int o = tf.offset[i] + bin;
    const double dd = dsub(x, t.d0[o]);
    double out = dadd(t.values[o], dmul(dd, t.dy[o]));
    const int wfi = tf.weight_field_id[i];
    if (wfi != -1) out = dmul(out, pick6(dvs, wfi));
    return out;
}

// FIT_eval_transfer (field_interpolation_tables.pxd:85-110).  B[] (transmittance) is only
// maintained when PARTIAL (multi-GPU affine partials).
template <bool PARTIAL>
__device__ __forceinline__ void eval_transfer64(const DevTF &tf, const TFShared64 &t, double dt,
                                                const double (&dvs)[6], double (&rgba)[4], double (&B)[4])
{
    double ist[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ist[i] = (i < tf.n_fits) ? fit_get_value64(tf, t, i, dvs) : 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i < tf.n_fits) {
            const int fid = tf.weight_table_id[i];
            if (fid != -1) ist[i] = dmul(ist[i], pick6(ist, fid));
        }
    }
    if (tf.grey) {
        const double ta = fmax(dsub(1.0, dmul(dt, pick6(ist, tf.field_table_id[3]))), 0.0);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double tc = pick6(ist, tf.field_table_id[c]);
            rgba[c] = dadd(dmul(dt, tc), dmul(ta, rgba[c]));
            if (PARTIAL) B[c] = dmul(B[c], ta);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double tc = pick6(ist, tf.field_table_id[c]);
            const double ta = fmax(dsub(1.0, dmul(dt, tc)), 0.0);
            rgba[c] = dadd(dmul(dt, tc), dmul(ta, rgba[c]));
            if (PARTIAL) B[c] = dmul(B[c], ta);
        }
    }
}

// ---- LightSourceRenderSampler pieces (fp64, reference evaluation order; not a tuned path) ----------------
// trilinear interpolation from the brick's stored type (fixed_interpolator.cpp:34-52)
template <typename T>
__device__ __forceinline__ double offset_interpolate_t(const T *d, int sx, int sy, const double (&dp)[3])
{
    double dv = dsub(1.0, dp[2]);
    double v0 = dadd(dmul(dv, (double)__ldg(d)),           dmul(dp[2], (double)__ldg(d + 1)));
    double v1 = dadd(dmul(dv, (double)__ldg(d + sy)),      dmul(dp[2], (double)__ldg(d + sy + 1)));
    double v2 = dadd(dmul(dv, (double)__ldg(d + sx)),      dmul(dp[2], (double)__ldg(d + sx + 1)));
    double v3 = dadd(dmul(dv, (double)__ldg(d + sx + sy)), dmul(dp[2], (double)__ldg(d + sx + sy + 1)));
    dv = dsub(1.0, dp[1]);
    v0 = dadd(dmul(dv, v0), dmul(dp[1], v1));
    v1 = dadd(dmul(dv, v2), dmul(dp[1], v3));
    dv = dsub(1.0, dp[0]);
    return dadd(dmul(dv, v0), dmul(dp[0], v1));
}

// eval_gradient (fixed_interpolator.cpp:116-148)
template <typename T>
__device__ __forceinline__ void eval_gradient_t(const T *d, int sx, int sy, double (&dp)[3], double (&grad)[3])
{
    double normval = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double backup = dp[i];
        double plus, minus;
        if (dp[i] >= 0.95) { plus = dp[i]; minus = dsub(dp[i], 0.05); }
        else if (dp[i] <= 0.05) { plus = dadd(dp[i], 0.05); minus = 0.0; }
        else { plus = dadd(dp[i], 0.05); minus = dsub(dp[i], 0.05); }
        const double denom = dsub(plus, minus);
        dp[i] = plus;
        double g = dadd(0.0, ddiv(offset_interpolate_t(d, sx, sy, dp), denom));
        dp[i] = minus;
        g = dsub(g, ddiv(offset_interpolate_t(d, sx, sy, dp), denom));
        dp[i] = backup;
        grad[i] = g;
        normval = dadd(normval, dmul(g, g));
    }
    if (normval != 0.0) {
        normval = __dsqrt_rn(normval);
#pragma unroll
        for (int i = 0; i < 3; i++) grad[i] = ddiv(grad[i], -normval);
    } else {
        grad[0] = grad[1] = grad[2] = 0.0;
    }
}

// The same two functions over the 8 corners of ONE cell held in registers (c[4 i + 2 j + k], fp64): the lit sampler
// interpolates field 0 seven times per sample (value + six gradient probes), 56 corner loads that are 8 per cell here.
__device__ __forceinline__ double cell_interpolate64(const double (&c)[8], const double (&dp)[3])
{
    double dv = dsub(1.0, dp[2]);
    double v0 = dadd(dmul(dv, c[0]), dmul(dp[2], c[1]));
    double v1 = dadd(dmul(dv, c[2]), dmul(dp[2], c[3]));
    double v2 = dadd(dmul(dv, c[4]), dmul(dp[2], c[5]));
    double v3 = dadd(dmul(dv, c[6]), dmul(dp[2], c[7]));
    dv = dsub(1.0, dp[1]);
    v0 = dadd(dmul(dv, v0), dmul(dp[1], v1));
    v1 = dadd(dmul(dv, v2), dmul(dp[1], v3));
    dv = dsub(1.0, dp[0]);
    return dadd(dmul(dv, v0), dmul(dp[0], v1));
}

__device__ __forceinline__ void cell_gradient64(const double (&c)[8], double (&dp)[3], double (&grad)[3])
{
    double normval = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double backup = dp[i];
        double plus, minus;
        if (dp[i] >= 0.95) { plus = dp[i]; minus = dsub(dp[i], 0.05); }
        else if (dp[i] <= 0.05) { plus = dadd(dp[i], 0.05); minus = 0.0; }
        else { plus = dadd(dp[i], 0.05); minus = dsub(dp[i], 0.05); }
        const double denom = dsub(plus, minus);
        dp[i] = plus;
        double g = dadd(0.0, ddiv(cell_interpolate64(c, dp), denom));
        dp[i] = minus;
        g = dsub(g, ddiv(cell_interpolate64(c, dp), denom));
        dp[i] = backup;
        grad[i] = g;
        normval = dadd(normval, dmul(g, g));
    }
    if (normval != 0.0) {
        normval = __dsqrt_rn(normval);
#pragma unroll
        for (int i = 0; i < 3; i++) grad[i] = ddiv(grad[i], -normval);
    } else {
        grad[0] = grad[1] = grad[2] = 0.0;
    }
}

// FIT_eval_transfer_with_light (field_interpolation_tables.pxd:115-143); B only when PARTIAL
template <bool PARTIAL>
__device__ __forceinline__ void eval_transfer64_light(const DevTF &tf, const TFShared64 &t, double dt, const double (&dvs)[6],
                                                      const double (&grad)[3], double (&rgba)[4], double (&B)[4])
{
    double dot_prod = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) dot_prod = dadd(dot_prod, dmul(tf.light_dir[i], grad[i]));
    double ist[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ist[i] = (i < tf.n_fits) ? fit_get_value64(tf, t, i, dvs) : 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i < tf.n_fits) {
            const int fid = tf.weight_table_id[i];
            if (fid != -1) ist[i] = dmul(ist[i], pick6(ist, fid));
        }
    }
    const double t0 = pick6(ist, tf.field_table_id[0]), t1 = pick6(ist, tf.field_table_id[1]), t2 = pick6(ist, tf.field_table_id[2]);
    const double tc[3] = {t0, t1, t2};
    const double ta_grey = fmax(dsub(1.0, dmul(dt, dadd(dadd(t0, t1), t2))), 0.0);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double ta = tf.grey ? ta_grey : fmax(dsub(1.0, dmul(dt, tc[c])), 0.0);
        const double lit = dadd(1., dmul(dot_prod, tf.light_rgba[c]));
        rgba[c] = dadd(dmul(dmul(dsub(1., ta), tc[c]), lit), dmul(ta, rgba[c]));
        if (PARTIAL) B[c] = dmul(B[c], ta);
    }
}

// ---- FAST (fp32) pieces ---------------------------------------------------------------------------
struct Corners {
    float c000, dz00, c010, dz01, c100, dz10, c110, dz11;  // low-z value and z-increment of the 4 z-edges
};

template <typename T>
__device__ __forceinline__ void load_corners(Corners &c, const T *d, int sx, int sy)
{
    const float a0 = (float)__ldg(d), a1 = (float)__ldg(d + 1);
    const float b0 = (float)__ldg(d + sy), b1 = (float)__ldg(d + sy + 1);
    const float e0 = (float)__ldg(d + sx), e1 = (float)__ldg(d + sx + 1);
    const float f0 = (float)__ldg(d + sx + sy), f1 = (float)__ldg(d + sx + sy + 1);
    c.c000 = a0; c.dz00 = a1 - a0;
    c.c010 = b0; c.dz01 = b1 - b0;
    c.c100 = e0; c.dz10 = e1 - e0;
    c.c110 = f0; c.dz11 = f1 - f0;
}

__device__ __forceinline__ float trilinear(const Corners &c, float x, float y, float z)
{
    const float v00 = fmaf(z, c.dz00, c.c000);
    const float v01 = fmaf(z, c.dz01, c.c010);
    const float v10 = fmaf(z, c.dz10, c.c100);
    const float v11 = fmaf(z, c.dz11, c.c110);
    const float v0 = fmaf(y, v01 - v00, v00);
    const float v1 = fmaf(y, v11 - v10, v10);
    return fmaf(x, v1 - v0, v0);
}

// One table lookup in fp32.  (x - b0) is formed with b0 split hi/lo so the bin coordinate keeps
// ~1e-7 relative accuracy even when |b0| >> (b1 - b0) (log fields).  Zero, NaN, inf and
// out-of-bounds values give 0 like FIT_get_value (field_interpolation_tables.pxd:70-71).
__device__ __forceinline__ float fit_get_value32(const DevTF &tf, const float2 *pairs, int i, const float (&dvs)[6])
{
    const float x = pick6(dvs, tf.field_id[i]);
    const float u = (x - tf.b0_hi[i]) - tf.b0_lo[i];
    if (!(u > 0.f) || !(x < tf.b1_f[i]) || x == 0.f) return 0.f;
    const float fb = u * tf.idbin_f[i];
    int bin = (int)fb;
    bin = max(0, min(bin, tf.nbins[i] - 2));
    const float2 p = pairs[tf.offset[i] + bin];
    float out = fmaf(fb - (float)bin, p.y, p.x);
    const int wfi = tf.weight_field_id[i];
    if (wfi != -1) out *= pick6(dvs, wfi);
    return out;
}

template <bool PARTIAL>
__device__ __forceinline__ void eval_transfer32(const DevTF &tf, const float2 *pairs, float dt,
                                                const float (&dvs)[6], float (&rgba)[4], float (&B)[4])
{
    float ist[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ist[i] = (i < tf.n_fits) ? fit_get_value32(tf, pairs, i, dvs) : 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i < tf.n_fits) {
            const int fid = tf.weight_table_id[i];
            if (fid != -1) ist[i] *= pick6(ist, fid);
        }
    }
    if (tf.grey) {
        const float ta = fmaxf(1.f - dt * pick6(ist, tf.field_table_id[3]), 0.f);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            rgba[c] = fmaf(ta, rgba[c], dt * pick6(ist, tf.field_table_id[c]));
            if (PARTIAL) B[c] *= ta;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float t = dt * pick6(ist, tf.field_table_id[c]);
            const float ta = fmaxf(1.f - t, 0.f);
            rgba[c] = fmaf(ta, rgba[c], t);
            if (PARTIAL) B[c] *= ta;
        }
    }
}

// FIT_eval_transfer_with_light in fp32 (FAST precision; the gradient comes in fp64)
template <bool PARTIAL>
__device__ __forceinline__ void eval_transfer32_light(const DevTF &tf, const float2 *pairs, float dt, const float (&dvs)[6],
                                                      const double (&grad)[3], float (&rgba)[4], float (&B)[4])
{
    const float dot_prod = (float)(tf.light_dir[0] * grad[0] + tf.light_dir[1] * grad[1] + tf.light_dir[2] * grad[2]);
    float ist[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ist[i] = (i < tf.n_fits) ? fit_get_value32(tf, pairs, i, dvs) : 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i < tf.n_fits) {
            const int fid = tf.weight_table_id[i];
            if (fid != -1) ist[i] *= pick6(ist, fid);
        }
    }
    const float tc[3] = {pick6(ist, tf.field_table_id[0]), pick6(ist, tf.field_table_id[1]), pick6(ist, tf.field_table_id[2])};
    const float ta_grey = fmaxf(1.f - dt * (tc[0] + tc[1] + tc[2]), 0.f);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float ta = tf.grey ? ta_grey : fmaxf(1.f - dt * tc[c], 0.f);
        rgba[c] = fmaf(ta, rgba[c], (1.f - ta) * tc[c] * (1.f + dot_prod * (float)tf.light_rgba[c]));
        if (PARTIAL) B[c] *= ta;
    }
}

// ColorTransferFunction fast path: 4 tables on field 0 with common bounds/nbins, RGB weighted by
// the alpha table (transfer_functions.py:385-406).  One bin computation, two 16-byte table reads.
//
// Shared-memory layout (ncu, round 1: the plain [bin] layout ran the LSU data pipe at 98 % with
// 26 wavefronts per warp-sample, 70 % of them bank conflicts, because 32 lanes hit 32 unrelated
// bins): the (r,g,b,a) row of every bin is replicated NCOPY (8) times, copy k living in 16-byte
// bank group k, and lane l reads copy l & 7.  The 8 lanes of each LDS.128 quarter-warp then
// touch 8 distinct bank groups whatever their bins are: 4 wavefronts per load, the minimum.
// Per-lane state of the colour fast path is just `base`; every other constant is read from the kernel
// parameters (constant bank operands cost no registers).
struct ColourTable {
    unsigned base;     // shared-space byte address of this lane's copy of row 0, pre-biased (see colour_cell_run)
    const DevTF *tf;   // kernel-parameter copy of the table description:
                       //   b0_hi[0]            fp32 head of the lower bound; its tail is folded into colour_c0
                       //   idbin_f[0], colour_c0   row coordinate of a value: fb = (val - b0_hi) * idbin + c0
                       //   colour_lo/hi        row coordinates of the bounds: a sample is valid iff lo < fb < hi
                       //   colour_mid/half     (lo + hi) / 2 and (hi - lo) / 2: valid iff |fb - mid| < half
                       //   colour_zrow         row coordinate of the value 0 (rejected when 0 lies inside the bounds)
                       //   colour_band         bound on the fp32 error of a row coordinate near the thresholds
                       //   colour_clamp_bound  upper bound of any table product that meets dt before a clamp
};
constexpr int COLOUR_COPIES = 8;                      // replication factor
constexpr unsigned COLOUR_ROW_BYTES = COLOUR_COPIES * 16;  // bytes between consecutive rows
// Row layout (per copy): [0, 0, y_0 ... y_{n-1}, y_{n-1}] -> n + 3 rows.  Row index = bin + 2, so that
// the two leading zero rows serve every rejected sample without a branch, and bin == n-1 (a value that
// rounds onto b1) interpolates between two copies of the last entry.
constexpr int COLOUR_PAD_ROWS = 3;
// Delta layout (tables of at most COLOUR_DELTA_MAX_BINS bins): a row is 256 bytes = the 8 copies of
// (r,g,b,a) followed by the 8 copies of the increments to the next bin, so the lerp is ONE packed FMA per
// pair and every LDS.128 quarter-warp still touches 8 distinct bank groups.  Rows 0, 1 and >= nbins + 1
// are zero: clamping the row coordinate to [0.5, nbins + 1.5] (two FMNMX, NaN -> 0.5) rejects samples.
constexpr int COLOUR_DELTA_MAX_BINS = 256;
constexpr unsigned COLOUR_DELTA_ROW_BYTES = 2 * COLOUR_ROW_BYTES;
constexpr float COLOUR_MAGIC = 8388608.f;             // 2^23: x + 2^23 rounded down == 2^23 + floor(x)

__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// sm_100 packed fp32 (FFMA2/FMUL2): two lanes of fp32 per issue slot.  The cast kernel is
// instruction-issue bound (ncu round 1: 82 % issue-active, 69 instr/sample), so halving the issue
// slots of the paired arithmetic is the lever.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 bcast2(float a) { return make_float2(a, a); }

// The 8 corners of a cell packed for paired z- and y-lerps.
struct CornersP {
    float2 p0, d0;  // (c000, c100) and their z-increments (c001-c000, c101-c100)
    float2 p1, d1;  // (c010, c110) and their z-increments
};

// Corners converted to table ROW COORDINATES at load time: trilinear interpolation commutes with the
// affine map val -> (val - b0) * idbin + 2, so the sample loop gets the row coordinate straight out of
// the interpolation (saves a subtract and an FMA per sample for ~10 packed ops per cell).  The subtract
// comes first so that |b0| >> (b1 - b0) (log fields) does not cancel catastrophically.
// The 8 raw corner values of a cell: v[0..3] = (c000, c100, c001, c101), v[4..7] = (c010, c110, c011, c111).
struct RawCorners {
    float a0, e0, a1, e1, b0, f0, b1, f1;
};

// Returns whether all 8 corners are exactly (+-)0.
__device__ __forceinline__ bool rows_from_raw(CornersP &c, const RawCorners &r, const ColourTable &ct)
{
    const float2 nb = bcast2(-ct.tf->b0_hi[0]), sc = bcast2(ct.tf->idbin_f[0]), c0 = bcast2(ct.tf->colour_c0);
    const float2 lo0 = fma2(add2(make_float2(r.a0, r.e0), nb), sc, c0), hi0 = fma2(add2(make_float2(r.a1, r.e1), nb), sc, c0);
    const float2 lo1 = fma2(add2(make_float2(r.b0, r.f0), nb), sc, c0), hi1 = fma2(add2(make_float2(r.b1, r.f1), nb), sc, c0);
    c.p0 = lo0; c.d0 = fma2(lo0, make_float2(-1.f, -1.f), hi0);
    c.p1 = lo1; c.d1 = fma2(lo1, make_float2(-1.f, -1.f), hi1);
    const unsigned bits = (__float_as_uint(r.a0) | __float_as_uint(r.a1) | __float_as_uint(r.b0)) |
                          (__float_as_uint(r.b1) | __float_as_uint(r.e0) | __float_as_uint(r.e1)) |
                          (__float_as_uint(r.f0) | __float_as_uint(r.f1));
    return (bits << 1) == 0u;
}

// Accumulators of one ray on the ColorTransferFunction path.  For an affine partial (PARTIAL: C -> A + B*C, the
// accumulators start at A = 0, B = 1) the transmittance needs no arithmetic of its own: each step is
// A <- e + a*A, B <- a*B with a = max(1 - e, 0), and while no clamp has fired a = 1 - e gives 1 - A' = (1 - e)(1 - A),
// i.e. B = 1 - A at every step (per channel; grey: B = 1 - A_alpha for all four).  A clamp that fires (e > 1) zeroes B
// for good; `fired` keeps one sticky bit per channel for that.  (Tracking B explicitly cost 2 of 35 instructions per
// sample and 4 registers in the multi-GPU casts.)
struct ColourAcc {
    float2 c01; float c2, c3;
    unsigned fired;
    float tb;   // ERT casts only: transmittance of the samples composited so far (front to back), see colour_cell_run
};

// All n_samples samples of one cell: trilinear value -> 4-table lookup -> composite
// (image_samplers.pyx:480-487 + field_interpolation_tables.pxd:66-110, fp32).
//
// Instruction diet (ncu r01a: 56 thread-instructions per sample, XU pipe 20 % busy with F2I/I2F):
//   * the corners are already in row coordinates, so the interpolated value IS the row coordinate;
//   * floor() and the row address come from one round-down add of 2^23 (FADD.RM): the integer part
//     sits in the mantissa bits, so `bits * 128 + base` is the row address (base is pre-biased) and
//     `(t - 2^23)` is the integer part as a float -- no F2I, no I2F, no clamps;
//   * rejected samples (out of bounds, NaN, inf) are steered between the two zero rows by ONE select on
//     the row coordinate instead of selects on fraction and alpha; exact zeros are rejected per cell;
//   * CLAMP=false drops the max(1 - dt*T, 0) clamps when the caller proved dt*T <= 1 for the cell.
// ERT (grey opacity only): the samples arrive in REVERSE composition order (the last sample the reference composites
// comes first), so the affine maps are applied front to back -- A <- A + B e, B <- B a, field_interpolation_tables.pxd:
// 103-110 read from the other end -- and B is the transmittance the termination test looks at.
// LINE (delta layout only): the rows hold (c, d) of y = d * fs + c in the (origin-shifted) row coordinate, see DevTF.
template <bool PARTIAL, bool GREY, bool CLAMP, bool EDGE, bool DELTA, bool ERT = false, bool LINE = false>
__device__ __forceinline__ float colour_cell_run(const CornersP &c, const ColourTable &ct, int n_samples, float x, float y, float z,
                                                float sx, float sy, float sz, float dt, ColourAcc &A)
{
    float closest = 3.0e38f;
    const float2 neg1 = make_float2(-1.f, -1.f), one2 = make_float2(1.f, 1.f);
    // FIT_get_value for the 4 tables at once + FIT_eval_transfer, for the sample whose row coordinate is fb
    auto sample = [&](const float fb) {
        // EDGE: keep the smallest distance to a bound of the cell; the caller compares it with the fp32 error band.
        if (EDGE) closest = fminf(closest, fabsf(fabsf(fb - ct.tf->colour_mid) - ct.tf->colour_half));
        float2 rg, ba;
        if (DELTA && LINE) {
            const float fs = fminf(fmaxf(fb, ct.tf->colour_clamp_lo), ct.tf->colour_clamp_hi);   // rejected samples land on zero rows
            const float t = __fadd_rd(fs, ct.tf->colour_magic);   // 2^23 + floor(fs + origin)
            const unsigned addr = (unsigned)__float_as_int(t) * COLOUR_DELTA_ROW_BYTES + ct.base;
            const float4 t0 = lds128(addr), t1 = lds128(addr + COLOUR_ROW_BYTES);
            const float2 fs2 = bcast2(fs);
            rg = fma2(fs2, make_float2(t1.x, t1.y), make_float2(t0.x, t0.y));
            ba = fma2(fs2, make_float2(t1.z, t1.w), make_float2(t0.z, t0.w));
        } else if (DELTA) {
            const float fs = fminf(fmaxf(fb, ct.tf->colour_clamp_lo), ct.tf->colour_clamp_hi);   // rejected samples land on zero rows
            const float t = __fadd_rd(fs, ct.tf->colour_magic);   // 2^23 + floor(fs)
            const float fr = fs - (t - ct.tf->colour_magic);
            const unsigned addr = (unsigned)__float_as_int(t) * COLOUR_DELTA_ROW_BYTES + ct.base;
            const float4 t0 = lds128(addr), t1 = lds128(addr + COLOUR_ROW_BYTES);
            const float2 fr2 = bcast2(fr);
            rg = fma2(fr2, make_float2(t1.x, t1.y), make_float2(t0.x, t0.y));
            ba = fma2(fr2, make_float2(t1.z, t1.w), make_float2(t0.z, t0.w));
        } else {
            // margin < 0 <=> b0 < value < b1 (false for NaN)
            const bool valid = fabsf(fb - ct.tf->colour_mid) - ct.tf->colour_half < 0.f;
            const float fs = valid ? fb : 0.5f;
            const float t = __fadd_rd(fs, COLOUR_MAGIC);   // 2^23 + floor(fs)
            const float fr = fs - (t - COLOUR_MAGIC);
            const unsigned addr = (unsigned)__float_as_int(t) * COLOUR_ROW_BYTES + ct.base;
            const float4 t0 = lds128(addr), t1 = lds128(addr + COLOUR_ROW_BYTES);
            const float2 fr2 = bcast2(fr);
            rg = fma2(fr2, fma2(make_float2(t0.x, t0.y), neg1, make_float2(t1.x, t1.y)), make_float2(t0.x, t0.y));
            ba = fma2(fr2, fma2(make_float2(t0.z, t0.w), neg1, make_float2(t1.z, t1.w)), make_float2(t0.z, t0.w));
        }
        const float adt = ba.y * dt;                   // dt * alpha-table value
        const float2 e01 = mul2(rg, bcast2(adt));      // dt * T_r, dt * T_g (RGB weighted by alpha)
        const float e2 = ba.x * adt;
        if (GREY && ERT) {
            float ta = 1.f - adt;
            if (CLAMP) ta = fmaxf(ta, 0.f);
            A.c01 = fma2(bcast2(A.tb), e01, A.c01);
            A.c2 = fmaf(A.tb, e2, A.c2);
            A.c3 = fmaf(A.tb, adt, A.c3);
            A.tb *= ta;
        } else if (GREY) {
            float ta = 1.f - adt;
            if (CLAMP) {
                if (PARTIAL && ta < 0.f) A.fired |= 15u;
                ta = fmaxf(ta, 0.f);
            }
            A.c01 = fma2(bcast2(ta), A.c01, e01);
            A.c2 = fmaf(ta, A.c2, e2);
            A.c3 = fmaf(ta, A.c3, adt);
        } else {
            float2 a01 = fma2(e01, neg1, one2);
            float a2 = 1.f - e2;
            if (CLAMP) {
                if (PARTIAL) A.fired |= (a01.x < 0.f ? 1u : 0u) | (a01.y < 0.f ? 2u : 0u) | (a2 < 0.f ? 4u : 0u);
                a01.x = fmaxf(a01.x, 0.f); a01.y = fmaxf(a01.y, 0.f); a2 = fmaxf(a2, 0.f);
            }
            A.c01 = fma2(a01, A.c01, e01);
            A.c2 = fmaf(a2, A.c2, e2);
        }
    };
#if defined(YTB_TRILINEAR_STEP)
    // one trilinear interpolation per sample, positions advanced by repeated addition (8 instructions / sample)
    float2 zy = make_float2(z, y);
    const float2 szy = make_float2(sz, sy);
#pragma unroll 2
    for (int s = 0; s < n_samples; s++) {
        const float2 zz = bcast2(zy.x), yy = bcast2(zy.y);
        const float2 va = fma2(zz, c.d0, c.p0);        // (v00, v10)
        const float2 vb = fma2(zz, c.d1, c.p1);        // (v01, v11)
        const float2 v = fma2(yy, fma2(va, neg1, vb), va);  // (v0, v1)
        sample(fmaf(x, v.y - v.x, v.x));               // row coordinate of the sample value
        zy = add2(zy, szy); x += sx;
    }
#else
    // Along the ray the trilinear interpolant (fixed_interpolator.cpp:34-52) is a CUBIC in the sample index t:
    // the z-lerps are linear in t, the y-lerp multiplies two linear factors, the x-lerp a linear and a
    // quadratic one.  19 instructions per cell build its coefficients; every pair of samples is then 3 packed
    // Horner steps (1.5 instructions per sample instead of 8 -- ncu r01d: the loop was FMA-pipe throttled).
    const float2 a0 = fma2(bcast2(z), c.d0, c.p0), a1 = mul2(bcast2(sz), c.d0);   // z-edges at (x0|x1, y0): a0 + a1 t
    const float2 b0 = fma2(bcast2(z), c.d1, c.p1), b1 = mul2(bcast2(sz), c.d1);   // z-edges at (x0|x1, y1)
    const float2 d0 = fma2(a0, neg1, b0), d1 = fma2(a1, neg1, b1);
    const float2 v0 = fma2(bcast2(y), d0, a0);                                    // faces x0|x1: v0 + v1 t + v2 t^2
    const float2 v1 = fma2(bcast2(sy), d0, fma2(bcast2(y), d1, a1));
    const float2 v2 = mul2(bcast2(sy), d1);
    const float e0 = v0.y - v0.x, e1 = v1.y - v1.x, e2 = v2.y - v2.x;
    const float2 hc0 = bcast2(fmaf(x, e0, v0.x));
    const float2 hc1 = bcast2(fmaf(sx, e0, fmaf(x, e1, v1.x)));
    const float2 hc2 = bcast2(fmaf(sx, e1, fmaf(x, e2, v2.x)));
    const float2 hc3 = bcast2(sx * e2);
    float2 tt = make_float2(0.f, 1.f);
    const float2 two2 = make_float2(2.f, 2.f);
#ifndef YTB_PAIR_UNROLL
#define YTB_PAIR_UNROLL 5   // n_samples = 10 (the reference default) is one trip; r01z9: 45.2 -> 44.3 ms against no unrolling
#endif
    constexpr int kPairUnroll = YTB_PAIR_UNROLL;
    const int pairs = n_samples >> 1;
#pragma unroll kPairUnroll
    for (int s2 = 0; s2 < pairs; s2++) {
        const float2 fb2 = fma2(fma2(fma2(hc3, tt, hc2), tt, hc1), tt, hc0);   // row coordinates of samples 2 s2, 2 s2 + 1
        tt = add2(tt, two2);
        sample(fb2.x);
        sample(fb2.y);
    }
    if (n_samples & 1) {
        const float t = tt.x;
        sample(fmaf(fmaf(fmaf(hc3.x, t, hc2.x), t, hc1.x), t, hc0.x));
    }
#endif
    return closest;
}

// One sample on the careful path: `fs` is the row coordinate (0.5 if the sample was rejected); same table
// lookup and compositing as colour_cell_run, with the clamps always on.
template <bool PARTIAL, bool GREY, bool ERT = false>
__device__ __forceinline__ void colour_sample(const ColourTable &ct, float fs, float dt, ColourAcc &A)
{
    const float t = __fadd_rd(fs, COLOUR_MAGIC);
    const float fr = fs - (t - COLOUR_MAGIC);
    const bool delta = ct.tf->colour_delta != 0;
    const unsigned addr = (unsigned)__float_as_int(t) * (delta ? COLOUR_DELTA_ROW_BYTES : COLOUR_ROW_BYTES) + ct.base;
    const float4 t0 = lds128(addr);
    float4 t1 = lds128(addr + COLOUR_ROW_BYTES);
    if (!delta) t1 = make_float4(t1.x - t0.x, t1.y - t0.y, t1.z - t0.z, t1.w - t0.w);
    // LINE-form rows hold (c, d) of y = d * (fs - origin) + c; this path keeps TRUE row coordinates (origin = magic - 2^23)
    const float w = ct.tf->colour_line ? fs - (ct.tf->colour_magic - COLOUR_MAGIC) : fr;
    const float r = fmaf(w, t1.x, t0.x), g = fmaf(w, t1.y, t0.y), b = fmaf(w, t1.z, t0.z);
    const float adt = fmaf(w, t1.w, t0.w) * dt;
    const float e0 = r * adt, e1 = g * adt, e2 = b * adt;
    if (GREY && ERT) {
        const float ta = fmaxf(1.f - adt, 0.f);
        A.c01.x = fmaf(A.tb, e0, A.c01.x); A.c01.y = fmaf(A.tb, e1, A.c01.y); A.c2 = fmaf(A.tb, e2, A.c2); A.c3 = fmaf(A.tb, adt, A.c3);
        A.tb *= ta;
    } else if (GREY) {
        if (PARTIAL && 1.f - adt < 0.f) A.fired |= 15u;
        const float ta = fmaxf(1.f - adt, 0.f);
        A.c01.x = fmaf(ta, A.c01.x, e0); A.c01.y = fmaf(ta, A.c01.y, e1); A.c2 = fmaf(ta, A.c2, e2); A.c3 = fmaf(ta, A.c3, adt);
    } else {
        if (PARTIAL) A.fired |= (1.f - e0 < 0.f ? 1u : 0u) | (1.f - e1 < 0.f ? 2u : 0u) | (1.f - e2 < 0.f ? 4u : 0u);
        const float a0 = fmaxf(1.f - e0, 0.f), a1 = fmaxf(1.f - e1, 0.f), a2 = fmaxf(1.f - e2, 0.f);
        A.c01.x = fmaf(a0, A.c01.x, e0); A.c01.y = fmaf(a1, A.c01.y, e1); A.c2 = fmaf(a2, A.c2, e2);
    }
}

// ---- shared by every cast kernel: CTA shape and ray generation ------------------------------------------------
#ifndef YTB_CTA_THREADS
#define YTB_CTA_THREADS 256
#endif
// a CTA's warps are 4 x 8 pixel patches laid two abreast (see cast_kernel): 16 columns, CTA_THREADS / 16 rows
constexpr int CTA_THREADS = YTB_CTA_THREADS, TILE_I = CTA_THREADS / 16, TILE_J = 16;
static_assert(CTA_THREADS % 64 == 0, "warps pair up across the tile");

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

// Does any brick of the order list reach this CTA's pixel box?  (CTA-uniform; every thread must call it.)  A tile no brick
// reaches reads neither its initial colours nor its z-buffer and stores nothing (a partial cast stores the identity map):
// with page-locked host arrays bound in place, that is PCIe traffic not spent on the empty part of the image.
__device__ __forceinline__ bool tile_has_bricks(const int4 *__restrict__ extent, const int4 *__restrict__ chunk_extent, int n_order,
                                                int ti0, int ti1, int tj0, int tj1)
{
    int hit = 0;
    for (int base = 0; base < n_order; base += CTA_THREADS) {
        const int4 ce = chunk_extent[base / CTA_THREADS];   // union footprint of this chunk of the list (chunk_extent_kernel)
        if (!(ce.x < ti1 && ce.y > ti0 && ce.z < tj1 && ce.w > tj0)) continue;
        const int k = base + (int)threadIdx.x;
        if (k < n_order) {
            const int4 e = extent[k];
            hit |= (e.x < ti1 && e.y > ti0 && e.z < tj1 && e.w > tj0) ? 1 : 0;
        }
    }
    return __syncthreads_or(hit) != 0;
}

// ---- partial images: store one pixel's map in any format; grid-wide completion signal ----------------------------
// out / Bout: the map as the cast kernels hold it at write-back (fp64 temporaries; B is exactly 0 where a clamp fired).
__device__ __forceinline__ void store_partial(void *buf, int fmt, size_t pix, const double (&A)[4], const double (&B)[4])
{
    if (fmt == PF_AB64) {
        double4 *dst = reinterpret_cast<double4 *>(buf) + pix * 2;
        dst[0] = make_double4(A[0], A[1], A[2], A[3]);
        dst[1] = make_double4(B[0], B[1], B[2], B[3]);
    } else if (fmt == PF_AB32) {
        float4 *dst = reinterpret_cast<float4 *>(buf) + pix * 2;
        dst[0] = make_float4((float)A[0], (float)A[1], (float)A[2], (float)A[3]);
        dst[1] = make_float4((float)B[0], (float)B[1], (float)B[2], (float)B[3]);
    } else if (fmt == PF_A32) {
        const unsigned bits = (B[0] == 0.0 ? 1u : 0u) | (B[1] == 0.0 ? 2u : 0u) | (B[2] == 0.0 ? 4u : 0u);
        reinterpret_cast<float4 *>(buf)[pix] = make_float4((float)A[0], (float)A[1], (float)A[2], (float)bits);
    } else if (fmt == PF_A32G) {
        const float a = fabsf((float)A[3]);
        reinterpret_cast<float4 *>(buf)[pix] = make_float4((float)A[0], (float)A[1], (float)A[2], B[3] == 0.0 ? -a : a);
    } else if (fmt == PF_SUM1) {
        reinterpret_cast<double *>(buf)[pix] = A[0];
    } else if (fmt == PF_SUM2) {
        reinterpret_cast<double2 *>(buf)[pix] = make_double2(A[0], A[1]);
    } else {
        reinterpret_cast<double4 *>(buf)[pix] = make_double4(A[0], A[1], A[2], A[3]);
    }
}

constexpr int SWAP_MAX_DST = 8;
struct SwapSignal { unsigned long long *flag[SWAP_MAX_DST]; int n; };

// Called by EVERY thread of EVERY CTA as the last thing a kernel does: when all CTAs have passed, the last one writes
// `seq` into the flag words (which may live in a peer's memory).  One fence per CTA, by thread 0 AFTER the barrier (the
// barrier orders the CTA's stores before the fence, the fence is cumulative -- the grid-sync idiom); a fence per thread
// cost 0.27 ms on a 2048^2 frame (r02j vs r02k: 5.80 against 5.54 ms for the same sub-volume).  The per-CTA fence is
// DEVICE scope: it orders the CTA's stores before its count at the L2, which is where a peer's NVLink reads arrive; the
// one CTA that counts last raises that to system scope before it publishes the flag.
__device__ __forceinline__ void grid_tail_signal(unsigned *counter, const SwapSignal &sg, unsigned long long seq)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(counter, 1u) == total - 1u) {
            *counter = 0u;   // the next kernel that counts on it is stream-ordered after this one
            __threadfence_system();
            for (int k = 0; k < sg.n; k++) *reinterpret_cast<volatile unsigned long long *>(sg.flag[k]) = seq;
            __threadfence_system();
        }
    }
}

}  // namespace ytb
