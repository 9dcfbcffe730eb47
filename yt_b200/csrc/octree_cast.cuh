// The octree cast: ImageSampler.cast_through_octree, image_samplers.pyx:202-299, over the
// parametric octree traversal of _octree_raytracing.hpp:199-371 (Revelles, Urena & Lastra).
//
// One thread owns one pixel: it walks the tree front to back and samples every terminal cell it meets with the
// sampler's per-cell body, so the reference's compositing order along a ray is kept by construction.  The
// reference recurses (proc_subtree); here the recursion is an explicit per-thread stack of (t0, t1, node, next
// child) frames for the ancestors that still have children to visit, the current node being kept in registers -- the same arithmetic in the same order (tm = (t0 + t1) * 0.5 per level, std::max / std::min
// written with the comparison libstdc++ uses), all fp64 without contraction: every operation is the reference's,
// so the image equals the reference's bit for bit (up to libm's sqrt, which is correctly rounded on both sides).
// That is the EXACT precision and what the projection kinds always use (they carry the 1e-6 bar); the volume render's
// FAST precision (the default) keeps the traversal in fp64 and samples in fp32 (oct_sample_cell_fast).
// General kernels, not a tuned path (the reference reaches it from OctreeVolumeSource only).
#pragma once
#include "raycast.cuh"

namespace ytb {

constexpr int OCT_MAX_LEVELS = 32;   // frames of the traversal stack: root + max_depth (<= 30) levels
constexpr int OCT_THREADS = 128, OCT_TILE_I = 8, OCT_TILE_J = 16;

struct OctParams {
    int nx, ny;
    const double *pos_px, *dir_px;   // (nx,ny,3) rays
    double *image;                   // (nx,ny,4) fp64, in/out
    const int *child;                // [n_nodes][8] child node index, -1 = absent (node 0 is the root)
    const int *leaf;                 // [n_nodes][2]: (terminal, key)
    double dle[3], dre[3], size[3];
    const double *cells;             // [n_cells][14]: 8 vertex values (x-major), left edge, right edge
    int n_samples;
    double inv_ns;                   // 1 / n_samples
    DevTF tf;
};

__device__ __forceinline__ double std_max(double a, double b) { return a < b ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return b < a ? b : a; }

// first_node / next_node, _octree_raytracing.hpp:346-371
__device__ __forceinline__ int oct_first_node(double tx0, double ty0, double tz0, double txm, double tym, double tzm)
{
    int index = 0;
    if (tx0 >= std_max(ty0, tz0)) { if (tym < tx0) index |= 2; if (tzm < tx0) index |= 1; }
    else if (ty0 >= std_max(tx0, tz0)) { if (txm < ty0) index |= 4; if (tzm < ty0) index |= 1; }
    else { if (txm < tz0) index |= 4; if (tym < tz0) index |= 2; }
    return index;
}
__device__ __forceinline__ int oct_next_node(double tx, double ty, double tz, int ix, int iy, int iz)
{
    if (tx < std_min(ty, tz)) return ix;
    else if (ty < std_min(tx, tz)) return iy;
    return iz;
}

// The sampler body for one terminal cell: a 1x1x1 VolumeContainer over the cell's 14 doubles, index (0,0,0), mask
// all ones (image_samplers.pyx:268-288 -> 327-343 / 372-402 / 453-487).
template <int KIND>
__device__ __forceinline__ void oct_sample_cell(const OctParams &p, const TFShared64 &t64, const double *cell,
                                                const double (&vp)[3], const double (&vd)[3], double enter_t, double exit_t,
                                                double (&acc)[4])
{
    if (KIND == KIND_PROJ) {
        acc[0] = dadd(acc[0], dmul(cell[0], dsub(exit_t, enter_t)));
        return;
    }
    double dp[3], ds[3], dvs[6], B64[4];
    const double dt = ddiv(dsub(exit_t, enter_t), (double)p.n_samples);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double le = cell[8 + i], dds = dsub(cell[11 + i], le), idds = ddiv(1.0, dds);
        dp[i] = dadd(dmul(dadd(enter_t, dmul(0.5, dt)), vd[i]), vp[i]);
        dp[i] = dsub(dp[i], dadd(dmul(0.0, dds), le));
        dp[i] = dmul(dp[i], idds);
        ds[i] = dmul(dmul(vd[i], idds), dt);
    }
#pragma unroll
    for (int j = 1; j < 6; j++) dvs[j] = 0.0;
    for (int s = 0; s < p.n_samples; s++) {
        dvs[0] = offset_interpolate64(cell, 4, 2, dp);
        if (KIND == KIND_VR) eval_transfer64<false>(p.tf, t64, dt, dvs, acc, B64);
        else acc[0] = dadd(acc[0], dmul(dvs[0], dt));
#pragma unroll
        for (int i = 0; i < 3; i++) dp[i] = dadd(dp[i], ds[i]);
    }
}

// FAST precision of the volume render (the product default, as for the kd path): the position of a cell's first
// sample is formed in fp64 (the cancellation pos - left_edge, H13) and rounded once; trilinear interpolation, the
// table lookups (fp32 (value, slope) pairs staged in shared memory, bounds split hi/lo) and the composite are fp32.
__device__ __forceinline__ void oct_sample_cell_fast(const OctParams &p, const float2 *pairs, const double *cell,
                                                     const double (&vp)[3], const double (&vd)[3], double enter_t, double exit_t,
                                                     float (&rgba)[4])
{
    const double dt64 = (exit_t - enter_t) * p.inv_ns;
    const double tmid = fma(0.5, dt64, enter_t);
    float x[3], st[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // 1 / dds to fp32 accuracy is enough here: it scales the already-formed difference pos - left_edge
        const double le = cell[8 + i], idds = (double)__frcp_rn((float)(cell[11 + i] - le));
        x[i] = (float)((fma(tmid, vd[i], vp[i]) - le) * idds);
        st[i] = (float)(vd[i] * idds * dt64);
    }
    Corners c;
    {
        const float a0 = (float)cell[0], a1 = (float)cell[1], b0 = (float)cell[2], b1 = (float)cell[3];
        const float e0 = (float)cell[4], e1 = (float)cell[5], f0 = (float)cell[6], f1 = (float)cell[7];
        c.c000 = a0; c.dz00 = a1 - a0; c.c010 = b0; c.dz01 = b1 - b0;
        c.c100 = e0; c.dz10 = e1 - e0; c.c110 = f0; c.dz11 = f1 - f0;
    }
    const float dt = (float)dt64;
    float dvs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, B[4];
    for (int s = 0; s < p.n_samples; s++) {
        dvs[0] = trilinear(c, x[0], x[1], x[2]);
        eval_transfer32<false>(p.tf, pairs, dt, dvs, rgba, B);
        x[0] += st[0]; x[1] += st[1]; x[2] += st[2];
    }
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(OCT_THREADS) octree_cast_kernel(const __grid_constant__ OctParams p)
{
    extern __shared__ float2 s_oct_pairs[];
    if (FAST) {
        for (int i = threadIdx.x; i < p.tf.total_bins; i += OCT_THREADS) s_oct_pairs[i] = p.tf.pairs[i];
        __syncthreads();
    }
    // a CTA is an 8 x 16 pixel tile, its warps 4 x 8 patches: neighbouring rays take nearly the same path through the tree
    const int tiles_j = (p.ny + OCT_TILE_J - 1) / OCT_TILE_J;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vi = (blockIdx.x / tiles_j) * OCT_TILE_I + (warp >> 1) * 4 + (lane >> 3);
    const int vj = (blockIdx.x % tiles_j) * OCT_TILE_J + (warp & 1) * 8 + (lane & 7);
    if (vi >= p.nx || vj >= p.ny) return;
    const long long px = (long long)vi * p.ny + vj;
    double vp[3], vd[3], o[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { vp[i] = p.pos_px[px * 3 + i]; vd[i] = p.dir_px[px * 3 + i]; }
    // Ray(o, d): the direction normalised (hpp:53-65); vp_dir_len of cast_through_octree (image_samplers.pyx:246-249)
    const double len = __dsqrt_rn(dadd(dadd(dmul(vd[0], vd[0]), dmul(vd[1], vd[1])), dmul(vd[2], vd[2])));
    // cast_ray (hpp:199-226): mirror the ray into the positive octant, remember the flips in `a`
    int a = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o[i] = vp[i];
        d[i] = ddiv(vd[i], len);
        if (d[i] < 0.0) { o[i] = dsub(p.size[i], o[i]); d[i] = -d[i]; a |= 4 >> i; }
    }
    const double inv_len = 1.0 / len;   // FAST only: t / |v_dir| as a product
    double st0[OCT_MAX_LEVELS][3], st1[OCT_MAX_LEVELS][3];
    int snode[OCT_MAX_LEVELS], snext[OCT_MAX_LEVELS];
#pragma unroll
    for (int i = 0; i < 3; i++) { st0[0][i] = ddiv(dsub(p.dle[i], o[i]), d[i]); st1[0][i] = ddiv(dsub(p.dre[i], o[i]), d[i]); }
    double mx = st0[0][0], mn = st1[0][0];
#pragma unroll
    for (int i = 1; i < 3; i++) { if (mx < st0[0][i]) mx = st0[0][i]; if (st1[0][i] < mn) mn = st1[0][i]; }
    if (!(mx < mn)) return;
    TFShared64 t64;
    t64.values = p.tf.values; t64.d0 = p.tf.d0; t64.dy = p.tf.dy;
    double acc[4];
    float acc32[4];
    bool touched = false;
    double *out = p.image + px * 4;
#pragma unroll
    for (int c = 0; c < 4; c++) { acc[c] = out[c]; acc32[c] = (float)acc[c]; }
    // The current node's frame lives in registers; the stack holds only those ancestors that still have a child to
    // visit (a node whose last child is being entered is not stacked: the recursion's tail call), so terminal cells --
    // most of the nodes a ray meets -- never touch local memory.
    double tx0 = st0[0][0], ty0 = st0[0][1], tz0 = st0[0][2], tx1 = st1[0][0], ty1 = st1[0][1], tz1 = st1[0][2];
    int n = 0, sp = 0, i = 0;
    for (;;) {
        bool restacked = false;
        {
            // proc_subtree prologue (hpp:270-289)
            bool done = tx1 < 0 || ty1 < 0 || tz1 < 0 || n < 0;
            if (!done && p.leaf[2 * n]) {
                const double t_in = std_max(std_max(tx0, ty0), tz0), t_out = std_min(std_min(tx1, ty1), tz1);
                const double *cell = p.cells + 14ll * p.leaf[2 * n + 1];
                if (FAST) { oct_sample_cell_fast(p, s_oct_pairs, cell, vp, vd, t_in * inv_len, t_out * inv_len, acc32); touched = true; }
                else oct_sample_cell<KIND>(p, t64, cell, vp, vd, ddiv(t_in, len), ddiv(t_out, len), acc);
                done = true;
            }
            if (done) {
                // return to the nearest ancestor with a child left
                if (sp == 0) break;
                sp--;
                tx0 = st0[sp][0]; ty0 = st0[sp][1]; tz0 = st0[sp][2];
                tx1 = st1[sp][0]; ty1 = st1[sp][1]; tz1 = st1[sp][2];
                n = snode[sp]; i = snext[sp];
                restacked = true;   // slot sp still holds this frame's t0 / t1 / node
            }
        }
        const double txm = dmul(dadd(tx0, tx1), 0.5), tym = dmul(dadd(ty0, ty1), 0.5), tzm = dmul(dadd(tz0, tz1), 0.5);
        if (!restacked) i = oct_first_node(tx0, ty0, tz0, txm, tym, tzm);
        // child i: lower half of an axis when its bit is clear, upper half when set; its successor along an axis is
        // i with that bit set, or 8 (leave the node) -- the eight cases of the switch at hpp:301-341
        const double c0x = (i & 4) ? txm : tx0, c1x = (i & 4) ? tx1 : txm;
        const double c0y = (i & 2) ? tym : ty0, c1y = (i & 2) ? ty1 : tym;
        const double c0z = (i & 1) ? tzm : tz0, c1z = (i & 1) ? tz1 : tzm;
        const int nxt = oct_next_node(c1x, c1y, c1z, (i & 4) ? 8 : (i | 4), (i & 2) ? 8 : (i | 2), (i & 1) ? 8 : (i | 1));
        if (nxt < 8) {
            if (sp >= OCT_MAX_LEVELS) break;   // deeper than any tree the host accepts
            if (!restacked) {
                st0[sp][0] = tx0; st0[sp][1] = ty0; st0[sp][2] = tz0;
                st1[sp][0] = tx1; st1[sp][1] = ty1; st1[sp][2] = tz1;
                snode[sp] = n;
            }
            snext[sp] = nxt;
            sp++;
        }
        n = p.child[8 * n + (i ^ a)];
        tx0 = c0x; ty0 = c0y; tz0 = c0z; tx1 = c1x; ty1 = c1y; tz1 = c1z;
    }
    if (FAST) {
        // channels the sampler never writes (alpha of a non-grey render, pixels whose ray met no cell) keep their fp64
        // initial value bit for bit
        if (touched) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < 3 || p.tf.grey) out[c] = (double)acc32[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) out[c] = acc[c];
    }
}

}  // namespace ytb
