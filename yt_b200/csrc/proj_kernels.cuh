// The ProjectionSampler cast (image_samplers.pyx:327-343 inside walk_volume, grid_traversal.pyx:26-140) as a
// SLAB WALK: sum over cells of value * (path length in t through the cell), the integral the reference's DDA
// accumulates, evaluated in a different order and arithmetic.  The projection carries the rtol 1e-6 bar, not bit
// parity, and it is order-free -- so nothing forces the reference's one-cell-at-a-time fp64 DDA (40 instructions and
// 11 fp64-pipe operations per cell, ncu r01zo: issue 67 %, fp64 pipe 34 %, DRAM 2.45x re-read).
//
//   * A ray advances along its DOMINANT axis a (largest |direction| in cell units) one cell plane at a time.  Inside
//     the slab between two planes it crosses at most one cell boundary on each minor axis (u, v), i.e. it visits at
//     most three cells: A (before both crossings), B (between them), C (after both).  The slab's contribution is
//     L * (vA*wA + vB*wB + vC*wC) with L the slab length in t and the weights the fractions of the slab spent in each.
//   * Positions on the minor axes are 32.32 FIXED POINT: `frac += step` is one add-with-carry, the carry IS "the ray
//     crossed a cell boundary in this slab" and the new fraction divided by the step is the share of the slab spent
//     after the crossing.  No fp64 in the loop, no comparisons of three tmax, ~30 instructions per SLAB (three cells).
//     The state is re-seeded from fp64 every 256 slabs, so the position error stays below 256 * 2^-33 = 3e-8 cells;
//     weights and values are fp32, slab sums are accumulated in fp64 (fp64 accumulation).
//   * The first and the last slab of a ray inside a brick are clipped by the entry / exit point (and by max_t): they
//     take a general fp64 path with clamped cell indices.  Rays the fast loop cannot guarantee in-bounds reads for
//     (closer than 1e-7 cells to a brick face at the end of their interior) take the general path for every slab.
//   * LOCKSTEP AND WARP SHAPE.  All lanes of a warp process the same slab index in the same iteration (the loop runs
//     over the warp's union of slab ranges; a lane outside its own range is predicated off), and the host picks the
//     lane -> pixel map: for a plane-parallel camera whose dominant axis is not z, a warp is a SHEARED ROW of 32
//     pixels along the image direction in which the rays' positions on a slab plane differ only in z (the contiguous
//     axis): 32 lanes then read 2-3 contiguous runs of one or two z-rows per load instead of 17 scattered sectors.
//   * Tiles are issued in bands so that the CTAs resident at one time cover a compact block of the image: cells
//     shared by neighbouring tiles are still in L2 when the neighbour reads them.
//
// Decisions that are discrete in the reference and not measure-zero are kept: a ray that does not move along an
// axis takes the cell floor((p - le) * idds) there and misses the brick when p lies on its upper face
// (grid_traversal.pyx:44-49, 104-117); on exact ties the higher axis crosses first (grid_traversal.pyx:122-131).
#pragma once
#include "proj.h"
#include "raycast.cuh"

namespace ytb {

#ifndef YTB_PROJ_UNROLL
#define YTB_PROJ_UNROLL 3      // slabs per group: 3 * UNROLL loads of one ray per register stage
                               // r02w/r02x A/B on cfg3 (3 CTAs/SM unless noted): U=2 2.22 ms, U=3 1.99 ms (80 registers, 32 B
                               // spilled), U=4 2.11; U=2 at 4 CTAs/SM (64 registers) 2.02, U=3 there 2.02, U=6 at 2 CTAs/SM 1.99
#endif
#ifndef YTB_PROJ_MINB
#define YTB_PROJ_MINB 3        // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef YTB_PROJ_PIPE
#define YTB_PROJ_PIPE 1        // 1: the loads of group g+1 are issued before group g is summed (two register stages)
                               // r02d A/B on cfg3: 2.79 ms unpipelined (U=4), 2.21 ms pipelined; U=4 + 2 CTAs/SM the same
#endif
#ifndef YTB_PROJ_PF
#define YTB_PROJ_PF 0          // > 0: prefetch.global.L2 of the A cell this many slabs ahead
#endif
constexpr int PROJ_CHUNK = 256;       // slabs between two re-seeds of the fixed-point state
constexpr double PROJ_FACE_MARGIN = 1e-7;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <int NF>
struct T_ptrs {   // the brick's field arrays (fp32 storage: the product path)
    const float *d[NF];
    __device__ __forceinline__ explicit T_ptrs(const DevBrick *b)
    {
#pragma unroll
        for (int f = 0; f < NF; f++) d[f] = static_cast<const float *>(b->data[f]);
    }
};

template <int NF, bool PARTIAL>
__global__ void __launch_bounds__(CTA_THREADS, YTB_PROJ_MINB) proj_kernel(const __grid_constant__ CastParams p, const __grid_constant__ ProjMap map)
{
    __shared__ int s_list[CTA_THREADS];
    __shared__ int s_warp_cnt[CTA_THREADS / 32];
    __shared__ double s_ray[6 * CTA_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- tile of this CTA (band order) and pixel of this thread -----------------------------------------
    int F, S;
    {
        const int per_band = map.n_ftiles * map.band;
        const int b = blockIdx.x / per_band, r = blockIdx.x - b * per_band;
        const int h = min(map.band, map.n_stiles - b * map.band);   // s-tiles in this band
        F = r / h;
        S = b * map.band + (r - F * h);
        if (F >= map.n_ftiles) {   // the last band is shorter: its surplus CTAs have no tile (they still count for the tail signal)
            if (PARTIAL && p.tail_flag != nullptr) {
                SwapSignal sg;
                sg.n = 1; sg.flag[0] = p.tail_flag;
                grid_tail_signal(p.tail_counter, sg, p.tail_seq);
            }
            return;
        }
    }
    int vi, vj, ti0, ti1, tj0, tj1;
    if (map.mode == 0) {
        S += map.s_first;
        ti0 = S * TILE_I; tj0 = F * TILE_J;
        ti1 = ti0 + TILE_I; tj1 = tj0 + TILE_J;
        vi = ti0 + (warp >> 1) * 4 + (lane >> 3);
        vj = tj0 + (warp & 1) * 8 + (lane & 7);
    } else {
        const int f = F * 32 + lane;
        const int s = map.s_lo + S * (CTA_THREADS / 32) + warp + __float2int_rn(map.slope * (float)lane);
        const int sh31 = __float2int_rn(map.slope * 31.f);
        const int f0 = F * 32, f1 = f0 + 32;
        const int s0 = map.s_lo + S * (CTA_THREADS / 32) + min(sh31, 0), s1 = map.s_lo + S * (CTA_THREADS / 32) + CTA_THREADS / 32 + max(sh31, 0);
        if (map.f_is_j) { vi = s; vj = f; ti0 = s0; ti1 = s1; tj0 = f0; tj1 = f1; }
        else            { vi = f; vj = s; ti0 = f0; ti1 = f1; tj0 = s0; tj1 = s1; }
    }
    const bool active = vi >= 0 && vj >= 0 && vi < p.cam.nx && vj < p.cam.ny;
    const size_t pix = (size_t)(active ? vi : 0) * p.cam.ny + (active ? vj : 0);

    if (!tile_has_bricks(p.extent, p.chunk_extent, p.n_order, ti0, ti1, tj0, tj1)) {
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

    double max_t = 1.0;
    double acc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) acc[f] = 0.0;
    {
        double vp[3] = {0.0, 0.0, 0.0}, vd[3] = {0.0, 0.0, 0.0};
        if (active) {
            make_ray(p.cam, vi, vj, vp, vd);
            if (p.zbuffer) { const double z = p.zbuffer[pix]; max_t = fmin(fmax(z, 0.0), 1.0); }
            if (!PARTIAL) {
                const double4 px = *reinterpret_cast<const double4 *>(p.image + pix * 4);
                acc[0] = px.x;
                if (NF > 1) acc[NF > 1 ? 1 : 0] = px.y;
                if (NF > 2) acc[NF > 2 ? 2 : 0] = px.z;
                if (NF > 3) acc[NF > 3 ? 3 : 0] = px.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) { s_ray[i * CTA_THREADS + tid] = vp[i]; s_ray[(3 + i) * CTA_THREADS + tid] = vd[i]; }
    }
    __syncthreads();

    // ---- the ordered brick list, culled against this tile's pixel box (order is irrelevant for a sum) ----------
    for (int lbase = 0; lbase < p.n_order; lbase += CTA_THREADS) {
        const int kk = lbase + tid;
        bool hitb = false;
        int b = -1;
        if (kk < p.n_order) {
            const int4 e = p.extent[kk];
            hitb = e.x < ti1 && e.y > ti0 && e.z < tj1 && e.w > tj0;
            b = p.order[kk];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hitb);
        if (lane == 0) s_warp_cnt[warp] = __popc(m);
        __syncthreads();
        int loff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CTA_THREADS / 32; w++) {
            const int cnt = s_warp_cnt[w];
            if (w < warp) loff += cnt;
            total += cnt;
        }
        if (hitb) s_list[loff + __popc(m & ((1u << lane) - 1u))] = b;
        __syncthreads();

        for (int q = 0; q < total; q++) {
            // every lane of the warp goes through the whole body (the slab loop is warp-collective); a lane without
            // a ray, or whose ray misses the brick, carries an empty slab range
            const DevBrick *br = p.bricks + s_list[q];
            double t0 = 0.0, t1 = max_t;
            bool hit = active;
            double vp[3], vd[3], rate[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                vp[c] = s_ray[c * CTA_THREADS + tid]; vd[c] = s_ray[(3 + c) * CTA_THREADS + tid];
                const double le = br->le[c], re = br->re[c];
                if (vd[c] == 0.0) {
                    // the reference takes the cell floor((p - le) * idds) and leaves when it is not one of the brick's
                    const int idx = __double2int_rd((vp[c] - le) * br->idds[c]);
                    hit = hit && le <= vp[c] && vp[c] < re && idx >= 0 && idx < br->dims[c];
                } else {
                    const double iv = 1.0 / vd[c];
                    const double ta = (le - vp[c]) * iv, tb = (re - vp[c]) * iv;
                    t0 = fmax(t0, fmin(ta, tb));
                    t1 = fmin(t1, fmax(ta, tb));
                }
                rate[c] = fabs(vd[c]) * br->idds[c];
            }
            hit = hit && t0 < t1;
            // dominant axis (ties keep z a minor axis: it is the contiguous one)
            const int a = (rate[0] >= rate[1] && rate[0] >= rate[2]) ? 0 : (rate[1] >= rate[2] ? 1 : 2);
            hit = hit && rate[0] + rate[1] + rate[2] > 0.0;
            const int ua = a == 0 ? 1 : 0, va = a == 2 ? 1 : 2;   // (a,u,v) = (0,1,2) (1,0,2) (2,0,1)
            // ray-oriented cell coordinates at t0: every coordinate increases (or stays) along the ray
            double qo[3];
            int dimc[3], strd[3];
            unsigned base = 0u;
            {
                const int px = br->pitch_x, py = br->pitch_y;
                const int st[3] = {px, py, 1};
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int d = br->dims[c];
                    double qc = ((vp[c] + t0 * vd[c]) - br->le[c]) * br->idds[c];
                    const bool neg = vd[c] < 0.0;
                    if (neg) { qc = (double)d - qc; base += (unsigned)(d - 1) * (unsigned)st[c]; }
                    // a moving coordinate is nudged by 1e-9 cells: rounding can then never put the ray one cell
                    // outside the face it has just entered through (a crossing moves by 1e-9 of a cell width)
                    if (vd[c] != 0.0) qc += 1e-9;
                    qo[c] = qc; dimc[c] = d; strd[c] = neg ? -st[c] : st[c];
                }
            }
            const double ra = a == 0 ? rate[0] : (a == 1 ? rate[1] : rate[2]);
            const double ira = hit ? 1.0 / ra : 0.0;
            const double su = (ua == 0 ? rate[0] : rate[1]) * ira, sv = (va == 1 ? rate[1] : rate[2]) * ira;
            const double A0 = a == 0 ? qo[0] : (a == 1 ? qo[1] : qo[2]);
            const double QU = ua == 0 ? qo[0] : qo[1], QV = va == 1 ? qo[1] : qo[2];
            const int da = a == 0 ? dimc[0] : (a == 1 ? dimc[1] : dimc[2]);
            const int du = ua == 0 ? dimc[0] : dimc[1], dv = va == 1 ? dimc[1] : dimc[2];
            const int Sa = a == 0 ? strd[0] : (a == 1 ? strd[1] : strd[2]);
            const int Su = ua == 0 ? strd[0] : strd[1], Sv = va == 1 ? strd[1] : strd[2];
            const double A1 = A0 + (t1 - t0) * ra;
            int k0 = 0, k1 = -1;
            if (hit) {
                k0 = clampi(__double2int_rd(A0), 0, da - 1);
                k1 = clampi(__double2int_ru(A1) - 1, k0, da - 1);
            }
            const T_ptrs<NF> dp(br);
            double bsum[NF];
#pragma unroll
            for (int f = 0; f < NF; f++) bsum[f] = 0.0;

            // ---- one slab on the general path: slab k, slab-local range [s0, s1] ---------------------------------
            auto general = [&](int k, double s0, double s1) {
                const double Uk = QU + ((double)k - A0) * su, Vk = QV + ((double)k - A0) * sv;
                const int iu0 = clampi(__double2int_rd(Uk + s0 * su), 0, du - 1), iv0 = clampi(__double2int_rd(Vk + s0 * sv), 0, dv - 1);
                const int iu1 = clampi(__double2int_rd(Uk + s1 * su), iu0, min(iu0 + 1, du - 1));
                const int iv1 = clampi(__double2int_rd(Vk + s1 * sv), iv0, min(iv0 + 1, dv - 1));
                double cu = s1, cv = s1;
                if (iu1 > iu0) cu = fmin(fmax(((double)iu1 - Uk) / su, s0), s1);
                if (iv1 > iv0) cv = fmin(fmax(((double)iv1 - Vk) / sv, s0), s1);
                const double fa = fmin(cu, cv), fb = fmax(cu, cv);
                const double wA = fa - s0, wB = fb - fa, wC = s1 - fb;
                const unsigned oK = base + (unsigned)k * (unsigned)Sa;
                const unsigned oA = oK + (unsigned)iu0 * (unsigned)Su + (unsigned)iv0 * (unsigned)Sv;
                const unsigned oC = oK + (unsigned)iu1 * (unsigned)Su + (unsigned)iv1 * (unsigned)Sv;
                const unsigned oB = cu < cv ? oK + (unsigned)iu1 * (unsigned)Su + (unsigned)iv0 * (unsigned)Sv
                                            : oK + (unsigned)iu0 * (unsigned)Su + (unsigned)iv1 * (unsigned)Sv;
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    // a cell the ray spends no time in is not read as a number (NaN * 0)
                    double v = 0.0;
                    if (wA > 0.0) v = (double)__ldg(dp.d[f] + oA) * wA;
                    if (wB > 0.0) v = fma((double)__ldg(dp.d[f] + oB), wB, v);
                    if (wC > 0.0) v = fma((double)__ldg(dp.d[f] + oC), wC, v);
                    bsum[f] += v;
                }
            };

            // ---- split: first / last slab general, interior by the fixed-point loop ----------------------------------
            int ks = 0x3fffffff, ke = -0x3fffffff;   // interior slabs [ks, ke)
            if (hit) {
                const double s0 = fmin(fmax(A0 - (double)k0, 0.0), 1.0), s1 = fmin(fmax(A1 - (double)k1, 0.0), 1.0);
                bool fast = k1 - k0 >= 2;
                if (fast) {
                    // the fixed-point loop reads cell (iu + 1) as soon as its own position crosses a boundary: keep it
                    // away from rays that end their interior within 1e-7 cells of a face they have not left through
                    const double Ue = QU + ((double)k1 - A0) * su, Ve = QV + ((double)k1 - A0) * sv;
                    if ((su > 0.0 && Ue > (double)du - PROJ_FACE_MARGIN) || (sv > 0.0 && Ve > (double)dv - PROJ_FACE_MARGIN)) fast = false;
                }
                for (int k = k0; k <= k1; k = (fast && k == k0) ? k1 : k + 1)
                    general(k, k == k0 ? s0 : 0.0, k == k1 ? s1 : 1.0);
                if (fast) { ks = k0 + 1; ke = k1; }
            }
            const int K0 = __reduce_min_sync(0xffffffffu, ks), K1 = __reduce_max_sync(0xffffffffu, ke);
            if (K0 < K1) {
                // per-lane constants of the fixed-point walk; a lane with an empty range gets harmless values
                const bool mine = ks < ke;
                const double su_ = mine ? su : 0.0, sv_ = mine ? sv : 0.0;
                const unsigned su32 = su_ >= 1.0 ? 0u : __double2uint_rn(su_ * 4294967296.0);
                const unsigned sv32 = sv_ >= 1.0 ? 0u : __double2uint_rn(sv_ * 4294967296.0);
                const unsigned alw_u = su_ >= 1.0 ? 1u : 0u, alw_v = sv_ >= 1.0 ? 1u : 0u;   // step == one cell: a crossing in every slab
                // share of the slab left after a crossing = new fraction / step
                const float inv_u = su_ > 0.0 ? (float)(1.0 / (su_ * 4294967296.0)) : 0.f;
                const float inv_v = sv_ > 0.0 ? (float)(1.0 / (sv_ * 4294967296.0)) : 0.f;
                const unsigned cnt = mine ? (unsigned)(ke - ks) : 0u;
                float fsum[NF];
                for (int kc = K0; kc < K1;) {
                    const int kend = min(K1, (kc | (PROJ_CHUNK - 1)) + 1);
                    // seed at slab kc (extrapolated along the ray for a lane whose own range starts later)
                    unsigned fu, fv, off;
                    {
                        const double U = mine ? QU + ((double)kc - A0) * su : 0.0, V = mine ? QV + ((double)kc - A0) * sv : 0.0;
                        const double Uf = floor(U), Vf = floor(V);
                        fu = __double2uint_rz((U - Uf) * 4294967296.0);
                        fv = __double2uint_rz((V - Vf) * 4294967296.0);
                        off = base + (unsigned)kc * (unsigned)Sa + (unsigned)(int)Uf * (unsigned)Su + (unsigned)(int)Vf * (unsigned)Sv;
                    }
#pragma unroll
                    for (int f = 0; f < NF; f++) fsum[f] = 0.f;
                    int k = kc;
                    constexpr int U = YTB_PROJ_UNROLL;
#if YTB_PROJ_PF > 0
                    // offset of the A cell YTB_PROJ_PF slabs ahead of the current one (to the nearest cell)
                    const int pf_delta = YTB_PROJ_PF * Sa + __double2int_rn(YTB_PROJ_PF * su_) * Su + __double2int_rn(YTB_PROJ_PF * sv_) * Sv;
#endif
                    // one slab: advance the fixed-point state, form the weights (as shares ru, rv of the slab left after
                    // the u / v crossing) and request the three cells
                    auto issue = [&](int kq, float (&v)[NF][3], float &ru, float &rv) {
                        const bool act = (unsigned)(kq - ks) < cnt;
                        unsigned cu, cv;
                        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %3, 0;" : "+r"(fu), "=r"(cu) : "r"(su32), "r"(alw_u));
                        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %3, 0;" : "+r"(fv), "=r"(cv) : "r"(sv32), "r"(alw_v));
                        ru = __uint2float_rn(fu * cu) * inv_u;
                        rv = __uint2float_rn(fv * cv) * inv_v;
                        const unsigned oU = off + cu * (unsigned)Su, oV = off + cv * (unsigned)Sv;
                        const unsigned oC = oU + cv * (unsigned)Sv;
                        const unsigned oB = ru > rv ? oU : oV;   // a tie lets the higher axis cross first, like the reference
#pragma unroll
                        for (int f = 0; f < NF; f++) {
                            v[f][0] = v[f][1] = v[f][2] = 0.f;
                            if (act) { v[f][0] = __ldg(dp.d[f] + off); v[f][1] = __ldg(dp.d[f] + oB); v[f][2] = __ldg(dp.d[f] + oC); }
                        }
#if YTB_PROJ_PF > 0
                        if (act) asm volatile("prefetch.global.L2 [%0];" :: "l"(dp.d[0] + (off + (unsigned)pf_delta)));
#endif
                        off = oC + (unsigned)Sa;
                    };
                    auto consume = [&](const float (&v)[NF][3], float ru, float rv) {
                        const float rmin = fminf(ru, rv), rmax = fmaxf(ru, rv);
                        const float wA = 1.f - rmax, wB = rmax - rmin;
#pragma unroll
                        for (int f = 0; f < NF; f++) fsum[f] = fmaf(v[f][2], rmin, fmaf(v[f][1], wB, fmaf(v[f][0], wA, fsum[f])));
                    };
                    auto flush = [&]() {
#pragma unroll
                        for (int f = 0; f < NF; f++) { bsum[f] += (double)fsum[f]; fsum[f] = 0.f; }
                    };
#if YTB_PROJ_PIPE
                    // two register stages: a lane past the end of the chunk (or of its own range) is predicated off, so the
                    // group count is simply rounded up
                    float va[U][NF][3], vb[U][NF][3], rua[U], rva[U], rub[U], rvb[U];
                    const int ngroups = (kend - kc + U - 1) / U;
                    auto issue_group = [&](int k0g, float (&v)[U][NF][3], float (&ru)[U], float (&rv)[U]) {
#pragma unroll
                        for (int j = 0; j < U; j++) issue(k0g + j < kend ? k0g + j : 0x7fffffff, v[j], ru[j], rv[j]);
                    };
                    auto consume_group = [&](const float (&v)[U][NF][3], const float (&ru)[U], const float (&rv)[U]) {
#pragma unroll
                        for (int j = 0; j < U; j++) consume(v[j], ru[j], rv[j]);
                        flush();
                    };
                    issue_group(k, va, rua, rva);
                    for (int g = 0; g < ngroups; g += 2) {
                        issue_group(k + (g + 1) * U, vb, rub, rvb);
                        consume_group(va, rua, rva);
                        issue_group(k + (g + 2) * U, va, rua, rva);
                        consume_group(vb, rub, rvb);
                    }
#else
                    float v[NF][3], ru, rv;
                    for (; k + U <= kend; k += U) {
#pragma unroll
                        for (int j = 0; j < U; j++) { issue(k + j, v, ru, rv); consume(v, ru, rv); }
                        flush();
                    }
                    for (; k < kend; k++) { issue(k, v, ru, rv); consume(v, ru, rv); }
                    flush();
#endif
                    kc = kend;
                }
            }
            if (hit) {
#pragma unroll
                for (int f = 0; f < NF; f++) acc[f] = fma(bsum[f], ira, acc[f]);
            }
        }
        __syncthreads();
    }

    // ---- write back -----------------------------------------------------------------------------------------
    if (active) {
        double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int f = 0; f < NF; f++) out[f] = acc[f];
        if (PARTIAL) {
            const double one[4] = {1.0, 1.0, 1.0, 1.0};
            store_partial(p.partial, p.partial_fmt, pix, out, one);
        } else {
            if (NF < 4) {
                // channels the sampler does not touch keep their value (image_samplers.pyx:337-343 sums min(4, n_fields))
                const double4 px = *reinterpret_cast<const double4 *>(p.image + pix * 4);
                if (NF < 2) out[1] = px.y;
                if (NF < 3) out[2] = px.z;
                out[3] = px.w;
            }
            *reinterpret_cast<double4 *>(p.image + pix * 4) = make_double4(out[0], out[1], out[2], out[3]);
        }
    }
    if (PARTIAL && p.tail_flag != nullptr) {
        SwapSignal sg;
        sg.n = 1; sg.flag[0] = p.tail_flag;
        grid_tail_signal(p.tail_counter, sg, p.tail_seq);
    }
}

}  // namespace ytb
