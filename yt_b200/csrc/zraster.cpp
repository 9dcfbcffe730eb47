// Opaque-source rasterisers: depth-tested lines and points painted into the (rgba, z) buffer that the volume
// cast then starts from (initial colours, image_samplers.pyx:183-184) and stops at (max_t, :185).
//
// What is reproduced, statement for statement: zlines and zpoints of yt/utilities/lib/misc_utilities.pyx:244-404
// (Bresenham stepping with C-int error terms, the depth increments dzx/dzy, the blend
// rgb = (a*premult + rgb*a_old*(1-a)) / a_new, the "opaque enough" threshold 1 - alpha < 1e-4 that writes the
// z-buffer, `thick` repeating the y-run once per unit of thickness exactly as the reference's unused loop index does).
// Both are order-dependent serial painters over a few thousand pixels -- host code in the reference (Cython) and
// here (C++); their OUTPUT is what the device consumes.  No CUDA in this file.
#include "../../include/ytb200.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace {

inline void paint(double *px, double *zb, const double alpha[4], double z0, bool lines)
{
    if (!(z0 < *zb)) return;
    if (alpha[3] != 1.0) {
        const double talpha = px[3];
        px[3] = alpha[3] + talpha * (1 - alpha[3]);
        for (int i = 0; i < 3; i++) {
            if (lines) {
                if (px[3] == 0.0) px[i] = 0.0;
                else px[i] = (alpha[3] * alpha[i] + px[i] * talpha * (1.0 - alpha[3])) / px[3];
            } else {
                px[i] = (alpha[3] * alpha[i] + px[i] * talpha * (1.0 - alpha[3])) / px[3];
                if (px[3] == 0.0) px[i] = 0.0;
            }
        }
    } else {
        for (int i = 0; i < 4; i++) px[i] = alpha[i];
    }
    if (1.0 - px[3] < 1.0e-4) {
        if (lines) px[3] = 1.0;   // zlines snaps alpha (misc_utilities.pyx:318-320); zpoints does not (:400-401)
        *zb = z0;
    }
}

}  // namespace

extern "C" int ytb200_zlines(double *image, double *zbuffer, int64_t nx_, int64_t ny_, const int64_t *xs, const int64_t *ys,
                             const double *zs, int64_t n_points, const double *colors, int64_t n_colors,
                             int points_per_color, int thick, int flip, int crop)
{
    if (!image || !zbuffer || (n_points > 0 && (!xs || !ys || !zs || !colors))) return -1;
    if (points_per_color < 1 || nx_ < 1 || ny_ < 1 || nx_ > 0x7fffffff || ny_ > 0x7fffffff) return -1;
    const int nx = (int)nx_, ny = (int)ny_;
    const int nl = (int)n_points;
    double alpha[4] = {0.0, 0.0, 0.0, 0.0};
    if (nl & 1) return -1;   // the reference would read past the arrays: refuse an odd point count
    for (int j = 0; j < nl; j += 2) {
        int64_t x0 = xs[j], y0 = ys[j];
        const int64_t x1 = xs[j + 1], y1 = ys[j + 1];
        double z0 = zs[j];
        const double z1 = zs[j + 1];
        const int dx = (int)std::llabs(x1 - x0), dy = (int)std::llabs(y1 - y0);
        const double dzx = (z1 - z0) / (dx * dx + dy * dy) * dx;
        const double dzy = (z1 - z0) / (dx * dx + dy * dy) * dy;
        int err = dx - dy;
        if (crop == 1 && (dx > nx / 2.0 || dy > ny / 2.0)) continue;
        const int c = j / points_per_color / 2;
        if (c < 0 || c >= n_colors) return -1;
        for (int i = 0; i < 3; i++) alpha[i] = colors[(size_t)c * 4 + i] * colors[(size_t)c * 4 + 3];
        alpha[3] = colors[(size_t)c * 4 + 3];
        const int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
        while (true) {
            if (x0 < thick && sx == -1) break;
            else if (x0 >= nx - thick + 1 && sx == 1) break;
            else if (y0 < thick && sy == -1) break;
            else if (y0 >= ny - thick + 1 && sy == 1) break;
            if (x0 >= thick && x0 < nx - thick && y0 >= thick && y0 < ny - thick) {
                for (int64_t rep = x0 - thick / 2; rep < x0 + (1 + thick) / 2; rep++) {
                    for (int64_t yi = y0 - thick / 2; yi < y0 + (1 + thick) / 2; yi++) {
                        const int64_t yi0 = flip ? ny - yi : yi;
                        if (yi0 < 0 || yi0 >= ny) continue;   // (the reference indexes out of bounds here with flip; we skip)
                        const size_t p = (size_t)x0 * ny + (size_t)yi0;
                        paint(image + p * 4, zbuffer + p, alpha, z0, true);
                    }
                }
            }
            if (x0 == x1 && y0 == y1) break;
            const int e2 = 2 * err;
            if (e2 > -dy) { err = err - dy; x0 += sx; z0 += dzx; }
            if (e2 < dx) { err = err + dx; y0 += sy; z0 += dzy; }
        }
    }
    return 0;
}

extern "C" int ytb200_zpoints(double *image, double *zbuffer, int64_t nx_, int64_t ny_, const int64_t *xs, const int64_t *ys,
                              const double *zs, int64_t n_points, const int64_t *order, const double *colors,
                              int64_t n_colors, const int64_t *radii, int points_per_color, int flip)
{
    if (!image || !zbuffer || (n_points > 0 && (!xs || !ys || !zs || !colors || !radii || !order))) return -1;
    if (points_per_color < 1 || nx_ < 1 || ny_ < 1) return -1;
    const int64_t nx = nx_, ny = ny_;
    double alpha[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t k = 0; k < n_points; k++) {
        const int64_t j = order[k];   // np.argsort(zs) of the caller: far-to-near painting (misc_utilities.pyx:362-364)
        if (j < 0 || j >= n_points) return -1;
        const int64_t r = radii[j];
        const int64_t r2 = (int64_t)((r + 0.3) * (r + 0.3));
        const double z0 = zs[j];
        for (int64_t kx = -r; kx <= r; kx++) {
            const int64_t x0 = xs[j] + kx;
            if (x0 < 0 || x0 >= nx) continue;
            for (int64_t ky = -r; ky <= r; ky++) {
                const int64_t y0 = ys[j] + ky;
                if (y0 < 0 || y0 >= ny) continue;
                if (kx * kx + ky * ky > r2) continue;
                const int64_t c = j / points_per_color;
                if (c >= n_colors) return -1;
                for (int i = 0; i < 3; i++) alpha[i] = colors[(size_t)c * 4 + i] * colors[(size_t)c * 4 + 3];
                alpha[3] = colors[(size_t)c * 4 + 3];
                const int64_t yi0 = flip ? ny - y0 : y0;
                if (yi0 < 0 || yi0 >= ny) continue;
                const size_t p = (size_t)x0 * (size_t)ny + (size_t)yi0;
                paint(image + p * 4, zbuffer + p, alpha, z0, false);
            }
        }
    }
    return 0;
}
