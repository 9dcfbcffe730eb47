// Host side of the slab-walk projection cast: lane -> pixel map, tile order, launch.
#include "proj.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "proj_kernels.cuh"

namespace ytb {

// Lane -> pixel map and tile order of the projection cast (proj_kernels.cuh).  For a plane-parallel camera whose
// dominant axis (in cell units of the first brick of the list) is x or y, warps become sheared rows of 32 pixels
// along the image direction in which the rays' crossing points of a slab plane differ only in z.
ProjMap make_proj_map(const DevCamera &c, const DevBrick *first, bool host_image)
{
    ProjMap m{};
    m.mode = 0;
    m.n_ftiles = (c.ny + TILE_J - 1) / TILE_J;
    m.n_stiles = (c.nx + TILE_I - 1) / TILE_I;
    m.band = 24;
    const char *env = getenv("YTB200_PROJ_MAP");
    if (env && atoi(env) == 0) return m;
    // a host image is read and written in place over PCIe: 16x16 tiles keep 256 contiguous bytes per 8 lanes
    if (host_image) return m;
    if (c.lens != 0 /* plane-parallel */ || c.nx < 2 || c.ny < 2 || !first) return m;
    const DevBrick &b = *first;
    double dc[3];
    for (int k = 0; k < 3; k++) dc[k] = c.vp_dir[k] * b.idds[k];
    const double r0 = fabs(dc[0]), r1 = fabs(dc[1]), r2 = fabs(dc[2]);
    const int a = (r0 >= r1 && r0 >= r2) ? 0 : (r1 >= r2 ? 1 : 2);
    if (a == 2 || !(fabs(dc[a]) > 0.0)) return m;
    const int u = a == 0 ? 1 : 0;
    double ei[3], ej[3];
    for (int k = 0; k < 3; k++) {
        ei[k] = c.vp_pos[k] * (c.width[0] / (c.nx - 1.0)) * b.idds[k];
        ej[k] = c.vp_pos[3 + k] * (c.width[1] / (c.ny - 1.0)) * b.idds[k];
    }
    // crossing point of a slab plane per pixel step: e - (e_a / d_a) d
    const double Ui = ei[u] - ei[a] / dc[a] * dc[u], Uj = ej[u] - ej[a] / dc[a] * dc[u];
    double slope;
    if (fabs(Uj) >= fabs(Ui) && Uj != 0.0) { m.f_is_j = 0; slope = -Ui / Uj; m.nf = c.nx; m.ns = c.ny; }
    else if (Ui != 0.0) { m.f_is_j = 1; slope = -Uj / Ui; m.nf = c.ny; m.ns = c.nx; }
    else return m;
    if (!std::isfinite(slope) || fabs(slope) > 1.0) return m;
    m.mode = 1;
    m.slope = (float)slope;
    const int sh31 = (int)lrintf(m.slope * 31.f);
    const int shmin = std::min(sh31, 0), shmax = std::max(sh31, 0);
    m.s_lo = -shmax;
    m.n_ftiles = (m.nf + 31) / 32;
    m.n_stiles = (m.ns + shmax - shmin + CTA_THREADS / 32 - 1) / (CTA_THREADS / 32);
    m.band = 48;
    return m;
}

template <int NF>
static cudaError_t launch_nf(const CastParams &p, const ProjMap &m, bool partial, cudaStream_t st)
{
    const int bands = (m.n_stiles + m.band - 1) / m.band;
    const unsigned grid = (unsigned)(m.n_ftiles * m.band * bands);
    if (partial) proj_kernel<NF, true><<<grid, CTA_THREADS, 0, st>>>(p, m);
    else proj_kernel<NF, false><<<grid, CTA_THREADS, 0, st>>>(p, m);
    return cudaGetLastError();
}

cudaError_t launch_proj(const CastParams &p, const ProjMap &m, int n_fields, bool partial, cudaStream_t st)
{
    switch (n_fields) {
    case 1: return launch_nf<1>(p, m, partial, st);
    case 2: return launch_nf<2>(p, m, partial, st);
    case 3: return launch_nf<3>(p, m, partial, st);
    default: return launch_nf<4>(p, m, partial, st);
    }
}

}  // namespace ytb
