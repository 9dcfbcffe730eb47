// libytb200.so -- C ABI (include/ytb200.h) over the cast kernels.  Host side only does what
// the reference's Cython glue does: argument marshalling, table initialisation, launches.
// No CPU compute path exists: every call needs a CUDA device and fails loudly otherwise.
#include "../../include/ytb200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "cast_kernels.cuh"
#include "octree_cast.cuh"
#include "ert.h"
#include "proj.h"

using namespace ytb;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(-100 - (int)e_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

enum { E_ARG = -1, E_STATE = -2, E_UNSUPPORTED = -3 };

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct ytb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    unsigned long long *counters = nullptr;  // device
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct ytb200_brickset {
    ytb200_ctx *ctx = nullptr;
    int64_t n = 0;
    int n_fields = 0, centring = 0, storage = 0;
    std::vector<DevBrick> host;
    std::vector<void *> owned;
    DevBrick *dev = nullptr;
    bool committed = false;
    bool any_mask = false;
    int64_t bytes = 0;
    // process-unique id of this set's committed contents: a new value at creation and at every commit, so that a
    // frame's cached footprints can never be taken for those of another set that reuses the address, nor for this
    // set's bricks before a re-upload (see prepare_order)
    uint64_t generation = 0;
};

static uint64_t next_generation()
{
    static std::atomic<uint64_t> g{0};
    return ++g;
}

struct ytb200_tf {
    ytb200_ctx *ctx = nullptr;
    DevTF d{};
    std::vector<void *> owned;
};

struct ytb200_frame {
    ytb200_ctx *ctx = nullptr;
    DevCamera cam{};
    double *image = nullptr, *zbuffer = nullptr;
    // page-locked host arrays bound as INPUTS (ytb200_frame_bind_host): device aliases read in place over PCIe
    const double *c0_in = nullptr, *zbuffer_in = nullptr;
    bool zbuffer_in_ones = false;
    double *pos_px = nullptr, *dir_px = nullptr;
    int *order_dev = nullptr;
    int4 *extent_dev = nullptr;
    int64_t order_cap = 0;
    std::vector<int> order_host;
    bool image_is_host = false;      // f->image currently aliases a page-locked host array (ytb200_render, in place)
    uint64_t order_generation = 0;   // ytb200_brickset::generation the cached order / footprints belong to (0: none)
    float ert_eps = 0.f;             // > 0: grey-opacity colour casts run the early-termination cast (ytb200_frame_set_early_termination)
    // staged, banded frames (ytb200_render of a projection): copy-in / copy-out streams and one event pair per band
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;
    cudaEvent_t ev_start = nullptr, ev_out = nullptr;
    int band_first = 0, band_tiles = -1;   // s-tile range of the launch being issued (-1: the whole image)
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------------
// library / context
// ------------------------------------------------------------------------------------------------
extern "C" int ytb200_abi_version(void) { return YTB200_ABI_VERSION; }
extern "C" const char *ytb200_last_error(void) { return g_err.c_str(); }

extern "C" int ytb200_device_count(int *count)
{
    if (!count) return fail(E_ARG, "count is NULL");
    CK(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int ytb200_ctx_create(int device, void *stream, ytb200_ctx **out)
{
    if (!out) return fail(E_ARG, "out is NULL");
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n == 0) return fail(E_STATE, "no CUDA device: libytb200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(E_ARG, "device %d out of range (0..%d)", device, n - 1);
    DeviceGuard g(device);
    auto *c = new ytb200_ctx;
    c->device = device;
    if (stream) c->stream = (cudaStream_t)stream;
    else { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaMalloc(&c->counters, 8 * sizeof(unsigned long long)));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    *out = c;
    return 0;
}

extern "C" int ytb200_ctx_destroy(ytb200_ctx *c)
{
    if (!c) return 0;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->counters);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int ytb200_ctx_synchronize(ytb200_ctx *c)
{
    if (!c) return fail(E_ARG, "ctx is NULL");
    DeviceGuard g(c->device);
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void *ytb200_ctx_stream(ytb200_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int ytb200_ctx_sm_count(ytb200_ctx *c) { return c ? c->sm_count : 0; }

// ------------------------------------------------------------------------------------------------
// bricks
// ------------------------------------------------------------------------------------------------
extern "C" int ytb200_bricks_create(ytb200_ctx *ctx, int64_t n_bricks, int n_fields, int centring, int storage,
                                    ytb200_brickset **out)
{
    if (!ctx || !out) return fail(E_ARG, "NULL argument");
    if (n_bricks < 0 || n_bricks > 0x7fffffff) return fail(E_ARG, "bad brick count %lld", (long long)n_bricks);
    if (n_fields < 1 || n_fields > YTB200_MAX_FIELDS) return fail(E_ARG, "n_fields must be 1..6, got %d", n_fields);
    if (centring != YTB200_CELL_CENTRED && centring != YTB200_VERTEX_CENTRED) return fail(E_ARG, "bad centring");
    if (storage != YTB200_STORE_F32 && storage != YTB200_STORE_F64) return fail(E_ARG, "bad storage");
    auto *s = new ytb200_brickset;
    s->ctx = ctx; s->n = n_bricks; s->n_fields = n_fields; s->centring = centring; s->storage = storage;
    s->host.resize((size_t)n_bricks);
    memset(s->host.data(), 0, sizeof(DevBrick) * (size_t)n_bricks);
    s->generation = next_generation();
    *out = s;
    return 0;
}

static int fill_geometry(ytb200_brickset *s, int64_t index, const int64_t dims[3], const double le[3], const double re[3])
{
    if (!s) return fail(E_ARG, "set is NULL");
    if (index < 0 || index >= s->n) return fail(E_ARG, "brick index %lld out of range", (long long)index);
    DevBrick &b = s->host[(size_t)index];
    for (int i = 0; i < 3; i++) {
        if (dims[i] < 1 || dims[i] > 0x7ffffff0) return fail(E_ARG, "bad dims[%d]=%lld", i, (long long)dims[i]);
        b.le[i] = le[i]; b.re[i] = re[i]; b.dims[i] = (int)dims[i];
        b.dds[i] = (b.re[i] - b.le[i]) / (double)dims[i];  // partitioned_grid.pyx:51-52
        b.idds[i] = 1.0 / b.dds[i];
    }
    // the kernels follow the walk with 32-bit running element offsets
    const double nelem = ((double)dims[0] + 1.0) * ((double)dims[1] + 1.0) * ((double)dims[2] + 1.0);
    if (nelem >= 4294967296.0)
        return fail(E_UNSUPPORTED, "brick %lld has %.3g elements per field; split it on cell boundaries below 2^32 "
                    "(sub-bricks reproduce the single-brick image, SURVEY.md finding 6)", (long long)index, nelem);
    b.n_fields = s->n_fields;
    const int c = s->centring == YTB200_VERTEX_CENTRED ? 1 : 0;
    b.pitch_y = (int)dims[2] + c;
    b.pitch_x = ((int)dims[1] + c) * b.pitch_y;
    s->committed = false;
    return 0;
}

// Row pitch (elements) for an array the library allocates itself: rows whose byte length is a multiple of 256 B are
// padded by one 32-byte sector (see DevBrick::pitch_y).
static int padded_row(int row_elems, size_t elem_bytes)
{
    if (((size_t)row_elems * elem_bytes) % 256 == 0) return row_elems + (int)(32 / elem_bytes);
    return row_elems;
}

static int set_pitch(ytb200_brickset *s, DevBrick &b, int pitch_y, int64_t index)
{
    const int c = s->centring == YTB200_VERTEX_CENTRED ? 1 : 0;
    const double total = ((double)b.dims[0] + c) * ((double)b.dims[1] + c) * (double)pitch_y;
    if (total >= 4294967296.0) return fail(E_UNSUPPORTED, "brick %lld has %.3g stored elements per field (>= 2^32)", (long long)index, total);
    b.pitch_y = pitch_y;
    b.pitch_x = (b.dims[1] + c) * pitch_y;
    return 0;
}

extern "C" int ytb200_brick_upload(ytb200_brickset *s, int64_t index, const int64_t dims[3], const double le[3],
                                   const double re[3], const double *const *data, const uint8_t *mask)
{
    if (!s || !dims || !le || !re || !data) return fail(E_ARG, "NULL argument");
    int rc = fill_geometry(s, index, dims, le, re);
    if (rc) return rc;
    DeviceGuard g(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    const int c = s->centring == YTB200_VERTEX_CENTRED ? 1 : 0;
    const size_t ncell = (size_t)dims[0] * (size_t)dims[1] * (size_t)dims[2];
    DevBrick &b = s->host[(size_t)index];
    const size_t esz = s->storage == YTB200_STORE_F64 ? sizeof(double) : sizeof(float);
    const int row = (int)dims[2] + c;
    rc = set_pitch(s, b, padded_row(row, esz), index);
    if (rc) return rc;
    const int pitch = b.pitch_y;
    const size_t nrows = (size_t)(dims[0] + c) * (size_t)(dims[1] + c);
    const size_t nstored = nrows * (size_t)pitch;
    // staging: whole rows, at most ~16 Mi elements (128 MiB of fp64) per chunk
    const size_t rows_per_chunk = std::max<size_t>(1, ((size_t)1 << 24) / (size_t)row);
    double *stage = nullptr;
    if (s->storage == YTB200_STORE_F32) CK(cudaMalloc(&stage, sizeof(double) * std::min(nrows, rows_per_chunk) * (size_t)row));
    for (int f = 0; f < s->n_fields; f++) {
        if (!data[f]) { cudaFree(stage); return fail(E_ARG, "data[%d] is NULL", f); }
        void *dst = nullptr;
        CK(cudaMalloc(&dst, nstored * esz));
        if (pitch != row) CK(cudaMemsetAsync(dst, 0, nstored * esz, st));
        if (s->storage == YTB200_STORE_F64) {
            CK(cudaMemcpy2DAsync(dst, (size_t)pitch * 8, data[f], (size_t)row * 8, (size_t)row * 8, nrows, cudaMemcpyHostToDevice, st));
        } else {
            for (size_t r0 = 0; r0 < nrows; r0 += rows_per_chunk) {
                const size_t m = std::min(rows_per_chunk, nrows - r0) * (size_t)row;
                CK(cudaMemcpyAsync(stage, data[f] + r0 * (size_t)row, m * sizeof(double), cudaMemcpyHostToDevice, st));
                repitch_kernel<float><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(stage, (float *)dst, (long long)m, row, pitch, (long long)r0);
                CK(cudaGetLastError());
            }
        }
        s->bytes += (int64_t)(nstored * esz);
        s->owned.push_back(dst);
        b.data[f] = dst;
    }
    b.mask = nullptr;
    if (mask) {
        bool all_one = true;
        for (size_t i = 0; i < ncell; i++) if (mask[i] != 1) { all_one = false; break; }
        if (!all_one) {
            void *dm = nullptr;
            CK(cudaMalloc(&dm, ncell));
            CK(cudaMemcpyAsync(dm, mask, ncell, cudaMemcpyHostToDevice, st));
            s->owned.push_back(dm);
            b.mask = (const unsigned char *)dm;
            s->any_mask = true;
            s->bytes += (int64_t)ncell;
        }
    }
    CK(cudaStreamSynchronize(st));
    if (stage) CK(cudaFree(stage));
    return 0;
}

// ---- vertex centring on the device -------------------------------------------------------------------
extern "C" int ytb200_cells_upload(ytb200_ctx *ctx, const double *cells_host, const int64_t dims[3], void **cells_dev_out)
{
    if (!ctx || !cells_host || !dims || !cells_dev_out) return fail(E_ARG, "NULL argument");
    for (int i = 0; i < 3; i++) if (dims[i] < 1) return fail(E_ARG, "bad dims[%d]=%lld", i, (long long)dims[i]);
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)dims[0] * (size_t)dims[1] * (size_t)dims[2];
    void *d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, cells_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(d); return fail(-100 - (int)e, "cells upload: %s", cudaGetErrorString(e)); }
    *cells_dev_out = d;
    return 0;
}

extern "C" int ytb200_cells_free(ytb200_ctx *ctx, void *cells_dev)
{
    if (!ctx) return fail(E_ARG, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (cells_dev) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(cells_dev)); }
    return 0;
}

static int brick_from_cells(ytb200_brickset *s, int64_t index, const int64_t grid_dims[3], const int64_t lo[3],
                            const int64_t hi[3], const double le[3], const double re[3],
                            const void *const *cells_dev, int cells_are_f32, const int32_t *log_fields,
                            const uint8_t *mask_host, ytb200_stats *stats, bool ghost)
{
    if (!s || !grid_dims || !lo || !hi || !le || !re || !cells_dev) return fail(E_ARG, "NULL argument");
    if (s->centring != YTB200_VERTEX_CENTRED) return fail(E_ARG, "vertex centring needs a VERTEX_CENTRED brick set");
    int64_t dims[3];
    for (int i = 0; i < 3; i++) {
        // ghost_zone_interpolate divides by (shape - 1): a one-cell axis is NaN in the reference
        if ((!ghost && grid_dims[i] < 2) || grid_dims[i] < 1 || grid_dims[i] > 0x7ffffff0) return fail(E_UNSUPPORTED, "grid_dims[%d]=%lld: no-ghost vertex centring needs >= 2 cells per axis", i, (long long)grid_dims[i]);
        if (lo[i] < 0 || hi[i] <= lo[i] || hi[i] > grid_dims[i]) return fail(E_ARG, "bad vertex range [%lld, %lld] on axis %d", (long long)lo[i], (long long)hi[i], i);
        dims[i] = hi[i] - lo[i];
    }
    int rc = fill_geometry(s, index, dims, le, re);
    if (rc) return rc;
    DeviceGuard g(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    DevBrick &b = s->host[(size_t)index];
    const int m0 = (int)dims[0] + 1, m1 = (int)dims[1] + 1, m2 = (int)dims[2] + 1;
    rc = set_pitch(s, b, padded_row(m2, s->storage == YTB200_STORE_F64 ? 8 : 4), index);
    if (rc) return rc;
    const int opy = b.pitch_y, opx = b.pitch_x;
    const size_t nelem = (size_t)m0 * (size_t)opx;
    const size_t ncell = (size_t)dims[0] * (size_t)dims[1] * (size_t)dims[2];
    if (m1 > 65535 || (m0 + VC_XCHUNK - 1) / VC_XCHUNK > 65535) return fail(E_UNSUPPORTED, "brick too large for one centring launch");
    const dim3 grid((unsigned)((m2 + VC_ZSPAN - 1) / VC_ZSPAN), (unsigned)m1, (unsigned)((m0 + VC_XCHUNK - 1) / VC_XCHUNK));
    if (stats) CK(cudaEventRecord(s->ctx->ev0, st));
    for (int f = 0; f < s->n_fields; f++) {
        if (!cells_dev[f]) return fail(E_ARG, "cells_dev[%d] is NULL", f);
        const int lg = log_fields ? log_fields[f] : 0;
        const bool f64 = s->storage == YTB200_STORE_F64;
        void *dst = nullptr;
        CK(cudaMalloc(&dst, nelem * (f64 ? sizeof(double) : sizeof(float))));
        if (opy != m2) CK(cudaMemsetAsync(dst, 0, nelem * (f64 ? sizeof(double) : sizeof(float)), st));
        s->owned.push_back(dst);
        s->bytes += (int64_t)(nelem * (f64 ? 8 : 4));
        const int nx = (int)grid_dims[0], ny = (int)grid_dims[1], nz = (int)grid_dims[2];
        const int l0 = (int)lo[0], l1 = (int)lo[1], l2 = (int)lo[2];
#define YTB_VC_LAUNCH(K, TI, TO) K<TI, TO><<<grid, VC_THREADS, 0, st>>>((const TI *)cells_dev[f], (TO *)dst, nx, ny, nz, l0, l1, l2, m0, m1, m2, lg, opx, opy)
        if (ghost) {
            if (cells_are_f32) { if (f64) YTB_VC_LAUNCH(vertex_centre_ghost_kernel, float, double); else YTB_VC_LAUNCH(vertex_centre_ghost_kernel, float, float); }
            else               { if (f64) YTB_VC_LAUNCH(vertex_centre_ghost_kernel, double, double); else YTB_VC_LAUNCH(vertex_centre_ghost_kernel, double, float); }
        } else {
            if (cells_are_f32) { if (f64) YTB_VC_LAUNCH(vertex_centre_kernel, float, double); else YTB_VC_LAUNCH(vertex_centre_kernel, float, float); }
            else               { if (f64) YTB_VC_LAUNCH(vertex_centre_kernel, double, double); else YTB_VC_LAUNCH(vertex_centre_kernel, double, float); }
        }
#undef YTB_VC_LAUNCH
        CK(cudaGetLastError());
        b.data[f] = dst;
    }
    if (stats) {
        CK(cudaEventRecord(s->ctx->ev1, st));
        CK(cudaEventSynchronize(s->ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ctx->ev0, s->ctx->ev1));
        stats->kernel_ms = ms; stats->n_launches = s->n_fields; stats->n_cells = 0; stats->n_samples = 0;
    }
    b.mask = nullptr;
    if (mask_host) {
        bool all_one = true;
        for (size_t i = 0; i < ncell; i++) if (mask_host[i] != 1) { all_one = false; break; }
        if (!all_one) {
            void *dm = nullptr;
            CK(cudaMalloc(&dm, ncell));
            CK(cudaMemcpyAsync(dm, mask_host, ncell, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            s->owned.push_back(dm);
            b.mask = (const unsigned char *)dm;
            s->any_mask = true;
            s->bytes += (int64_t)ncell;
        }
    }
    return 0;
}

extern "C" int ytb200_brick_from_cells(ytb200_brickset *s, int64_t index, const int64_t grid_dims[3], const int64_t lo[3],
                                       const int64_t hi[3], const double le[3], const double re[3],
                                       const void *const *cells_dev, int cells_are_f32, const int32_t *log_fields,
                                       const uint8_t *mask_host, ytb200_stats *stats)
{
    return brick_from_cells(s, index, grid_dims, lo, hi, le, re, cells_dev, cells_are_f32, log_fields, mask_host, stats, false);
}

extern "C" int ytb200_brick_from_ghost_cells(ytb200_brickset *s, int64_t index, const int64_t grid_dims[3], const int64_t lo[3],
                                             const int64_t hi[3], const double le[3], const double re[3],
                                             const void *const *cells_dev, int cells_are_f32, const int32_t *log_fields,
                                             const uint8_t *mask_host, ytb200_stats *stats)
{
    return brick_from_cells(s, index, grid_dims, lo, hi, le, re, cells_dev, cells_are_f32, log_fields, mask_host, stats, true);
}

extern "C" int ytb200_brick_attach(ytb200_brickset *s, int64_t index, const int64_t dims[3], const double le[3],
                                   const double re[3], const void *const *dev_data, const uint8_t *dev_mask)
{
    if (!s || !dims || !le || !re || !dev_data) return fail(E_ARG, "NULL argument");
    int rc = fill_geometry(s, index, dims, le, re);
    if (rc) return rc;
    DevBrick &b = s->host[(size_t)index];
    const int c = s->centring == YTB200_VERTEX_CENTRED ? 1 : 0;
    const size_t nelem = (size_t)(dims[0] + c) * (size_t)(dims[1] + c) * (size_t)(dims[2] + c);
    for (int f = 0; f < s->n_fields; f++) {
        if (!dev_data[f]) return fail(E_ARG, "dev_data[%d] is NULL", f);
        b.data[f] = dev_data[f];
        s->bytes += (int64_t)(nelem * (s->storage == YTB200_STORE_F64 ? 8 : 4));
    }
    b.mask = dev_mask;
    if (dev_mask) s->any_mask = true;
    return 0;
}

extern "C" int ytb200_brick_attach_pitched(ytb200_brickset *s, int64_t index, const int64_t dims[3], const double le[3],
                                           const double re[3], const void *const *dev_data, const uint8_t *dev_mask,
                                           int64_t row_pitch)
{
    int rc = ytb200_brick_attach(s, index, dims, le, re, dev_data, dev_mask);
    if (rc) return rc;
    const int c = s->centring == YTB200_VERTEX_CENTRED ? 1 : 0;
    if (row_pitch < dims[2] + c || row_pitch > 0x7fffffff) return fail(E_ARG, "row_pitch %lld shorter than a row", (long long)row_pitch);
    return set_pitch(s, s->host[(size_t)index], (int)row_pitch, index);
}

extern "C" int ytb200_brick_pitch(const ytb200_brickset *s, int64_t index, int64_t *plane_pitch, int64_t *row_pitch)
{
    if (!s || !plane_pitch || !row_pitch) return fail(E_ARG, "NULL argument");
    if (index < 0 || index >= s->n) return fail(E_ARG, "brick index %lld out of range", (long long)index);
    *plane_pitch = s->host[(size_t)index].pitch_x;
    *row_pitch = s->host[(size_t)index].pitch_y;
    return 0;
}

extern "C" int ytb200_bricks_commit(ytb200_brickset *s)
{
    if (!s) return fail(E_ARG, "set is NULL");
    DeviceGuard g(s->ctx->device);
    for (int64_t i = 0; i < s->n; i++)
        if (!s->host[(size_t)i].data[0]) return fail(E_STATE, "brick %lld has no data", (long long)i);
    if (!s->dev && s->n) CK(cudaMalloc(&s->dev, sizeof(DevBrick) * (size_t)s->n));
    if (s->n) {
        CK(cudaMemcpyAsync(s->dev, s->host.data(), sizeof(DevBrick) * (size_t)s->n, cudaMemcpyHostToDevice, s->ctx->stream));
        CK(cudaStreamSynchronize(s->ctx->stream));
    }
    s->committed = true;
    s->generation = next_generation();
    return 0;
}

extern "C" int ytb200_bricks_destroy(ytb200_brickset *s)
{
    if (!s) return 0;
    DeviceGuard g(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    for (void *p : s->owned) cudaFree(p);
    cudaFree(s->dev);
    delete s;
    return 0;
}

extern "C" int ytb200_brick_device_ptrs(const ytb200_brickset *s, int64_t index, void **data_out, void **mask_out)
{
    if (!s || !data_out || !mask_out) return fail(E_ARG, "NULL argument");
    if (index < 0 || index >= s->n) return fail(E_ARG, "brick index %lld out of range", (long long)index);
    const DevBrick &b = s->host[(size_t)index];
    for (int f = 0; f < s->n_fields; f++) data_out[f] = const_cast<void *>(b.data[f]);
    *mask_out = const_cast<unsigned char *>(b.mask);
    return 0;
}

extern "C" int64_t ytb200_bricks_device_bytes(const ytb200_brickset *s) { return s ? s->bytes : 0; }

// ------------------------------------------------------------------------------------------------
// transfer function
// ------------------------------------------------------------------------------------------------
template <typename T>
static int to_device(ytb200_tf *t, const std::vector<T> &h, const T **out)
{
    void *d = nullptr;
    CK(cudaMalloc(&d, sizeof(T) * (h.size() ? h.size() : 1)));
    CK(cudaMemcpyAsync(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, t->ctx->stream));
    t->owned.push_back(d);
    *out = (const T *)d;
    return 0;
}

// LINE form of a delta-layout colour table (DevTF::colour_line): row r = bin + 2 holds (c, d) with d the fp32 increment as
// stored in `col` and c = y - (r - origin) * d, so the product in c is exact and the error of a bin against the fp64 table is
// |fl(c) - c| + |fl(d) - d| (the latter times the fractional position, <= 1).  Returns the worst ratio, over all bins and
// channels, of that error to its allowance -- 5e-6 of the bin's mean magnitude plus 1e-9 of the channel's maximum -- for the
// integer origin that minimises it; `*line_out` gets the rows for that origin.  A ratio <= 1 means the form may be used:
// a relative error eps per lookup moves a pixel by at most (2 + tau) * eps of its value (emission twice -- colour times alpha --
// and the optical depth tau once), 6e-5 at tau = 10, where less than 5e-5 of what lies behind shows through anyway.
static double line_form_rows(const double *const y[4], int nb, const std::vector<float4> &col, int *origin_out, std::vector<float4> *line_out)
{
    double cmax[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < 4; c++)
        for (int j = 0; j < nb; j++) cmax[c] = fmax(cmax[c], fabs(y[c][j]));
    auto worst_bin = [&](int o, std::vector<float4> *out) {
        double worst = 0.0;
        for (int c = 0; c < 4; c++) {
            for (int j = 0; j + 1 < nb; j++) {
                const size_t r = (size_t)j + 2;
                const float df = (&col[2 * r + 1].x)[c];
                const double d64 = y[c][j + 1] - y[c][j];
                const double c64 = y[c][j] - ((double)r - (double)o) * (double)df;
                const float cf = (float)c64;
                if (out) { (&(*out)[2 * r].x)[c] = cf; (&(*out)[2 * r + 1].x)[c] = df; }
                const double abs_err = fabs((double)cf - c64) + fabs((double)df - d64);
                const double lim = 5e-6 * 0.5 * (fabs(y[c][j]) + fabs(y[c][j + 1])) + 1e-9 * cmax[c];
                if (!(abs_err <= lim)) worst = fmax(worst, (lim > 0.0 && abs_err == abs_err) ? abs_err / lim : 1e300);   // NaN fails
                else if (lim > 0.0) worst = fmax(worst, abs_err / lim);
            }
        }
        return worst;
    };
    double best = 1e300;
    int origin = nb / 2 + 2;
    for (int o = 2; o < nb + 2; o++) {
        const double w = worst_bin(o, nullptr);
        if (w < best) { best = w; origin = o; }
    }
    if (line_out) {
        line_out->assign(col.size(), make_float4(0.f, 0.f, 0.f, 0.f));
        best = fmax(best, worst_bin(origin, line_out));
    }
    if (origin_out) *origin_out = origin;
    return best;
}

// The (value, increment) rows of the delta layout: row = bin + 2, rows 0, 1 and >= nbins + 1 are zero.
static std::vector<float4> delta_rows(const double *const y[4], int nb)
{
    std::vector<float4> col(2 * ((size_t)nb + 2), make_float4(0.f, 0.f, 0.f, 0.f));
    auto row = [&](int j) { return make_float4((float)y[0][j], (float)y[1][j], (float)y[2][j], (float)y[3][j]); };
    for (int j = 0; j + 1 < nb; j++) {
        const float4 a = row(j), b = row(j + 1);
        col[2 * ((size_t)j + 2)] = a;
        col[2 * ((size_t)j + 2) + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w - a.w);
    }
    return col;
}

// Host-only probe of the decision ytb200_tf_create takes for a ColorTransferFunction's four tables (no device needed; the
// CPU tests pin it): returns 1 if the table would be sampled through LINE-form rows, 0 if it keeps (value, increment) rows.
extern "C" int ytb200_tf_line_form_probe(int nbins, const double *r, const double *g, const double *b, const double *a,
                                         int *origin_out, double *worst_ratio_out)
{
    if (nbins < 2 || !r || !g || !b || !a) return fail(E_ARG, "bad argument");
    const double *y4[4] = {r, g, b, a};
    int origin = 0;
    const double worst = nbins <= COLOUR_DELTA_MAX_BINS ? line_form_rows(y4, nbins, delta_rows(y4, nbins), &origin, nullptr) : 1e300;
    if (origin_out) *origin_out = origin;
    if (worst_ratio_out) *worst_ratio_out = worst;
    return worst <= 1.0 ? 1 : 0;
}

extern "C" int ytb200_tf_create(ytb200_ctx *ctx, int n_tables, const int32_t *nbins, const double *const *y,
                                const double *bounds, const int32_t field_ids[6], const int32_t weight_field_ids[6],
                                const int32_t weight_table_ids[6], const int32_t field_table_ids[6], int grey_opacity,
                                ytb200_tf **out)
{
    if (!ctx || !out || !nbins || !y || !bounds || !field_ids || !weight_field_ids || !weight_table_ids || !field_table_ids)
        return fail(E_ARG, "NULL argument");
    if (n_tables < 1 || n_tables > YTB200_MAX_TABLES) return fail(E_ARG, "n_tables must be 1..6, got %d", n_tables);
    DeviceGuard g(ctx->device);
    auto *t = new ytb200_tf;
    t->ctx = ctx;
    DevTF &d = t->d;
    d.n_fits = n_tables;
    d.grey = grey_opacity ? 1 : 0;
    int total = 0;
    for (int i = 0; i < n_tables; i++) {
        if (nbins[i] < 2) { delete t; return fail(E_ARG, "table %d needs >= 2 bins", i); }
        if (field_ids[i] < 0 || field_ids[i] >= 6 || weight_field_ids[i] < -1 || weight_field_ids[i] >= 6 ||
            weight_table_ids[i] < -1 || weight_table_ids[i] >= n_tables) {
            delete t;
            return fail(E_ARG, "table %d: field/weight ids out of range", i);
        }
        d.nbins[i] = nbins[i]; d.offset[i] = total; total += nbins[i];
        d.field_id[i] = field_ids[i]; d.weight_field_id[i] = weight_field_ids[i]; d.weight_table_id[i] = weight_table_ids[i];
    }
    for (int i = 0; i < 6; i++) {
        if (field_table_ids[i] < 0 || field_table_ids[i] >= 6) { delete t; return fail(E_ARG, "field_table_ids[%d] out of range", i); }
        d.field_table_id[i] = field_table_ids[i];
    }
    d.total_bins = total;
    std::vector<double> values(total), d0(total, 0.0), dy(total, 0.0);
    std::vector<float2> pairs(total);
    for (int i = 0; i < n_tables; i++) {
        // FIT_initialize_table, field_interpolation_tables.pxd:44-61
        const double b0 = bounds[2 * i], b1 = bounds[2 * i + 1];
        const double dbin = (b1 - b0) / (nbins[i] - 1), idbin = 1.0 / dbin;
        d.b0[i] = b0; d.b1[i] = b1; d.idbin[i] = idbin;
        d.b0_hi[i] = (float)b0; d.b0_lo[i] = (float)(b0 - (double)d.b0_hi[i]);
        d.b0_f[i] = (float)b0; d.b1_f[i] = (float)b1; d.idbin_f[i] = (float)idbin;
        for (int j = 0; j < nbins[i]; j++) {
            const int o = d.offset[i] + j;
            values[o] = y[i][j];
            if (j < nbins[i] - 1) {
                d0[o] = b0 + j * dbin;
                dy[o] = (y[i][j + 1] - y[i][j]) * idbin;
                pairs[o] = make_float2((float)y[i][j], (float)(y[i][j + 1] - y[i][j]));
            } else {
                pairs[o] = make_float2((float)y[i][j], 0.f);
            }
        }
    }
    // ColorTransferFunction layout (transfer_functions.py:385-406)
    bool colour = n_tables == 4;
    for (int i = 0; colour && i < 4; i++) {
        colour = nbins[i] == nbins[0] && bounds[2 * i] == bounds[0] && bounds[2 * i + 1] == bounds[1] &&
                 field_ids[i] == 0 && weight_field_ids[i] == -1 && weight_table_ids[i] == (i < 3 ? 3 : -1) &&
                 field_table_ids[i] == i;
    }
    d.colour_fast = colour ? 1 : 0;
    int rc = 0;
    rc |= to_device(t, values, &d.values);
    rc |= to_device(t, d0, &d.d0);
    rc |= to_device(t, dy, &d.dy);
    rc |= to_device(t, pairs, &d.pairs);
    if (colour) {
        // shared-memory row layouts of the fast path (raycast.cuh); row = bin + 2, rows 0 and 1 are zero:
        //   delta layout  (nbins <= 256): row r holds (y_{r-2}, y_{r-1} - y_{r-2}); rows >= nbins + 1 are zero
        //   value layout  (nbins <= 512): rows [0, 0, y_0 .. y_{n-1}, y_{n-1}], the lerp reads rows r and r + 1
        const int nb = nbins[0];
        d.colour_delta = nb <= COLOUR_DELTA_MAX_BINS ? 1 : 0;
        std::vector<float4> col;
        double rgb_max = 0.0, a_max = 0.0;
        auto row = [&](int j) { return make_float4((float)y[0][j], (float)y[1][j], (float)y[2][j], (float)y[3][j]); };
        for (int j = 0; j < nb; j++) {
            rgb_max = fmax(rgb_max, fmax(fabs(y[0][j]), fmax(fabs(y[1][j]), fabs(y[2][j]))));
            a_max = fmax(a_max, fabs(y[3][j]));
        }
        if (d.colour_delta) {
            col.assign(2 * ((size_t)nb + 2), make_float4(0.f, 0.f, 0.f, 0.f));
            for (int j = 0; j + 1 < nb; j++) {
                const float4 a = row(j), b = row(j + 1);
                col[2 * ((size_t)j + 2)] = a;
                col[2 * ((size_t)j + 2) + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w - a.w);
            }
            d.colour_rows = nb + 2;
        } else {
            col.assign((size_t)nb + COLOUR_PAD_ROWS, make_float4(0.f, 0.f, 0.f, 0.f));
            for (int j = 0; j < nb; j++) col[(size_t)j + 2] = row(j);
            col[(size_t)nb + 2] = col[(size_t)nb + 1];
            d.colour_rows = nb + COLOUR_PAD_ROWS;
        }
        // does a sample flipping across a bound change the result?  Only if the alpha table (which also
        // weights R, G, B) does not vanish there.
        const double edge = fmax(fabs(y[3][0]), fabs(y[3][nb - 1])) * fmax(1.0, rgb_max);
        d.colour_edge = edge > 1e-30 ? 1 : 0;
        d.colour_line = 0; d.colour_clamp_lo = 0.5f; d.colour_magic = COLOUR_MAGIC;
        int origin = 0;
        if (d.colour_delta && !getenv("YTB200_NO_LINE")) {
            // LINE form (DevTF): row r holds (c, d) with c = y - (r - origin) * d, so that y(fs) = d * (fs - origin) + c needs
            // no fractional part (two instructions fewer per sample).  What it costs is accuracy: c is rounded to fp32 at a
            // magnitude of up to |r - origin| * |d|, so the form is used only if EVERY bin keeps a relative accuracy of 5e-6 of
            // its own mean magnitude (plus 1e-9 of the channel's maximum) -- computed here, per table, against the fp64
            // table; a table that fails keeps the (value, increment) rows.  d is the fp32 increment as stored, so the product
            // in c is exact and the error of a bin is |fl(c) - c| + |fl(d) - d| (the latter times the fractional position,
            // <= 1).  The origin is the integer row that minimises the worst bin.
            std::vector<float4> line;
            const double *y4[4] = {y[0], y[1], y[2], y[3]};
            const bool ok = line_form_rows(y4, nb, col, &origin, &line) <= 1.0;
            if (ok) {
                col.swap(line);
                d.colour_line = 1;
                d.colour_clamp_lo = 0.5f - (float)origin;
                d.colour_magic = COLOUR_MAGIC + (float)origin;
            } else origin = 0;
        }
        rc |= to_device(t, col, &d.colour);
        {
            // the shared-memory image of the table, staged by a cast CTA with ONE bulk async copy (cp.async.bulk)
            std::vector<float4> image(col.size() * COLOUR_COPIES);
            for (size_t i = 0; i < image.size(); i++) image[i] = col[i / COLOUR_COPIES];
            rc |= to_device(t, image, &d.colour_image);
        }
        // every row is replicated 8x in shared memory (bank-conflict-free): 64.4 KiB (value layout, 512 bins)
        // or 64.5 KiB (delta layout, 256 bins)
        d.colour_copy_shift = 3;
        const double b0 = bounds[0], b1 = bounds[1];
        const double idbin = d.idbin[0];
        // fb = (val - b0_hi) * idbin + c0 with c0 = 2 - b0_lo * idbin  (== (val - b0) * idbin + 2)
        d.colour_c0 = (float)(2.0 - (b0 - (double)d.b0_hi[0]) * idbin - (double)origin);
        // every product that meets dt before a clamp: alpha (grey) or alpha * colour (RGB weighted by alpha)
        d.colour_clamp_bound = (float)(1.0000002 * a_max * fmax(1.0, rgb_max));
        d.colour_zero_inside = (b0 < 0.0 && b1 > 0.0) ? 1 : 0;
        d.colour_nonneg = 1;   // alpha never negative: a ray's alpha accumulator keeps its sign bit free (PF_A32G)
        for (int j = 0; j < nb; j++) if (!(y[3][j] >= 0.0)) d.colour_nonneg = 0;
        d.colour_lo = 2.f; d.colour_hi = (float)(nb + 1);
        d.colour_clamp_hi = (float)nb + 1.5f - (float)origin;
        d.colour_mid = 0.5f * (float)(nb + 3) - (float)origin; d.colour_half = 0.5f * (float)(nb - 1);   // origin != 0: LINE form (delta layout)
        d.colour_band = 4e-6f * (float)(nb + 512);
        // row coordinate of the value 0, with the very operations load_corners_rows applies to a corner
        d.colour_zrow = fmaf(0.f - d.b0_hi[0], d.idbin_f[0], d.colour_c0);
        if (nb > 512 || !std::isfinite(d.colour_clamp_bound) || !std::isfinite(d.colour_c0)) d.colour_fast = 0;
    }
    if (rc) { ytb200_tf_destroy(t); return rc; }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ytb200_tf_destroy(t); return fail(-100 - (int)e, "tf upload: %s", cudaGetErrorString(e)); }
    *out = t;
    return 0;
}

extern "C" int ytb200_tf_set_light(ytb200_tf *t, const double *light_dir, const double *light_rgba)
{
    if (!t) return fail(E_ARG, "tf is NULL");
    if (!light_dir || !light_rgba) { t->d.light_on = 0; return 0; }
    for (int i = 0; i < 3; i++) t->d.light_dir[i] = light_dir[i];
    for (int i = 0; i < 4; i++) t->d.light_rgba[i] = light_rgba[i];
    t->d.light_on = 1;
    return 0;
}

extern "C" int ytb200_tf_destroy(ytb200_tf *t)
{
    if (!t) return 0;
    DeviceGuard g(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    for (void *p : t->owned) cudaFree(p);
    delete t;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// frame
// ------------------------------------------------------------------------------------------------
static bool invert3(const double m[9], double inv[9])
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double scale = fabs(a * e * i) + fabs(b * f * g) + fabs(c * d * h) + fabs(c * e * g) + fabs(b * d * i) + fabs(a * f * h);
    if (!(fabs(det) > 1e-12 * scale) || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    inv[0] = (e * i - f * h) * id; inv[1] = (c * h - b * i) * id; inv[2] = (b * f - c * e) * id;
    inv[3] = (f * g - d * i) * id; inv[4] = (a * i - c * g) * id; inv[5] = (c * d - a * f) * id;
    inv[6] = (d * h - e * g) * id; inv[7] = (b * g - a * h) * id; inv[8] = (a * e - b * d) * id;
    return true;
}

extern "C" int ytb200_frame_create(ytb200_ctx *ctx, const ytb200_camera *cam, ytb200_frame **out)
{
    if (!ctx || !cam || !out) return fail(E_ARG, "NULL argument");
    if (cam->nx < 1 || cam->ny < 1) return fail(E_ARG, "bad image shape %d x %d", cam->nx, cam->ny);
    if (cam->lens < 0 || cam->lens > 2) return fail(E_ARG, "bad lens %d", cam->lens);
    if (cam->lens != YTB200_LENS_PLANE_PARALLEL && (!cam->vp_pos_px || !cam->vp_dir_px))
        return fail(E_ARG, "Bad lens shape / direction: per-pixel vp_pos/vp_dir required for this lens");
    DeviceGuard g(ctx->device);
    auto *f = new ytb200_frame;
    f->ctx = ctx;
    DevCamera &c = f->cam;
    c.lens = cam->lens; c.nx = cam->nx; c.ny = cam->ny;
    memcpy(c.vp_pos, cam->vp_pos, sizeof c.vp_pos);
    memcpy(c.vp_dir, cam->vp_dir, sizeof c.vp_dir);
    memcpy(c.width, cam->width, sizeof c.width);
    c.ext_mode = 0;
    const size_t npx = (size_t)cam->nx * cam->ny;
    if (cam->lens == YTB200_LENS_PLANE_PARALLEL) {
        // a point X maps to image-plane coordinates by solving X - back_center = e*px + n*py + d*t
        const double m[9] = {c.vp_pos[0], c.vp_pos[3], c.vp_dir[0], c.vp_pos[1], c.vp_pos[4], c.vp_dir[1],
                             c.vp_pos[2], c.vp_pos[5], c.vp_dir[2]};
        double inv[9];
        if (invert3(m, inv) && cam->nx > 1 && cam->ny > 1 && cam->width[0] != 0.0 && cam->width[1] != 0.0) {
            memcpy(c.ext_rows, inv, sizeof(double) * 6);
            for (int i = 0; i < 3; i++) c.ext_origin[i] = c.vp_pos[9 + i];
            c.ext_mode = 1;
        }
    } else {
        cudaError_t e;
        if ((e = cudaMalloc(&f->pos_px, npx * 3 * sizeof(double))) != cudaSuccess ||
            (e = cudaMalloc(&f->dir_px, npx * 3 * sizeof(double))) != cudaSuccess) {
            ytb200_frame_destroy(f);
            return fail(-100 - (int)e, "ray buffers: %s", cudaGetErrorString(e));
        }
        cudaMemcpyAsync(f->pos_px, cam->vp_pos_px, npx * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(f->dir_px, cam->vp_dir_px, npx * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        c.pos_px = f->pos_px; c.dir_px = f->dir_px;
        if (cam->lens == YTB200_LENS_PERSPECTIVE && cam->has_camera_data && cam->nx > 1 && cam->ny > 1) {
            memcpy(c.persp, cam->camera_data, sizeof c.persp);
            for (int i = 0; i < 3; i++) c.ext_origin[i] = cam->camera_data[i];
            if (cam->camera_data[3] != 0.0 && cam->camera_data[5] != 0.0) c.ext_mode = 2;
        }
    }
    // the device image / z-buffer are allocated on first use (ensure_buffers): a frame that only ever renders page-locked
    // host arrays in place (ytb200_render, what a Scene does every frame) never needs them, and 167 MB of cudaMalloc +
    // cudaFree per frame were 21 ms of a 2048^2 scene render (tools/profile_scene.py)
    (void)npx;
    *out = f;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Device image (zeros) and z-buffer (ones) of a frame, allocated and initialised on first use.
static int ensure_buffers(ytb200_frame *f)
{
    if (f->image && f->zbuffer) return 0;
    if (f->image_is_host) return fail(E_STATE, "frame is rendering a host image in place");
    DeviceGuard g(f->ctx->device);
    const size_t npx = (size_t)f->cam.nx * f->cam.ny;
    cudaError_t e;
    if ((!f->image && (e = cudaMalloc(&f->image, npx * 4 * sizeof(double))) != cudaSuccess) ||
        (!f->zbuffer && (e = cudaMalloc(&f->zbuffer, npx * sizeof(double))) != cudaSuccess))
        return fail(-100 - (int)e, "frame buffers: %s", cudaGetErrorString(e));
    CK(cudaMemsetAsync(f->image, 0, npx * 4 * sizeof(double), f->ctx->stream));
    fill_f64_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, f->ctx->stream>>>(f->zbuffer, 1.0, (long long)npx);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ytb200_frame_destroy(ytb200_frame *f)
{
    if (!f) return 0;
    DeviceGuard g(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    cudaFree(f->image); cudaFree(f->zbuffer); cudaFree(f->pos_px); cudaFree(f->dir_px);
    cudaFree(f->order_dev); cudaFree(f->extent_dev);
    if (f->s_in) cudaStreamDestroy(f->s_in);
    if (f->s_out) cudaStreamDestroy(f->s_out);
    for (cudaEvent_t e : f->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : f->ev_done) cudaEventDestroy(e);
    if (f->ev_start) cudaEventDestroy(f->ev_start);
    if (f->ev_out) cudaEventDestroy(f->ev_out);
    delete f;
    return 0;
}

extern "C" int ytb200_frame_upload(ytb200_frame *f, const double *image_host, const double *zbuffer_host)
{
    if (!f || !image_host) return fail(E_ARG, "NULL argument");
    { int rc = ensure_buffers(f); if (rc) return rc; }
    DeviceGuard g(f->ctx->device);
    const size_t npx = (size_t)f->cam.nx * f->cam.ny;
    CK(cudaMemcpyAsync(f->image, image_host, npx * 4 * sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
    if (zbuffer_host) {
        CK(cudaMemcpyAsync(f->zbuffer, zbuffer_host, npx * sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
    } else {
        fill_f64_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, f->ctx->stream>>>(f->zbuffer, 1.0, (long long)npx);
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int ytb200_frame_clear(ytb200_frame *f)
{
    if (!f) return fail(E_ARG, "frame is NULL");
    if (!f->image || !f->zbuffer) return ensure_buffers(f);   // a fresh allocation is a cleared one
    DeviceGuard g(f->ctx->device);
    const size_t npx = (size_t)f->cam.nx * f->cam.ny;
    CK(cudaMemsetAsync(f->image, 0, npx * 4 * sizeof(double), f->ctx->stream));
    fill_f64_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, f->ctx->stream>>>(f->zbuffer, 1.0, (long long)npx);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ytb200_frame_download(ytb200_frame *f, double *image_host)
{
    if (!f || !image_host) return fail(E_ARG, "NULL argument");
    { int rc = ensure_buffers(f); if (rc) return rc; }
    DeviceGuard g(f->ctx->device);
    const size_t npx = (size_t)f->cam.nx * f->cam.ny;
    CK(cudaMemcpyAsync(image_host, f->image, npx * 4 * sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
    CK(cudaStreamSynchronize(f->ctx->stream));
    return 0;
}

extern "C" void *ytb200_frame_device_image(ytb200_frame *f) { return (f && ensure_buffers(f) == 0) ? (void *)f->image : nullptr; }

// Upload the order list and compute per-brick pixel footprints (cached while unchanged).
static int prepare_order(ytb200_frame *f, const ytb200_brickset *set, const int64_t *order, int64_t n_order, int *launches)
{
    if (!set->committed) return fail(E_STATE, "brick set not committed (call ytb200_bricks_commit)");
    if (set->ctx->device != f->ctx->device) return fail(E_ARG, "brick set and frame live on different devices");
    if (n_order < 0 || n_order > set->n * 64 + 1024) return fail(E_ARG, "bad n_order %lld", (long long)n_order);
    bool same = f->order_generation != 0 && f->order_generation == set->generation && (int64_t)f->order_host.size() == n_order;
    if (same)
        for (int64_t i = 0; i < n_order; i++)
            if (f->order_host[(size_t)i] != (int)order[i]) { same = false; break; }
    if (same) return 0;
    f->order_host.resize((size_t)n_order);
    for (int64_t i = 0; i < n_order; i++) {
        if (order[i] < 0 || order[i] >= set->n) return fail(E_ARG, "order[%lld]=%lld out of range", (long long)i, (long long)order[i]);
        f->order_host[(size_t)i] = (int)order[i];
    }
    f->order_generation = 0;
    if (n_order > f->order_cap) {
        CK(cudaStreamSynchronize(f->ctx->stream));
        cudaFree(f->order_dev); cudaFree(f->extent_dev);
        f->order_dev = nullptr; f->extent_dev = nullptr;
        CK(cudaMalloc(&f->order_dev, sizeof(int) * (size_t)n_order));
        // footprints of the entries, then the union footprints of every CTA_THREADS consecutive entries (chunk_extent_kernel)
        CK(cudaMalloc(&f->extent_dev, sizeof(int4) * ((size_t)n_order + (size_t)n_order / CTA_THREADS + 1)));
        f->order_cap = n_order;
    }
    if (n_order) {
        CK(cudaMemcpyAsync(f->order_dev, f->order_host.data(), sizeof(int) * (size_t)n_order, cudaMemcpyHostToDevice, f->ctx->stream));
        extent_kernel<<<(unsigned)((n_order + 127) / 128), 128, 0, f->ctx->stream>>>(f->cam, set->dev, f->order_dev, (int)n_order, f->extent_dev);
        CK(cudaGetLastError());
        chunk_extent_kernel<<<(unsigned)((n_order + CTA_THREADS - 1) / CTA_THREADS), CTA_THREADS, 0, f->ctx->stream>>>(f->extent_dev, (int)n_order, f->extent_dev + f->order_cap);
        CK(cudaGetLastError());
        if (launches) (*launches) += 2;
    }
    f->order_generation = set->generation;
    return 0;
}

template <int KIND, bool EXACT, int TFM, bool PARTIAL, bool NF1, typename T>
static cudaError_t launch_one(const CastParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    void (*k)(const CastParams) = cast_kernel<KIND, EXACT, TFM, PARTIAL, NF1, T>;
    // long brick lists (AMR) take the MANY instantiation of the colour volume render (see cast_kernel)
    if constexpr (KIND == KIND_VR && !EXACT && tfm_is_colour(TFM) && std::is_same<T, float>::value) {
        if (p.n_order > 32) k = cast_kernel<KIND, EXACT, TFM, PARTIAL, NF1, T, true>;
    }
    if (smem > 32 * 1024) {  // static shared memory counts against the 48 KiB default too
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, CTA_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

template <int KIND, int TFM, bool PARTIAL, bool NF1>
static cudaError_t launch_prec(const CastParams &p, unsigned grid, size_t smem, cudaStream_t st, bool exact, bool f64)
{
    if (exact) return launch_one<KIND, true, (KIND == KIND_VR ? TFM_GENERAL : TFM_NONE), PARTIAL, NF1, double>(p, grid, smem, st);
    if (f64) return launch_one<KIND, false, TFM, PARTIAL, NF1, double>(p, grid, smem, st);
    return launch_one<KIND, false, TFM, PARTIAL, NF1, float>(p, grid, smem, st);
}

template <int KIND, int TFM>
static cudaError_t launch_kind(const CastParams &p, unsigned grid, size_t smem, cudaStream_t st, bool exact, bool f64,
                               bool partial, bool nf1)
{
    constexpr int TFM_MULTI = tfm_is_colour(TFM) ? TFM_GENERAL : TFM;  // the colour fast path is single-field
    if (partial) {
        if (nf1) return launch_prec<KIND, TFM, true, true>(p, grid, smem, st, exact, f64);
        return launch_prec<KIND, TFM_MULTI, true, false>(p, grid, smem, st, exact, f64);
    }
    if (nf1) return launch_prec<KIND, TFM, false, true>(p, grid, smem, st, exact, f64);
    return launch_prec<KIND, TFM_MULTI, false, false>(p, grid, smem, st, exact, f64);
}

static int cast_common(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order, int64_t n_order,
                       const ytb200_tf *tf, int n_samples, int precision, double *partial, ytb200_stats *stats,
                       int partial_fmt = PF_AB64, unsigned *tail_counter = nullptr, unsigned long long *tail_flag = nullptr,
                       unsigned long long tail_seq = 0)
{
    if (!f || !set || (!order && n_order > 0)) return fail(E_ARG, "NULL argument");
    if (kind < 0 || kind > 2) return fail(E_ARG, "bad sampler kind %d", kind);
    if (precision != YTB200_PREC_FAST && precision != YTB200_PREC_EXACT) return fail(E_ARG, "bad precision %d", precision);
    if (kind == YTB200_VOLUME_RENDER && !tf) return fail(E_ARG, "VolumeRenderSampler needs a transfer function");
    if (kind != YTB200_PROJECTION && n_samples < 1) return fail(E_ARG, "n_samples must be >= 1");
    if (kind == YTB200_PROJECTION && set->centring != YTB200_CELL_CENTRED)
        return fail(E_ARG, "ProjectionSampler needs cell-centred bricks");
    if (kind != YTB200_PROJECTION && set->centring != YTB200_VERTEX_CENTRED)
        return fail(E_ARG, "interpolating samplers need vertex-centred bricks");
    const bool exact = precision == YTB200_PREC_EXACT;
    if (exact && set->storage != YTB200_STORE_F64) return fail(E_ARG, "EXACT precision needs fp64 brick storage");
    if (kind == YTB200_VOLUME_RENDER) {
        for (int i = 0; i < tf->d.n_fits; i++)
            if (tf->d.field_id[i] >= set->n_fields || tf->d.weight_field_id[i] >= set->n_fields)
                return fail(E_ARG, "transfer function table %d refers to field beyond n_fields=%d", i, set->n_fields);
    }
    DeviceGuard g(f->ctx->device);
    cudaStream_t st = f->ctx->stream;
    int launches = 0;
    if (!f->image_is_host) { int rcb = ensure_buffers(f); if (rcb) return rcb; }
    if (stats) CK(cudaEventRecord(f->ctx->ev0, st));
    // early ray termination: the composition order is walked from its far end, so the list goes in reversed
    const bool ert = f->ert_eps > 0.f && kind == YTB200_VOLUME_RENDER && !exact && set->n_fields == 1 && set->storage == YTB200_STORE_F32 &&
                     tf->d.colour_fast && tf->d.grey && !tf->d.light_on;
    std::vector<int64_t> reversed;
    if (ert && n_order > 1) {
        reversed.assign(order, order + n_order);
        std::reverse(reversed.begin(), reversed.end());
        order = reversed.data();
    }
    int rc = prepare_order(f, set, order, n_order, &launches);
    if (rc) return rc;
    if (n_order > 0 || partial) {
        CastParams p{};
        p.cam = f->cam;
        p.bricks = set->dev; p.order = f->order_dev; p.extent = f->extent_dev; p.chunk_extent = f->extent_dev + f->order_cap; p.n_order = (int)n_order;
        p.n_samples = n_samples < 1 ? 1 : n_samples;
        p.inv_ns = 1.0 / (double)p.n_samples;
        {
            const double nn = (double)p.n_samples, s1 = 0.5 * nn * (nn - 1.0);
            p.iproj_s0 = (float)nn; p.iproj_s1 = (float)s1; p.iproj_s2 = (float)(nn * (nn - 1.0) * (2.0 * nn - 1.0) / 6.0); p.iproj_s3 = (float)(s1 * s1);
        }
        p.image = f->image; p.zbuffer = f->zbuffer; p.partial = partial; p.partial_fmt = partial_fmt;
        p.tail_counter = tail_counter; p.tail_flag = tail_flag; p.tail_seq = tail_seq;
        if (partial && f->zbuffer_in) p.zbuffer = f->zbuffer_in;
        if (partial && f->zbuffer_in_ones) p.zbuffer = nullptr;
        if (tf) p.tf = tf->d;
        p.counters = f->ctx->counters;
        const unsigned grid = (unsigned)(((f->cam.nx + TILE_I - 1) / TILE_I) * ((f->cam.ny + TILE_J - 1) / TILE_J));
        const bool nf1 = set->n_fields == 1;
        const bool f64 = set->storage == YTB200_STORE_F64;
        size_t smem = 0;
        cudaError_t e;
        if (kind == YTB200_VOLUME_RENDER) {
            const bool colour = !exact && nf1 && tf->d.colour_fast && !tf->d.light_on;
            if (exact) smem = (size_t)tf->d.total_bins * 3 * sizeof(double);
            else if (colour) smem = ((size_t)tf->d.colour_rows << tf->d.colour_copy_shift) * sizeof(float4) * (tf->d.colour_delta ? 2 : 1);
            else smem = (size_t)tf->d.total_bins * sizeof(float2);
            if (smem > 200 * 1024) return fail(E_UNSUPPORTED, "transfer function tables (%zu B) exceed shared memory", smem);
            if (colour && ert) {
                p.ert_eps = f->ert_eps;
                if (stats) CK(cudaMemsetAsync(f->ctx->counters + 1, 0, sizeof(unsigned long long), st));
                e = launch_ert(p, grid, smem, (tf->d.colour_zero_inside ? TFM_ZERO_INSIDE : 0) | (tf->d.colour_edge ? TFM_EDGE : 0),
                               partial != nullptr, p.n_order > 32, st);
            } else if (colour) {
                const int mode = (tf->d.grey ? TFM_GREY : 0) | (tf->d.colour_zero_inside ? TFM_ZERO_INSIDE : 0) | (tf->d.colour_edge ? TFM_EDGE : 0);
                const bool pt = partial != nullptr;
                switch (mode) {
                case 0: e = launch_kind<KIND_VR, TFM_COLOUR + 0>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 1: e = launch_kind<KIND_VR, TFM_COLOUR + 1>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 2: e = launch_kind<KIND_VR, TFM_COLOUR + 2>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 3: e = launch_kind<KIND_VR, TFM_COLOUR + 3>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 4: e = launch_kind<KIND_VR, TFM_COLOUR + 4>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 5: e = launch_kind<KIND_VR, TFM_COLOUR + 5>(p, grid, smem, st, exact, f64, pt, nf1); break;
                case 6: e = launch_kind<KIND_VR, TFM_COLOUR + 6>(p, grid, smem, st, exact, f64, pt, nf1); break;
                default: e = launch_kind<KIND_VR, TFM_COLOUR + 7>(p, grid, smem, st, exact, f64, pt, nf1); break;
                }
            } else e = launch_kind<KIND_VR, TFM_GENERAL>(p, grid, smem, st, exact, f64, partial != nullptr, nf1);
        } else if (kind == YTB200_INTERPOLATED_PROJECTION) {
            e = launch_kind<KIND_IPROJ, TFM_NONE>(p, grid, 0, st, exact, f64, partial != nullptr, nf1);
        } else if (!exact && !f64 && !getenv("YTB200_PROJ_DDA")) {
            // fp32 storage, FAST: the slab walk (proj_kernels.cuh); image_samplers.pyx:337-343 sums min(4, n_fields) fields
            ProjMap pm = make_proj_map(f->cam, n_order > 0 ? &set->host[(size_t)order[0]] : nullptr, f->image_is_host || f->band_tiles >= 0);
            if (f->band_tiles >= 0) { pm.s_first = f->band_first; pm.n_stiles = f->band_tiles; }
            e = launch_proj(p, pm, std::min(4, set->n_fields), partial != nullptr, st);
        } else {
            e = launch_kind<KIND_PROJ, TFM_NONE>(p, grid, 0, st, exact, f64, partial != nullptr, nf1);
        }
        if (e != cudaSuccess) return fail(-100 - (int)e, "cast launch: %s", cudaGetErrorString(e));
        launches++;
    }
    if (stats) {
        CK(cudaEventRecord(f->ctx->ev1, st));
        CK(cudaEventSynchronize(f->ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, f->ctx->ev0, f->ctx->ev1));
        stats->kernel_ms = ms;
        stats->n_launches = launches;
        stats->n_cells = 0;
        stats->n_samples = 0;
        if (ert && (n_order > 0 || partial)) {
            // cells the terminated rays actually sampled (the algorithmic count is ytb200_frame_count's)
            unsigned long long done = 0;
            CK(cudaMemcpyAsync(&done, f->ctx->counters + 1, sizeof(done), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            stats->n_cells = (int64_t)done;
            stats->n_samples = (int64_t)done * (n_samples < 1 ? 1 : n_samples);
        }
    }
    return 0;
}

extern "C" int ytb200_frame_set_early_termination(ytb200_frame *f, double eps)
{
    if (!f) return fail(E_ARG, "frame is NULL");
    if (!(eps >= 0.0) || eps >= 1.0) return fail(E_ARG, "early termination threshold %g not in [0, 1)", eps);
    f->ert_eps = (float)eps;
    return 0;
}

extern "C" int ytb200_frame_cast(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                 int64_t n_order, const ytb200_tf *tf, int n_samples, int precision, ytb200_stats *stats)
{
    return cast_common(f, kind, set, order, n_order, tf, n_samples, precision, nullptr, stats);
}

extern "C" int ytb200_frame_cast_partial(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                         int64_t n_order, const ytb200_tf *tf, int n_samples, int precision,
                                         double *partial_dev, ytb200_stats *stats)
{
    if (!partial_dev) return fail(E_ARG, "partial_dev is NULL");
    return cast_common(f, kind, set, order, n_order, tf, n_samples, precision, partial_dev, stats);
}

extern "C" int ytb200_frame_cast_partial_f32(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                             int64_t n_order, const ytb200_tf *tf, int n_samples, int precision,
                                             float *partial_dev, ytb200_stats *stats)
{
    if (!partial_dev) return fail(E_ARG, "partial_dev is NULL");
    if (precision != YTB200_PREC_FAST) return fail(E_ARG, "fp32 partial images are for FAST precision only");
    return cast_common(f, kind, set, order, n_order, tf, n_samples, precision, reinterpret_cast<double *>(partial_dev), stats, PF_AB32);
}

extern "C" int ytb200_frame_count(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                  int64_t n_order, int n_samples, ytb200_stats *stats)
{
    if (!f || !set || !stats || (!order && n_order > 0)) return fail(E_ARG, "NULL argument");
    if (kind < 0 || kind > 2) return fail(E_ARG, "bad sampler kind %d", kind);
    DeviceGuard g(f->ctx->device);
    cudaStream_t st = f->ctx->stream;
    int launches = 0;
    int rc = prepare_order(f, set, order, n_order, &launches);
    if (rc) return rc;
    CK(cudaMemsetAsync(f->ctx->counters, 0, 8 * sizeof(unsigned long long), st));
    CastParams p{};
    p.cam = f->cam;
    p.bricks = set->dev; p.order = f->order_dev; p.extent = f->extent_dev; p.chunk_extent = f->extent_dev + f->order_cap; p.n_order = (int)n_order;
    p.zbuffer = f->zbuffer; p.counters = f->ctx->counters;
    const unsigned grid = (unsigned)(((f->cam.nx + TILE_I - 1) / TILE_I) * ((f->cam.ny + TILE_J - 1) / TILE_J));
    CK(cudaEventRecord(f->ctx->ev0, st));
    count_kernel<<<grid, CTA_THREADS, 0, st>>>(p, kind);
    CK(cudaGetLastError());
    CK(cudaEventRecord(f->ctx->ev1, st));
    unsigned long long n = 0;
    CK(cudaMemcpyAsync(&n, f->ctx->counters, sizeof n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, f->ctx->ev0, f->ctx->ev1));
    stats->kernel_ms = ms;
    stats->n_launches = launches + 1;
    stats->n_cells = (int64_t)n;
    stats->n_samples = (int64_t)n * (kind == YTB200_PROJECTION ? 1 : (n_samples < 1 ? 1 : n_samples));
    return 0;
}

extern "C" int ytb200_partial_compose(ytb200_ctx *ctx, const double *earlier, const double *later, double *out, int64_t n)
{
    if (!ctx || !earlier || !later || !out || n < 0) return fail(E_ARG, "bad argument");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    compose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const double4 *)earlier, (const double4 *)later, (double4 *)out, (long long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ytb200_partial_apply(ytb200_ctx *ctx, const double *partial, double *image, int64_t n)
{
    if (!ctx || !partial || !image || n < 0) return fail(E_ARG, "bad argument");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const double4 *)partial, (double4 *)image, (long long)n);
    CK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// octree path: Octree<int> (_octree_raytracing.hpp) built on the host, flattened for the device
// ------------------------------------------------------------------------------------------------
struct ytb200_octree {
    ytb200_ctx *ctx = nullptr;
    int max_depth = 0;
    double dle[3], dre[3], size[3];
    std::vector<int> child;   // [n_nodes][8]
    std::vector<int> leaf;    // [n_nodes][2] (terminal, key)
    int max_key = -1;
    bool dirty = true;
    int *child_dev = nullptr, *leaf_dev = nullptr;
    double *cells_dev = nullptr;
    int64_t n_cells = 0;
};

static int oct_new_node(ytb200_octree *o)
{
    const int n = (int)(o->leaf.size() / 2);
    o->child.insert(o->child.end(), 8, -1);
    o->leaf.push_back(0); o->leaf.push_back(0);
    return n;
}

extern "C" int ytb200_octree_create(ytb200_ctx *ctx, int max_depth, const double dle[3], const double dre[3], ytb200_octree **out)
{
    if (!ctx || !dle || !dre || !out) return fail(E_ARG, "NULL argument");
    if (max_depth < 1 || max_depth > 30) return fail(E_ARG, "octree depth %d outside [1, 30]", max_depth);
    auto *o = new ytb200_octree;
    o->ctx = ctx; o->max_depth = max_depth;
    for (int i = 0; i < 3; i++) { o->dle[i] = dle[i]; o->dre[i] = dre[i]; o->size[i] = dre[i] - dle[i]; }
    oct_new_node(o);
    *out = o;
    return 0;
}

extern "C" int ytb200_octree_destroy(ytb200_octree *o)
{
    if (!o) return 0;
    DeviceGuard g(o->ctx->device);
    cudaFree(o->child_dev); cudaFree(o->leaf_dev); cudaFree(o->cells_dev);
    delete o;
    return 0;
}

// insert_node (hpp:137-170): one level per bit of ipos, from bit max_depth-1 down to bit max_depth-lvl
extern "C" int ytb200_octree_add_nodes(ytb200_octree *o, int64_t n, const int32_t *ipos, const int32_t *lvl, const int32_t *key)
{
    if (!o || (n > 0 && (!ipos || !lvl || !key))) return fail(E_ARG, "NULL argument");
    for (int64_t c = 0; c < n; c++) {
        if (lvl[c] < 0 || lvl[c] > o->max_depth) return fail(E_ARG, "cell %lld: level %d beyond the octree depth %d", (long long)c, lvl[c], o->max_depth);
        if (key[c] < 0) return fail(E_ARG, "cell %lld: negative key", (long long)c);
        uint64_t mask = (uint64_t)1 << (o->max_depth - 1);
        int node = 0;
        for (int ibit = o->max_depth - 1; ibit >= o->max_depth - lvl[c]; --ibit) {
            int iflat = 0;
            for (int d = 0; d < 3; d++) iflat = (iflat << 1) | (((uint64_t)(int64_t)ipos[3 * c + d] & mask) ? 1 : 0);
            mask >>= 1;
            if (o->child[8 * (size_t)node + iflat] < 0) {
                const int nn = oct_new_node(o);
                o->child[8 * (size_t)node + iflat] = nn;
            }
            node = o->child[8 * (size_t)node + iflat];
        }
        o->leaf[2 * (size_t)node] = 1;
        o->leaf[2 * (size_t)node + 1] = key[c];
        if (key[c] > o->max_key) o->max_key = key[c];
    }
    o->dirty = true;
    return 0;
}

extern "C" int64_t ytb200_octree_node_count(const ytb200_octree *o) { return o ? (int64_t)(o->leaf.size() / 2) : 0; }

extern "C" int ytb200_octree_set_cells(ytb200_octree *o, int64_t n_cells, const double *cells_host)
{
    if (!o || n_cells < 1 || !cells_host) return fail(E_ARG, "bad cell array");
    DeviceGuard g(o->ctx->device);
    if (n_cells != o->n_cells) {
        cudaFree(o->cells_dev); o->cells_dev = nullptr; o->n_cells = 0;
        CK(cudaMalloc(&o->cells_dev, (size_t)n_cells * 14 * sizeof(double)));
        o->n_cells = n_cells;
    }
    CK(cudaMemcpyAsync(o->cells_dev, cells_host, (size_t)n_cells * 14 * sizeof(double), cudaMemcpyHostToDevice, o->ctx->stream));
    CK(cudaStreamSynchronize(o->ctx->stream));
    return 0;
}

extern "C" int ytb200_frame_cast_octree(ytb200_frame *f, int kind, ytb200_octree *o, const ytb200_tf *tf, int n_samples,
                                        int precision, ytb200_stats *stats)
{
    if (!f || !o) return fail(E_ARG, "NULL argument");
    if (kind < 0 || kind > 2) return fail(E_ARG, "bad sampler kind %d", kind);
    if (precision != YTB200_PREC_FAST && precision != YTB200_PREC_EXACT) return fail(E_ARG, "bad precision %d", precision);
    if (f->ctx != o->ctx) return fail(E_ARG, "frame and octree belong to different contexts");
    if (!f->pos_px || !f->dir_px) return fail(E_ARG, "cast_through_octree reads per-pixel vp_pos / vp_dir: the frame has none (plane-parallel description)");
    if (!o->cells_dev) return fail(E_STATE, "ytb200_octree_set_cells first");
    if (o->max_key >= o->n_cells) return fail(E_ARG, "octree key %d beyond the %lld cells", o->max_key, (long long)o->n_cells);
    if (kind == YTB200_VOLUME_RENDER) {
        if (!tf) return fail(E_ARG, "VolumeRenderSampler needs a transfer function");
        if (tf->d.light_on) return fail(E_UNSUPPORTED, "the lit sampler has no octree cast");
        for (int i = 0; i < tf->d.n_fits; i++)
            if (tf->d.field_id[i] >= 1 || tf->d.weight_field_id[i] >= 1)
                return fail(E_ARG, "transfer function table %d refers to a field beyond the octree's single field", i);
    }
    if (kind != YTB200_PROJECTION && n_samples < 1) return fail(E_ARG, "n_samples must be >= 1");
    DeviceGuard g(f->ctx->device);
    cudaStream_t st = f->ctx->stream;
    if (o->dirty) {
        cudaFree(o->child_dev); cudaFree(o->leaf_dev); o->child_dev = o->leaf_dev = nullptr;
        CK(cudaMalloc(&o->child_dev, o->child.size() * sizeof(int)));
        CK(cudaMalloc(&o->leaf_dev, o->leaf.size() * sizeof(int)));
        CK(cudaMemcpyAsync(o->child_dev, o->child.data(), o->child.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(o->leaf_dev, o->leaf.data(), o->leaf.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        o->dirty = false;
    }
    { int rcb = ensure_buffers(f); if (rcb) return rcb; }
    if (stats) CK(cudaEventRecord(f->ctx->ev0, st));
    OctParams p{};
    p.nx = f->cam.nx; p.ny = f->cam.ny;
    p.pos_px = f->pos_px; p.dir_px = f->dir_px; p.image = f->image;
    p.child = o->child_dev; p.leaf = o->leaf_dev; p.cells = o->cells_dev;
    for (int i = 0; i < 3; i++) { p.dle[i] = o->dle[i]; p.dre[i] = o->dre[i]; p.size[i] = o->size[i]; }
    p.n_samples = n_samples < 1 ? 1 : n_samples;
    p.inv_ns = 1.0 / (double)p.n_samples;
    if (tf) p.tf = tf->d;
    const unsigned grid = (unsigned)(((p.nx + OCT_TILE_I - 1) / OCT_TILE_I) * ((p.ny + OCT_TILE_J - 1) / OCT_TILE_J));
    if (kind == YTB200_VOLUME_RENDER && precision == YTB200_PREC_FAST) {
        const size_t smem = (size_t)tf->d.total_bins * sizeof(float2);
        if (smem > 200 * 1024) return fail(E_UNSUPPORTED, "transfer function tables (%zu B) exceed shared memory", smem);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(octree_cast_kernel<KIND_VR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        octree_cast_kernel<KIND_VR, true><<<grid, OCT_THREADS, smem, st>>>(p);
    } else if (kind == YTB200_VOLUME_RENDER) octree_cast_kernel<KIND_VR, false><<<grid, OCT_THREADS, 0, st>>>(p);
    else if (kind == YTB200_INTERPOLATED_PROJECTION) octree_cast_kernel<KIND_IPROJ, false><<<grid, OCT_THREADS, 0, st>>>(p);
    else octree_cast_kernel<KIND_PROJ, false><<<grid, OCT_THREADS, 0, st>>>(p);
    CK(cudaGetLastError());
    if (stats) {
        CK(cudaEventRecord(f->ctx->ev1, st));
        CK(cudaEventSynchronize(f->ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, f->ctx->ev0, f->ctx->ev1));
        stats->kernel_ms = ms; stats->n_launches = 1; stats->n_cells = 0; stats->n_samples = 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// binary swap over NVLink peer memory (replaces AMRKDTree.reduce_tree_images, amr_kdtree.py:297-323)
// ------------------------------------------------------------------------------------------------
// Per rank ONE cudaMalloc'd slab, shared with the peers through a CUDA IPC handle:
//   [0, 512)            64 x u64: flag[j] (j < 16) "partner_j's buffer_j of frame `seq` is complete" (pushed by that
//                       partner's kernel tail), final[w] at 16 + w "rank w has stored its strip of frame `seq`",
//                       error word at 56, CTA counter of the tail signals at 57
//   two parity sets of buffers: buffer_0 (n_pixels, the cast's output), buffer_{j+1} (ceil(cap_j / 2)) ...
//   the final region (n_pixels, partial format): every rank's last round stores its composed strip into the destination
//   ranks' copy of this region; the destination applies it to the initial colours -- frames can come and go, the slab stays
// The layout is identical on every rank, so a peer's buffer address is its slab base + the same offset.
struct SwapRound { int partner; int64_t p0, p1, k0, k1; int earlier; };

struct ytb200_swap {
    ytb200_ctx *ctx = nullptr;
    int64_t npx = 0;
    int rank = 0, world = 1, fmt = PF_AB32, bpp = 32, k = 0;
    unsigned char *slab = nullptr;
    size_t slab_bytes = 0, set_bytes = 0;
    std::vector<size_t> buf_off;             // offset of buffer_j inside one parity set
    std::vector<unsigned char *> peer_slab;  // [world], own slab at [rank]
    size_t final_off = 0;                    // offset of the final region
    std::vector<SwapRound> rounds;
    unsigned long long seq = 0;              // frames cast so far
    unsigned long long composited = 0;       // frames composited so far
    // host-image frames: this rank's strip of the initial colours is fetched on a side stream WHILE the cast runs
    cudaStream_t side = nullptr;
    cudaEvent_t ev_c0 = nullptr, ev_frame = nullptr;
    double4 *c0_stage = nullptr;             // device copy of the strip
    int64_t stage_cap = 0;
    bool c0_staged = false;                  // the current frame's strip is (being) staged
};

static const size_t SWAP_HDR = 512;
enum { SWAP_FLAG_FINAL = 16, SWAP_WORD_ERROR = 56, SWAP_WORD_COUNTER = 57 };

static unsigned long long *swap_flag(unsigned char *slab, int idx) { return reinterpret_cast<unsigned long long *>(slab) + idx; }

extern "C" int ytb200_partial_format(int kind, const ytb200_tf *tf, int n_fields, int precision)
{
    if (kind == YTB200_VOLUME_RENDER) {
        if (!tf) return fail(E_ARG, "VolumeRenderSampler needs a transfer function");
        if (precision == YTB200_PREC_EXACT) return PF_AB64;
        if (n_fields == 1 && tf->d.colour_fast && !tf->d.light_on && tf->d.colour_nonneg) return tf->d.grey ? PF_A32G : PF_A32;
        return PF_AB32;
    }
    // image_samplers.pyx:337-343 / 393-401: the first min(4, n) / min(3, n) fields are summed
    const int n = kind == YTB200_PROJECTION ? std::min(4, n_fields) : std::min(3, n_fields);
    return n <= 1 ? PF_SUM1 : (n == 2 ? PF_SUM2 : PF_SUM4);
}

extern "C" int ytb200_partial_bytes(int format) { return (format >= 0 && format < PF_COUNT) ? partial_bytes(format) : 0; }

extern "C" int ytb200_swap_create(ytb200_ctx *ctx, int64_t n_pixels, int rank, int world, int format, ytb200_swap **out)
{
    if (!ctx || !out) return fail(E_ARG, "NULL argument");
    if (world < 1 || world > SWAP_MAX_DST || (world & (world - 1))) return fail(E_ARG, "world must be a power of two <= %d", SWAP_MAX_DST);
    if (rank < 0 || rank >= world || n_pixels < 1) return fail(E_ARG, "bad rank / n_pixels");
    if (format < 0 || format >= PF_COUNT) return fail(E_ARG, "bad partial format %d", format);
    DeviceGuard g(ctx->device);
    auto *s = new ytb200_swap;
    s->ctx = ctx; s->npx = n_pixels; s->rank = rank; s->world = world; s->fmt = format; s->bpp = partial_bytes(format);
    int k = 0;
    while ((1 << k) < world) k++;
    s->k = k;
    size_t off = 0;
    int64_t cap = n_pixels;
    for (int j = 0; j <= k; j++) {
        s->buf_off.push_back(off);
        off += ((size_t)cap * (size_t)s->bpp + 255) / 256 * 256;
        cap = (cap + 1) / 2;
    }
    s->set_bytes = off;
    s->final_off = SWAP_HDR + 2 * off;
    s->slab_bytes = s->final_off + (size_t)n_pixels * (size_t)s->bpp;
    CK(cudaMalloc(&s->slab, s->slab_bytes));
    CK(cudaMemsetAsync(s->slab, 0, SWAP_HDR, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s->peer_slab.assign((size_t)world, nullptr);
    s->peer_slab[(size_t)rank] = s->slab;
    CK(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s->ev_c0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->ev_frame, cudaEventDisableTiming));
    *out = s;
    return 0;
}

extern "C" int ytb200_swap_format(const ytb200_swap *s) { return s ? s->fmt : -1; }

extern "C" int ytb200_swap_destroy(ytb200_swap *s)
{
    if (!s) return 0;
    DeviceGuard g(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    for (int w = 0; w < s->world; w++) {
        if (w == s->rank) continue;
        if (s->peer_slab[(size_t)w]) cudaIpcCloseMemHandle(s->peer_slab[(size_t)w]);
    }
    cudaFree(s->slab);
    if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
    if (s->ev_c0) cudaEventDestroy(s->ev_c0);
    if (s->ev_frame) cudaEventDestroy(s->ev_frame);
    cudaFree(s->c0_stage);
    delete s;
    return 0;
}

extern "C" int ytb200_swap_handle(ytb200_swap *s, void *handle64)
{
    if (!s || !handle64) return fail(E_ARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(s->ctx->device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->slab));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int ytb200_swap_open(ytb200_swap *s, int peer_rank, const void *slab_handle64)
{
    if (!s || !slab_handle64) return fail(E_ARG, "NULL argument");
    if (peer_rank < 0 || peer_rank >= s->world || peer_rank == s->rank) return fail(E_ARG, "bad peer rank %d", peer_rank);
    DeviceGuard g(s->ctx->device);
    cudaIpcMemHandle_t h;
    void *p = nullptr;
    memcpy(&h, slab_handle64, 64);
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_slab[(size_t)peer_rank] = static_cast<unsigned char *>(p);
    return 0;
}

extern "C" int ytb200_swap_plan(ytb200_swap *s, int n_rounds, const int32_t *partner, const int64_t *p0, const int64_t *p1,
                                const int64_t *k0, const int64_t *k1, const int32_t *mine_is_earlier)
{
    if (!s || (n_rounds > 0 && (!partner || !p0 || !p1 || !k0 || !k1 || !mine_is_earlier))) return fail(E_ARG, "NULL argument");
    if (n_rounds != s->k) return fail(E_ARG, "%d rounds given, log2(world) = %d", n_rounds, s->k);
    s->rounds.clear();
    int64_t cap = s->npx, q0 = 0, q1 = s->npx;
    for (int j = 0; j < n_rounds; j++) {
        if (partner[j] < 0 || partner[j] >= s->world || partner[j] == s->rank) return fail(E_ARG, "round %d: bad partner %d", j, partner[j]);
        if (p0[j] != q0 || p1[j] != q1) return fail(E_ARG, "round %d: region [%lld,%lld) does not continue the previous keep", j, (long long)p0[j], (long long)p1[j]);
        if (k0[j] < p0[j] || k1[j] > p1[j] || k1[j] < k0[j] || (k0[j] != p0[j] && k1[j] != p1[j])) return fail(E_ARG, "round %d: bad keep range", j);
        cap = (cap + 1) / 2;
        if (k1[j] - k0[j] > cap) return fail(E_ARG, "round %d: keep range longer than its buffer", j);
        s->rounds.push_back({partner[j], p0[j], p1[j], k0[j], k1[j], mine_is_earlier[j] ? 1 : 0});
        q0 = k0[j]; q1 = k1[j];
    }
    return 0;
}

// This rank's partial cast of the next frame, straight into the swap's buffer_0 and in the swap's format; the cast
// kernel's last CTA tells the round-0 partner.
extern "C" int ytb200_swap_cast(ytb200_swap *s, ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                int64_t n_order, const ytb200_tf *tf, int n_samples, int precision, ytb200_stats *stats)
{
    if (!s || !f || !set) return fail(E_ARG, "NULL argument");
    if ((int64_t)f->cam.nx * f->cam.ny != s->npx) return fail(E_ARG, "frame has %lld pixels, swap was created for %lld", (long long)f->cam.nx * f->cam.ny, (long long)s->npx);
    if (f->ctx != s->ctx) return fail(E_ARG, "frame and swap use different contexts");
    if ((int)s->rounds.size() != s->k) return fail(E_STATE, "ytb200_swap_plan has not been called");
    if (s->seq != s->composited) return fail(E_STATE, "the previous frame has been cast but not composited");
    const int want = ytb200_partial_format(kind, tf, set->n_fields, precision);
    if (want < 0) return want;
    // a swap created for a wider format of the same family can carry a narrower map; anything else is a mismatch
    const bool ok = want == s->fmt || (want == PF_AB32 && s->fmt == PF_AB64) ||
                    ((want == PF_A32 || want == PF_A32G) && (s->fmt == PF_AB32 || s->fmt == PF_AB64)) ||
                    (want == PF_SUM1 && (s->fmt == PF_SUM2 || s->fmt == PF_SUM4)) || (want == PF_SUM2 && s->fmt == PF_SUM4);
    if (!ok) return fail(E_ARG, "swap was created for partial format %d, this cast produces format %d (ytb200_partial_format)", s->fmt, want);
    if (s->fmt == PF_AB32 && precision != YTB200_PREC_FAST) return fail(E_ARG, "fp32 partial images are for FAST precision only");
    for (int j = 0; j < s->k; j++)
        if (!s->peer_slab[(size_t)s->rounds[(size_t)j].partner]) return fail(E_STATE, "peer %d not opened", s->rounds[(size_t)j].partner);
    const unsigned long long seq = s->seq + 1;
    s->c0_staged = false;
    if (f->c0_in && s->k > 0) {
        // the initial colours of this rank's final strip, read over PCIe on the side stream while the cast runs
        // (after the previous frame's stores into the same host image have completed)
        DeviceGuard g(s->ctx->device);
        const SwapRound &lr = s->rounds[(size_t)s->k - 1];
        const int64_t n = lr.k1 - lr.k0;
        if (n > s->stage_cap) {
            CK(cudaStreamSynchronize(s->side));
            cudaFree(s->c0_stage); s->c0_stage = nullptr; s->stage_cap = 0;
            CK(cudaMalloc(&s->c0_stage, (size_t)n * sizeof(double4)));
            s->stage_cap = n;
        }
        if (n > 0) {
            CK(cudaEventRecord(s->ev_frame, s->ctx->stream));
            CK(cudaStreamWaitEvent(s->side, s->ev_frame, 0));
            CK(cudaMemcpyAsync(s->c0_stage, reinterpret_cast<const double4 *>(f->c0_in) + lr.k0, (size_t)n * sizeof(double4), cudaMemcpyDefault, s->side));
            CK(cudaEventRecord(s->ev_c0, s->side));
            s->c0_staged = true;
        }
    }
    unsigned char *buf0 = s->slab + SWAP_HDR + (size_t)(seq & 1) * s->set_bytes + s->buf_off[0];
    unsigned long long *flag = s->k > 0 ? swap_flag(s->peer_slab[(size_t)s->rounds[0].partner], 0) : nullptr;
    int rc = cast_common(f, kind, set, order, n_order, tf, n_samples, precision, reinterpret_cast<double *>(buf0), stats, s->fmt,
                         reinterpret_cast<unsigned *>(swap_flag(s->slab, SWAP_WORD_COUNTER)), flag, seq);
    if (rc) return rc;
    s->seq = seq;
    return 0;
}

template <int FMT>
static int swap_run(ytb200_swap *s, ytb200_frame *f, int broadcast, double *host_image)
{
    cudaStream_t st = s->ctx->stream;
    const unsigned long long seq = s->seq;
    const size_t set = SWAP_HDR + (size_t)(seq & 1) * s->set_bytes;
    const size_t bpp = (size_t)s->bpp;
    int *err = reinterpret_cast<int *>(swap_flag(s->slab, SWAP_WORD_ERROR));
    unsigned *counter = reinterpret_cast<unsigned *>(swap_flag(s->slab, SWAP_WORD_COUNTER));
    auto buf = [&](unsigned char *slab, int j) { return slab + set + s->buf_off[(size_t)j]; };
    { int rcb = ensure_buffers(f); if (rcb) return rcb; }
    const double4 *c0 = reinterpret_cast<const double4 *>(f->c0_in ? f->c0_in : f->image);
    const bool dest = broadcast || s->rank == 0;
    auto blocks = [&](long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)s->ctx->sm_count * 8)); };
    for (int w = 0; w < s->world; w++)
        if ((broadcast || w == 0) && !s->peer_slab[(size_t)w]) return fail(E_STATE, "destination rank %d not opened", w);
    int64_t f0 = 0, f1 = s->npx;
    for (int j = 0; j < s->k; j++) {
        const SwapRound &r = s->rounds[(size_t)j];
        const bool last = j + 1 == s->k;
        SwapRoundArgs a{};
        a.mine = buf(s->slab, j) + (size_t)(r.k0 - r.p0) * bpp;
        a.theirs = buf(s->peer_slab[(size_t)r.partner], j) + (size_t)(r.k0 - r.p0) * bpp;
        a.n = (long long)(r.k1 - r.k0);
        a.mine_is_earlier = r.earlier;
        a.wait_flag = swap_flag(s->slab, j); a.seq = seq; a.error = err;
        a.counter = counter;
        f0 = r.k0; f1 = r.k1;
        if (!last) {
            a.out.n = 1; a.out.buf[0] = buf(s->slab, j + 1);
            a.done.n = 1; a.done.flag[0] = swap_flag(s->peer_slab[(size_t)s->rounds[(size_t)j + 1].partner], j + 1);
            swap_round_kernel<FMT, false><<<blocks(a.n), 256, 0, st>>>(a);
        } else if (host_image) {
            // finished fp64 pixels of this rank's strip into the caller's page-locked image over this GPU's PCIe link
            a.c0 = c0 + f0; a.image = reinterpret_cast<double4 *>(host_image) + f0;
            if (s->c0_staged && c0 == reinterpret_cast<const double4 *>(host_image)) {
                // The strip of initial colours was staged in HBM during the cast: finish the pixels in place there and let
                // the copy engine carry them out -- one bulk PCIe write instead of 32-byte stores from the SMs (r02j: the
                // in-kernel stores of eight ranks together sustained 48 GB/s, 2.8 ms for the 134 MB image) -- then tell
                // the destination rank from a one-thread kernel that is stream-ordered behind the copy.
                CK(cudaStreamWaitEvent(st, s->ev_c0, 0));
                a.c0 = s->c0_stage; a.image = s->c0_stage;
                a.done.n = 0;
                swap_round_kernel<FMT, true><<<blocks(a.n), 256, 0, st>>>(a);
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(reinterpret_cast<double4 *>(host_image) + f0, s->c0_stage, (size_t)a.n * sizeof(double4), cudaMemcpyDefault, st));
                swap_signal_kernel<<<1, 1, 0, st>>>(swap_flag(s->peer_slab[0], SWAP_FLAG_FINAL + s->rank), seq);
            } else {
                a.done.n = 1; a.done.flag[0] = swap_flag(s->peer_slab[0], SWAP_FLAG_FINAL + s->rank);
                swap_round_kernel<FMT, true><<<blocks(a.n), 256, 0, st>>>(a);
            }
        } else {
            for (int w = 0; w < s->world; w++) {
                if (!broadcast && w != 0) continue;
                a.out.buf[a.out.n++] = s->peer_slab[(size_t)w] + s->final_off + (size_t)f0 * bpp;
                a.done.flag[a.done.n++] = swap_flag(s->peer_slab[(size_t)w], SWAP_FLAG_FINAL + s->rank);
            }
            swap_round_kernel<FMT, false><<<blocks(a.n), 256, 0, st>>>(a);
        }
        CK(cudaGetLastError());
    }
    if (host_image) {
        if (s->k == 0) {
            SwapApplyArgs a{};
            a.partial = buf(s->slab, 0); a.c0 = c0; a.image = reinterpret_cast<double4 *>(host_image); a.n = (long long)s->npx;
            swap_apply_kernel<FMT><<<blocks(a.n), 256, 0, st>>>(a);
        } else if (s->rank == 0) {
            swap_wait_kernel<<<1, 32, 0, st>>>(swap_flag(s->slab, SWAP_FLAG_FINAL), s->world, seq, err);
        }
    } else if (dest) {
        SwapApplyArgs a{};
        a.partial = s->k == 0 ? buf(s->slab, 0) : s->slab + s->final_off;
        a.c0 = c0; a.image = reinterpret_cast<double4 *>(f->image); a.n = (long long)s->npx;
        a.wait_flags = swap_flag(s->slab, SWAP_FLAG_FINAL); a.n_wait = s->k == 0 ? 0 : s->world; a.seq = seq; a.error = err;
        swap_apply_kernel<FMT><<<blocks(a.n), 256, 0, st>>>(a);
    }
    CK(cudaGetLastError());
    s->composited = seq;
    return 0;
}

static int swap_dispatch(ytb200_swap *s, ytb200_frame *f, int broadcast, double *host_image)
{
    switch (s->fmt) {
    case PF_AB64: return swap_run<PF_AB64>(s, f, broadcast, host_image);
    case PF_AB32: return swap_run<PF_AB32>(s, f, broadcast, host_image);
    case PF_A32: return swap_run<PF_A32>(s, f, broadcast, host_image);
    case PF_A32G: return swap_run<PF_A32G>(s, f, broadcast, host_image);
    case PF_SUM1: return swap_run<PF_SUM1>(s, f, broadcast, host_image);
    case PF_SUM2: return swap_run<PF_SUM2>(s, f, broadcast, host_image);
    default: return swap_run<PF_SUM4>(s, f, broadcast, host_image);
    }
}

static int swap_check(ytb200_swap *s, ytb200_frame *f)
{
    if (!s || !f) return fail(E_ARG, "NULL argument");
    if ((int64_t)f->cam.nx * f->cam.ny != s->npx) return fail(E_ARG, "frame has %lld pixels, swap was created for %lld", (long long)f->cam.nx * f->cam.ny, (long long)s->npx);
    if ((int)s->rounds.size() != s->k) return fail(E_STATE, "ytb200_swap_plan has not been called");
    if (f->ctx != s->ctx) return fail(E_ARG, "frame and swap use different contexts");
    if (s->seq == s->composited) return fail(E_STATE, "ytb200_swap_cast first: no frame waits to be composited");
    return 0;
}

extern "C" int ytb200_swap_composite(ytb200_swap *s, ytb200_frame *f, int broadcast)
{
    int rc = swap_check(s, f);
    if (rc) return rc;
    DeviceGuard g(s->ctx->device);
    return swap_dispatch(s, f, broadcast, nullptr);
}

static double *mapped_host_alias(const void *host);

// Same composite, but every rank's last kernel stores the FINISHED pixels of its strip into `image_host`: a
// page-locked array mapped into every rank's process (yt_b200.parallel.shared_pinned_empty), holding the initial
// colours on entry.  The image leaves the box over all the PCIe links at once; rank 0's stream completes last.
extern "C" int ytb200_swap_composite_host(ytb200_swap *s, ytb200_frame *f, double *image_host)
{
    int rc = swap_check(s, f);
    if (rc) return rc;
    double *img = mapped_host_alias(image_host);
    if (!img) return fail(E_ARG, "image_host is not page-locked and mapped in this process (ytb200_host_alloc / ytb200_host_register)");
    DeviceGuard g(s->ctx->device);
    const double *keep = f->c0_in;
    f->c0_in = img;   // the initial colours are read in place
    rc = swap_dispatch(s, f, 0, img);
    f->c0_in = keep;
    return rc;
}

extern "C" int ytb200_swap_error(ytb200_swap *s, int *error_out)
{
    if (!s || !error_out) return fail(E_ARG, "NULL argument");
    DeviceGuard g(s->ctx->device);
    CK(cudaStreamSynchronize(s->ctx->stream));
    CK(cudaMemcpy(error_out, swap_flag(s->slab, SWAP_WORD_ERROR), sizeof(int), cudaMemcpyDeviceToHost));
    // reported once: a later frame that completes must not inherit this frame's failure
    if (*error_out) CK(cudaMemset(swap_flag(s->slab, SWAP_WORD_ERROR), 0, sizeof(unsigned long long)));
    return 0;
}

// Device-visible alias of a page-locked host allocation (cudaHostAlloc / cudaHostRegister), or nullptr.
static double *mapped_host_alias(const void *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return static_cast<double *>(a.devicePointer);
}

extern "C" int ytb200_frame_bind_host(ytb200_frame *f, const double *image_host, const double *zbuffer_host)
{
    if (!f) return fail(E_ARG, "frame is NULL");
    f->c0_in = nullptr; f->zbuffer_in = nullptr; f->zbuffer_in_ones = false;
    if (!image_host) return 0;   // unbind
    const double *img = mapped_host_alias(image_host);
    if (!img) return fail(E_ARG, "image_host is not page-locked (ytb200_host_alloc / cudaHostRegister)");
    if (zbuffer_host) {
        const double *zb = mapped_host_alias(zbuffer_host);
        if (!zb) return fail(E_ARG, "zbuffer_host is not page-locked");
        f->zbuffer_in = zb;
    } else f->zbuffer_in_ones = true;
    f->c0_in = img;
    return 0;
}

// A cast SHORTER than its PCIe traffic (the slab-walk projection: 2 ms of kernel for 302 MB of fp64 image and z-buffer at
// 2048^2) gains nothing from touching the host arrays in place -- SM-issued PCIe accesses sustain ~34 GB/s, the frame took
// 10.8 ms.  Such a frame is staged instead, in ROW BANDS through three streams: band k's rows of the image and the z-buffer
// come in with the copy engine while band k-1 is cast and band k-2 goes out, so the frame costs about the longer of the
// two transfers.  (A volume render is the other way round: 41 ms of kernel hide the same traffic, in place.)
static int render_projection_banded(ytb200_frame *f, const ytb200_brickset *set, const int64_t *order, int64_t n_order,
                                    int n_samples, int precision, const double *zbuffer_host, double *image_host, ytb200_stats *stats)
{
    DeviceGuard g(f->ctx->device);
    const int nx = f->cam.nx, ny = f->cam.ny;
    const int n_stiles = (nx + TILE_I - 1) / TILE_I;
    const int n_bands = std::min(8, n_stiles);
    if (!f->s_in) {
        CK(cudaStreamCreateWithFlags(&f->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&f->s_out, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&f->ev_start, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&f->ev_out, cudaEventDisableTiming));
    }
    while ((int)f->ev_in.size() < n_bands) {
        cudaEvent_t a, b;
        CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        f->ev_in.push_back(a); f->ev_done.push_back(b);
    }
    cudaStream_t st = f->ctx->stream;
    { int rcb = ensure_buffers(f); if (rcb) return rcb; }
    if (stats) CK(cudaEventRecord(f->ctx->ev0, st));
    // footprints once, on the compute stream (cached while the list is unchanged); the copy-in stream starts behind
    // whatever the compute stream was doing with the frame's buffers
    int launches = 0;
    int rc = prepare_order(f, set, order, n_order, &launches);
    if (rc) return rc;
    CK(cudaEventRecord(f->ev_start, st));
    CK(cudaStreamWaitEvent(f->s_in, f->ev_start, 0));
    CK(cudaStreamWaitEvent(f->s_out, f->ev_start, 0));
    const size_t row_px = (size_t)ny;
    for (int b = 0; b < n_bands; b++) {
        const int s0 = (int)((long long)n_stiles * b / n_bands), s1 = (int)((long long)n_stiles * (b + 1) / n_bands);
        const size_t r0 = (size_t)s0 * TILE_I, r1 = std::min((size_t)nx, (size_t)s1 * TILE_I);
        if (r1 <= r0) continue;
        const size_t px0 = r0 * row_px, npx = (r1 - r0) * row_px;
        CK(cudaMemcpyAsync(f->image + px0 * 4, image_host + px0 * 4, npx * 4 * sizeof(double), cudaMemcpyHostToDevice, f->s_in));
        if (zbuffer_host) CK(cudaMemcpyAsync(f->zbuffer + px0, zbuffer_host + px0, npx * sizeof(double), cudaMemcpyHostToDevice, f->s_in));
        CK(cudaEventRecord(f->ev_in[(size_t)b], f->s_in));
        CK(cudaStreamWaitEvent(st, f->ev_in[(size_t)b], 0));
        if (!zbuffer_host) {
            fill_f64_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(f->zbuffer + px0, 1.0, (long long)npx);
            CK(cudaGetLastError());
        }
        f->band_first = s0; f->band_tiles = s1 - s0;
        rc = cast_common(f, YTB200_PROJECTION, set, order, n_order, nullptr, n_samples, precision, nullptr, nullptr);
        f->band_first = 0; f->band_tiles = -1;
        if (rc) return rc;
        CK(cudaEventRecord(f->ev_done[(size_t)b], st));
        CK(cudaStreamWaitEvent(f->s_out, f->ev_done[(size_t)b], 0));
        CK(cudaMemcpyAsync(image_host + px0 * 4, f->image + px0 * 4, npx * 4 * sizeof(double), cudaMemcpyDeviceToHost, f->s_out));
    }
    CK(cudaEventRecord(f->ev_out, f->s_out));
    CK(cudaStreamWaitEvent(st, f->ev_out, 0));
    if (stats) {
        CK(cudaEventRecord(f->ctx->ev1, st));
        CK(cudaEventSynchronize(f->ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, f->ctx->ev0, f->ctx->ev1));
        stats->kernel_ms = ms; stats->n_launches = n_bands + launches; stats->n_cells = 0; stats->n_samples = 0;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ytb200_render(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order, int64_t n_order,
                             const ytb200_tf *tf, int n_samples, int precision, const double *zbuffer_host,
                             double *image_host_inout, ytb200_stats *stats)
{
    if (!f || !image_host_inout) return fail(E_ARG, "NULL argument");
    // Page-locked host arrays are read and written by the cast kernel itself: every CTA loads the initial colours
    // and max_t of its 16x16 tile when it starts and stores the result when it ends, so the PCIe traffic of the
    // frame is spread over the whole cast instead of bracketing it with two serial copies (r01e: 5.5 of 50.8 ms).
    // The frame's device image is not touched on this route.  Pageable arrays take the copy-in / copy-out route.
    double *img = mapped_host_alias(image_host_inout);
    double *zb = zbuffer_host ? mapped_host_alias(zbuffer_host) : nullptr;
    if (img && (zb || !zbuffer_host) && set && kind == YTB200_PROJECTION && precision == YTB200_PREC_FAST && set->storage == YTB200_STORE_F32 &&
        f->cam.nx >= 4 * TILE_I && !getenv("YTB200_PROJ_DDA") && !getenv("YTB200_NO_BANDS"))
        return render_projection_banded(f, set, order, n_order, n_samples, precision, zbuffer_host, image_host_inout, stats);
    if (img && (zb || !zbuffer_host) && !getenv("YTB200_NO_ZEROCOPY")) {
        double *keep_img = f->image, *keep_zb = f->zbuffer;
        f->image = img;
        f->zbuffer = zb;   // nullptr == ones
        f->image_is_host = true;
        int rc = cast_common(f, kind, set, order, n_order, tf, n_samples, precision, nullptr, stats);
        f->image = keep_img;
        f->zbuffer = keep_zb;
        f->image_is_host = false;
        if (rc) return rc;
        DeviceGuard g(f->ctx->device);
        CK(cudaStreamSynchronize(f->ctx->stream));
        return 0;
    }
    int rc = ytb200_frame_upload(f, image_host_inout, zbuffer_host);
    if (rc) return rc;
    rc = ytb200_frame_cast(f, kind, set, order, n_order, tf, n_samples, precision, stats);
    if (rc) return rc;
    return ytb200_frame_download(f, image_host_inout);
}

extern "C" int ytb200_host_alloc(int64_t bytes, void **out)
{
    if (!out || bytes <= 0) return fail(E_ARG, "bad argument");
    CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}

extern "C" int ytb200_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return 0;
}

extern "C" int ytb200_host_register(void *p, int64_t bytes)
{
    if (!p || bytes <= 0) return fail(E_ARG, "bad argument");
    CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}

extern "C" int ytb200_host_unregister(void *p)
{
    if (p) CK(cudaHostUnregister(p));
    return 0;
}
