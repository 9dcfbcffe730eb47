/*
 * ytb200.h -- C ABI of libytb200.so, the B200 (sm_100a) ray caster that replaces
 * yt's Cython ImageSampler path.  Plain pointers and sizes only; no torch, no
 * Python, no C++ types.  Every entry point returns 0 on success and a negative
 * code on failure; ytb200_last_error() then holds a human-readable message
 * (thread-local).  There is NO CPU fallback: every call needs a CUDA device.
 *
 * Each entry point names the reference interface (paths relative to the yt
 * checkout) it stands in for.  INTEGRATION.md shows the ctypes binding a yt
 * maintainer would add in yt/visualization/volume_rendering/utils.py.
 */
#ifndef YTB200_H
#define YTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YTB200_ABI_VERSION 1
#define YTB200_MAX_FIELDS 6   /* dvs[6]: yt/utilities/lib/image_samplers.pyx:384,474 */
#define YTB200_MAX_TABLES 6   /* fits[6]: yt/utilities/lib/image_samplers.pyx:425-430 */

/* sampler kinds == the three ImageSampler subclasses named by the north star */
enum {
    YTB200_PROJECTION = 0,              /* ProjectionSampler             image_samplers.pyx:327-343 */
    YTB200_INTERPOLATED_PROJECTION = 1, /* InterpolatedProjectionSampler image_samplers.pyx:346-402 */
    YTB200_VOLUME_RENDER = 2            /* VolumeRenderSampler           image_samplers.pyx:405-494 */
};

/* lens modes == the (extent_function, vector_function) pairs of image_samplers.pyx:85-103 */
enum {
    YTB200_LENS_PLANE_PARALLEL = 0, /* lenses.pyx:22-55 + 179-192 */
    YTB200_LENS_PERSPECTIVE = 1,    /* lenses.pyx:60-161 + 196-207 (needs camera_data) */
    YTB200_LENS_GENERIC = 2         /* lenses.pyx:169-175 + 196-207 (per-pixel rays, full image) */
};

/* data centring of a brick set */
enum { YTB200_CELL_CENTRED = 0, YTB200_VERTEX_CENTRED = 1 };
/* device storage of brick data */
enum { YTB200_STORE_F32 = 0, YTB200_STORE_F64 = 1 };
/* arithmetic mode: FAST = fp64 traversal + fp32 sampling/compositing (the product path);
 * EXACT = everything fp64 in the reference's evaluation order (needs STORE_F64). */
enum { YTB200_PREC_FAST = 0, YTB200_PREC_EXACT = 1 };

typedef struct ytb200_ctx ytb200_ctx;           /* one device + one stream */
typedef struct ytb200_brickset ytb200_brickset; /* device-resident PartitionedGrid list */
typedef struct ytb200_tf ytb200_tf;             /* device-resident FieldInterpolationTables */
typedef struct ytb200_frame ytb200_frame;       /* device image + z-buffer + rays of one sampler */

/* ImageSampler.__init__ arguments (image_samplers.pyx:59-130), flattened. */
typedef struct ytb200_camera {
    int32_t lens;            /* YTB200_LENS_* */
    int32_t nx, ny;          /* image.shape[0], image.shape[1] */
    double vp_pos[12];       /* plane-parallel: lens.py:89-94 packing (inv_mat.ravel("F"), back_center) */
    double vp_dir[3];        /* plane-parallel: box_vectors[2] */
    const double *vp_pos_px; /* other lenses: HOST pointer, C-order (nx, ny, 3) */
    const double *vp_dir_px; /* other lenses: HOST pointer, C-order (nx, ny, 3) */
    double center[3];
    double bounds[4];
    double x_vec[3], y_vec[3];
    double width[3];
    int32_t has_camera_data; /* kwargs["camera_data"] given (camera.py:278-280) */
    double camera_data[15];  /* rows: position, width, east, north, normal */
} ytb200_camera;

/* Per-cast statistics, all measured on the device. */
typedef struct ytb200_stats {
    double kernel_ms;    /* CUDA-event time of the cast kernels on the context stream (0 if not timed) */
    int64_t n_launches;  /* kernels launched by the call */
    int64_t n_cells;     /* cells handed to sample() (ytb200_frame_count; an early-termination cast: the cells it sampled) */
    int64_t n_samples;   /* n_cells * n_samples for the sampling kinds, n_cells for PROJECTION */
} ytb200_stats;

/* ---- library ---------------------------------------------------------------------------- */
int ytb200_abi_version(void);
const char *ytb200_last_error(void);
int ytb200_device_count(int *count);

/* ---- context: device + stream.  `stream` is a cudaStream_t (or NULL to create one). --------
 * Replaces the OpenMP parallel region / num_threads of cast_through_kdtree (image_samplers.pyx:172). */
int ytb200_ctx_create(int device, void *stream, ytb200_ctx **out);
int ytb200_ctx_destroy(ytb200_ctx *ctx);
int ytb200_ctx_synchronize(ytb200_ctx *ctx);
void *ytb200_ctx_stream(ytb200_ctx *ctx);
int ytb200_ctx_sm_count(ytb200_ctx *ctx);

/* ---- bricks: PartitionedGrid.__cinit__ (partitioned_grid.pyx:28-63) ------------------------ */
int ytb200_bricks_create(ytb200_ctx *ctx, int64_t n_bricks, int n_fields, int centring, int storage,
                         ytb200_brickset **out);
/* data[f] are HOST fp64 C-contiguous arrays of shape dims (+1 per axis when vertex-centred);
 * mask is a HOST u8 array of shape dims or NULL for "all ones". Converts to `storage` and uploads. */
int ytb200_brick_upload(ytb200_brickset *set, int64_t index, const int64_t dims[3], const double left_edge[3],
                        const double right_edge[3], const double *const *data, const uint8_t *mask);
/* Same, but data[f] / mask are DEVICE pointers already in the set's storage type; not copied, not owned. */
int ytb200_brick_attach(ytb200_brickset *set, int64_t index, const int64_t dims[3], const double left_edge[3],
                        const double right_edge[3], const void *const *dev_data, const uint8_t *dev_mask);
/* As ytb200_brick_attach, for device arrays whose z-rows are `row_pitch` elements apart (>= the row length; the plane
 * pitch is rows-per-plane * row_pitch).  Arrays the library allocates itself (brick_upload, brick_from_cells) pad a
 * row whose byte length is a multiple of 256 B by one 32-byte sector: power-of-two row strides put all rows a ray
 * packet touches into the same L1 sets (1024^3 fp32 cells cast 24 % slower than 1032-pitched ones). */
int ytb200_brick_attach_pitched(ytb200_brickset *set, int64_t index, const int64_t dims[3], const double left_edge[3],
                                const double right_edge[3], const void *const *dev_data, const uint8_t *dev_mask,
                                int64_t row_pitch);
/* Element strides of brick `index`'s data arrays (plane pitch, row pitch). */
int ytb200_brick_pitch(const ytb200_brickset *set, int64_t index, int64_t *plane_pitch, int64_t *row_pitch);

/* ---- vertex centring on the device (replaces the host passes that feed PartitionedGrid) -------
 * grid.get_vertex_centered_data(fields, no_ghost=True)  grid_patch.py:270-296 -> ghost_zone_interpolate,
 *                                                       interpolators.pyx:177-227 (rf=1, input_left=0.5)
 * log10 with values <= 0 -> NaN                         amr_kdtree.py:346-352
 * data[li:ri+1] slices of the kd leaf                   amr_kdtree.py:333-335,362-365
 * Cell data of a grid go to the device once (fp64, C-order (nx,ny,nz)) ... */
int ytb200_cells_upload(ytb200_ctx *ctx, const double *cells_host, const int64_t dims[3], void **cells_dev_out);
int ytb200_cells_free(ytb200_ctx *ctx, void *cells_dev);
/* ... and every kd leaf of that grid becomes brick `index`: vertices lo..hi (inclusive, hi - lo = the leaf's
 * cell dims) of the grid's vertex-centred field, in the set's storage type, log10'd where log_fields[f] != 0
 * (NULL = none).  cells_dev[f] are DEVICE pointers to the grid's cell-centred fields (fp64, or fp32 when
 * cells_are_f32); mask_host as in ytb200_brick_upload.  fp64 storage is bit-identical to the reference for
 * linear fields; log10 is CUDA's (<= 2 ulp from libm).  Needs >= 2 cells per grid axis (the reference yields NaN
 * below that).  `stats` (optional) returns the CUDA-event time of the centring kernels. */
int ytb200_brick_from_cells(ytb200_brickset *set, int64_t index, const int64_t grid_dims[3], const int64_t lo[3],
                            const int64_t hi[3], const double left_edge[3], const double right_edge[3],
                            const void *const *cells_dev, int cells_are_f32, const int32_t *log_fields,
                            const uint8_t *mask_host, ytb200_stats *stats);
/* The ghost-zone flavour (grid.get_vertex_centered_data(no_ghost=False), grid_patch.py:298-310): cells_dev[f] hold the
 * grid PADDED with one layer of ghost zones, shape grid_dims + 2 (the smoothed covering grid of retrieve_ghost_zones,
 * assembled by yt_b200.loaders.smoothed_covering_grid), and every vertex is 0.125 times the sum of its 8 cells in the
 * reference's order -- bit-identical in fp64.  Otherwise as ytb200_brick_from_cells. */
int ytb200_brick_from_ghost_cells(ytb200_brickset *set, int64_t index, const int64_t grid_dims[3], const int64_t lo[3],
                                  const int64_t hi[3], const double left_edge[3], const double right_edge[3],
                                  const void *const *cells_dev, int cells_are_f32, const int32_t *log_fields,
                                  const uint8_t *mask_host, ytb200_stats *stats);
/* Must be called after the last upload/attach and before the first cast (uploads the brick table). */
int ytb200_bricks_commit(ytb200_brickset *set);
int ytb200_bricks_destroy(ytb200_brickset *set);
/* Device pointers of brick `index` (data_out[n_fields], *mask_out may be NULL = all ones), so that a
 * brick uploaded once can be attached to other sets without a copy (bricks are cached across frames
 * like node.data in amr_kdtree.py:326-327). */
int ytb200_brick_device_ptrs(const ytb200_brickset *set, int64_t index, void **data_out, void **mask_out);
int64_t ytb200_bricks_device_bytes(const ytb200_brickset *set);

/* ---- transfer function: VolumeRenderSampler.__cinit__ + FIT_initialize_table
 * (image_samplers.pyx:405-447, field_interpolation_tables.pxd:44-61) ------------------------- */
int ytb200_tf_create(ytb200_ctx *ctx, int n_tables, const int32_t *nbins, const double *const *y,
                     const double *bounds /* n_tables x 2 */, const int32_t field_ids[6],
                     const int32_t weight_field_ids[6], const int32_t weight_table_ids[6],
                     const int32_t field_table_ids[6], int grey_opacity, ytb200_tf **out);
/* LightSourceRenderSampler.__cinit__ (image_samplers.pyx:496-549): switches the volume render of this table set to
 * the lit sampler -- gradient of field 0 by eval_gradient (fixed_interpolator.cpp:116-148), composite by
 * FIT_eval_transfer_with_light (field_interpolation_tables.pxd:115-143), no cell-mask test (image_samplers.pyx:551-589).
 * light_dir must already be normalised (the reference normalises it in place); NULL switches the light off. */
int ytb200_tf_set_light(ytb200_tf *tf, const double light_dir[3], const double light_rgba[4]);
int ytb200_tf_destroy(ytb200_tf *tf);

/* ---- frame: the image / z-buffer / ray set one ImageSampler instance owns ------------------- */
int ytb200_frame_create(ytb200_ctx *ctx, const ytb200_camera *cam, ytb200_frame **out);
int ytb200_frame_destroy(ytb200_frame *f);
/* H2D of the caller's (nx,ny,4) fp64 image (initial colours, image_samplers.pyx:183-184) and (nx,ny)
 * fp64 z-buffer (NULL = ones, image_samplers.pyx:78-80).  Asynchronous on the context stream. */
int ytb200_frame_upload(ytb200_frame *f, const double *image_host, const double *zbuffer_host);
/* Zero the device image and set the z-buffer to ones without touching the host. */
int ytb200_frame_clear(ytb200_frame *f);
/* Cast `n_order` bricks of `set`, in the given order (the order AMRKDTree.traverse yields them,
 * render_source.py:594-597), into the frame's device image.  Asynchronous.
 * == for b in order: sampler(brick_b)  (image_samplers.pyx:132-197, grid_traversal.pyx:26-140). */
int ytb200_frame_cast(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                      int64_t n_order, const ytb200_tf *tf, int n_samples, int precision, ytb200_stats *stats);
/* Same traversal, but only counts cells/samples (exact denominator of the Gsamples/s metric). */
int ytb200_frame_count(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                       int64_t n_order, int n_samples, ytb200_stats *stats);
/* D2H of the image into the caller's fp64 (nx,ny,4) array (sampler.aimage); synchronises. */
int ytb200_frame_download(ytb200_frame *f, double *image_host);
/* Device pointer of the fp64 (nx,ny,4) image, for zero-copy consumers (torch, NCCL). */
void *ytb200_frame_device_image(ytb200_frame *f);

/* ---- partial images for object-space multi-GPU rendering (replaces reduce_tree_images,
 * amr_kdtree.py:297-323 / amr_kdtools.py:6-50).  A partial is the per-channel affine map
 * C -> A + B*C of all samples a rank owns, stored as fp64 (nx,ny,8): A[4],B[4]. -- */
int ytb200_frame_cast_partial(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                              int64_t n_order, const ytb200_tf *tf, int n_samples, int precision,
                              double *partial_dev /* device (nx,ny,8) */, ytb200_stats *stats);
/* out = later o earlier on n_pixels pixels: A = A2 + B2*A1, B = B2*B1.  `out` may alias either input. */
int ytb200_partial_compose(ytb200_ctx *ctx, const double *earlier_dev, const double *later_dev,
                           double *out_dev, int64_t n_pixels);
/* image = A + B * image on n_pixels pixels (applies the composed map to the initial colours). */
int ytb200_partial_apply(ytb200_ctx *ctx, const double *partial_dev, double *image_dev, int64_t n_pixels);

/* Same cast with the partial stored as fp32 (nx,ny,8): half the bytes to exchange; FAST precision only. */
int ytb200_frame_cast_partial_f32(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                                  int64_t n_order, const ytb200_tf *tf, int n_samples, int precision,
                                  float *partial_dev /* device (nx,ny,8) */, ytb200_stats *stats);

/* ---- binary swap over NVLink peer memory: AMRKDTree.reduce_tree_images (amr_kdtree.py:297-323) -------
 * One process per GPU.  Every rank creates a swap object (one cudaMalloc'd slab holding the partial-image buffers of
 * every round, a final region and the flag words), passes its CUDA IPC handle to the others ONCE, and then per frame
 *   ytb200_swap_cast(swap, frame, ...);  ytb200_swap_composite(swap, frame, broadcast);
 * The cast writes this rank's partial straight into the slab, in the smallest exact format of the affine map
 * C -> A + B*C (ytb200_partial_format: 16 B per pixel for the ColorTransferFunction render, 8 B per summed field for
 * the projections, the general (A,B) pair otherwise), and its last CTA tells the round-0 partner.  Each swap round is
 * ONE kernel: it waits for the partner's flag, reads the partner's half straight out of its HBM through the peer
 * mapping, composes and signals.  The last round stores every rank's composed strip into the destination ranks'
 * slabs (rank 0, or every rank with `broadcast`), whose apply kernel writes A + B*C0 into the frame's image.  No
 * NCCL call, no host round trip: a frame is log2(world) + 2 launches on a destination rank.
 * ytb200_swap_composite_host instead lets every rank store the FINISHED pixels of its strip into one page-locked
 * image mapped into every rank's process (ytb200_host_register of the same shared memory): the image leaves the box
 * over all PCIe links at once.  Peers must be on one host, with peer access (<= 8 ranks, power of two);
 * yt_b200.parallel.make_compositor falls back to NCCL send/recv (BinarySwap) otherwise.
 * Host inputs: ytb200_frame_bind_host lets the partial casts read max_t, and the composite read its strip of the
 * initial colours, in place from page-locked arrays. */
/* Multi-GPU frames with host inputs: bind the caller's PAGE-LOCKED image (initial colours) and z-buffer (NULL = ones)
 * as inputs of the frame; the partial casts then read max_t, and the composite the initial colours of the rank's final
 * strip, in place over PCIe instead of from uploaded copies (a rank reads 1/P of the image instead of uploading all of
 * it).  image_host = NULL unbinds. */
int ytb200_frame_bind_host(ytb200_frame *f, const double *image_host, const double *zbuffer_host);
/* Early ray termination (ours; the reference has none).  With eps > 0, volume renders through a GREY-OPACITY
 * ColorTransferFunction (fp32 storage, FAST precision, one field) run the reversed cast: FIT_eval_transfer's composition
 * C <- e + a C (field_interpolation_tables.pxd:103-110) is evaluated from its far end, front to back, and a ray stops once
 * the transmittance of what it has composited is below eps -- the pixel then differs from the full composition by at most
 * eps times the brightest colour the skipped part could have reached.  Other tables and samplers ignore the setting (a
 * non-grey table absorbs each channel with its own emission, and a ray saturates only where all three channels do).
 * A timed cast reports the cells the rays actually sampled in ytb200_stats.n_cells / n_samples; the algorithmic count
 * stays ytb200_frame_count's.  eps = 0 (the default) disables. */
int ytb200_frame_set_early_termination(ytb200_frame *f, double eps);
/* Host-only probe (no device needed; ours): would ytb200_tf_create sample a ColorTransferFunction's four tables -- the
 * (r, g, b, a) arrays of FIT_initialize_table, field_interpolation_tables.pxd:27-50 -- through LINE-form rows (y = d*fs + c
 * in the table's row coordinate, two instructions fewer per sample), or keep (value, increment) rows?  The decision is an
 * accuracy guard computed against the fp64 table: every bin must keep 5e-6 of its own mean magnitude (+ 1e-9 of the channel
 * maximum) for the best integer origin.  Returns 1 / 0; *origin_out and *worst_ratio_out (error / allowance of the worst bin)
 * may be NULL. */
int ytb200_tf_line_form_probe(int nbins, const double *r, const double *g, const double *b, const double *a,
                              int *origin_out, double *worst_ratio_out);
typedef struct ytb200_swap ytb200_swap;
#define YTB200_PARTIAL_AB64 0   /* fp64 (A[4], B[4]): 64 B per pixel -- EXACT precision */
#define YTB200_PARTIAL_AB32 1   /* fp32 (A[4], B[4]): 32 B -- FAST volume render, any transfer function */
#define YTB200_PARTIAL_A32 2    /* fp32 A[4], B = 1 - A or 0 where flagged: 16 B -- FAST ColorTransferFunction render */
#define YTB200_PARTIAL_A32G 3   /* the same with grey opacity (flag in the sign of alpha) */
#define YTB200_PARTIAL_SUM1 4   /* fp64 sums, B == 1: 8 / 16 / 32 B -- the projection samplers */
#define YTB200_PARTIAL_SUM2 5
#define YTB200_PARTIAL_SUM4 6
int ytb200_partial_format(int kind, const ytb200_tf *tf, int n_fields, int precision);   /* < 0: error */
int ytb200_partial_bytes(int format);
int ytb200_swap_create(ytb200_ctx *ctx, int64_t n_pixels, int rank, int world, int format, ytb200_swap **out);
int ytb200_swap_destroy(ytb200_swap *swap);
int ytb200_swap_format(const ytb200_swap *swap);
int ytb200_swap_handle(ytb200_swap *swap, void *handle64_out);
int ytb200_swap_open(ytb200_swap *swap, int peer_rank, const void *slab_handle64);
/* The schedule of this rank (yt_b200.parallel.plan_binary_swap): per round the partner, the pixel region [p0,p1) both
 * hold, the part [k0,k1) this rank keeps, and whether this rank's content is composited first. */
int ytb200_swap_plan(ytb200_swap *swap, int n_rounds, const int32_t *partner, const int64_t *p0, const int64_t *p1,
                     const int64_t *k0, const int64_t *k1, const int32_t *mine_is_earlier);
int ytb200_swap_cast(ytb200_swap *swap, ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order,
                     int64_t n_order, const ytb200_tf *tf, int n_samples, int precision, ytb200_stats *stats);
int ytb200_swap_composite(ytb200_swap *swap, ytb200_frame *f, int broadcast);
int ytb200_swap_composite_host(ytb200_swap *swap, ytb200_frame *f, double *image_host_inout);
int ytb200_swap_error(ytb200_swap *swap, int *error_out);   /* synchronises; 1 = a peer did not arrive within 10 s */

/* ---- opaque sources: zlines / zpoints (yt/utilities/lib/misc_utilities.pyx:244-404) -------------------
 * Depth-tested painters of line segments (consecutive point pairs) and discs into the HOST (nx,ny,4) image and
 * (nx,ny) z-buffer that the cast then starts from / stops at (render_source.py:1062-1252, scene.py:581-625).
 * Order-dependent serial host code, as in the reference; statement-for-statement equal results.
 * `order` (zpoints) is np.argsort(zs) of the caller.  Returns 0 or -1 (bad argument). */
int ytb200_zlines(double *image, double *zbuffer, int64_t nx, int64_t ny, const int64_t *xs, const int64_t *ys,
                  const double *zs, int64_t n_points, const double *colors /* n_colors x 4 */, int64_t n_colors,
                  int points_per_color, int thick, int flip, int crop);
int ytb200_zpoints(double *image, double *zbuffer, int64_t nx, int64_t ny, const int64_t *xs, const int64_t *ys,
                   const double *zs, int64_t n_points, const int64_t *order, const double *colors, int64_t n_colors,
                   const int64_t *radii, int points_per_color, int flip);

/* ---- one-shot convenience: everything above for host buffers --------------------------------
 * upload image+zbuffer -> cast -> download.  This is the call the e2e benchmark times.
 * When `image_host_inout` (and `zbuffer_host`, if given) are page-locked (ytb200_host_alloc, cudaHostAlloc,
 * cudaHostRegister) the cast kernel reads and writes them in place over PCIe -- no staging copies, and the
 * frame's device image is left untouched; pageable arrays are staged through the frame's device buffers.
 * Either way the result is in `image_host_inout` when the call returns (image_samplers.pyx:111,192). */
int ytb200_render(ytb200_frame *f, int kind, const ytb200_brickset *set, const int64_t *order, int64_t n_order,
                  const ytb200_tf *tf, int n_samples, int precision, const double *zbuffer_host,
                  double *image_host_inout, ytb200_stats *stats);

/* ---- octree path (OctreeVolumeSource) ---------------------------------------------------------
 * Replaces yt/utilities/lib/_octree_raytracing.pyx:15-34 (_OctreeRayTracing: Octree<int> of
 * _octree_raytracing.hpp, built with insert_node per cell) and ImageSampler.cast_through_octree,
 * image_samplers.pyx:202-299: every pixel's ray (per-pixel vp_pos / vp_dir) walks the tree front to back
 * (cast_ray / proc_subtree, hpp:199-343) and the sampler's per-cell body runs on each terminal cell met, over
 * the cell's 14 doubles -- 8 vertex values, left edge, right edge (render_source.py:669-680).  The z-buffer is
 * not consulted (as in the reference). */
typedef struct ytb200_octree ytb200_octree;
/* _OctreeRayTracing.__init__(LE, RE, depth) */
int ytb200_octree_create(ytb200_ctx *ctx, int max_depth, const double left_edge[3], const double right_edge[3],
                         ytb200_octree **out);
int ytb200_octree_destroy(ytb200_octree *oct);
/* _OctreeRayTracing.add_nodes(ipos (n,3) int32, lvl (n) int32, key (n) int32) */
int ytb200_octree_add_nodes(ytb200_octree *oct, int64_t n, const int32_t *ipos, const int32_t *lvl, const int32_t *key);
int64_t ytb200_octree_node_count(const ytb200_octree *oct);
/* the PartitionedGrid payload of OctreeVolumeSource.render: HOST (n_cells, 14) fp64, uploaded once */
int ytb200_octree_set_cells(ytb200_octree *oct, int64_t n_cells, const double *cells_host);
/* sampler(pg, oct=octree) with volume_method == "Octree": casts into the frame's device image.  Asynchronous.
 * precision: YTB200_PREC_EXACT = everything fp64 in the reference's order; YTB200_PREC_FAST = the volume render samples in
 * fp32 (traversal and first-sample position stay fp64); the projection kinds are fp64 either way. */
int ytb200_frame_cast_octree(ytb200_frame *f, int kind, ytb200_octree *oct, const ytb200_tf *tf, int n_samples,
                             int precision, ytb200_stats *stats);

/* ---- AMR kd-tree: brick decomposition + view-order traversal, host side -----------------------
 * Replaces yt/utilities/lib/amr_kdtools.pyx (Node.add_grids/insert_grids/split_grids/
 * kdtree_get_choices 136-458,756-823; kd_traverse/viewpoint_traverse 102-111,612-622,708-755) and
 * the Tree wrapper of yt/utilities/amr_kdtree/amr_kdtree.py:63-118.  Node numbering, split choices
 * and leaf order are the reference's; one call returns the whole ordered leaf list. */
typedef struct ytb200_kdtree ytb200_kdtree;
const char *ytb200_kdtree_last_error(void);
/* Tree.__init__: trunk = Node(None, None, None, left, right, -1, 1); (rank, size) as comm_rank/comm_size */
int ytb200_kdtree_create(const double left[3], const double right[3], int rank, int size, ytb200_kdtree **out);
int ytb200_kdtree_destroy(ytb200_kdtree *t);
/* Tree.add_grids for one refinement level: gles/gres are (n,3) C-order, gids the grid ids */
int ytb200_kdtree_add_grids(ytb200_kdtree *t, int64_t n, const double *gles, const double *gres, const int64_t *gids);
int64_t ytb200_kdtree_node_count(const ytb200_kdtree *t);
/* Dump every node (arrays of node_count entries; any pointer may be NULL): heap node_id, grid id
 * (-1 none), child indices (-1 none), split (dim -1 / pos NaN for leaves), edges (n,3). */
int ytb200_kdtree_nodes(const ytb200_kdtree *t, int64_t *node_id, int64_t *grid, int32_t *left, int32_t *right,
                        int32_t *split_dim, double *split_pos, double *left_edge, double *right_edge);
/* Node.kd_traverse(viewpoint): indices (into the node dump) of the leaves that carry a grid, farthest
 * from `viewpoint` first; viewpoint NULL = depth-first left-to-right.  Returns the count (or <0);
 * `out` may be NULL to query the count. */
int64_t ytb200_kdtree_traverse(const ytb200_kdtree *t, const double *viewpoint, int32_t *out, int64_t cap);

/* Pinned host memory for the image / z-buffer (cudaHostAlloc). */
int ytb200_host_alloc(int64_t bytes, void **out);
int ytb200_host_free(void *p);
/* Page-lock and map memory the caller owns (e.g. POSIX shared memory every rank of a multi-GPU render has mapped):
 * afterwards it is accepted wherever a page-locked array is (ytb200_render, ytb200_frame_bind_host,
 * ytb200_swap_composite_host). */
int ytb200_host_register(void *p, int64_t bytes);
int ytb200_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif /* YTB200_H */
