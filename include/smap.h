/*
 * smap.h -- C ABI of the B200 (sm_100a) semantic-mapping hot path.
 *
 * The reference (AutonomousVehicleLaboratory/vision_semantic_segmentation) is pure
 * Python/numpy and has no FFI of its own; the boundary it offers is
 * the Python mapping API.  This header is the C ABI that sits directly under that
 * API: every entry point names the reference statements it replaces (paths relative
 * to the reference root).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - return value: 0 = SMAP_OK, negative = error; smap_last_error() gives the text
 *     (thread-local).  Nothing throws across this boundary.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers
 *     are host memory.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Calls only ENQUEUE work unless documented as synchronising.
 *   - one handle per GPU; a handle is not thread-safe.
 *   - the BEV grid is (map_height, map_width, num_classes) float64, C-contiguous,
 *     axis 0 <- world x, axis 1 <- world y (src/mapping_replay.py:80,181).
 */
#ifndef SMAP_B200_H_
#define SMAP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SMAP_API __attribute__((visibility("default")))
#else
#define SMAP_API
#endif

#define SMAP_ABI_VERSION 5
#define SMAP_MAX_CLASSES 31 /* C class bits + 1 intensity-boost bit in a 32-bit cell mask */
#define SMAP_MAX_CAMERAS 8

enum {
    SMAP_OK = 0,
    SMAP_ERR_INVALID = -1,  /* bad argument */
    SMAP_ERR_CUDA = -2,     /* a CUDA runtime call failed */
    SMAP_ERR_NOMEM = -3,    /* allocation failed */
    SMAP_ERR_STATE = -4,    /* call order: classes / camera not set */
    SMAP_ERR_NO_DEVICE = -5, /* no usable sm_100 device */
    SMAP_ERR_COMM = -6       /* NCCL: library not found, or a call failed */
};

/* point-cloud layouts accepted by the kernels */
enum {
    SMAP_PTS_F32X4 = 0,  /* N x {x, y, z, intensity} float32, 16-byte aligned (PointCloud2 / float4) */
    SMAP_PTS_F64_SOA = 1 /* (4, N) float64 rows x, y, z, intensity with row stride `ld` (the reference's pcd) */
};

/* what smap_frame.image_dev holds */
enum {
    SMAP_IMG_RGB = 0,      /* (H, W, 3) uint8 colour-coded label image: what the segmentation node publishes
                            * (src/vision_semantic_segmentation_node.py:113-121) and project_pcd indexes */
    SMAP_IMG_CLASS_IDS = 1 /* (h, w) uint8 class-id plane: the network's own output BEFORE the node's
                            * cv2.resize(..., INTER_NEAREST) and apply_color_map
                            * (src/vision_semantic_segmentation_node.py:109-113,
                            * src/network/deeplab_v3_plus/data/utils/mapillary_visualization.py:70-89).
                            * The fused kernel applies the same nearest-neighbour index map and the palette
                            * (smap_set_label_palette) on the fly: same grid, bit for bit, as painting and
                            * upscaling first; 1 byte per network pixel instead of 3 per camera pixel. */
};

typedef struct smap_handle smap_handle;

/* Replaces the constructor arithmetic of SemanticMapping.__init__ (src/mapping_replay.py:86-94). */
typedef struct smap_config {
    int32_t map_height;      /* int((B[0][1]-B[0][0]) / res) */
    int32_t map_width;       /* int((B[1][1]-B[1][0]) / res) */
    int32_t num_classes;     /* len(cfg.LABELS), 1..31 */
    int32_t use_intensity;   /* cfg.MAPPING.PCD.USE_INTENSITY */
    int32_t lane_index;      /* index of the class named "lane", -1 if none (src/mapping_replay.py:288) */
    int32_t device;          /* CUDA device ordinal */
    double boundary_x_min;   /* cfg.MAPPING.BOUNDARY[0][0] */
    double boundary_y_min;   /* cfg.MAPPING.BOUNDARY[1][0] */
    double resolution;       /* cfg.MAPPING.RESOLUTION */
    double origin_offset_x;  /* 1369.0496826171875  (src/mapping_replay.py:261) */
    double origin_offset_y;  /* 562.84814453125 */
    double range_max;        /* cfg.MAPPING.PCD.RANGE_MAX */
    void *map_dev;           /* caller-owned grid (MH*MW*C doubles) or NULL: the handle allocates it */
    int32_t map_is_zero;     /* caller-owned grid is all zeros right now (enables the atomic count update, below) */
    int32_t reserved;
} smap_config;

/* One frame of the batched fused path. */
typedef struct smap_frame {
    const void *points_dev;   /* cloud, layout below */
    int64_t n_points;
    int64_t ld;               /* row stride for SMAP_PTS_F64_SOA, ignored for F32X4 */
    int32_t layout;           /* SMAP_PTS_* */
    int32_t camera;           /* slot set with smap_set_camera */
    const uint8_t *image_dev; /* label image, see image_format */
    int32_t image_width;      /* W, H of the camera image the cloud is projected into (the frustum cull uses them); */
    int32_t image_height;     /* for SMAP_IMG_RGB also the shape of image_dev */
    int32_t has_transform;    /* 0: cloud already in the velodyne frame (pcd_frame_id == "velodyne") */
    int32_t image_format;     /* SMAP_IMG_* (0 = RGB, the reference's format) */
    double world_to_velodyne[16]; /* row-major 4x4 = inv(T_base_to_origin(pose) @ T_velodyne_to_baselink) */
    int32_t ids_width;        /* SMAP_IMG_CLASS_IDS: shape (ids_height, ids_width) of the class-id plane; 0 = W, H. */
    int32_t ids_height;       /* Pixel (u, v) reads ids[min(floor(v * fy), h - 1), min(floor(u * fx), w - 1)] with
                               * fx = 1.0 / ((double)W / w): cv2.resize's INTER_NEAREST index map, in double as OpenCV.
                               * The handle caches the map of the most recent (W, H, w, h); a new shape costs a host
                               * tabulation and, when the map needs its table form, a device synchronisation. */
} smap_frame;

/* counters of the most recent frame(s); see smap_get_stats */
typedef struct smap_stats {
    int64_t frames;          /* frames integrated since create / clear */
    int64_t points;          /* points read */
    int64_t touched_cells;   /* (cell, frame) pairs updated by the most recent ORDERED update launch (K for a single
                              * frame); -1 after a count update of a float4 cloud (tags, nothing is replayed) */
    int64_t kernel_launches; /* kernels launched by this handle */
    /* filled while smap_set_profiling(h, 1) is active: device time (CUDA events on the caller's stream, launches
     * serialised on it) of the streaming kernels and of the apply kernels, and the frames they covered */
    int64_t profiled_frames;
    double stream_kernel_ms;
    double apply_kernel_ms;
} smap_stats;

SMAP_API int smap_abi_version(void);
SMAP_API const char *smap_last_error(void);

/* Number of visible CUDA devices (<0 on error) and name / SM count / compute capability of one. */
SMAP_API int smap_device_count(void);
SMAP_API int smap_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor);

SMAP_API int smap_create(const smap_config *cfg, smap_handle **out);
SMAP_API int smap_destroy(smap_handle *h);

/* Camera projection matrix P = K [R t] (3x4 row-major, src/camera.py:28) for a slot. */
SMAP_API int smap_set_camera(smap_handle *h, int camera, const double P_host[12]);

/* cfg.LABEL_COLORS (C x 3 uint8) and the update matrix (C x C float64 row-major): np.eye(C) for the
 * count update or ConfusionMatrix.get_submatrix(LABELS, True, True) (src/mapping_replay.py:104-109).
 * Column i is added to a cell when class i is observed there (src/mapping_replay.py:281). */
SMAP_API int smap_set_classes(smap_handle *h, const uint8_t *colors_host, const double *cm_host);

/* Palette of the segmentation network: colour of class id i = rgb[3 i .. 3 i + 2], n_ids <= 256 (the "labels" of the
 * dataset config, src/network/deeplab_v3_plus/data/utils/mapillary_visualization.py:70-89).  Needed before a frame
 * with image_format == SMAP_IMG_CLASS_IDS is integrated.  Ids >= n_ids are painted black, as apply_color_map leaves
 * them (its canvas is np.zeros), so they match the classes whose colour has R == G == 0.  The handle folds palette and
 * cfg.LABEL_COLORS into a 256-entry id -> class-bit table that reproduces the reference's R,G-only colour compare
 * (src/mapping_replay.py:276), collisions included.  Synchronises (small upload). */
SMAP_API int smap_set_label_palette(smap_handle *h, const uint8_t *rgb_host, int n_ids);

/* ---- parity API: SemanticMapping.project_pcd (src/mapping_replay.py:214-246) -------------------
 * Transform, project, cull (range + frustum), stable compaction, label gather.
 * Outputs have room for n points with row stride out_ld (>= n):
 *   out_pcd   (4, out_ld) float64  masked_pcd = pcd[:, mask]       (original coordinates + intensity)
 *   out_label (3, out_ld) uint8    label = image[v, u].T
 *   out_uv    (2, out_ld) int32    image_idx = IXY[:, mask]        (may be NULL)
 *   out_keep  (n) uint8            the mask itself                 (may be NULL)
 * SYNCHRONISES the stream to return M in *m_host. */
SMAP_API int smap_project(smap_handle *h, const smap_frame *frame, double *out_pcd_dev, uint8_t *out_label_dev,
                 int32_t *out_uv_dev, uint8_t *out_keep_dev, int64_t out_ld, int64_t *m_host, void *stream);

/* ---- parity API: SemanticMapping.update_map (src/mapping_replay.py:248-301) --------------------
 * pcd (4, M) float64 row stride ld, label (3, M) uint8 row stride ldl; updates `map_dev` (NULL = the
 * handle's own grid) in place, de-duplicated per (cell, class) within the call, classes applied in
 * ascending order -> bit-exact with the reference, log-likelihood mode included. */
SMAP_API int smap_update(smap_handle *h, double *map_dev, const double *pcd_dev, int64_t ld, const uint8_t *label_dev,
                int64_t ldl, int64_t m, void *stream);

/* ---- fused path: project_pcd + update_map of ONE frame, nothing materialised -------------------
 * (src/mapping_replay.py:184-192 loop body).  Bit-exact with the reference for any update matrix (count or
 * log-likelihood): a fused kernel ORs the frame's class bits into a per-frame cell mask (the per-frame
 * (cell, class) de-duplication), a second kernel adds the matrix columns to the touched cells in ascending
 * class order and clears the mask.
 * Count update shortcut: when the matrix is exactly np.eye(C) AND the grid is known to hold integer-valued
 * counts (zero-initialised -- by the handle, by smap_clear, or declared with cfg.map_is_zero -- and since then
 * only updated by count updates of this handle), the fused kernel itself adds 1.0 per newly observed
 * (cell, class) and 2.0 per lane boost with float64 atomics: sums of small integers, exact in any order, so the
 * result is the same bits.  "Newly observed" is decided with per-(cell, class) frame tags (up to 7 classes; the
 * handle then allocates 4 x cells x (C + 1) uint32 on first use) or with the per-frame cell masks (more classes).
 * smap_upload / smap_notify_map_modified switch back to the ordered update.
 * float4 clouds take the fast kernel (float32 decisions with rigorous error bounds, float64 only for the few
 * points they cannot decide); label images of 2^28 pixels or more are rejected there.
 * Class-id planes (SMAP_IMG_CLASS_IDS) are accepted by the smap_integrate* calls for float4 clouds; smap_project
 * returns RGB labels and therefore takes RGB images only. */
SMAP_API int smap_integrate(smap_handle *h, const smap_frame *frame, void *stream);

/* Same rule for n_frames frames IN ORDER.  Frames are queued together in chunks: one fused kernel per frame, four
 * frames' kernels independent of each other (kernel i follows kernel i - 4), the whole chunk handed to `stream` as ONE
 * CUDA graph launch -- instantiated once per (update mode, label format, frames) and re-parameterised per chunk; with
 * SMAP_FUSE_GRAPH=0 in the environment, or for chunks of fewer than four frames, as per-frame launches on four internal
 * streams forked from and joined back into `stream`.  Ordered update and count update through the cell masks: chunks
 * of up to 16 frames, each frame scatters into its own mask slot and the apply kernel replays the slots in frame
 * order per cell (the apply of one chunk runs on an internal stream beside the scatter kernels of the next chunk of the
 * same call; `stream` joins it before the call returns).  Tagged count update (identity matrix, C + 1 <= 8, grid of
 * counts): no per-frame state, chunks of up to 64 frames.  Either way the result is bit-identical to n_frames calls of
 * smap_integrate.  Frames of one chunk must share a point layout. */
SMAP_API int smap_integrate_batch(smap_handle *h, const smap_frame *frames_host, int n_frames, void *stream);

/* Same as smap_integrate with HOST buffers: frame->points_dev / image_dev hold HOST pointers here.  Points (layout as
 * in frame) and image are copied through the handle's two-stage device ring on an INTERNAL copy stream, and the
 * kernels -- queued on `stream` -- wait for their frame's copies only: frame i + 1 crosses PCIe while frame i is
 * integrated, nothing synchronises the host.  Pass pinned memory (pageable memory makes cudaMemcpyAsync block) and keep
 * the buffers unchanged until `stream` has passed this call (synchronising `stream` implies the copies are done). */
SMAP_API int smap_integrate_host(smap_handle *h, const smap_frame *frame_with_host_ptrs, void *stream);

/* The reference's cloud layout, (4, N) float64 rows with stride ld (SMAP_PTS_F64_SOA), converted on the device to the
 * float4 layout of the fast kernel.  *flag_dev (int32, zeroed by the CALLER) is OR-ed with 1 when some coordinate or
 * intensity is not float32-representable -- the caller then keeps the float64 layout for that cloud, so results never
 * change.  PointCloud2 fields are FLOAT32 (src/mapping.py:178-180 copies them into a float64 buffer): recorded clouds
 * normally convert losslessly. */
SMAP_API int smap_cloud_to_f32x4(const double *soa_dev, int64_t ld, int64_t n_points, void *out_f32x4_dev,
                                 int32_t *flag_dev, int device, void *stream);

/* ---- rendering: free functions of src/renderer.py, any grid ------------------------------------ */
/* apply_filter: cv2.filter2D 3x3 box, BORDER_REFLECT_101 (src/renderer.py:175-189). src != dst. */
SMAP_API int smap_apply_filter(const double *src_dev, int mh, int mw, int c, double *dst_dev, int device, void *stream);
/* render_bev_map: argmax colour, zero-sum cells black (src/renderer.py:32-59). */
SMAP_API int smap_render(const double *map_dev, int mh, int mw, int c, const uint8_t *colors_host, uint8_t *rgb_dev,
                int device, void *stream);
/* apply_filter + render_bev_map in one pass (src/mapping_replay.py:198-200); filtered_dev may be NULL. */
SMAP_API int smap_filter_render(const double *map_dev, int mh, int mw, int c, const uint8_t *colors_host,
                       uint8_t *rgb_dev, double *filtered_dev, int device, void *stream);
/* render_bev_map_with_thresholds (src/renderer.py:131-172): priority (C ints, low -> high),
 * thresholds (C doubles, indexed by paint step as in the reference). */
SMAP_API int smap_render_thresholds(const double *map_dev, int mh, int mw, int c, const uint8_t *colors_host,
                           const int32_t *priority_host, const double *thresholds_host, uint8_t *rgb_dev,
                           int device, void *stream);

/* ---- evaluation: the step after rendering ------------------------------------------------------
 * convert_labels (test/test_semantic_mapping.py:6-18) of the rendered map fused with the sums of Test.iou
 * (:127-161) against the ground-truth label map slice truth[shift_rows : shift_rows + mh, shift_cols : shift_cols + mw]
 * (Test.test_single_map, :117-125).  truth: (truth_rows, truth_cols) uint8 labels 0 unknown, 1 road, 2 crosswalk,
 * 3 lane; mask: optional (mask_rows, mask_cols) uint8 validity mask applied to the map as convert_labels does
 * (NULL: none).  counts_dev receives 12 int64 (zeroed by the call):
 *   [0..2] intersection, [3..5] ground-truth pixels, [6..8] map pixels of classes 1, 2, 3;
 *   [9] known ground truth, [10] known and mapped, [11] map == ground truth where known.
 * Integer sums, exact; IoU = [k] / ([3+k] + [6+k] - [k]), missing rate = 1 - [10] / [9], accuracy = [11] / [9]. */
SMAP_API int smap_eval_counts(const uint8_t *rgb_dev, int mh, int mw, const uint8_t *truth_dev, int truth_rows,
                     int truth_cols, int shift_rows, int shift_cols, const uint8_t *mask_dev, int mask_rows,
                     int mask_cols, int64_t *counts_dev, int device, void *stream);

/* ---- multi-GPU: frames sharded over the ranks, grids summed ------------------------
 * update_map only ever ADDS frame-determined constants to the grid (src/mapping_replay.py:281,294), so the ranks of
 * a job integrate disjoint blocks of frames into their own full-size grids and sum them once.  One handle per GPU /
 * process; NCCL is resolved at run time (the libnccl.so.2 already loaded in the process, else the loader's; override
 * with the environment variable SMAP_NCCL_LIB), so single-GPU users never need it.
 *
 * The exchange moves only the union window of the cells any rank touched since its last smap_clear (tracked on the
 * device by the update kernels), and a grid of counts (count update: small non-negative integers) travels packed --
 * two uint16 per word while the global sum provably stays below 2^16 (3 per integrated frame), else uint32 -- which
 * is EXACT and 4x / 2x fewer bytes than float64; log-likelihood grids travel as float64 (summation order across
 * ranks differs from the sequential reference: <= 1e-5 relative by north_star, ~1e-15 in practice). */
#define SMAP_COMM_ID_BYTES 128
typedef struct smap_comm_info {
    int32_t n_ranks, rank;
    int32_t window[4];   /* last exchange: rows window[0]..window[1], columns window[2]..window[3] (empty: [1] < [0]) */
    int32_t pack;        /* last exchange: 0 = uint16 pairs, 1 = uint32, 2 = float64 */
    int32_t reserved;
    int64_t bytes;       /* last exchange: payload bytes this rank handed to NCCL */
    int64_t grid_bytes;  /* size of the whole float64 grid, for comparison */
    int64_t exchanges;   /* exchanges so far */
    /* streaming exchange, last data phase (CUDA events on the internal stream; smap_comm_get_info waits for them): */
    double pack_ms, reduce_ms, add_ms;
    double host_wait_ms; /* host time the last smap_exchange_async spent waiting for the ranks' agreement */
} smap_comm_info;
/* rank 0: a fresh NCCL unique id, to be distributed to all ranks by the caller (MPI, torch.distributed, a file...). */
SMAP_API int smap_comm_unique_id(uint8_t id_out[SMAP_COMM_ID_BYTES]);
/* Collective over all ranks: creates the handle's communicator (ncclCommInitRank on the handle's device). */
SMAP_API int smap_comm_init(smap_handle *h, int n_ranks, int rank, const uint8_t id[SMAP_COMM_ID_BYTES]);
/* Or: use an ncclComm_t the caller owns (passed as void*; not destroyed by the handle). */
SMAP_API int smap_comm_attach(smap_handle *h, void *nccl_comm);
SMAP_API int smap_comm_destroy(smap_handle *h);
/* Collective: every rank's grid becomes the sum of all ranks' grids.  Enqueues on `stream` but SYNCHRONISES it once
 * (the ranks first agree on the window with an 8-int all-reduce whose result sizes the exchange). */
SMAP_API int smap_allreduce(smap_handle *h, void *stream);
/* Collective, for maps too large to filter / render on every rank: the summed grid is scattered by rows.  With
 * per = ceil(MH / n_ranks), rank r receives rows [r0, r1) = [r per, min((r+1) per, MH)) plus one halo row from each
 * neighbouring tile (top / bottom = 0 or 1), written to tile_dev as (top + r1 - r0 + bottom, MW, C) float64 (rows in
 * grid order; tile_rows_cap >= per + 2 is enough); the handle's own grid is left as it was.  The 3x3 filter of the
 * tile then sees real neighbours across the seams and BORDER_REFLECT_101 only at true map edges. */
SMAP_API int smap_reduce_scatter_rows(smap_handle *h, double *tile_dev, int64_t tile_rows_cap, int32_t *r0, int32_t *r1,
                                      int32_t *top, int32_t *bottom, void *stream);
/* Streaming exchange -- the same sum, but overlapped with the integration of the next frames (a replay that is
 * summed every few hundred microseconds cannot afford to stop for the collective).  smap_comm_streaming(h, 1) makes the
 * update kernels accumulate into one of two internal buffers of LOCAL increments instead of the grid.
 * smap_exchange_async (collective, same call sequence on every rank) closes the current buffer, starts the ranks'
 * agreement on its window behind the work queued on `stream` so far, and queues the DATA phase of the PREVIOUS call's
 * buffer -- pack + zero, NCCL all-reduce, add to the grid -- on an internal high-priority stream: it runs while the
 * frames integrated after the call fill the other buffer.  The host blocks only until that previous agreement (8
 * ints) has arrived; `stream` never waits for a collective.  smap_exchange_flush queues the last data phase and makes
 * `stream` wait for it: afterwards the grid of every rank holds the sum of everything handed to
 * smap_exchange_async (frames integrated after the last smap_exchange_async are still local).  The grid is written
 * by the exchanges only; smap_map_ptr / download / render see exchanged increments only.  Counts stay exact
 * (integer sums); log-likelihood grids are summed in exchange order (<= 1e-5 relative by north_star). */
SMAP_API int smap_comm_streaming(smap_handle *h, int on, void *stream);
SMAP_API int smap_exchange_async(smap_handle *h, void *stream);
SMAP_API int smap_exchange_flush(smap_handle *h, void *stream);
SMAP_API int smap_comm_get_info(smap_handle *h, smap_comm_info *out);

/* ---- the planar update (cfg.MAPPING.DEPTH_METHOD other than points_map / points_raw) ------------------------------
 * src/mapping.py:446-488 warps the label image onto the map plane and then compares the warped uint8 image with the
 * label NAMES (:474): never equal, so no cell is ever incremented; what does run is map_local[map_local < 0] = 0
 * (:481).  That is the whole observable behaviour, and this is it (any float64 grid; -0.0 and NaN are left alone, as
 * numpy's mask leaves them). */
SMAP_API int smap_clamp_negative(double *map_dev, int64_t n_elements, int device, void *stream);

/* The warp itself: cv2.warpPerspective(image, h, (dst_w, dst_h)) as generate_homography calls it
 * (src/homography.py:53-55: default flags -- INTER_LINEAR, BORDER_CONSTANT 0) for an 8-bit image of 1..4 interleaved
 * channels, bit for bit (OpenCV's fixed-point arithmetic, restated in oracle/warp_port.py).  h_host: the 3 x 3 homography
 * source -> destination, row-major, as cv2.findHomography returns it.  src_dev and dst_dev must not overlap. */
SMAP_API int smap_warp_perspective(const uint8_t *src_dev, int src_h, int src_w, int channels, const double h_host[9],
                                   uint8_t *dst_dev, int dst_h, int dst_w, int device, void *stream);

/* ---- generate_convex_hull (src/semantic_convex_hull.py:17-91): the per-pixel part -----------------------------------
 * smap_hull_components: mask = (img == index), cv2.erode(mask, ones(3, 3)), 8-connected components of what is left.
 * labels_dev[p] = pixel index of the raster-first pixel of p's component (-1 outside the mask); areas_dev[root] = pixels
 * of the component (0 elsewhere).  img_dev: (h, w) uint8, as cv2.erode requires of the reference's input; scratch_dev:
 * h * w bytes.  h * w < 2^31.
 * smap_hull_row_extremes: per image row the smallest / largest column of component `root`, with the component's first
 * pixel left out as the reference leaves it out (:70); rowmin_dev[y] = INT_MAX, rowmax_dev[y] = -1 for rows it misses.
 * The convex hull of those <= 2 h points is the hull cv2.convexHull returns for the whole component (the host wrapper
 * semantic_convex_hull.py finishes there). */
SMAP_API int smap_hull_components(const uint8_t *img_dev, int h, int w, int index, uint8_t *scratch_dev, int32_t *labels_dev,
                                  int32_t *areas_dev, int device, void *stream);
SMAP_API int smap_hull_row_extremes(const int32_t *labels_dev, int h, int w, int root, int32_t *rowmin_dev,
                                    int32_t *rowmax_dev, int device, void *stream);

/* Dev builds with -DSMAP_DEBUG_BOUNDS (tools/bounds_check.sh): out[0] = index violations the kernels counted since the
 * library was loaded (stack pushes, label loads, mask / tag / grid updates, staged tiles), out[1] = code of the first
 * one.  Synchronises the device.  Returns SMAP_ERR_STATE in a normal build (nothing is checked there). */
SMAP_API int smap_debug_bounds(unsigned long long out[2]);

/* ---- grid access ------------------------------------------------------------------------------- */
SMAP_API int smap_map_ptr(smap_handle *h, double **map_dev, int64_t *n_elements);
SMAP_API int smap_clear(smap_handle *h, void *stream);            /* self.map = np.zeros(...)  src/mapping_replay.py:181 */
/* The caller wrote into the grid through its own pointer (map_dev / smap_map_ptr): the grid may no longer hold
 * integer-valued counts, use the ordered update from now on. */
SMAP_API int smap_notify_map_modified(smap_handle *h);
SMAP_API int smap_download(smap_handle *h, double *map_host);     /* synchronises */
SMAP_API int smap_upload(smap_handle *h, const double *map_host); /* synchronises */
SMAP_API int smap_get_stats(smap_handle *h, smap_stats *out);     /* synchronises the handle's last stream */
/* on != 0: time the kernels of every smap_integrate* call with CUDA events (the per-frame launches are then
 * issued on the caller's stream only, not over the internal streams); read the totals with smap_get_stats. */
SMAP_API int smap_set_profiling(smap_handle *h, int on);
/* Test hook: sets the counter the count update draws its per-frame tags from (uint32, strictly increasing; the tag
 * planes are re-zeroed and the counter restarts when it would overflow).  Lets a test reach that path without
 * integrating 2^32 frames.  Only moves the counter forward. */
SMAP_API int smap_debug_set_frame_tag(smap_handle *h, uint32_t value);
/* Test hook, host only (no device needed): the per-frame constants of the float32 decision path of the fused kernel
 * (csrc/smap_fuse.cuh, struct Fast32, in declaration order, uint32 fields as exact doubles; 69 values) for a
 * configuration, a frame description (pointers ignored) and a camera matrix.  tests/test_fast32_bounds.py emulates
 * the kernel's float32 arithmetic with them on the CPU and checks every certified decision against the oracle. */
SMAP_API int smap_debug_fast32(const smap_config *cfg, const smap_frame *frame, const double P_host[12], double *out69);

/* Test hook, host only (no device needed): the nearest-neighbour index map of one image axis (dst entries into
 * tab_out) as the library tabulates it for SMAP_IMG_CLASS_IDS frames, and the multiply-shift it hands to the kernel
 * when one reproduces the whole table, (x * *mul) >> *shift; *shift == 0: none, the kernel reads the table. */
SMAP_API int smap_debug_nearest_map(int dst, int src, uint16_t *tab_out, uint32_t *mul, uint32_t *shift);

#ifdef __cplusplus
}
#endif
#endif /* SMAP_B200_H_ */
