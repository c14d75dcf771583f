// C ABI (include/smap.h) of the B200 semantic-mapping path: host-side launch logic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/smap.h"
#include "smap_kernels.cuh"
#include "smap_fuse.cuh"
#include "smap_comm.cuh"
#include "smap_warp.cuh"
#include "smap_hull.cuh"

#ifndef SMAP_AUX_STREAMS
#define SMAP_AUX_STREAMS 4      // internal streams the per-frame k_fuse launches of a batch alternate over
#endif
#ifndef SMAP_TAG_MAX_PLANES
#ifndef SMAP_APPLY_OVERLAP
#define SMAP_APPLY_OVERLAP 1   // 0: dev switch, k_apply / k_clear_masks on the caller's stream, nothing beside them
#endif
#ifndef SMAP_RENDER_BULK
#define SMAP_RENDER_BULK 1   // 0: dev switch, every grid through the register-staged render kernel
#endif
#define SMAP_TAG_MAX_PLANES 8   // count update: per-(cell, class) tags up to this many tags per cell (C + 1), masks beyond
#endif
#ifndef SMAP_FUSE_GRID_DIV
#define SMAP_FUSE_GRID_DIV 2    // > 1: inside a batch a frame's launch fills only 1/DIV of the resident block slots, so
#endif                          // that the launches of DIV frames (on different internal streams) run side by side

using namespace smap;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) return fail(SMAP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace

struct smap_handle {
    smap_config cfg;
    GridParams gp;
    double* map = nullptr;
    bool own_map = false;
    int64_t cells = 0;
    // per-frame cell masks: n_slots slots of `cells` words, all zero between launches
    uint32_t* mask = nullptr;
    int n_slots = 0;
    // count update (k_fuse MODE 1): one plane of (cells, C + 1) uint32 tags per launching stream; a frame's tag is
    // larger than every tag written before it in its plane
    uint32_t* tags = nullptr;
    int n_tag_planes = 0;
    uint32_t frame_tag = 0;
    bool fuse_attr_set = false;
    bool last_update_counted = true;   // the most recent update went through k_apply (which counts touched cells)
    int64_t slot_words = 0;   // cells rounded up to a multiple of 4: k_apply reads the slots with 16-byte loads
    // double-buffered per-slot bounding boxes + touched counters (k_stream writes [parity], k_apply resets [parity^1])
    FrameBox* boxes = nullptr;                // [2][kMaxBatch]
    unsigned long long* touched = nullptr;    // [2]
    int parity = 0;
    // Two sets of mask slots (set = parity): while k_apply / k_clear_masks of chunk k runs on apply_stream over set p,
    // the k_fuse launches of chunk k + 1 scatter into set p ^ 1 (one smap_integrate_batch call of several chunks).
    cudaStream_t apply_stream = nullptr;
    cudaEvent_t ev_fused = nullptr;           // the chunk's scatter launches have joined the caller's stream
    cudaEvent_t ev_applied[2] = {};           // the apply / clear over set p has finished
    bool applied_pending[2] = {false, false};
    bool applied_writes_grid[2] = {false, false};   // the pending kernel is a k_apply (k_clear_masks only touches masks)
    int sm_count = 148;
    bool identity_cm = false;   // update matrix is exactly np.eye(C): the count update
    bool integer_grid = false;  // the grid is known to hold integer-valued counts (zeroed by us, then only count updates)
    // k_fuse launches of one batch alternate over a few internal streams (fork / join with events around the
    // batch): the frames are independent, so the ramp-up and tail of one launch overlap the next one's body
    static constexpr int kAux = SMAP_AUX_STREAMS;
    cudaStream_t aux[kAux > 0 ? kAux : 1] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_join[kAux > 0 ? kAux : 1] = {};
    FuseLaunch fuse_one;                  // parameter block of the next k_fuse launch
    // The k_fuse launches of a chunk as ONE CUDA graph launch: the same kernels in the same four-lane order (node i
    // depends on node i - lanes), instantiated once per (mode, label format, frames, lanes) and re-parameterised per
    // chunk (cudaGraphExecKernelNodeSetParams), so the host pays one graph launch per chunk instead of one kernel
    // launch + events per frame (6.6 us per frame on one GPU, 13 - 19 us with eight processes on one host).
    struct FuseGraph {
        int mode = -1, fmt = 0, n = 0, lanes = 0;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::vector<cudaGraphNode_t> nodes;
        uint64_t last_use = 0;
    };
    std::vector<FuseGraph> fuse_graphs;
    uint64_t graph_clock = 0;
    bool use_graph = true;                // SMAP_FUSE_GRAPH=0 in the environment: per-frame launches on the internal streams
    std::vector<FuseLaunch> graph_params;
    std::vector<int64_t> graph_gx;
    // union window of every cell touched since the last clear (device; maintained by k_fuse / k_apply / k_clear_masks):
    // what a multi-GPU exchange has to move.  ubox_full: the grid was written from outside, assume all of it.
    FrameBox* ubox = nullptr;
    bool ubox_full = false;
    int64_t value_bound = 0;              // no grid element exceeds this while integer_grid holds (3 per frame)
    // Where the update kernels accumulate: the grid itself, or -- streaming exchange (smap_comm_streaming) -- the
    // current one of two buffers of LOCAL increments that smap_exchange_async hands to the other ranks while the next
    // frames are integrated into the other buffer; abox = the union window of what `acc` holds.
    double* acc = nullptr;
    FrameBox* abox = nullptr;
    bool streaming = false;
    double* delta[2] = {};
    FrameBox* dbox[2] = {};
    bool d_integer[2] = {true, true};     // the buffer holds integer-valued counts
    int64_t d_bound[2] = {0, 0};          // no element of the buffer exceeds this
    int cur = 0;
    cudaStream_t comm_stream = nullptr;   // agreement, pack, NCCL, unpack of the streaming exchange
    cudaEvent_t ev_chunk = nullptr;       // the caller's stream has finished the chunk being exchanged
    cudaEvent_t ev_agreed[2] = {};        // the ranks' agreement on buffer b has reached the host
    cudaEvent_t ev_packed[2] = {};        // buffer b has been packed and zeroed: free for new increments
    bool packed_recorded[2] = {false, false};
    cudaEvent_t ev_exchanged = nullptr;   // the last queued exchange has been added to the grid
    bool pend_active = false;             // an agreement is in flight (its data phase is queued by the next call)
    int pend_buf = 0;
    int* xch_dev2[2] = {};                // agreement words of the streaming exchange, per buffer
    cudaEvent_t ev_phase[4] = {};         // timing of the last data phase: start, packed, reduced, added
    bool phase_recorded = false;
    double host_wait_ms = 0.0;            // host time the last smap_exchange_async spent waiting for the agreement
    int* xch_host2[2] = {};
    // multi-GPU exchange (smap_comm_* / smap_allreduce / smap_reduce_scatter_rows)
    ncclComm_t comm = nullptr;
    bool own_comm = false;
    int n_ranks = 1, rank = 0;
    void* xbuf = nullptr;                 // packed window
    size_t xbuf_cap = 0;
    void* xbuf2 = nullptr;                // reduce-scatter: the received tile, all-gather of the halo rows
    size_t xbuf2_cap = 0;
    int* xch_dev = nullptr;               // kCommWords ints the ranks agree on before an exchange
    int* xch_host = nullptr;              // pinned
    smap_comm_info comm_last = {};
    // class tables
    double* cm_dev = nullptr;
    uint8_t colors[SMAP_MAX_CLASSES * 3];
    bool classes_set = false;
    double P[SMAP_MAX_CAMERAS][12];
    bool cam_set[SMAP_MAX_CAMERAS] = {};
    // class-id planes (SMAP_IMG_CLASS_IDS): the network's palette, the id -> class-bit table folded from it and
    // cfg.LABEL_COLORS, and the most recent nearest-neighbour index map (see FuseFrame in smap_fuse.cuh)
    uint8_t palette[256 * 3] = {};
    bool palette_set = false;
    uint32_t* id_lut_dev = nullptr;
    struct NearestMap {
        int W = 0, H = 0, w = 0, h = 0;   // key
        uint32_t mx = 0, my = 0, sx = 0, sy = 0;
        bool use_tab = false;
        uint16_t* tab_dev = nullptr;      // W + H entries, only when no multiply-shift reproduces the map
        size_t tab_cap = 0;
    } nn;
    // project_pcd scratch
    uint8_t* keep = nullptr;
    int32_t* iu = nullptr;
    int32_t* iv = nullptr;
    uint32_t* blk_count = nullptr;
    int64_t* blk_offset = nullptr;
    int64_t* total_dev = nullptr;
    int64_t scratch_cap = 0;
    // host-buffer staging ring (smap_integrate_host)
    static constexpr int kStages = 2;
    void* stage_pts[kStages] = {};
    uint8_t* stage_img[kStages] = {};
    size_t stage_pts_cap[kStages] = {};
    size_t stage_img_cap[kStages] = {};
    cudaEvent_t stage_done[kStages] = {};
    cudaEvent_t stage_copied[kStages] = {};
    cudaStream_t copy_stream = nullptr;
    int stage_next = 0;
    // profiling (smap_set_profiling): three events per smap_integrate_batch chunk, harvested lazily
    bool profiling = false;
    struct ProfRec { cudaEvent_t e[3]; int frames; };
    ProfRec prof_pending[64];
    int n_prof_pending = 0;
    // bookkeeping
    smap_stats stats = {};
    cudaStream_t last_stream = nullptr;
};

namespace {

// properties of the accumulation target (the grid, or the current buffer of the streaming exchange)
inline bool& acc_integer(smap_handle* h) { return h->streaming ? h->d_integer[h->cur] : h->integer_grid; }
inline int64_t& acc_bound(smap_handle* h) { return h->streaming ? h->d_bound[h->cur] : h->value_bound; }

int fill_frame_params(const smap_handle* h, const smap_frame* f, FrameParams& fp) {
    if (!f) return fail(SMAP_ERR_INVALID, "frame is NULL");
    if (f->camera < 0 || f->camera >= SMAP_MAX_CAMERAS || !h->cam_set[f->camera])
        return fail(SMAP_ERR_STATE, "camera slot not set (smap_set_camera)");
    if (f->n_points < 0) return fail(SMAP_ERR_INVALID, "n_points < 0");
    if (f->layout != SMAP_PTS_F32X4 && f->layout != SMAP_PTS_F64_SOA) return fail(SMAP_ERR_INVALID, "unknown point layout");
    if (f->layout == SMAP_PTS_F64_SOA && f->ld < f->n_points) return fail(SMAP_ERR_INVALID, "ld < n_points");
    if (f->image_width <= 0 || f->image_height <= 0) return fail(SMAP_ERR_INVALID, "empty label image");
    if (f->image_format != SMAP_IMG_RGB && f->image_format != SMAP_IMG_CLASS_IDS) return fail(SMAP_ERR_INVALID, "unknown image format");
    if (f->image_format == SMAP_IMG_CLASS_IDS) {
        if (f->layout != SMAP_PTS_F32X4) return fail(SMAP_ERR_INVALID, "class-id planes need the float4 cloud layout");
        if (f->ids_width < 0 || f->ids_height < 0 || f->ids_width > 65535 || f->ids_height > 65535)
            return fail(SMAP_ERR_INVALID, "bad class-id plane shape");
    }
    if (f->image_width > 65535 || f->image_height > 32767) return fail(SMAP_ERR_INVALID, "label image larger than 65535 x 32767");
    if (f->layout == SMAP_PTS_F64_SOA && f->n_points >= ((int64_t)1 << 32)) return fail(SMAP_ERR_INVALID, "more than 2^32 points in a float64 cloud");
    if (f->n_points > 0 && (!f->points_dev || !f->image_dev)) return fail(SMAP_ERR_INVALID, "NULL points / image");
    if (f->layout == SMAP_PTS_F32X4 && (reinterpret_cast<uintptr_t>(f->points_dev) & 15u))
        return fail(SMAP_ERR_INVALID, "float4 cloud must be 16-byte aligned");
    memcpy(fp.T, f->world_to_velodyne, sizeof fp.T);
    memcpy(fp.P, h->P[f->camera], sizeof fp.P);
    fp.range_max = h->cfg.range_max;
    {   // float32 pre-cull matrix: row 0 = velodyne-x row of T, rows 1..3 = P * T (double product, rounded once)
        double Tm[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if (f->has_transform) memcpy(Tm, f->world_to_velodyne, sizeof Tm);
        for (int j = 0; j < 4; ++j) fp.Mf[j] = (float)Tm[j];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 4; ++j) {
                double acc = 0.0;
                for (int k = 0; k < 4; ++k) acc += fp.P[4 * r + k] * Tm[4 * k + j];
                fp.Mf[4 * (r + 1) + j] = (float)acc;
            }
        for (int r = 0; r < 4; ++r) {
            float a = 0.f;
            for (int j = 0; j < 3; ++j) a = fmaxf(a, fabsf(fp.Mf[4 * r + j]));
            // next float up so that the bound survives the rounding of these two products
            fp.Ea[r] = nextafterf(kCullSlack * a, INFINITY);
            fp.Eb[r] = nextafterf(kCullSlack * fabsf(fp.Mf[4 * r + 3]), INFINITY);
        }
        fp.range_hi = (float)fp.range_max * (1.0f + kCullSlack);
        fp.img_wf = (float)f->image_width;
        fp.img_hf = (float)f->image_height;
        // certified fast projection: M = P * T and the error bounds of smap_device.cuh (fast_project)
        const double u64 = 64.0 * 1.1102230246251565e-16;  // 64 * 2^-53
        double e[3];
        for (int r = 0; r < 3; ++r) {
            double axyz = 0.0, aw = 0.0;
            for (int j = 0; j < 4; ++j) {
                double acc = 0.0, aabs = 0.0;
                for (int k = 0; k < 4; ++k) {
                    acc += fp.P[4 * r + k] * Tm[4 * k + j];
                    aabs += fabs(fp.P[4 * r + k]) * fabs(Tm[4 * k + j]);
                }
                fp.M[4 * r + j] = acc;
                if (j < 3) axyz = fmax(axyz, aabs); else aw = aabs;
            }
            e[r] = u64 * (axyz * 3.0 * kCoordBound + aw);
        }
        const double umax = (double)(f->image_width > f->image_height ? f->image_width : f->image_height) + 2.0;
        fp.e3x4 = 4.0 * e[2];
        fp.cgu = (4.0 / 3.0) * (e[0] + umax * e[2]) * (1.0 + 1e-9);
        fp.cgv = (4.0 / 3.0) * (e[1] + umax * e[2]) * (1.0 + 1e-9);
        fp.c0 = umax * 5.6843418860808015e-14;  // 2^-44
        fp.img_wd = (double)f->image_width;
        fp.img_hd = (double)f->image_height;
        fp.img_wd1 = fp.img_wd + 1.0;
        fp.img_hd1 = fp.img_hd + 1.0;
        // non-finite bounds (absurd matrices) switch the fast path off: q2 > inf never holds
        if (!(fp.cgu == fp.cgu) || !(fp.cgv == fp.cgv) || !(fp.e3x4 == fp.e3x4)) fp.e3x4 = INFINITY;
    }
    fp.has_T = f->has_transform ? 1 : 0;
    fp.img_w = f->image_width;
    fp.img_h = f->image_height;
    fp.pad = 0;
    return SMAP_OK;
}

// float next above / below (the thresholds of the float32 path are rounded away from the certified side)
inline float f_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}
inline float f_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
}

// Constants of the float32 path of k_fuse (error analysis: header of smap_fuse.cuh).  Everything is composed in
// double here; whenever a precondition of the analysis cannot be met (absurd matrices, a vehicle millions of cells
// away from the grid, ...) the float32 decisions are switched off (coord_l < 0): every survivor of the cull is then
// decided by the float64 path, which is always valid.
void fill_fast32(const smap_handle* h, const smap_frame* f, const FrameParams& fp, Fast32& k) {
    memset(&k, 0, sizeof k);
    const double u53 = 1.1102230246251565e-16, p21 = 4.76837158203125e-07 /* 2^-21 */, slack = 1.0 + 9.765625e-4;
    const double p22 = 0.5 * p21, p23 = 0.25 * p21;
    double Tm[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (f->has_transform) memcpy(Tm, f->world_to_velodyne, sizeof Tm);
    // rows: 0..2 = M = P T, 3 = velodyne x (row 0 of T); abs = |P| |T| resp. |T row 0|
    double row[4][4], arow[4][4];
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0, aabs = 0.0;
            for (int q = 0; q < 4; ++q) {
                acc += fp.P[4 * r + q] * Tm[4 * q + j];
                aabs += fabs(fp.P[4 * r + q]) * fabs(Tm[4 * q + j]);
            }
            row[r][j] = acc;
            arow[r][j] = aabs;
        }
    for (int j = 0; j < 4; ++j) { row[3][j] = Tm[j]; arow[3][j] = fabs(Tm[j]); }
    const double W = (double)f->image_width, H = (double)f->image_height, R = fp.range_max;
    const double umax = (W > H ? W : H) + 2.0;
    const bool r_ok = (R == R) && fabs(R) < 1e30;

    // re-centring point: the velodyne origin in world coordinates (any float32 point is valid; this one keeps
    // the local coordinates small)
    float cf[3] = {0.f, 0.f, 0.f};
    if (f->has_transform)
        for (int j = 0; j < 3; ++j) {
            double c = 0.0;
            for (int i = 0; i < 3; ++i) c -= Tm[4 * i + j] * Tm[4 * i + 3];
            const float cc = (float)c;
            cf[j] = (cc == cc && fabsf(cc) < 1e30f) ? cc : 0.f;
        }
    const double cd[3] = {(double)cf[0], (double)cf[1], (double)cf[2]};
    const double cmax = fmax(fmax(fabs(cd[0]), fabs(cd[1])), fabs(cd[2]));

    // ---------------- conservative cull in world coordinates, constant error bounds for |x|,|y|,|z| <= bw
    {
        k.c_bw = f_down(cmax + 1024.0);
        const double bw = (double)k.c_bw * (1.0 + 1e-6);
        double S[4], E[4];
        for (int r = 0; r < 4; ++r) {
            double s = fabs(row[r][3]), sa = arow[r][3];
            for (int j = 0; j < 3; ++j) { s += fabs(row[r][j]) * bw; sa += arow[r][j] * bw; }
            S[r] = s;
            E[r] = (p21 * s + 64.0 * u53 * sa) * slack;
        }
        for (int j = 0; j < 4; ++j) {
            k.c_dc[j] = make_float2((float)row[3][j], (float)row[2][j]);
            k.c_ab[j] = make_float2((float)row[0][j], (float)row[1][j]);
        }
        k.c_wh = make_float2((float)W, (float)H);
        k.c_depth = f_up(E[2]);
        k.c_lo_u = f_down(-((E[0] + E[2]) + p23 * (S[0] + S[2])) * slack);
        k.c_hi_u = f_down(-((E[0] + W * E[2]) + p23 * (S[0] + W * S[2])) * slack);
        k.c_lo_v = f_down(-((E[1] + E[2]) + p23 * (S[1] + S[2])) * slack);
        k.c_hi_v = f_down(-((E[1] + H * E[2]) + p23 * (S[1] + H * S[2])) * slack);
        if (r_ok) {
            k.c_rh = (float)(0.5 * R);
            k.c_rthr = f_up((0.5 * R + E[3] + p22 * fabs(R) + p23 * S[3]) * slack);
        } else {   // non-finite RANGE_MAX: no range cull here, the float64 path applies the reference's own test
            k.c_rh = 0.f;
            k.c_rthr = (R == R) ? INFINITY : NAN;
        }
        // a bound that is not finite cannot certify anything: let every point through
        const double chk = E[0] + E[1] + W * E[2] + H * E[2] + E[3];
        if (!(chk == chk) || chk > 1e30) {
            k.c_bw = -1.f;
        }
    }

    // ---------------- certified float32 decisions in re-centred coordinates
    const double L = 512.0;
    const double pE = 0.75 * p21;   // 6 u: the row bound of the analysis is 5.03 u (1 + 3 u)
    k.n_ctr_xy = make_float2(-cf[0], -cf[1]);
    k.n_ctr_z = -cf[2];
    bool ok = r_ok;
    for (int r = 0; r < 4; ++r)
        for (int j = 0; j < 4; ++j)
            if (!(fabs(row[r][j]) < 1e30)) ok = false;   // NaN / inf in the matrices: fmax() below would skip them
    {
        const double bloc = cmax + L;
        // local rows: 0 = q0 - q2/2, 1 = q1 - q2/2, 2 = q2, 3 = velodyne x;  beta = row . (c, 1)
        double A[4][3], beta[4], e[4], amax[4];
        for (int r = 0; r < 4; ++r) {
            double rr[4], ra[4];
            for (int j = 0; j < 4; ++j) {
                rr[j] = row[r][j]; ra[j] = arow[r][j];
                if (r < 2) { rr[j] -= 0.5 * row[2][j]; ra[j] += 0.5 * arow[2][j]; }
            }
            double b = rr[3], sa = ra[3];
            amax[r] = 0.0;
            for (int j = 0; j < 3; ++j) {
                A[r][j] = rr[j];
                b += rr[j] * cd[j];
                sa += ra[j] * bloc;
                amax[r] = fmax(amax[r], fabs(rr[j]));
            }
            beta[r] = b;
            e[r] = 128.0 * u53 * sa;   // reference chain + host composition of beta
        }
        for (int j = 0; j < 3; ++j) {
            k.d_uv[j] = make_float2((float)A[0][j], (float)A[1][j]);
            k.d_cd[j] = make_float2((float)A[2][j], (float)A[3][j]);
        }
        k.d_uv[3] = make_float2((float)beta[0], (float)beta[1]);
        k.d_cd[3] = make_float2((float)beta[2], (float)beta[3]);
        const double k1 = (4.0 / 3.0) * pE * (fmax(amax[0], amax[1]) + umax * amax[2]) * slack;
        const double k0 = (4.0 / 3.0) * (fmax(pE * fabs(beta[0]) + e[0], pE * fabs(beta[1]) + e[1]) +
                                         umax * (pE * fabs(beta[2]) + e[2])) * slack;
        const double c0 = umax * p21 + 2.0 * p21;
        k.g_k1 = f_up(k1);
        k.g_k0 = f_up(k0);
        k.g_hc = f_down(0.5 - c0);
        k.r_h = (float)(0.5 * R);
        k.r_kd = f_up(pE * amax[3] * slack);
        k.r_thr0 = f_down((0.5 * R - (pE * fabs(beta[3]) + e[3] + 2.0 * p21 * fabs(R))) * (1.0 - 1e-6));
        k.mid_uv = make_float2((float)(0.5 * (W - 2.0)), (float)(0.5 * (H - 2.0)));
        k.half_uv = make_float2((float)(0.5 * W), (float)(0.5 * H));
        if (!(k1 < 1e10) || !(k0 < 1e10) || !(c0 < 0.25)) ok = false;   // also keeps MUFU.RCP away from flush-to-zero
    }
    {
        const smap_config& c = h->cfg;
        const double rinv = 1.0 / c.resolution;
        const double g0[2] = {((cd[0] + c.origin_offset_x) - c.boundary_x_min) / c.resolution,
                              ((cd[1] + c.origin_offset_y) - c.boundary_y_min) / c.resolution};
        const double i0[2] = {rint(g0[0]), rint(g0[1])};
        const double dims[2] = {(double)c.map_height, (double)c.map_width};
        const double lim = 2097152.0;   // 2^21
        if (!(fabs(i0[0]) < lim && fabs(i0[1]) < lim && L * fabs(rinv) < lim && dims[0] < lim && dims[1] < lim)) ok = false;
        const double span = cmax + L + fabs(c.origin_offset_x) + fabs(c.origin_offset_y) + fabs(c.boundary_x_min) + fabs(c.boundary_y_min);
        const double cc = p22 + 9.094947017729282e-13 * (span * fabs(rinv) + 1.0);
        k.cell_rf = (float)rinv;
        k.cell_kc = f_up(0.875 * p22 * fabs(rinv) * slack);   // 3.5 u: the analysis gives 3.1 u (1 + 3 u)
        k.cell_hg0 = f_down(0.5 - cc - p22);
        if (!(cc < 0.25)) ok = false;
        if (ok) {
            k.cell_f0 = make_float2((float)(g0[0] - i0[0] - 0.5), (float)(g0[1] - i0[1] - 0.5));
            k.mid_c = make_float2((float)(0.5 * (dims[0] - 2.0) - i0[0]), (float)(0.5 * (dims[1] - 2.0) - i0[1]));
            k.half_c = make_float2((float)(0.5 * dims[0]), (float)(0.5 * dims[1]));
            k.clamp_c = make_float2((float)(12582912.0 - i0[0]), (float)(12582912.0 - i0[1]));
            const int64_t mb = 0x4B400000ll;   // bit pattern of 1.5 * 2^23
            k.cell_k = (uint32_t)(uint64_t)(((int64_t)i0[0] - mb) * (int64_t)c.map_width + ((int64_t)i0[1] - mb));
            k.pix_k = (uint32_t)(uint64_t)(-mb * (int64_t)f->image_width - mb);
        }
    }
    k.coord_l = ok ? (float)L : -1.f;
}

int ensure_scratch(smap_handle* h, int64_t n) {
    if (n <= h->scratch_cap) return SMAP_OK;
    CK(cudaDeviceSynchronize());
    cudaFree(h->keep); cudaFree(h->iu); cudaFree(h->iv); cudaFree(h->blk_count); cudaFree(h->blk_offset);
    h->keep = nullptr; h->iu = nullptr; h->iv = nullptr; h->blk_count = nullptr; h->blk_offset = nullptr;
    h->scratch_cap = 0;
    const int64_t cap = n + n / 4 + 1024;
    const int64_t nb = ceil_div(cap, kCompactTile);
    CK(cudaMalloc(&h->keep, (size_t)cap));
    CK(cudaMalloc(&h->iu, sizeof(int32_t) * (size_t)cap));
    CK(cudaMalloc(&h->iv, sizeof(int32_t) * (size_t)cap));
    CK(cudaMalloc(&h->blk_count, sizeof(uint32_t) * (size_t)nb));
    CK(cudaMalloc(&h->blk_offset, sizeof(int64_t) * (size_t)nb));
    h->scratch_cap = cap;
    return SMAP_OK;
}

// Make room for `want` mask slots (slot 0 always exists).  New memory is zero = "never written".
// slot i of the current set (set = parity)
inline uint32_t* mask_slot(const smap_handle* h, int i) {
    return h->mask + ((size_t)h->parity * (size_t)h->n_slots + (size_t)i) * (size_t)h->slot_words;
}

// n_slots = slots per set; two sets.  Every slot is all zero between chunks, so growing needs no copy.
int ensure_slots(smap_handle* h, int want) {
    if (want <= h->n_slots) return SMAP_OK;
    uint32_t* fresh = nullptr;
    const size_t slot_bytes = sizeof(uint32_t) * (size_t)h->slot_words;
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&fresh, slot_bytes * 2 * want));
    CK(cudaMemset(fresh, 0, slot_bytes * 2 * want));
    if (h->mask) CK(cudaFree(h->mask));
    h->mask = fresh;
    h->n_slots = want;
    return SMAP_OK;
}

int ensure_apply_stream(smap_handle* h) {
    if (h->apply_stream) return SMAP_OK;
    CK(cudaStreamCreateWithFlags(&h->apply_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fused, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&h->ev_applied[i], cudaEventDisableTiming));
    return SMAP_OK;
}

// Before anything scatters into the current slot set: the apply / clear that last used the set has finished (its masks
// are zero again), the set's boxes are empty and its touched-cell counter is zero.
int begin_chunk(smap_handle* h, cudaStream_t st) {
    if (h->applied_pending[h->parity]) {
        CK(cudaStreamWaitEvent(st, h->ev_applied[h->parity], 0));
        h->applied_pending[h->parity] = false;
    }
    k_reset_slot_state<<<1, 32, 0, st>>>(h->boxes + (size_t)h->parity * kMaxBatch, h->touched + h->parity);
    CK(cudaGetLastError());
    return SMAP_OK;
}

// The caller's stream waits for every apply / clear still running on apply_stream (end of an API call: the
// stream-ordered contract -- whatever follows on `st` sees the updated grid).
int join_applies(smap_handle* h, cudaStream_t st) {
    for (int p = 0; p < 2; ++p) {
        if (!h->applied_pending[p]) continue;
        CK(cudaStreamWaitEvent(st, h->ev_applied[p], 0));
        h->applied_pending[p] = false;
    }
    return SMAP_OK;
}

// Stream the apply / clear of the current set runs on: apply_stream behind the chunk's scatter launches when `overlap`
// (more chunks follow), the caller's stream -- behind the previous chunk's apply, the grid updates are ordered --
// otherwise.
int apply_stream_for(smap_handle* h, cudaStream_t st, bool overlap, cudaStream_t* out) {
    if (overlap) {
        int rc = ensure_apply_stream(h);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_fused, st));
        CK(cudaStreamWaitEvent(h->apply_stream, h->ev_fused, 0));
        *out = h->apply_stream;
    } else {
        int rc = join_applies(h, st);
        if (rc) return rc;
        *out = st;
    }
    return SMAP_OK;
}

int applied(smap_handle* h, cudaStream_t as, bool overlap, bool writes_grid) {
    if (overlap) {
        CK(cudaEventRecord(h->ev_applied[h->parity], as));
        h->applied_pending[h->parity] = true;
        h->applied_writes_grid[h->parity] = writes_grid;
    }
    h->parity ^= 1;
    h->stats.kernel_launches += 1;
    return SMAP_OK;
}

// K3b launch: ordered apply of the `n_slots_used` mask slots of the current set (boxes[parity]); flips parity.
int launch_apply(smap_handle* h, double* map, int n_slots_used, cudaStream_t st, bool overlap = false) {
    const int c = h->cfg.num_classes;
    ApplyParams ap;
    memset(&ap, 0, sizeof ap);
    ap.n_frames = n_slots_used;
    for (int i = 0; i < n_slots_used; ++i) ap.mask[i] = mask_slot(h, i);
    const size_t smem = sizeof(double) * c * c;
    FrameBox* boxes = h->boxes + (size_t)h->parity * kMaxBatch;
    unsigned long long* tt = h->touched + h->parity;
    const unsigned grid = (unsigned)h->sm_count * 8;
    const int lane = h->cfg.lane_index, mw = h->cfg.map_width;
    cudaStream_t as = st;
    int rc = apply_stream_for(h, st, overlap, &as);
    if (rc) return rc;
#define SMAP_LAUNCH_APPLY(NJ) k_apply<NJ><<<grid, kThreads, smem, as>>>(map, ap, boxes, nullptr, tt, nullptr, h->abox, h->cm_dev, c, lane, mw)
    if (c <= 8) SMAP_LAUNCH_APPLY(1);
    else if (c <= 16) SMAP_LAUNCH_APPLY(2);
    else if (c <= 24) SMAP_LAUNCH_APPLY(3);
    else SMAP_LAUNCH_APPLY(4);
#undef SMAP_LAUNCH_APPLY
    CK(cudaGetLastError());
    return applied(h, as, overlap, true);
}

// After a count update through the masks (k_fuse MODE 2): zero the slots inside the frames' boxes; flips parity.
int launch_clear(smap_handle* h, int n_slots_used, cudaStream_t st, bool overlap = false) {
    ApplyParams ap;
    memset(&ap, 0, sizeof ap);
    ap.n_frames = n_slots_used;
    for (int i = 0; i < n_slots_used; ++i) ap.mask[i] = mask_slot(h, i);
    const dim3 grid((unsigned)h->sm_count, kMaxBatch);
    cudaStream_t as = st;
    int rc = apply_stream_for(h, st, overlap, &as);
    if (rc) return rc;
    k_clear_masks<<<grid, kThreads, 0, as>>>(ap, h->boxes + (size_t)h->parity * kMaxBatch, nullptr, nullptr, h->abox,
                                             h->cfg.map_width);
    CK(cudaGetLastError());
    return applied(h, as, overlap, false);
}

// Queue one k_stream_soa launch per non-empty (4, N) float64 frame; frame i of the non-empty ones scatters into
// mask slot i.  Returns the number of slots used in *slots_used.
int launch_stream(smap_handle* h, const smap_frame* frames, const FrameParams* fps, int n_frames, cudaStream_t st,
                  int* slots_used) {
    int used = 0;
    for (int i = 0; i < n_frames; ++i) {
        if (frames[i].n_points == 0) continue;
        if (frames[i].layout != SMAP_PTS_F64_SOA) return fail(SMAP_ERR_INVALID, "frames of one batch must share a point layout");
        StreamParams sp;
        sp.fp = fps[i];
        sp.pts = frames[i].points_dev;
        sp.image = frames[i].image_dev;
        sp.mask = mask_slot(h, used);
        sp.n = frames[i].n_points;
        sp.ld = frames[i].ld;
        // persistent grid: as many blocks as stay resident, never more than the cloud has rounds
        int64_t grid = (int64_t)h->sm_count * SMAP_STREAM_MINB;
        const int64_t rounds = ceil_div(sp.n, kBlockRoundPts);
        if (grid > rounds) grid = rounds;
        k_stream_soa<<<(unsigned)grid, kThreads, 0, st>>>(sp, h->gp, h->boxes + (size_t)h->parity * kMaxBatch + used);
        CK(cudaGetLastError());
        h->stats.kernel_launches += 1;
        ++used;
    }
    *slots_used = used;
    return SMAP_OK;
}

// Tag planes of the count update (k_fuse MODE 1); zero = "never written", frame tags start at 1.
//   one launch per frame:  one plane per internal stream (launches on one stream are serialised)
//   persistent launch:     one plane per frame of a launch, interleaved per element; as many as a batch has frames when
//                          memory allows (a quarter of what is free at the first call), never fewer than one
int ensure_tags(smap_handle* h, int want) {
    if (h->n_tag_planes > 0) return SMAP_OK;
    const size_t plane_bytes = sizeof(uint32_t) * (size_t)h->cells * (size_t)(h->cfg.num_classes + 1);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    int planes = want;
    while (planes > 1 && plane_bytes * (size_t)planes > free_b / 4) planes /= 2;
    CK(cudaMalloc(&h->tags, plane_bytes * planes));
    CK(cudaMemset(h->tags, 0, plane_bytes * planes));
    h->n_tag_planes = planes;
    h->frame_tag = 0;
    return SMAP_OK;
}

// id -> class bits: bit i set when the palette colour of the id matches cfg.LABEL_COLORS[i] in R and G
// (src/mapping_replay.py:276 compares only those two channels); ids beyond the palette are black, as
// apply_color_map's np.zeros canvas leaves them (mapillary_visualization.py:80-87).
int upload_id_lut(smap_handle* h) {
    if (!h->palette_set || !h->classes_set) return SMAP_OK;
    uint32_t lut[256];
    for (int id = 0; id < 256; ++id) {
        uint32_t bits = 0;
        for (int i = 0; i < h->cfg.num_classes; ++i)
            if (h->palette[3 * id] == h->colors[3 * i] && h->palette[3 * id + 1] == h->colors[3 * i + 1]) bits |= 1u << i;
        lut[id] = bits;
    }
    if (!h->id_lut_dev) CK(cudaMalloc(&h->id_lut_dev, sizeof lut));
    CK(cudaDeviceSynchronize());  // a previous frame may still be reading the table
    CK(cudaMemcpy(h->id_lut_dev, lut, sizeof lut, cudaMemcpyHostToDevice));
    return SMAP_OK;
}

// cv2.resize(..., INTER_NEAREST) index map of one axis, in double as OpenCV computes it:
// x_ofs[x] = min(cvFloor(x * ifx), src - 1), ifx = 1. / ((double)dst / src).
void nearest_table(int dst, int src, uint16_t* out) {
    const double inv_scale = (double)dst / (double)src;
    const double ifx = 1.0 / inv_scale;
    for (int x = 0; x < dst; ++x) {
        int sx = (int)floor((double)x * ifx);
        out[x] = (uint16_t)(sx < src - 1 ? sx : src - 1);
    }
}

// (m, s) with (x * m) >> s == tab[x] for every x < dst in 32-bit arithmetic, if there is one
bool nearest_mulshift(int dst, int src, const uint16_t* tab, uint32_t* m_out, uint32_t* s_out) {
    int bits = 1;
    while (((int64_t)1 << bits) < dst) ++bits;          // x < 2^bits
    for (int s = 31 - bits; s >= 1; --s) {               // m <= 2^s (src <= dst) keeps x * m below 2^32
        const uint64_t base = ((uint64_t)src << s) / (uint64_t)dst;
        for (uint64_t m = base; m <= base + 1; ++m) {
            if (m == 0 || (uint64_t)(dst - 1) * m >= ((uint64_t)1 << 32)) continue;
            bool ok = true;
            for (int x = 0; x < dst && ok; ++x) ok = (uint32_t)(((uint64_t)x * m) >> s) == tab[x];
            if (ok) { *m_out = (uint32_t)m; *s_out = (uint32_t)s; return true; }
        }
    }
    return false;
}

int ensure_nearest_map(smap_handle* h, int W, int H, int w, int hh) {
    smap_handle::NearestMap& nn = h->nn;
    if (nn.W == W && nn.H == H && nn.w == w && nn.h == hh) return SMAP_OK;
    std::vector<uint16_t> tab((size_t)W + (size_t)H);
    nearest_table(W, w, tab.data());
    nearest_table(H, hh, tab.data() + W);
    const bool ms = w <= W && hh <= H && nearest_mulshift(W, w, tab.data(), &nn.mx, &nn.sx) &&
                    nearest_mulshift(H, hh, tab.data() + W, &nn.my, &nn.sy);
    if (!ms) {
        // frames still in flight may be reading the previous table
        CK(cudaDeviceSynchronize());
        if (tab.size() > nn.tab_cap) {
            cudaFree(nn.tab_dev);
            nn.tab_dev = nullptr; nn.tab_cap = 0;
            CK(cudaMalloc(&nn.tab_dev, tab.size() * sizeof(uint16_t)));
            nn.tab_cap = tab.size();
        }
        CK(cudaMemcpy(nn.tab_dev, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    nn.use_tab = !ms;
    nn.W = W; nn.H = H; nn.w = w; nn.h = hh;
    return SMAP_OK;
}

inline int ids_w(const smap_frame* f) { return f->ids_width > 0 ? f->ids_width : f->image_width; }
inline int ids_h(const smap_frame* f) { return f->ids_height > 0 ? f->ids_height : f->image_height; }

// Everything about the float4 frames of a chunk that can be checked WITHOUT launching anything: run for the whole
// chunk before its first launch, so that a bad frame never leaves half a batch integrated (mask slots and boxes are
// only consistent between complete batches).
int validate_fuse_frame(const smap_handle* h, const smap_frame* fr) {
    if (fr->layout != SMAP_PTS_F32X4) return fail(SMAP_ERR_INVALID, "frames of one batch must share a point layout");
    if ((int64_t)fr->image_width * fr->image_height >= ((int64_t)1 << 28))
        return fail(SMAP_ERR_INVALID, "label image has 2^28 pixels or more");
    if (fr->n_points >= ((int64_t)1 << 40)) return fail(SMAP_ERR_INVALID, "cloud too large for one launch");
    if (fr->image_format == SMAP_IMG_CLASS_IDS) {
        if (!h->palette_set) return fail(SMAP_ERR_STATE, "class-id plane without a palette (smap_set_label_palette)");
        const int64_t w = fr->ids_width > 0 ? fr->ids_width : fr->image_width;
        const int64_t hh = fr->ids_height > 0 ? fr->ids_height : fr->image_height;
        if (w * hh >= ((int64_t)1 << 28)) return fail(SMAP_ERR_INVALID, "class-id plane has 2^28 pixels or more");
    }
    return SMAP_OK;
}

int fill_fuse_frame(smap_handle* h, const smap_frame* fr, const FrameParams& fp, int mode, int slot, FuseFrame& f) {
    f.nn_tab = nullptr;
    f.nn_mx = f.nn_my = f.nn_sx = f.nn_sy = 0;
    f.src_w = fr->image_width;
    f.pad = 0;
    if (fr->image_format == SMAP_IMG_CLASS_IDS) {
        int rc = ensure_nearest_map(h, fr->image_width, fr->image_height, ids_w(fr), ids_h(fr));
        if (rc) return rc;
        f.nn_tab = h->nn.use_tab ? h->nn.tab_dev : nullptr;
        f.nn_mx = h->nn.mx; f.nn_my = h->nn.my; f.nn_sx = h->nn.sx; f.nn_sy = h->nn.sy;
        f.src_w = ids_w(fr);
    }
    f.fp = fp;
    fill_fast32(h, fr, fp, f.fk);
    f.pts = static_cast<const float4*>(fr->points_dev);
    f.image = fr->image_dev;
    f.mask = mode == 1 ? nullptr : mask_slot(h, slot);
    f.fk.tag = ++h->frame_tag;   // larger than every tag written to this frame's plane before
    f.n = fr->n_points;
    f.img64 = ((reinterpret_cast<uintptr_t>(fr->image_dev) & 7u) == 0u &&
               ((int64_t)fr->image_width * fr->image_height * 3) % 8 == 0) ? 1 : 0;
    return SMAP_OK;
}

// persistent grid: as many blocks as stay resident, never more than the cloud has block-rounds
int64_t fuse_grid(const smap_handle* h, FuseFrame* f, int div) {
    int64_t gx = (int64_t)h->sm_count * SMAP_FUSE_MINB / div;
    const int64_t rounds = ceil_div(f->n, kFBlockRoundPts);
    if (gx > rounds) gx = rounds;
    if (gx < 1) gx = 1;
    f->per_warp = (int32_t)ceil_div(f->n, gx * kFWarps);   // n < 2^40 (validate_fuse_frame): fits
    return gx;
}

template <int MODE, int FMT>
cudaError_t set_fuse_smem() {
    return cudaFuncSetAttribute(k_fuse<MODE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, fuse_block_smem(MODE));
}

template <int MODE, int FMT>
void launch_k_fuse(const FuseLaunch& fb, const GridParams& gp, FrameBox* boxes, double* map, int64_t gx, cudaStream_t ls) {
    k_fuse<MODE, FMT><<<(unsigned)gx, kFThreads, fuse_block_smem(MODE), ls>>>(fb, gp, boxes, map);
}

const void* fuse_kernel(int mode, bool ids) {
    if (mode == 1) return ids ? (const void*)k_fuse<1, 1> : (const void*)k_fuse<1, 0>;
    if (mode == 2) return ids ? (const void*)k_fuse<2, 1> : (const void*)k_fuse<2, 0>;
    return ids ? (const void*)k_fuse<0, 1> : (const void*)k_fuse<0, 0>;
}

constexpr int kGraphMinFrames = 4;     // shorter chunks: plain launches
constexpr size_t kGraphCacheCap = 12;  // instantiated graphs kept per handle

void destroy_fuse_graph(smap_handle::FuseGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr; g.graph = nullptr; g.nodes.clear(); g.mode = -1;
}

// The chunk's k_fuse launches through a cached graph.  params[i] / gx[i] / boxes(i): launch i of the chunk.
// Returns SMAP_OK after the graph has been launched into `st`; any other value: nothing was launched (the caller falls
// back to per-frame launches).
int launch_fuse_graph(smap_handle* h, int mode, bool ids, int n, int lanes, cudaStream_t st) {
    smap_handle::FuseGraph* g = nullptr;
    for (auto& c : h->fuse_graphs)
        if (c.mode == mode && c.fmt == (ids ? 1 : 0) && c.n == n && c.lanes == lanes) { g = &c; break; }
    const void* func = fuse_kernel(mode, ids);
    double* map = h->acc;
    GridParams gp = h->gp;
    auto node_params = [&](int i, FrameBox** boxes_slot, void** args, cudaKernelNodeParams* np) {
        *boxes_slot = mode == 1 ? h->abox : h->boxes + (size_t)h->parity * kMaxBatch + i;
        args[0] = &h->graph_params[i];
        args[1] = &gp;
        args[2] = boxes_slot;
        args[3] = &map;
        memset(np, 0, sizeof *np);
        np->func = const_cast<void*>(func);
        np->gridDim = dim3((unsigned)h->graph_gx[i], 1, 1);
        np->blockDim = dim3(kFThreads, 1, 1);
        np->sharedMemBytes = (unsigned)fuse_block_smem(mode);
        np->kernelParams = args;
        np->extra = nullptr;
    };
    FrameBox* boxes_slot = nullptr;
    void* args[4];
    cudaKernelNodeParams np;
    if (!g) {
        if (h->fuse_graphs.size() >= kGraphCacheCap) {   // evict the least recently used one (none is being updated now)
            size_t victim = 0;
            for (size_t k = 1; k < h->fuse_graphs.size(); ++k)
                if (h->fuse_graphs[k].last_use < h->fuse_graphs[victim].last_use) victim = k;
            destroy_fuse_graph(h->fuse_graphs[victim]);
            h->fuse_graphs.erase(h->fuse_graphs.begin() + (long)victim);
        }
        smap_handle::FuseGraph fresh;
        fresh.mode = mode; fresh.fmt = ids ? 1 : 0; fresh.n = n; fresh.lanes = lanes;
        if (cudaGraphCreate(&fresh.graph, 0) != cudaSuccess) { cudaGetLastError(); return SMAP_ERR_CUDA; }
        fresh.nodes.resize((size_t)n);
        for (int i = 0; i < n; ++i) {
            node_params(i, &boxes_slot, args, &np);
            const cudaGraphNode_t* dep = i >= lanes ? &fresh.nodes[(size_t)(i - lanes)] : nullptr;
            if (cudaGraphAddKernelNode(&fresh.nodes[(size_t)i], fresh.graph, dep, dep ? 1 : 0, &np) != cudaSuccess) {
                cudaGetLastError();
                destroy_fuse_graph(fresh);
                return SMAP_ERR_CUDA;
            }
        }
        if (cudaGraphInstantiate(&fresh.exec, fresh.graph, 0) != cudaSuccess) {
            cudaGetLastError();
            destroy_fuse_graph(fresh);
            return SMAP_ERR_CUDA;
        }
        h->fuse_graphs.push_back(std::move(fresh));
        g = &h->fuse_graphs.back();
    } else {
        for (int i = 0; i < n; ++i) {
            node_params(i, &boxes_slot, args, &np);
            if (cudaGraphExecKernelNodeSetParams(g->exec, g->nodes[(size_t)i], &np) != cudaSuccess) {
                cudaGetLastError();
                // the executable graph may be half updated: drop it, the caller launches this chunk the plain way
                destroy_fuse_graph(*g);
                h->fuse_graphs.erase(h->fuse_graphs.begin() + (g - h->fuse_graphs.data()));
                return SMAP_ERR_CUDA;
            }
        }
    }
    g->last_use = ++h->graph_clock;
    if (cudaGraphLaunch(g->exec, st) != cudaSuccess) { cudaGetLastError(); return SMAP_ERR_CUDA; }
    return SMAP_OK;
}

// Queue the float4 frames of a batch (validated by the caller), one launch per frame, alternating over the internal
// streams (fork / join with events around the batch).  Count update: nothing else to do afterwards; otherwise frame i
// of the non-empty ones scatters into mask slot i and *slots_used tells k_apply / k_clear_masks how many there are.
// The join is queued even when a launch fails, and *slots_used is valid then too, so that the caller can restore the
// "slots all zero between batches" invariant before it reports the error.
// mode: 0 ordered update (masks, k_apply afterwards), 1 count update with tags, 2 count update through the masks
// (k_clear_masks afterwards)
int launch_fuse(smap_handle* h, const smap_frame* frames, const FrameParams* fps, int n_frames, int mode,
                cudaStream_t st, int* slots_used) {
    int used = 0;
    *slots_used = 0;
    const bool count_atomics = mode == 1;
    if (count_atomics) {
        int rc = ensure_tags(h, smap_handle::kAux > 0 ? smap_handle::kAux : 1);
        if (rc) return rc;
    }
    {
        // the tag counter is advanced by every launch, whatever the mode: start over before it can wrap
        if (h->frame_tag > 0xffffffffu - (uint32_t)n_frames - 1u) {
            CK(cudaDeviceSynchronize());
            if (h->tags)
                CK(cudaMemset(h->tags, 0, sizeof(uint32_t) * (size_t)h->cells * (size_t)(h->cfg.num_classes + 1) * h->n_tag_planes));
            h->frame_tag = 0;
        }
    }
    if (!h->fuse_attr_set) {   // the per-warp stacks need more than the default 48 KB per block
        CK((set_fuse_smem<0, 0>())); CK((set_fuse_smem<1, 0>())); CK((set_fuse_smem<2, 0>()));
        CK((set_fuse_smem<0, 1>())); CK((set_fuse_smem<1, 1>())); CK((set_fuse_smem<2, 1>()));
        h->fuse_attr_set = true;
    }
    int n_nonempty = 0;
    for (int i = 0; i < n_frames; ++i) n_nonempty += frames[i].n_points > 0;
    const bool fork = smap_handle::kAux > 0 && n_nonempty > 1 && !h->profiling;
    const size_t plane_words = (size_t)h->cells * (size_t)(h->cfg.num_classes + 1);
    if (fork && h->use_graph && n_nonempty >= kGraphMinFrames) {
        // ---- one graph launch for the chunk (frames of one label format)
        bool same_fmt = true;
        int first_fmt = -1;
        for (int i = 0; i < n_frames; ++i) {
            if (frames[i].n_points == 0) continue;
            if (first_fmt < 0) first_fmt = frames[i].image_format;
            same_fmt = same_fmt && frames[i].image_format == first_fmt;
        }
        if (same_fmt) {
            const int lanes = (mode == 1 && h->n_tag_planes < smap_handle::kAux) ? h->n_tag_planes : smap_handle::kAux;
            const uint32_t tag0 = h->frame_tag;
            h->graph_params.resize((size_t)n_nonempty);
            h->graph_gx.resize((size_t)n_nonempty);
            int k = 0, rc = SMAP_OK;
            for (int i = 0; i < n_frames && !rc; ++i) {
                if (frames[i].n_points == 0) continue;
                FuseLaunch& fb = h->graph_params[(size_t)k];
                rc = fill_fuse_frame(h, frames + i, fps[i], mode, k, fb.f);
                if (rc) break;
                h->graph_gx[(size_t)k] = fuse_grid(h, &fb.f, SMAP_FUSE_GRID_DIV);
                fb.tags = count_atomics ? h->tags + plane_words * (size_t)(k % lanes) : nullptr;
                fb.id_lut = h->id_lut_dev;
                ++k;
            }
            if (rc) { h->frame_tag = tag0; return rc; }   // nothing has been launched
            if (launch_fuse_graph(h, mode, first_fmt == SMAP_IMG_CLASS_IDS, n_nonempty, lanes, st) == SMAP_OK) {
                h->stats.kernel_launches += n_nonempty;
                *slots_used = n_nonempty;
                return SMAP_OK;
            }
            h->frame_tag = tag0;   // the graph could not be built / launched: the plain launches below redo the chunk
        }
    }
    if (fork) {
        if (!h->ev_fork) {
            CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            for (int a = 0; a < smap_handle::kAux; ++a) {
                CK(cudaStreamCreateWithFlags(&h->aux[a], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&h->ev_join[a], cudaEventDisableTiming));
            }
        }
        CK(cudaEventRecord(h->ev_fork, st));
        for (int a = 0; a < smap_handle::kAux; ++a) CK(cudaStreamWaitEvent(h->aux[a], h->ev_fork, 0));
    }
    FuseLaunch* fb = &h->fuse_one;
    int rc = SMAP_OK;
    for (int i = 0; i < n_frames && !rc; ++i) {
        if (frames[i].n_points == 0) continue;
        // tagged count update: never more streams than tag planes (two frames on one plane must not overlap)
        const int n_streams = (mode == 1 && h->n_tag_planes < smap_handle::kAux) ? h->n_tag_planes : smap_handle::kAux;
        const int lane_stream = fork ? used % n_streams : 0;
        cudaStream_t ls = fork ? h->aux[lane_stream] : st;
        rc = fill_fuse_frame(h, frames + i, fps[i], mode, used, fb->f);
        if (rc) break;
        const int64_t gx = fuse_grid(h, &fb->f, fork ? SMAP_FUSE_GRID_DIV : 1);
        // one tag plane per launching stream; a frame's tag is larger than every tag written to its plane before
        fb->tags = count_atomics ? h->tags + plane_words * (size_t)lane_stream : nullptr;
        fb->id_lut = h->id_lut_dev;
        // mode 1: the union window itself; otherwise the frame's own box (folded into the window by k_apply / k_clear_masks)
        FrameBox* boxes = mode == 1 ? h->abox : h->boxes + (size_t)h->parity * kMaxBatch + used;
        const bool ids = frames[i].image_format == SMAP_IMG_CLASS_IDS;
        if (mode == 1) ids ? launch_k_fuse<1, 1>(*fb, h->gp, boxes, h->acc, gx, ls) : launch_k_fuse<1, 0>(*fb, h->gp, boxes, h->acc, gx, ls);
        else if (mode == 2) ids ? launch_k_fuse<2, 1>(*fb, h->gp, boxes, h->acc, gx, ls) : launch_k_fuse<2, 0>(*fb, h->gp, boxes, h->acc, gx, ls);
        else ids ? launch_k_fuse<0, 1>(*fb, h->gp, boxes, h->acc, gx, ls) : launch_k_fuse<0, 0>(*fb, h->gp, boxes, h->acc, gx, ls);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fail(SMAP_ERR_CUDA, "k_fuse launch: %s", cudaGetErrorString(e)); break; }
        h->stats.kernel_launches += 1;
        ++used;
    }
    *slots_used = used;
    if (fork) {
        for (int a = 0; a < smap_handle::kAux; ++a) {
            if (cudaEventRecord(h->ev_join[a], h->aux[a]) != cudaSuccess || cudaStreamWaitEvent(st, h->ev_join[a], 0) != cudaSuccess) {
                cudaGetLastError();
                if (!rc) rc = fail(SMAP_ERR_CUDA, "joining the internal streams failed");
            }
        }
    }
    return rc;
}

int harvest_profile(smap_handle* h) {
    for (int i = 0; i < h->n_prof_pending; ++i) {
        smap_handle::ProfRec& r = h->prof_pending[i];
        CK(cudaEventSynchronize(r.e[2]));
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, r.e[0], r.e[1]));
        CK(cudaEventElapsedTime(&b, r.e[1], r.e[2]));
        h->stats.stream_kernel_ms += a;
        h->stats.apply_kernel_ms += b;
        h->stats.profiled_frames += r.frames;
        for (int k = 0; k < 3; ++k) cudaEventDestroy(r.e[k]);
    }
    h->n_prof_pending = 0;
    return SMAP_OK;
}

int render_common_checks(const void* map, int mh, int mw, int c) {
    if (!map) return fail(SMAP_ERR_INVALID, "map is NULL");
    if (mh <= 0 || mw <= 0) return fail(SMAP_ERR_INVALID, "empty grid");
    if (c < 1 || c > 32) return fail(SMAP_ERR_INVALID, "num_classes must be in 1..32 for rendering");
    return SMAP_OK;
}

template <bool FILTER>
int launch_render(const double* map, int mh, int mw, int c, const uint8_t* colors_host, uint8_t* rgb,
                  double* filtered, cudaStream_t st) {
    RenderColors rc;
    memset(&rc, 0, sizeof rc);
    if (colors_host) memcpy(rc.rgb, colors_host, (size_t)c * 3);
    const int cs = c | 1;                                    // cells padded to an odd number of doubles in shared memory
    const uint32_t div_c = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)c - 1) / (uint64_t)c);   // umulhi(e, div_c) = e / c, e < 2^16
    // fewer than 8 classes: strips of 4 cells, one running class sum; otherwise strips of 2 and numpy's eight partial sums
    const bool small = c < 8;
    const int r = small ? kRSmall : 2;
    dim3 grid((unsigned)ceil_div(mw, kRX), (unsigned)ceil_div(mh, render_tile_rows(r)));
    if (grid.y > 65535u) return fail(SMAP_ERR_INVALID, "grid has too many rows for the render kernel");
    // tile rows staged by bulk copies: cells of an odd number of doubles, even number of columns, 16-byte aligned grid
    // (smap_render.cuh, k_render_bulk); every other grid goes through the register-staged kernel
    const size_t bulk_smem = render_bulk_smem_bytes(FILTER, r, c);
    if (SMAP_RENDER_BULK && (c & 1) && !(mw & 1) && (reinterpret_cast<uintptr_t>(map) & 15u) == 0 && bulk_smem <= 200 * 1024) {
#define SMAP_LAUNCH_RENDER_BULK(RR, NPS, CT)                                                                               \
    do {                                                                                                                  \
        if (FILTER && filtered) {                                                                                         \
            CK(cudaFuncSetAttribute(k_render_bulk<FILTER, FILTER, RR, NPS, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem)); \
            k_render_bulk<FILTER, FILTER, RR, NPS, CT><<<grid, kRThreads, bulk_smem, st>>>(map, mh, mw, c, rc, rgb, filtered); \
        } else {                                                                                                          \
            CK(cudaFuncSetAttribute(k_render_bulk<FILTER, false, RR, NPS, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem)); \
            k_render_bulk<FILTER, false, RR, NPS, CT><<<grid, kRThreads, bulk_smem, st>>>(map, mh, mw, c, rc, rgb, filtered); \
        }                                                                                                                 \
    } while (0)
        // the reference's two class counts (cfg.LABELS: 5 by default, 19 in full) as compile-time constants
        if (c == 5) SMAP_LAUNCH_RENDER_BULK(kRSmall, 1, 5);
        else if (c == 19) SMAP_LAUNCH_RENDER_BULK(2, 8, 19);
        else if (small) SMAP_LAUNCH_RENDER_BULK(kRSmall, 1, 0);
        else SMAP_LAUNCH_RENDER_BULK(2, 8, 0);
#undef SMAP_LAUNCH_RENDER_BULK
        CK(cudaGetLastError());
        return SMAP_OK;
    }
    const size_t smem = render_smem_bytes(FILTER, r, cs);
    if (small) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_render<FILTER, kRSmall, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_render<FILTER, kRSmall, 1><<<grid, kRThreads, smem, st>>>(map, mh, mw, c, cs, div_c, rc, rgb, filtered);
    } else {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_render<FILTER, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_render<FILTER, 2, 8><<<grid, kRThreads, smem, st>>>(map, mh, mw, c, cs, div_c, rc, rgb, filtered);
    }
    CK(cudaGetLastError());
    return SMAP_OK;
}

#define NCCLCK(expr)                                                                                     \
    do {                                                                                                 \
        ncclResult_t r__ = (expr);                                                                       \
        if (r__ != ncclSuccess) return fail(SMAP_ERR_COMM, "%s: %s", #expr, nccl_api().GetErrorString(r__)); \
    } while (0)

int ensure_bytes(void** buf, size_t* cap, size_t need) {
    if (need <= *cap) return SMAP_OK;
    if (*buf) { CK(cudaDeviceSynchronize()); cudaFree(*buf); *buf = nullptr; *cap = 0; }
    const size_t want = need + need / 8 + 256;
    CK(cudaMalloc(buf, want));
    *cap = want;
    return SMAP_OK;
}

// The ranks agree on the exchange: union window, value bound, integer or not (element-wise MAX of kCommWords ints).
// Synchronises `st` (the host sizes the NCCL calls from the result).
int comm_agree(smap_handle* h, cudaStream_t st, int out[kCommWords]) {
    NcclApi& N = nccl_api();
    const int vb = h->value_bound > 0x7fffffffll ? 0x7fffffff : (int)h->value_bound;
    k_comm_prep<<<1, 32, 0, st>>>(h->ubox, h->ubox_full ? 1 : 0, h->cfg.map_height, h->cfg.map_width, vb,
                                  h->integer_grid ? 0 : 1, h->xch_dev);
    CK(cudaGetLastError());
    NCCLCK(N.AllReduce(h->xch_dev, h->xch_dev, kCommWords, ncclInt32, ncclMax, h->comm, st));
    CK(cudaMemcpyAsync(h->xch_host, h->xch_dev, sizeof(int) * kCommWords, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < kCommWords; ++i) out[i] = h->xch_host[i];
    return SMAP_OK;
}

int pick_pack(const smap_handle* h, const int agreed[kCommWords], int64_t* bound_total) {
    *bound_total = (int64_t)h->n_ranks * (int64_t)agreed[4];
    if (agreed[5]) return kPackF64;
    if (*bound_total < 65536) return kPackU16;
    if (*bound_total < ((int64_t)1 << 32)) return kPackU32;
    return kPackF64;
}

inline dim3 window_grid(int units, int rows) {
    int gx = (int)ceil_div(units, (int64_t)kThreads * 4);
    if (gx < 1) gx = 1;
    if (gx > 32) gx = 32;
    return dim3((unsigned)gx, (unsigned)rows);
}

// rows [row0, row0 + rows) of the packed buffer <- grid rows of the same index, columns [y0, y0 + cols);
// grid rows outside [wx0, wx1] are written as zeros
int launch_pack(const smap_handle* h, double* src, bool clear, int pack, int row0, int rows, int wx0, int wx1, int y0,
                int cols, void* out, cudaStream_t st) {
    const int c = h->cfg.num_classes, run = cols * c, run_words = pack == kPackU16 ? (run + 1) / 2 : run;
    const size_t row_bytes = pack == kPackF64 ? (size_t)run * 8 : (size_t)run_words * 4;
    for (int done = 0; done < rows; done += 32768) {
        const int n = rows - done < 32768 ? rows - done : 32768;
        void* o = static_cast<char*>(out) + (size_t)done * row_bytes;
        const dim3 g = window_grid(pack == kPackU16 ? 2 * run_words : run, n);
#define SMAP_PACK(P, CL) k_pack_window<P, CL><<<g, kThreads, 0, st>>>(src, wx0, wx1, h->cfg.map_width, c, row0 + done, y0, run, run_words, o, 0ll)
        if (pack == kPackU16) { if (clear) SMAP_PACK(kPackU16, true); else SMAP_PACK(kPackU16, false); }
        else if (pack == kPackU32) { if (clear) SMAP_PACK(kPackU32, true); else SMAP_PACK(kPackU32, false); }
        else { if (clear) SMAP_PACK(kPackF64, true); else SMAP_PACK(kPackF64, false); }
#undef SMAP_PACK
        CK(cudaGetLastError());
    }
    return SMAP_OK;
}

// rows [dst_row0, dst_row0 + rows) of a (*, dst_mw, C) float64 array <- rows [src_row0, ...) of the packed buffer
int launch_unpack(const smap_handle* h, int pack, double* dst, int dst_mw, int dst_row0, int rows, int y0, int cols,
                  const void* in, int src_row0, cudaStream_t st, bool add = false) {
    const int c = h->cfg.num_classes, run = cols * c, run_words = pack == kPackU16 ? (run + 1) / 2 : run;
    for (int done = 0; done < rows; done += 32768) {
        const int n = rows - done < 32768 ? rows - done : 32768;
        const dim3 g = window_grid(run, n);
#define SMAP_UNPACK(P, AD) k_unpack_window<P, AD><<<g, kThreads, 0, st>>>(dst, dst_mw, c, dst_row0 + done, y0, run, run_words, in, src_row0 + done)
        if (pack == kPackU16) { if (add) SMAP_UNPACK(kPackU16, true); else SMAP_UNPACK(kPackU16, false); }
        else if (pack == kPackU32) { if (add) SMAP_UNPACK(kPackU32, true); else SMAP_UNPACK(kPackU32, false); }
        else { if (add) SMAP_UNPACK(kPackF64, true); else SMAP_UNPACK(kPackF64, false); }
#undef SMAP_UNPACK
        CK(cudaGetLastError());
    }
    return SMAP_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int smap_abi_version(void) { return SMAP_ABI_VERSION; }

const char* smap_last_error(void) { return g_err.c_str(); }

int smap_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return fail(SMAP_ERR_NO_DEVICE, "cudaGetDeviceCount failed");
    }
    return n;
}

int smap_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (name && name_len > 0) {
        strncpy(name, p.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return SMAP_OK;
}

int smap_create(const smap_config* cfg, smap_handle** out) {
    if (!cfg || !out) return fail(SMAP_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->map_height <= 0 || cfg->map_width <= 0) return fail(SMAP_ERR_INVALID, "empty grid");
    if (cfg->num_classes < 1 || cfg->num_classes > SMAP_MAX_CLASSES)
        return fail(SMAP_ERR_INVALID, "num_classes must be in 1..31");
    if ((int64_t)cfg->map_height * cfg->map_width >= (int64_t)1 << 31)
        return fail(SMAP_ERR_INVALID, "grid has more than 2^31 cells");
    if (cfg->lane_index >= cfg->num_classes) return fail(SMAP_ERR_INVALID, "lane_index out of range");
    if (!(cfg->resolution == cfg->resolution) || cfg->resolution == 0.0)
        return fail(SMAP_ERR_INVALID, "resolution must be non-zero");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return fail(SMAP_ERR_NO_DEVICE, "no such CUDA device");
    }
    DeviceGuard guard(cfg->device);
    if (!guard.ok) return fail(SMAP_ERR_CUDA, "cudaSetDevice failed");
    smap_handle* h = new (std::nothrow) smap_handle();
    if (!h) return fail(SMAP_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    h->cells = (int64_t)cfg->map_height * cfg->map_width;
    GridParams& g = h->gp;
    memset(&g, 0, sizeof g);
    g.off_x = cfg->origin_offset_x; g.off_y = cfg->origin_offset_y;
    g.bx0 = cfg->boundary_x_min; g.by0 = cfg->boundary_y_min;
    g.res = cfg->resolution;
    g.mh = cfg->map_height; g.mw = cfg->map_width; g.c = cfg->num_classes;
    g.lane = cfg->lane_index < 0 ? -1 : cfg->lane_index;
    g.use_intensity = cfg->use_intensity ? 1 : 0;
    g.rinv = 1.0 / cfg->resolution;
    g.mh_d1 = (double)cfg->map_height + 1.0;
    g.mw_d1 = (double)cfg->map_width + 1.0;
    const size_t map_bytes = sizeof(double) * (size_t)h->cells * cfg->num_classes;
    cudaError_t e = cudaSuccess;
    if (cfg->map_dev) {
        h->map = static_cast<double*>(cfg->map_dev);
        h->integer_grid = cfg->map_is_zero != 0;
    } else {
        e = cudaMalloc(&h->map, map_bytes);
        if (e == cudaSuccess) { h->own_map = true; e = cudaMemset(h->map, 0, map_bytes); }
        h->integer_grid = true;
    }
    h->acc = h->map;
    {
        const char* eg = getenv("SMAP_FUSE_GRAPH");
        h->use_graph = !(eg && eg[0] == '0');
    }
    h->slot_words = (h->cells + 3) / 4 * 4;
    if (e == cudaSuccess) e = cudaMalloc(&h->mask, sizeof(uint32_t) * (size_t)h->slot_words * 2);   // two sets of one slot
    if (e == cudaSuccess) e = cudaMemset(h->mask, 0, sizeof(uint32_t) * (size_t)h->slot_words * 2);
    if (e == cudaSuccess) h->n_slots = 1;
    if (e == cudaSuccess) e = cudaMalloc(&h->boxes, sizeof(FrameBox) * (2 * kMaxBatch + 3));
    if (e == cudaSuccess) {
        FrameBox init[2 * kMaxBatch + 3];
        for (int i = 0; i < 2 * kMaxBatch + 3; ++i) { init[i].x0 = 0x7fffffff; init[i].x1 = -1; init[i].y0 = 0x7fffffff; init[i].y1 = -1; }
        e = cudaMemcpy(h->boxes, init, sizeof init, cudaMemcpyHostToDevice);
        h->ubox = h->boxes + 2 * kMaxBatch;
        h->dbox[0] = h->ubox + 1;
        h->dbox[1] = h->ubox + 2;
        h->abox = h->ubox;
    }
    // a caller-owned grid that is not declared zero may hold anything anywhere
    h->ubox_full = cfg->map_dev && !cfg->map_is_zero;
    if (e == cudaSuccess) e = cudaMalloc(&h->touched, sizeof(unsigned long long) * 2);
    if (e == cudaSuccess) e = cudaMemset(h->touched, 0, sizeof(unsigned long long) * 2);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->cm_dev, sizeof(double) * SMAP_MAX_CLASSES * SMAP_MAX_CLASSES);
    if (e == cudaSuccess) e = cudaMalloc(&h->total_dev, sizeof(int64_t));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fail(e == cudaErrorMemoryAllocation ? SMAP_ERR_NOMEM : SMAP_ERR_CUDA, "smap_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        smap_destroy(h);
        return e == cudaErrorMemoryAllocation ? SMAP_ERR_NOMEM : SMAP_ERR_CUDA;
    }
    *out = h;
    return SMAP_OK;
}

int smap_destroy(smap_handle* h) {
    if (!h) return SMAP_OK;
    DeviceGuard guard(h->cfg.device);
    cudaDeviceSynchronize();
    harvest_profile(h);
    if (h->own_map) cudaFree(h->map);
    for (auto& g : h->fuse_graphs) destroy_fuse_graph(g);
    h->fuse_graphs.clear();
    if (h->apply_stream) cudaStreamDestroy(h->apply_stream);
    if (h->ev_fused) cudaEventDestroy(h->ev_fused);
    for (int i = 0; i < 2; ++i) if (h->ev_applied[i]) cudaEventDestroy(h->ev_applied[i]);
    cudaFree(h->mask); cudaFree(h->tags); cudaFree(h->boxes); cudaFree(h->touched); cudaFree(h->cm_dev); cudaFree(h->total_dev);
    cudaFree(h->id_lut_dev); cudaFree(h->nn.tab_dev);
    if (h->comm && h->own_comm && nccl_api().lib) nccl_api().CommDestroy(h->comm);
    cudaFree(h->xbuf); cudaFree(h->xbuf2); cudaFree(h->xch_dev);
    if (h->xch_host) cudaFreeHost(h->xch_host);
    for (int b = 0; b < 2; ++b) {
        cudaFree(h->delta[b]); cudaFree(h->xch_dev2[b]);
        if (h->xch_host2[b]) cudaFreeHost(h->xch_host2[b]);
        if (h->ev_agreed[b]) cudaEventDestroy(h->ev_agreed[b]);
        if (h->ev_packed[b]) cudaEventDestroy(h->ev_packed[b]);
    }
    for (int i = 0; i < 4; ++i)
        if (h->ev_phase[i]) cudaEventDestroy(h->ev_phase[i]);
    if (h->ev_chunk) cudaEventDestroy(h->ev_chunk);
    if (h->ev_exchanged) cudaEventDestroy(h->ev_exchanged);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    cudaFree(h->keep); cudaFree(h->iu); cudaFree(h->iv); cudaFree(h->blk_count); cudaFree(h->blk_offset);
    if (h->ev_fork) {
        cudaEventDestroy(h->ev_fork);
        for (int a = 0; a < smap_handle::kAux; ++a) {
            cudaEventDestroy(h->ev_join[a]);
            cudaStreamDestroy(h->aux[a]);
        }
    }
    for (int i = 0; i < smap_handle::kStages; ++i) {
        cudaFree(h->stage_pts[i]);
        cudaFree(h->stage_img[i]);
        if (h->stage_done[i]) cudaEventDestroy(h->stage_done[i]);
        if (h->stage_copied[i]) cudaEventDestroy(h->stage_copied[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaGetLastError();
    delete h;
    return SMAP_OK;
}

int smap_set_camera(smap_handle* h, int camera, const double P_host[12]) {
    if (!h || !P_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (camera < 0 || camera >= SMAP_MAX_CAMERAS) return fail(SMAP_ERR_INVALID, "camera slot out of range");
    memcpy(h->P[camera], P_host, sizeof(double) * 12);
    h->cam_set[camera] = true;
    return SMAP_OK;
}

int smap_set_classes(smap_handle* h, const uint8_t* colors_host, const double* cm_host) {
    if (!h || !colors_host || !cm_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    const int c = h->cfg.num_classes;
    memcpy(h->colors, colors_host, (size_t)c * 3);
    for (int i = 0; i < c; ++i) {
        h->gp.col_r[i] = colors_host[3 * i];
        h->gp.col_g[i] = colors_host[3 * i + 1];
    }
    CK(cudaDeviceSynchronize());  // a previous frame may still be reading the table
    CK(cudaMemcpy(h->cm_dev, cm_host, sizeof(double) * c * c, cudaMemcpyHostToDevice));
    h->identity_cm = true;
    for (int i = 0; i < c; ++i)
        for (int j = 0; j < c; ++j)
            if (cm_host[i * c + j] != (i == j ? 1.0 : 0.0)) h->identity_cm = false;
    h->classes_set = true;
    return upload_id_lut(h);
}

int smap_set_label_palette(smap_handle* h, const uint8_t* rgb_host, int n_ids) {
    if (!h || (n_ids > 0 && !rgb_host)) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (n_ids < 0 || n_ids > 256) return fail(SMAP_ERR_INVALID, "a palette has between 0 and 256 colours");
    DeviceGuard guard(h->cfg.device);
    memset(h->palette, 0, sizeof h->palette);
    if (n_ids > 0) memcpy(h->palette, rgb_host, (size_t)n_ids * 3);
    h->palette_set = true;
    return upload_id_lut(h);
}

int smap_project(smap_handle* h, const smap_frame* frame, double* out_pcd, uint8_t* out_label, int32_t* out_uv,
                 uint8_t* out_keep, int64_t out_ld, int64_t* m_host, void* stream) {
    if (!h || !m_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    FrameParams fp;
    int rc = fill_frame_params(h, frame, fp);
    if (rc) return rc;
    if (frame->image_format != SMAP_IMG_RGB) return fail(SMAP_ERR_INVALID, "smap_project returns RGB labels: it takes RGB label images only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n = frame->n_points;
    *m_host = 0;
    if (n == 0) return SMAP_OK;
    if (!out_pcd || !out_label) return fail(SMAP_ERR_INVALID, "NULL output");
    if (out_ld < n) return fail(SMAP_ERR_INVALID, "out_ld < n_points");
    rc = ensure_scratch(h, n);
    if (rc) return rc;
    const int64_t nb = ceil_div(n, kCompactTile);
    if (frame->layout == SMAP_PTS_F32X4)
        k_project_flags<SMAP_PTS_F32X4><<<(unsigned)nb, kThreads, 0, st>>>(frame->points_dev, n, frame->ld, fp, h->keep, h->iu, h->iv, h->blk_count);
    else
        k_project_flags<SMAP_PTS_F64_SOA><<<(unsigned)nb, kThreads, 0, st>>>(frame->points_dev, n, frame->ld, fp, h->keep, h->iu, h->iv, h->blk_count);
    CK(cudaGetLastError());
    k_scan_blocks<<<1, 1024, 0, st>>>(h->blk_count, h->blk_offset, nb, h->total_dev);
    CK(cudaGetLastError());
    if (frame->layout == SMAP_PTS_F32X4)
        k_compact<SMAP_PTS_F32X4><<<(unsigned)nb, kThreads, 0, st>>>(frame->points_dev, n, frame->ld, frame->image_dev, fp.img_w, h->keep, h->iu, h->iv, h->blk_offset, out_pcd, out_label, out_uv, out_ld);
    else
        k_compact<SMAP_PTS_F64_SOA><<<(unsigned)nb, kThreads, 0, st>>>(frame->points_dev, n, frame->ld, frame->image_dev, fp.img_w, h->keep, h->iu, h->iv, h->blk_offset, out_pcd, out_label, out_uv, out_ld);
    CK(cudaGetLastError());
    if (out_keep) CK(cudaMemcpyAsync(out_keep, h->keep, (size_t)n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(m_host, h->total_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->stats.kernel_launches += 3;
    h->last_stream = st;
    return SMAP_OK;
}

int smap_update(smap_handle* h, double* map_dev, const double* pcd, int64_t ld, const uint8_t* label, int64_t ldl,
                int64_t m, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (!h->classes_set) return fail(SMAP_ERR_STATE, "classes not set (smap_set_classes)");
    if (m < 0 || ld < m || ldl < m) return fail(SMAP_ERR_INVALID, "bad point count / strides");
    if (m > 0 && (!pcd || !label)) return fail(SMAP_ERR_INVALID, "NULL pcd / label");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (m == 0) return SMAP_OK;
    int64_t grid = ceil_div(m, kThreads);
    if (grid > (int64_t)h->sm_count * 16) grid = (int64_t)h->sm_count * 16;
    {
        int rc = begin_chunk(h, st);
        if (rc) return rc;
    }
    k_update_scatter<<<(unsigned)grid, kThreads, 0, st>>>(pcd, ld, label, ldl, m, h->gp, mask_slot(h, 0),
                                                         h->boxes + (size_t)h->parity * kMaxBatch);
    CK(cudaGetLastError());
    h->stats.kernel_launches += 1;
    h->last_stream = st;
    double* target = map_dev ? map_dev : h->acc;
    if (target == h->acc) {
        acc_integer(h) = acc_integer(h) && h->identity_cm;
        acc_bound(h) += 3;
    } else if (target == h->map) {   // streaming: the caller named the grid itself
        h->integer_grid = h->integer_grid && h->identity_cm;
        h->value_bound += 3;
        h->ubox_full = true;          // k_apply folds the frame's box into the accumulation window, not the grid's
    }
    h->last_update_counted = true;
    return launch_apply(h, target, 1, st);
}

int smap_integrate_batch(smap_handle* h, const smap_frame* frames, int n_frames, void* stream) {
    if (!h || (n_frames > 0 && !frames)) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (n_frames < 0) return fail(SMAP_ERR_INVALID, "n_frames < 0");
    if (!h->classes_set) return fail(SMAP_ERR_STATE, "classes not set (smap_set_classes)");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    h->last_stream = st;
    // The tagged count update keeps no per-frame state (no mask slot, no box), so its chunks -- one fork / join of the
    // internal streams each -- may be longer than the kMaxBatch frames the mask paths apply in one pass.
    constexpr int kTagChunk = 4 * kMaxBatch;
    std::vector<FrameParams> fps_store((size_t)(n_frames < kTagChunk ? (n_frames > 0 ? n_frames : 1) : kTagChunk));
    FrameParams* const fps = fps_store.data();
    const bool tags_possible = h->identity_cm && h->cfg.num_classes + 1 <= SMAP_TAG_MAX_PLANES &&
                               h->cells * (int64_t)(h->cfg.num_classes + 1) < ((int64_t)1 << 32);
    for (int begin = 0; begin < n_frames;) {
        // the first non-empty frame decides the kernel (empty frames carry no layout that matters)
        int first_ne = begin;
        while (first_ne < n_frames && frames[first_ne].n_points <= 0) ++first_ne;
        const bool f4_ahead = first_ne >= n_frames || frames[first_ne].layout == SMAP_PTS_F32X4;
        const int cap = (tags_possible && f4_ahead && acc_integer(h)) ? kTagChunk : kMaxBatch;
        const int chunk = (n_frames - begin < cap) ? n_frames - begin : cap;
        // ---- validate the whole chunk before anything is launched: a bad frame must not leave half a batch behind
        int first = -1;
        for (int i = 0; i < chunk; ++i) {
            int rc = fill_frame_params(h, frames + begin + i, fps[i]);
            if (rc) return rc;
            if (first < 0 && frames[begin + i].n_points > 0) first = i;
        }
        const bool f4 = first < 0 || frames[begin + first].layout == SMAP_PTS_F32X4;
        for (int i = 0; i < chunk; ++i) {
            const smap_frame* fr = frames + begin + i;
            if (fr->n_points == 0) continue;
            if (f4) {
                int rc = validate_fuse_frame(h, fr);
                if (rc) return rc;
            } else if (fr->layout != SMAP_PTS_F64_SOA) {
                return fail(SMAP_ERR_INVALID, "frames of one batch must share a point layout");
            }
        }
        // the count update on a grid of integer-valued counts may add its increments with float64 atomics (exact
        // in any order): de-duplicated with per-(cell, class) frame tags when there are few classes, through the
        // per-frame cell masks otherwise; everything else goes through the ordered apply
        const bool count_atomics = f4 && h->identity_cm && acc_integer(h) &&
                                   h->cells * (int64_t)(h->cfg.num_classes + 1) < ((int64_t)1 << 32);
        const int mode = !count_atomics ? 0 : (h->cfg.num_classes + 1 <= SMAP_TAG_MAX_PLANES ? 1 : 2);
        if (!h->identity_cm) acc_integer(h) = false;
        int rc = mode == 1 ? SMAP_OK : ensure_slots(h, chunk);
        if (rc) return rc;
        if (mode != 0 && ((h->applied_pending[0] && h->applied_writes_grid[0]) || (h->applied_pending[1] && h->applied_writes_grid[1]))) {
            // the count update's scatter kernels add to the grid themselves: not beside a k_apply that rewrites its rows
            rc = join_applies(h, st);
            if (rc) return rc;
        }
        if (mode != 1) {
            rc = begin_chunk(h, st);
            if (rc) { join_applies(h, st); return rc; }
        }
        // mask paths: the apply / clear of this chunk runs beside the scatter launches of the next one (second slot set)
        const bool overlap = SMAP_APPLY_OVERLAP && mode != 1 && begin + chunk < n_frames && !h->profiling;
        int used = 0;
        smap_handle::ProfRec* pr = nullptr;
        if (h->profiling) {
            if (h->n_prof_pending == 64) rc = harvest_profile(h);
            if (rc) return rc;
            pr = &h->prof_pending[h->n_prof_pending++];
            pr->frames = 0;
            for (int k = 0; k < 3; ++k) CK(cudaEventCreate(&pr->e[k]));
            CK(cudaEventRecord(pr->e[0], st));
        }
        rc = f4 ? launch_fuse(h, frames + begin, fps, chunk, mode, st, &used)
                : launch_stream(h, frames + begin, fps, chunk, st, &used);
        if (pr) { pr->frames = used; cudaEventRecord(pr->e[1], st); }
        // also after a failed launch: the frames queued so far are applied / their slots cleared, so that the slots are
        // all zero and the boxes reset when the call returns (the error is still reported)
        int rc2 = SMAP_OK;
        if (used > 0 && mode == 0) rc2 = launch_apply(h, h->acc, used, st, overlap && !rc);
        if (used > 0 && mode == 2) rc2 = launch_clear(h, used, st, overlap && !rc);
        h->last_update_counted = mode == 0;
        if (pr) cudaEventRecord(pr->e[2], st);
        for (int i = 0; i < chunk; ++i) {
            h->stats.frames += 1;
            h->stats.points += frames[begin + i].n_points;
        }
        acc_bound(h) += 3 * (int64_t)used;
        if (rc || rc2) {
            join_applies(h, st);
            return rc ? rc : rc2;
        }
        begin += chunk;
    }
    return join_applies(h, st);
}

int smap_integrate(smap_handle* h, const smap_frame* frame, void* stream) {
    if (!h || !frame) return fail(SMAP_ERR_INVALID, "NULL argument");
    return smap_integrate_batch(h, frame, 1, stream);
}

int smap_integrate_host(smap_handle* h, const smap_frame* f, void* stream) {
    if (!h || !f) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (f->n_points < 0 || f->image_width <= 0 || f->image_height <= 0) return fail(SMAP_ERR_INVALID, "bad frame sizes");
    if (f->layout != SMAP_PTS_F32X4 && f->layout != SMAP_PTS_F64_SOA) return fail(SMAP_ERR_INVALID, "unknown point layout");
    const size_t pts_bytes = f->layout == SMAP_PTS_F32X4 ? (size_t)f->n_points * 16 : (size_t)f->ld * 4 * sizeof(double);
    const size_t img_bytes = f->image_format == SMAP_IMG_CLASS_IDS ? (size_t)ids_w(f) * (size_t)ids_h(f)
                                                                   : (size_t)f->image_width * f->image_height * 3;
    if (f->n_points > 0 && (!f->points_dev || !f->image_dev)) return fail(SMAP_ERR_INVALID, "NULL points / image");
    const int s = h->stage_next;
    h->stage_next = (s + 1) % smap_handle::kStages;
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < smap_handle::kStages; ++i) {
            CK(cudaEventCreateWithFlags(&h->stage_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->stage_copied[i], cudaEventDisableTiming));
        }
    }
    // The copies run on an internal stream: frame i + 1 crosses PCIe while frame i's kernel runs on `stream`.  The
    // kernels of the frame that last used this stage must be finished before it is overwritten.
    CK(cudaStreamWaitEvent(h->copy_stream, h->stage_done[s], 0));
    if (pts_bytes > h->stage_pts_cap[s]) {
        CK(cudaEventSynchronize(h->stage_done[s]));
        cudaFree(h->stage_pts[s]);
        h->stage_pts[s] = nullptr; h->stage_pts_cap[s] = 0;
        CK(cudaMalloc(&h->stage_pts[s], pts_bytes + pts_bytes / 4));
        h->stage_pts_cap[s] = pts_bytes + pts_bytes / 4;
    }
    if (img_bytes > h->stage_img_cap[s]) {
        CK(cudaEventSynchronize(h->stage_done[s]));
        cudaFree(h->stage_img[s]);
        h->stage_img[s] = nullptr; h->stage_img_cap[s] = 0;
        CK(cudaMalloc(&h->stage_img[s], img_bytes));
        h->stage_img_cap[s] = img_bytes;
    }
    if (pts_bytes) CK(cudaMemcpyAsync(h->stage_pts[s], f->points_dev, pts_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->stage_img[s], f->image_dev, img_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->stage_copied[s], h->copy_stream));
    CK(cudaStreamWaitEvent(st, h->stage_copied[s], 0));
    smap_frame dev = *f;
    dev.points_dev = h->stage_pts[s];
    dev.image_dev = h->stage_img[s];
    int rc = smap_integrate(h, &dev, stream);
    // recorded even after a failed launch: the stage must become reusable
    CK(cudaEventRecord(h->stage_done[s], st));
    return rc;
}

// The reference's own cloud layout -> the float4 layout of the fast kernel, with the check that makes it legal:
// PointCloud2 fields are FLOAT32 (src/mapping.py:178-180 fills a float64 buffer from them), so the (4, N) float64
// rows of a recorded frame normally hold float32-representable values and the conversion loses nothing; *flag is
// raised (atomicOr 1) when some value is not, and the caller keeps the float64 path for that cloud.
__global__ void __launch_bounds__(kThreads)
k_soa_to_f32x4(const double* __restrict__ soa, int64_t ld, int64_t n, float4* __restrict__ out, int* __restrict__ flag) {
    bool bad = false;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double x = __ldcs(soa + k), y = __ldcs(soa + ld + k), z = __ldcs(soa + 2 * ld + k), w = __ldcs(soa + 3 * ld + k);
        const float4 p = make_float4((float)x, (float)y, (float)z, (float)w);
        // NaN stays NaN (dropped by either kernel, as by the reference); everything else must survive the round trip
        bad |= !(((double)p.x == x) | (x != x)) | !(((double)p.y == y) | (y != y)) | !(((double)p.z == z) | (z != z)) |
               !(((double)p.w == w) | (w != w));
        out[k] = p;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int smap_cloud_to_f32x4(const double* soa_dev, int64_t ld, int64_t n_points, void* out_f32x4_dev, int32_t* flag_dev,
                        int device, void* stream) {
    if (n_points < 0 || ld < n_points) return fail(SMAP_ERR_INVALID, "bad point count / stride");
    if (n_points > 0 && (!soa_dev || !out_f32x4_dev)) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (!flag_dev) return fail(SMAP_ERR_INVALID, "NULL flag");
    if (reinterpret_cast<uintptr_t>(out_f32x4_dev) & 15u) return fail(SMAP_ERR_INVALID, "float4 cloud must be 16-byte aligned");
    if (n_points == 0) return SMAP_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SMAP_ERR_CUDA, "cudaSetDevice failed");
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t grid = ceil_div(n_points, kThreads);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    k_soa_to_f32x4<<<(unsigned)grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        soa_dev, ld, n_points, static_cast<float4*>(out_f32x4_dev), flag_dev);
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_apply_filter(const double* src, int mh, int mw, int c, double* dst, int device, void* stream) {
    int rc = render_common_checks(src, mh, mw, c);
    if (rc) return rc;
    if (!dst || dst == src) return fail(SMAP_ERR_INVALID, "dst must be a distinct buffer");
    DeviceGuard guard(device);
    return launch_render<true>(src, mh, mw, c, nullptr, nullptr, dst, static_cast<cudaStream_t>(stream));
}

int smap_render(const double* map, int mh, int mw, int c, const uint8_t* colors_host, uint8_t* rgb, int device,
                void* stream) {
    int rc = render_common_checks(map, mh, mw, c);
    if (rc) return rc;
    if (!rgb || !colors_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    return launch_render<false>(map, mh, mw, c, colors_host, rgb, nullptr, static_cast<cudaStream_t>(stream));
}

int smap_filter_render(const double* map, int mh, int mw, int c, const uint8_t* colors_host, uint8_t* rgb,
                       double* filtered, int device, void* stream) {
    int rc = render_common_checks(map, mh, mw, c);
    if (rc) return rc;
    if (!rgb || !colors_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (filtered == map) return fail(SMAP_ERR_INVALID, "filtered must be a distinct buffer");
    DeviceGuard guard(device);
    return launch_render<true>(map, mh, mw, c, colors_host, rgb, filtered, static_cast<cudaStream_t>(stream));
}

int smap_render_thresholds(const double* map, int mh, int mw, int c, const uint8_t* colors_host,
                           const int32_t* priority_host, const double* thresholds_host, uint8_t* rgb, int device,
                           void* stream) {
    int rc = render_common_checks(map, mh, mw, c);
    if (rc) return rc;
    if (!rgb || !colors_host || !priority_host || !thresholds_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    RenderColors rcol;
    ThresholdParams tp;
    memset(&rcol, 0, sizeof rcol);
    memset(&tp, 0, sizeof tp);
    memcpy(rcol.rgb, colors_host, (size_t)c * 3);
    for (int i = 0; i < c; ++i) {
        if (priority_host[i] < 0 || priority_host[i] >= c) return fail(SMAP_ERR_INVALID, "priority entry out of range");
        tp.priority[i] = priority_host[i];
        tp.thresholds[i] = thresholds_host[i];
    }
    const int64_t cells = (int64_t)mh * mw;
    k_render_thresholds<<<(unsigned)ceil_div(cells, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        map, cells, c, rcol, tp, rgb);
    CK(cudaGetLastError());
    return SMAP_OK;
}

// map[map < 0] = 0 -- the one observable effect of the reference's planar update (src/mapping.py:481): its per-class
// increments never fire (the warped uint8 image is compared with label NAMES, :474), the clamp always runs.
__global__ void __launch_bounds__(kThreads) k_clamp_negative(double* __restrict__ map, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = map[i];
        if (v < 0.0) map[i] = 0.0;   // NaN and -0.0 stay, as numpy's boolean mask leaves them
    }
}

int smap_warp_perspective(const uint8_t* src_dev, int src_h, int src_w, int channels, const double h_host[9],
                          uint8_t* dst_dev, int dst_h, int dst_w, int device, void* stream) {
    if (!src_dev || !dst_dev || !h_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return fail(SMAP_ERR_INVALID, "empty image");
    if (channels < 1 || channels > 4) return fail(SMAP_ERR_INVALID, "1 to 4 channels");
    if (src_h >= 32768 || src_w >= 32768 || dst_h > 65535) return fail(SMAP_ERR_INVALID, "image too large (OpenCV keeps source coordinates in int16)");
    DeviceGuard guard(device);
    WarpParams p;
    // cv::invert of a 3 x 3 double matrix: the closed adjugate formula
    const double* s = h_host;
    double d = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
    if (d == 0.0) {
        for (int i = 0; i < 9; ++i) p.m[i] = 0.0;
    } else {
        d = 1.0 / d;
        p.m[0] = (s[4] * s[8] - s[5] * s[7]) * d;
        p.m[1] = (s[2] * s[7] - s[1] * s[8]) * d;
        p.m[2] = (s[1] * s[5] - s[2] * s[4]) * d;
        p.m[3] = (s[5] * s[6] - s[3] * s[8]) * d;
        p.m[4] = (s[0] * s[8] - s[2] * s[6]) * d;
        p.m[5] = (s[2] * s[3] - s[0] * s[5]) * d;
        p.m[6] = (s[3] * s[7] - s[4] * s[6]) * d;
        p.m[7] = (s[1] * s[6] - s[0] * s[7]) * d;
        p.m[8] = (s[0] * s[4] - s[1] * s[3]) * d;
    }
    p.src_h = src_h; p.src_w = src_w; p.cn = channels; p.dst_h = dst_h; p.dst_w = dst_w;
    // WarpPerspectiveInvoker's block width: BLOCK_SZ = 32, bh0 = min(16, h), bw0 = min(1024 / bh0, w)
    const int bh0 = dst_h < 16 ? dst_h : 16;
    p.bw0 = (1024 / bh0 < dst_w) ? 1024 / bh0 : dst_w;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)ceil_div(dst_w, 256), (unsigned)dst_h);
    switch (channels) {
        case 1: k_warp_perspective<1><<<grid, 256, 0, st>>>(p, src_dev, dst_dev); break;
        case 2: k_warp_perspective<2><<<grid, 256, 0, st>>>(p, src_dev, dst_dev); break;
        case 3: k_warp_perspective<3><<<grid, 256, 0, st>>>(p, src_dev, dst_dev); break;
        default: k_warp_perspective<4><<<grid, 256, 0, st>>>(p, src_dev, dst_dev); break;
    }
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_hull_components(const uint8_t* img_dev, int h, int w, int index, uint8_t* scratch_dev, int32_t* labels_dev,
                         int32_t* areas_dev, int device, void* stream) {
    if (!img_dev || !scratch_dev || !labels_dev || !areas_dev) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (h <= 0 || w <= 0) return fail(SMAP_ERR_INVALID, "empty image");
    if ((int64_t)h * w >= ((int64_t)1 << 31) || h > 65535) return fail(SMAP_ERR_INVALID, "image too large");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)h * w;
    const dim3 grid2((unsigned)ceil_div(w, 256), (unsigned)h);
    const unsigned grid1 = (unsigned)ceil_div(n, 256);
    k_hull_erode<<<grid2, 256, 0, st>>>(img_dev, h, w, index, scratch_dev);
    k_ccl_init<<<grid1, 256, 0, st>>>(scratch_dev, n, labels_dev, areas_dev);
    k_ccl_merge<<<grid2, 256, 0, st>>>(scratch_dev, h, w, labels_dev);
    k_ccl_flatten<<<grid1, 256, 0, st>>>(n, labels_dev, areas_dev);
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_hull_row_extremes(const int32_t* labels_dev, int h, int w, int root, int32_t* rowmin_dev, int32_t* rowmax_dev,
                           int device, void* stream) {
    if (!labels_dev || !rowmin_dev || !rowmax_dev) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (h <= 0 || w <= 0 || h > 65535) return fail(SMAP_ERR_INVALID, "bad image size");
    if (root < 0 || (int64_t)root >= (int64_t)h * w) return fail(SMAP_ERR_INVALID, "root outside the image");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    k_fill_i32<<<(unsigned)ceil_div(h, 256), 256, 0, st>>>(rowmin_dev, h, 0x7fffffff);
    k_fill_i32<<<(unsigned)ceil_div(h, 256), 256, 0, st>>>(rowmax_dev, h, -1);
    const dim3 grid2((unsigned)ceil_div(w, 256), (unsigned)h);
    k_hull_row_extremes<<<grid2, 256, 0, st>>>(labels_dev, h, w, root, rowmin_dev, rowmax_dev);
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_debug_bounds(unsigned long long out[2]) {
    if (!out) return fail(SMAP_ERR_INVALID, "NULL argument");
#ifdef SMAP_DEBUG_BOUNDS
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, g_bounds, sizeof(unsigned long long) * 2));
    return SMAP_OK;
#else
    out[0] = out[1] = 0ull;
    return fail(SMAP_ERR_STATE, "not a -DSMAP_DEBUG_BOUNDS build");
#endif
}

int smap_clamp_negative(double* map_dev, int64_t n_elements, int device, void* stream) {
    if (n_elements < 0 || (n_elements > 0 && !map_dev)) return fail(SMAP_ERR_INVALID, "bad grid");
    if (n_elements == 0) return SMAP_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SMAP_ERR_CUDA, "cudaSetDevice failed");
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t grid = ceil_div(n_elements, kThreads);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    k_clamp_negative<<<(unsigned)grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(map_dev, n_elements);
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_map_ptr(smap_handle* h, double** map_dev, int64_t* n_elements) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (map_dev) *map_dev = h->map;
    if (n_elements) *n_elements = h->cells * h->cfg.num_classes;
    return SMAP_OK;
}

int smap_clear(smap_handle* h, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    DeviceGuard guard(h->cfg.device);
    if (h->streaming) {
        // exchanges still in flight add to the grid on the communication stream: let them land first; increments
        // not yet handed to smap_exchange_async are dropped with the rest
        int rc = smap_exchange_flush(h, stream);
        if (rc) return rc;
        if (h->d_bound[h->cur] != 0 || !h->d_integer[h->cur]) {
            CK(cudaMemsetAsync(h->delta[h->cur], 0, sizeof(double) * (size_t)h->cells * h->cfg.num_classes, static_cast<cudaStream_t>(stream)));
            k_box_set<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->dbox[h->cur], 0x7fffffff, -1, 0x7fffffff, -1);
            h->d_bound[h->cur] = 0;
            h->d_integer[h->cur] = true;
        }
    }
    CK(cudaMemsetAsync(h->map, 0, sizeof(double) * (size_t)h->cells * h->cfg.num_classes, static_cast<cudaStream_t>(stream)));
    k_box_set<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->ubox, 0x7fffffff, -1, 0x7fffffff, -1);
    CK(cudaGetLastError());
    h->ubox_full = false;
    h->value_bound = 0;
    h->stats.frames = 0;
    h->stats.points = 0;
    h->integer_grid = true;
    h->last_stream = static_cast<cudaStream_t>(stream);
    return SMAP_OK;
}

int smap_set_profiling(smap_handle* h, int on) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    DeviceGuard guard(h->cfg.device);
    int rc = harvest_profile(h);
    if (rc) return rc;
    h->profiling = on != 0;
    if (on) { h->stats.profiled_frames = 0; h->stats.stream_kernel_ms = 0.0; h->stats.apply_kernel_ms = 0.0; }
    return SMAP_OK;
}

int smap_debug_set_frame_tag(smap_handle* h, uint32_t value) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (value < h->frame_tag) return fail(SMAP_ERR_INVALID, "the frame tag only moves forward");
    h->frame_tag = value;
    return SMAP_OK;
}

int smap_debug_fast32(const smap_config* cfg, const smap_frame* frame, const double P_host[12], double* out) {
    if (!cfg || !frame || !P_host || !out) return fail(SMAP_ERR_INVALID, "NULL argument");
    smap_handle* h = new (std::nothrow) smap_handle();
    if (!h) return fail(SMAP_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    memcpy(h->P[0], P_host, sizeof(double) * 12);
    h->cam_set[0] = true;
    smap_frame f = *frame;
    f.camera = 0;
    f.points_dev = nullptr;
    f.image_dev = nullptr;
    f.n_points = 0;
    f.layout = SMAP_PTS_F32X4;
    FrameParams fp;
    int rc = fill_frame_params(h, &f, fp);
    if (!rc) {
        Fast32 k;
        fill_fast32(h, &f, fp, k);
        int n = 0;
        auto put2 = [&](float2 v) { out[n++] = v.x; out[n++] = v.y; };
        for (int j = 0; j < 4; ++j) put2(k.c_dc[j]);
        for (int j = 0; j < 4; ++j) put2(k.c_ab[j]);
        put2(k.c_wh);
        out[n++] = k.c_bw; out[n++] = k.c_rh; out[n++] = k.c_rthr; out[n++] = k.c_depth;
        out[n++] = k.c_lo_u; out[n++] = k.c_hi_u; out[n++] = k.c_lo_v; out[n++] = k.c_hi_v;
        put2(k.n_ctr_xy);
        out[n++] = k.n_ctr_z; out[n++] = k.coord_l;
        for (int j = 0; j < 4; ++j) put2(k.d_uv[j]);
        for (int j = 0; j < 4; ++j) put2(k.d_cd[j]);
        out[n++] = k.g_k1; out[n++] = k.g_k0; out[n++] = k.g_hc;
        out[n++] = k.r_h; out[n++] = k.r_kd; out[n++] = k.r_thr0;
        put2(k.mid_uv); put2(k.half_uv); put2(k.cell_f0);
        out[n++] = k.cell_rf; out[n++] = k.cell_kc; out[n++] = k.cell_hg0;
        put2(k.mid_c); put2(k.half_c); put2(k.clamp_c);
        out[n++] = (double)k.pix_k; out[n++] = (double)k.cell_k;
    }
    delete h;
    return rc;
}

int smap_eval_counts(const uint8_t* rgb, int mh, int mw, const uint8_t* truth, int truth_rows, int truth_cols,
                     int shift_rows, int shift_cols, const uint8_t* mask, int mask_rows, int mask_cols, int64_t* counts,
                     int device, void* stream) {
    if (!rgb || !truth || !counts) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (mh <= 0 || mw <= 0) return fail(SMAP_ERR_INVALID, "empty map");
    if (shift_rows < 0 || shift_cols < 0 || (int64_t)shift_rows + mh > truth_rows || (int64_t)shift_cols + mw > truth_cols)
        return fail(SMAP_ERR_INVALID, "the shifted map does not lie inside the ground-truth label map");
    if (mask && (mask_rows < mh || mask_cols < mw)) return fail(SMAP_ERR_INVALID, "the mask is smaller than the map");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SMAP_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EvalParams p;
    p.mh = mh; p.mw = mw; p.truth_ld = truth_cols; p.shift_r = shift_rows; p.shift_c = shift_cols;
    p.mask_ld = mask ? mask_cols : 0;
    CK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * 12, st));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t grid = ceil_div((int64_t)mh * mw, kThreads);
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;   // grid-stride: 8 resident blocks of 256 threads per SM
    k_eval_counts<<<(unsigned)grid, kThreads, 0, st>>>(rgb, truth, mask, p, reinterpret_cast<unsigned long long*>(counts));
    CK(cudaGetLastError());
    return SMAP_OK;
}

int smap_debug_nearest_map(int dst, int src, uint16_t* tab_out, uint32_t* mul, uint32_t* shift) {
    if (!tab_out || !mul || !shift) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (dst <= 0 || src <= 0 || dst > 65535 || src > 65535) return fail(SMAP_ERR_INVALID, "sizes must be in 1..65535");
    nearest_table(dst, src, tab_out);
    *mul = 0; *shift = 0;
    if (src <= dst) nearest_mulshift(dst, src, tab_out, mul, shift);
    return SMAP_OK;
}

int smap_notify_map_modified(smap_handle* h) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    h->integer_grid = false;
    h->ubox_full = true;
    return SMAP_OK;
}

int smap_download(smap_handle* h, double* map_host) {
    if (!h || !map_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(map_host, h->map, sizeof(double) * (size_t)h->cells * h->cfg.num_classes, cudaMemcpyDeviceToHost));
    return SMAP_OK;
}

int smap_upload(smap_handle* h, const double* map_host) {
    if (!h || !map_host) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h->map, map_host, sizeof(double) * (size_t)h->cells * h->cfg.num_classes, cudaMemcpyHostToDevice));
    h->integer_grid = false;
    h->ubox_full = true;
    return SMAP_OK;
}


// ---- multi-GPU exchange --------------------------------------------------------------------------------------------
int smap_comm_unique_id(uint8_t id_out[SMAP_COMM_ID_BYTES]) {
    if (!id_out) return fail(SMAP_ERR_INVALID, "NULL argument");
    static_assert(SMAP_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "smap.h and nccl.h disagree about the id size");
    NcclApi& N = nccl_api();
    if (!N.load()) return fail(SMAP_ERR_COMM, "libnccl.so.2 not found (set SMAP_NCCL_LIB)");
    ncclUniqueId id;
    NCCLCK(N.GetUniqueId(&id));
    memcpy(id_out, id.internal, SMAP_COMM_ID_BYTES);
    return SMAP_OK;
}

static int comm_scratch(smap_handle* h) {
    if (!h->xch_dev) CK(cudaMalloc(&h->xch_dev, sizeof(int) * kCommWords));
    if (!h->xch_host) CK(cudaHostAlloc(&h->xch_host, sizeof(int) * kCommWords, cudaHostAllocDefault));
    return SMAP_OK;
}

int smap_comm_init(smap_handle* h, int n_ranks, int rank, const uint8_t id[SMAP_COMM_ID_BYTES]) {
    if (!h || !id) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SMAP_ERR_INVALID, "bad rank / n_ranks");
    if (h->comm) return fail(SMAP_ERR_STATE, "the handle already has a communicator (smap_comm_destroy first)");
    NcclApi& N = nccl_api();
    if (!N.load()) return fail(SMAP_ERR_COMM, "libnccl.so.2 not found (set SMAP_NCCL_LIB)");
    DeviceGuard guard(h->cfg.device);
    int rc = comm_scratch(h);
    if (rc) return rc;
    ncclUniqueId uid;
    memcpy(uid.internal, id, SMAP_COMM_ID_BYTES);
    NCCLCK(N.CommInitRank(&h->comm, n_ranks, uid, rank));
    h->own_comm = true;
    h->n_ranks = n_ranks;
    h->rank = rank;
    return SMAP_OK;
}

int smap_comm_attach(smap_handle* h, void* nccl_comm) {
    if (!h || !nccl_comm) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (h->comm) return fail(SMAP_ERR_STATE, "the handle already has a communicator (smap_comm_destroy first)");
    NcclApi& N = nccl_api();
    if (!N.load()) return fail(SMAP_ERR_COMM, "libnccl.so.2 not found (set SMAP_NCCL_LIB)");
    DeviceGuard guard(h->cfg.device);
    int rc = comm_scratch(h);
    if (rc) return rc;
    ncclComm_t c = static_cast<ncclComm_t>(nccl_comm);
    int n = 0, r = 0;
    NCCLCK(N.CommCount(c, &n));
    NCCLCK(N.CommUserRank(c, &r));
    h->comm = c;
    h->own_comm = false;
    h->n_ranks = n;
    h->rank = r;
    return SMAP_OK;
}

int smap_comm_destroy(smap_handle* h) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (h->comm && h->own_comm) {
        DeviceGuard guard(h->cfg.device);
        cudaDeviceSynchronize();
        NCCLCK(nccl_api().CommDestroy(h->comm));
    }
    h->comm = nullptr;
    h->own_comm = false;
    h->n_ranks = 1;
    h->rank = 0;
    return SMAP_OK;
}

int smap_comm_get_info(smap_handle* h, smap_comm_info* out) {
    if (!h || !out) return fail(SMAP_ERR_INVALID, "NULL argument");
    *out = h->comm_last;
    out->host_wait_ms = h->host_wait_ms;
    out->pack_ms = out->reduce_ms = out->add_ms = 0.0;
    if (h->phase_recorded && cudaEventSynchronize(h->ev_phase[3]) == cudaSuccess) {
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&a, h->ev_phase[0], h->ev_phase[1]);
        cudaEventElapsedTime(&b, h->ev_phase[1], h->ev_phase[2]);
        cudaEventElapsedTime(&c, h->ev_phase[2], h->ev_phase[3]);
        out->pack_ms = a; out->reduce_ms = b; out->add_ms = c;
    }
    out->n_ranks = h->n_ranks;
    out->rank = h->rank;
    out->grid_bytes = (int64_t)sizeof(double) * h->cells * h->cfg.num_classes;
    return SMAP_OK;
}

int smap_allreduce(smap_handle* h, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (!h->comm) return fail(SMAP_ERR_STATE, "no communicator (smap_comm_init / smap_comm_attach)");
    if (h->streaming) return fail(SMAP_ERR_STATE, "streaming exchange active: smap_comm_streaming(h, 0) first");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    h->last_stream = st;
    NcclApi& N = nccl_api();
    int ag[kCommWords];
    int rc = comm_agree(h, st, ag);
    if (rc) return rc;
    const int x0 = -ag[0], x1 = ag[1], y0 = -ag[2], y1 = ag[3];
    int64_t bound_total = 0;
    const int pack = pick_pack(h, ag, &bound_total);
    h->comm_last.window[0] = x0; h->comm_last.window[1] = x1; h->comm_last.window[2] = y0; h->comm_last.window[3] = y1;
    h->comm_last.pack = pack;
    h->comm_last.bytes = 0;
    h->comm_last.exchanges += 1;
    if (ag[5]) h->integer_grid = false;          // some rank holds non-integer values: so will every rank's sum
    h->value_bound = bound_total;
    if (x1 < x0 || y1 < y0) return SMAP_OK;      // nobody touched anything
    const int rows = x1 - x0 + 1, cols = y1 - y0 + 1, c = h->cfg.num_classes;
    const int64_t run = (int64_t)cols * c;
    if (run >= ((int64_t)1 << 31)) return fail(SMAP_ERR_INVALID, "grid row too long for the packed exchange");
    const int64_t run_words = pack == kPackU16 ? (run + 1) / 2 : run;
    const size_t count = (size_t)rows * (size_t)(pack == kPackF64 ? run : run_words);
    const size_t bytes = count * (pack == kPackF64 ? 8 : 4);
    rc = ensure_bytes(&h->xbuf, &h->xbuf_cap, bytes);
    if (rc) return rc;
    rc = launch_pack(h, h->map, false, pack, x0, rows, x0, x1, y0, cols, h->xbuf, st);
    if (rc) return rc;
    NCCLCK(N.AllReduce(h->xbuf, h->xbuf, count, pack == kPackF64 ? ncclFloat64 : ncclUint32, ncclSum, h->comm, st));
    rc = launch_unpack(h, pack, h->map, h->cfg.map_width, x0, rows, y0, cols, h->xbuf, 0, st);
    if (rc) return rc;
    // every rank now holds the total inside the window, and only there
    k_box_set<<<1, 32, 0, st>>>(h->ubox, x0, x1, y0, y1);
    CK(cudaGetLastError());
    h->ubox_full = false;
    h->comm_last.bytes = (int64_t)bytes;
    h->stats.kernel_launches += 3;
    return SMAP_OK;
}

int smap_reduce_scatter_rows(smap_handle* h, double* tile_dev, int64_t tile_rows_cap, int32_t* r0_out, int32_t* r1_out,
                             int32_t* top_out, int32_t* bottom_out, void* stream) {
    if (!h || !tile_dev || !r0_out || !r1_out || !top_out || !bottom_out) return fail(SMAP_ERR_INVALID, "NULL argument");
    if (!h->comm) return fail(SMAP_ERR_STATE, "no communicator (smap_comm_init / smap_comm_attach)");
    if (h->streaming)   // every rank's grid already holds the global sum: take the rows from it
        return fail(SMAP_ERR_STATE, "streaming exchange active: smap_comm_streaming(h, 0) first");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    h->last_stream = st;
    NcclApi& N = nccl_api();
    const int mh = h->cfg.map_height, mw = h->cfg.map_width, c = h->cfg.num_classes, n = h->n_ranks;
    const int per = (mh + n - 1) / n;
    const int r0 = h->rank * per < mh ? h->rank * per : mh;
    const int r1 = (h->rank + 1) * per < mh ? (h->rank + 1) * per : mh;
    const int top = (r0 > 0 && r1 > r0) ? 1 : 0, bottom = (r1 < mh && r1 > r0) ? 1 : 0;
    if (tile_rows_cap < (int64_t)(r1 - r0) + top + bottom) return fail(SMAP_ERR_INVALID, "tile too small (per + 2 rows are enough)");
    *r0_out = r0; *r1_out = r1; *top_out = top; *bottom_out = bottom;
    int ag[kCommWords];
    int rc = comm_agree(h, st, ag);
    if (rc) return rc;
    const int x0 = -ag[0], x1 = ag[1], y0 = -ag[2], y1 = ag[3];
    int64_t bound_total = 0;
    const int pack = pick_pack(h, ag, &bound_total);
    h->comm_last.window[0] = x0; h->comm_last.window[1] = x1; h->comm_last.window[2] = y0; h->comm_last.window[3] = y1;
    h->comm_last.pack = pack;
    h->comm_last.bytes = 0;
    h->comm_last.exchanges += 1;
    const int tile_rows = r1 - r0 + top + bottom;
    if (tile_rows > 0) CK(cudaMemsetAsync(tile_dev, 0, sizeof(double) * (size_t)tile_rows * mw * c, st));
    if (x1 < x0 || y1 < y0) return SMAP_OK;
    const int cols = y1 - y0 + 1;
    const int64_t run = (int64_t)cols * c;
    if (run >= ((int64_t)1 << 31)) return fail(SMAP_ERR_INVALID, "grid row too long for the packed exchange");
    const int64_t run_words = pack == kPackU16 ? (run + 1) / 2 : run;
    const size_t row_count = (size_t)(pack == kPackF64 ? run : run_words);
    const size_t esize = pack == kPackF64 ? 8 : 4;
    const ncclDataType_t dt = pack == kPackF64 ? ncclFloat64 : ncclUint32;
    // send buffer: per * n rows (the grid padded with zero rows), window columns; afterwards reused for the halo rows
    size_t need = (size_t)per * n * row_count * esize;
    if (need < (size_t)2 * n * row_count * esize) need = (size_t)2 * n * row_count * esize;
    rc = ensure_bytes(&h->xbuf, &h->xbuf_cap, need);
    if (rc) return rc;
    rc = ensure_bytes(&h->xbuf2, &h->xbuf2_cap, (size_t)(per + 2) * row_count * esize);
    if (rc) return rc;
    rc = launch_pack(h, h->map, false, pack, 0, per * n, x0, x1, y0, cols, h->xbuf, st);
    if (rc) return rc;
    NCCLCK(N.ReduceScatter(h->xbuf, h->xbuf2, (size_t)per * row_count, dt, ncclSum, h->comm, st));
    // halo exchange: everybody publishes the first and the last row of its tile (rows of an empty tile: whatever)
    char* own = static_cast<char*>(h->xbuf2);
    char* edge = own + (size_t)per * row_count * esize;
    const int last = r1 > r0 ? r1 - r0 - 1 : 0;
    CK(cudaMemcpyAsync(edge, own, row_count * esize, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(edge + row_count * esize, own + (size_t)last * row_count * esize, row_count * esize, cudaMemcpyDeviceToDevice, st));
    NCCLCK(N.AllGather(edge, h->xbuf, 2 * row_count, dt, h->comm, st));
    if (top) {   // last row of the tile above
        rc = launch_unpack(h, pack, tile_dev, mw, 0, 1, y0, cols, h->xbuf, 2 * ((r0 - 1) / per) + 1, st);
        if (rc) return rc;
    }
    if (r1 > r0) {
        rc = launch_unpack(h, pack, tile_dev, mw, top, r1 - r0, y0, cols, h->xbuf2, 0, st);
        if (rc) return rc;
    }
    if (bottom) {   // first row of the tile below
        rc = launch_unpack(h, pack, tile_dev, mw, top + (r1 - r0), 1, y0, cols, h->xbuf, 2 * (r1 / per), st);
        if (rc) return rc;
    }
    h->comm_last.bytes = (int64_t)((size_t)per * n * row_count * esize);
    h->stats.kernel_launches += 4;
    return SMAP_OK;
}

// ---- streaming exchange: the ranks' increments are summed into every rank's grid while integration goes on ---------
namespace {

// Data phase of the exchange whose agreement is in flight: pack + zero the buffer's window, all-reduce, add to the
// grid -- all on the internal communication stream.  Blocks the HOST until the agreement has arrived (it was queued
// one call earlier, behind a chunk that has normally long finished), never the caller's stream.
int finish_pending(smap_handle* h) {
    if (!h->pend_active) return SMAP_OK;
    NcclApi& N = nccl_api();
    const int b = h->pend_buf;
    cudaStream_t cs = h->comm_stream;
    {
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        CK(cudaEventSynchronize(h->ev_agreed[b]));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        h->host_wait_ms = 1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    h->pend_active = false;
    const int* ag = h->xch_host2[b];
    const int x0 = -ag[0], x1 = ag[1], y0 = -ag[2], y1 = ag[3];
    int64_t bound_total = 0;
    const int pack = pick_pack(h, ag, &bound_total);
    h->comm_last.window[0] = x0; h->comm_last.window[1] = x1; h->comm_last.window[2] = y0; h->comm_last.window[3] = y1;
    h->comm_last.pack = pack;
    h->comm_last.bytes = 0;
    h->comm_last.exchanges += 1;
    if (ag[5]) h->integer_grid = false;
    h->value_bound += bound_total;
    h->d_integer[b] = true;      // zeroed below
    h->d_bound[b] = 0;
    if (x1 >= x0 && y1 >= y0) {
        const int rows = x1 - x0 + 1, cols = y1 - y0 + 1, c = h->cfg.num_classes;
        const int64_t run = (int64_t)cols * c;
        if (run >= ((int64_t)1 << 31)) return fail(SMAP_ERR_INVALID, "grid row too long for the packed exchange");
        const int64_t run_words = pack == kPackU16 ? (run + 1) / 2 : run;
        const size_t count = (size_t)rows * (size_t)(pack == kPackF64 ? run : run_words);
        const size_t bytes = count * (pack == kPackF64 ? 8 : 4);
        int rc = ensure_bytes(&h->xbuf, &h->xbuf_cap, bytes);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_phase[0], cs));
        rc = launch_pack(h, h->delta[b], true, pack, x0, rows, x0, x1, y0, cols, h->xbuf, cs);
        if (rc) return rc;
        k_box_set<<<1, 32, 0, cs>>>(h->dbox[b], 0x7fffffff, -1, 0x7fffffff, -1);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev_packed[b], cs));
        h->packed_recorded[b] = true;
        CK(cudaEventRecord(h->ev_phase[1], cs));
        NCCLCK(N.AllReduce(h->xbuf, h->xbuf, count, pack == kPackF64 ? ncclFloat64 : ncclUint32, ncclSum, h->comm, cs));
        CK(cudaEventRecord(h->ev_phase[2], cs));
        rc = launch_unpack(h, pack, h->map, h->cfg.map_width, x0, rows, y0, cols, h->xbuf, 0, cs, true);
        if (rc) return rc;
        k_box_fold<<<1, 32, 0, cs>>>(h->ubox, x0, x1, y0, y1);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev_phase[3], cs));
        h->phase_recorded = true;
        h->comm_last.bytes = (int64_t)bytes;
        h->stats.kernel_launches += 4;
    } else {
        CK(cudaEventRecord(h->ev_packed[b], cs));
        h->packed_recorded[b] = true;
    }
    CK(cudaEventRecord(h->ev_exchanged, cs));
    return SMAP_OK;
}

}  // namespace

int smap_comm_streaming(smap_handle* h, int on, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!on) {
        if (!h->streaming) return SMAP_OK;
        int rc = smap_exchange_flush(h, stream);
        if (rc) return rc;
        if (h->d_bound[h->cur] != 0) return fail(SMAP_ERR_STATE, "frames integrated since the last smap_exchange_async would be lost");
        h->streaming = false;
        h->acc = h->map;
        h->abox = h->ubox;
        return SMAP_OK;
    }
    if (h->streaming) return SMAP_OK;
    if (!h->comm) return fail(SMAP_ERR_STATE, "no communicator (smap_comm_init / smap_comm_attach)");
    const size_t map_bytes = sizeof(double) * (size_t)h->cells * h->cfg.num_classes;
    if (!h->delta[0]) {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, greatest));
        CK(cudaEventCreateWithFlags(&h->ev_chunk, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_exchanged, cudaEventDisableTiming));
        for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&h->ev_phase[i]));
        for (int b = 0; b < 2; ++b) {
            CK(cudaEventCreateWithFlags(&h->ev_agreed[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_packed[b], cudaEventDisableTiming));
            CK(cudaMalloc(&h->xch_dev2[b], sizeof(int) * kCommWords));
            CK(cudaHostAlloc(&h->xch_host2[b], sizeof(int) * kCommWords, cudaHostAllocDefault));
            CK(cudaMalloc(&h->delta[b], map_bytes));
            CK(cudaMemset(h->delta[b], 0, map_bytes));   // device-synchronous: done before anything queued later
        }
    }
    // the buffers are zero and their boxes empty between streaming sessions (every exchange leaves them so)
    h->cur = 0;
    h->d_integer[0] = h->d_integer[1] = true;
    h->d_bound[0] = h->d_bound[1] = 0;
    h->packed_recorded[0] = h->packed_recorded[1] = false;
    h->streaming = true;
    h->acc = h->delta[0];
    h->abox = h->dbox[0];
    h->last_stream = st;
    return SMAP_OK;
}

int smap_exchange_async(smap_handle* h, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (!h->streaming) return fail(SMAP_ERR_STATE, "smap_comm_streaming(h, 1) first");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream), cs = h->comm_stream;
    h->last_stream = st;
    NcclApi& N = nccl_api();
    // 1. the previous exchange's data phase (its agreement has normally arrived long ago)
    int rc = finish_pending(h);
    if (rc) return rc;
    // 2. this exchange's agreement, behind the chunk the caller has queued so far
    const int b = h->cur;
    CK(cudaEventRecord(h->ev_chunk, st));
    CK(cudaStreamWaitEvent(cs, h->ev_chunk, 0));
    const int vb = h->d_bound[b] > 0x7fffffffll ? 0x7fffffff : (int)h->d_bound[b];
    k_comm_prep<<<1, 32, 0, cs>>>(h->dbox[b], 0, h->cfg.map_height, h->cfg.map_width, vb, h->d_integer[b] ? 0 : 1, h->xch_dev2[b]);
    CK(cudaGetLastError());
    NCCLCK(N.AllReduce(h->xch_dev2[b], h->xch_dev2[b], kCommWords, ncclInt32, ncclMax, h->comm, cs));
    CK(cudaMemcpyAsync(h->xch_host2[b], h->xch_dev2[b], sizeof(int) * kCommWords, cudaMemcpyDeviceToHost, cs));
    CK(cudaEventRecord(h->ev_agreed[b], cs));
    h->pend_active = true;
    h->pend_buf = b;
    // 3. the next frames go to the other buffer -- once its previous content has been packed and zeroed
    h->cur ^= 1;
    h->acc = h->delta[h->cur];
    h->abox = h->dbox[h->cur];
    if (h->packed_recorded[h->cur]) CK(cudaStreamWaitEvent(st, h->ev_packed[h->cur], 0));
    return SMAP_OK;
}

int smap_exchange_flush(smap_handle* h, void* stream) {
    if (!h) return fail(SMAP_ERR_INVALID, "NULL handle");
    if (!h->streaming) return SMAP_OK;
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    h->last_stream = st;
    const bool had = h->pend_active;
    int rc = finish_pending(h);
    if (rc) return rc;
    if (had || h->comm_last.exchanges > 0) CK(cudaStreamWaitEvent(st, h->ev_exchanged, 0));
    return SMAP_OK;
}

#ifdef SMAP_FUSE_STATS
// diagnostic builds only: read and reset the routing counters of k_fuse
__attribute__((visibility("default"))) int smap_debug_fuse_stats(unsigned long long out[4]) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, g_fuse_stats, sizeof(unsigned long long) * 4));
    unsigned long long zero[4] = {0, 0, 0, 0};
    CK(cudaMemcpyToSymbol(g_fuse_stats, zero, sizeof zero));
    return SMAP_OK;
}
#endif

int smap_get_stats(smap_handle* h, smap_stats* out) {
    if (!h || !out) return fail(SMAP_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->cfg.device);
    CK(cudaStreamSynchronize(h->last_stream));
    {
        int rc = harvest_profile(h);
        if (rc) return rc;
    }
    // after launch_apply flipped the parity, the finished launch's counter sits in touched[parity ^ 1]
    unsigned long long k = 0;
    CK(cudaMemcpy(&k, h->touched + (h->parity ^ 1), sizeof k, cudaMemcpyDeviceToHost));
    h->stats.touched_cells = h->last_update_counted ? (int64_t)k : -1;
    *out = h->stats;
    return SMAP_OK;
}

}  // extern "C"
