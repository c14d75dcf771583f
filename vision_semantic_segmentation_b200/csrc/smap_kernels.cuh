// Kernels of the semantic-mapping path (sm_100a).  See DESIGN.md for the data layout and the
// roofline of each kernel.  None of this is GEMM-shaped: the work is HBM-bound streaming of the cloud,
// a sector-granular gather from the label image and a scatter of per-cell class masks, so the design
// rules that matter are coalescing (float4 / contiguous rows), keeping atomics off single hot
// addresses, and grid sizes that fill 148 SMs.
#pragma once
#include "smap_device.cuh"
#include "smap_render.cuh"

namespace smap {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// Per-frame cell masks.
// One 32-bit word per BEV cell and frame slot: bit i = "class i observed in this cell by this frame", bit C =
// "lane point with a strong/weak LiDAR return" (the intensity boost).  OR-ing bits is the per-frame
// (cell, class) de-duplication of the reference's fancy-index "+=" (src/mapping_replay.py:281,294).  The
// scatter only issues fire-and-forget RED.OR (no returned atomics); k_apply reads the words inside the
// frame's bounding box, adds the update-matrix columns in frame and class order, and writes the words back
// to zero, so a slot is always clean between launches.  (The count update on float4 clouds does not use the
// masks at all: k_fuse MODE 1, smap_fuse.cuh.)
// ------------------------------------------------------------------------------------------------

// Bounding box (in cells, inclusive) of everything a frame touched; written by the scatter kernels.
struct FrameBox {
    int x0, x1, y0, y1;   // empty when x1 < x0
};

__device__ __forceinline__ void box_reset(int* b) { b[0] = 0x7fffffff; b[1] = -1; b[2] = 0x7fffffff; b[3] = -1; }

// One frame = one launch: every per-frame constant is then a kernel parameter at a fixed offset, i.e. a
// constant-bank operand of the FP64 instructions (no register, no load).  A batched launch indexed by blockIdx.y
// was tried: the run-time frame index turned every constant into a register-indexed LDC.
struct StreamParams {
    FrameParams fp;
    const void* pts;
    const uint8_t* image;
    uint32_t* mask;       // this frame's mask slot (MH*MW words)
    int64_t n;
    int64_t ld;
};

// tuning knobs (overridable at compile time for the variant sweeps recorded in profiles/)
#ifndef SMAP_STREAM_ROUND
#define SMAP_STREAM_ROUND 4     // 32-point chunks per round
#endif
#ifndef SMAP_STREAM_MINB
#define SMAP_STREAM_MINB 3      // resident blocks per SM the register allocation aims for
#endif
constexpr int kWarps = kThreads / 32;
constexpr int kRound = SMAP_STREAM_ROUND;
constexpr int kRoundPts = 32 * kRound;                // points a warp handles per round
constexpr int kBlockRoundPts = kWarps * kRoundPts;    // points a block handles per round
constexpr int kQueueCap = kRoundPts + 32;             // per-warp survivor stack: < 32 left over + one round

// The reference's own rounding chain for one point; only reached when the certified fast path cannot decide.
// Everything by value (a pointer to a caller's local would force it onto the stack in the hot loop).
__device__ __noinline__ int exact_project_slow(const FrameParams* f, double x, double y, double z) {
    int iu, iv;
    if (!project_point(*f, x, y, z, iu, iv)) return kDrop;
    return (iv << 16) | iu;
}

__device__ __noinline__ long long exact_cell_slow(const GridParams* g, double x, double y) {
    int cx, cy;
    if (!cell_xy(*g, x, y, cx, cy)) return -1;
    return ((long long)cx << 32) | (unsigned int)cy;
}

// ------------------------------------------------------------------------------------------------
// k_stream_soa: the fused path for the reference's own cloud layout, (4, N) float64 rows
// (project -> cull -> label lookup -> class bits -> cell -> mask scatter; src/mapping_replay.py:223-244,
// :261-277, :288-290).  float4 clouds -- the fast path -- go through k_fuse (smap_fuse.cuh).
//
// Persistent and warp-autonomous: the blocks walk the cloud in rounds of kBlockRoundPts points; inside a round
// every warp owns a contiguous slice and never synchronises with the other warps:
//   stream   conservative float32 cull (precull_pass) on the coordinates rounded to float; the indices of the
//            survivors (~35 %) go on the warp's private stack in shared memory (ballot + popc, no atomics);
//   drain    whenever >= 32 survivors are stacked, pop 32 - one per lane, all lanes busy: float64 certified fast
//            projection (exact fallback), label gather, class bits from the shared colour tables, certified
//            fast cell index, one RED.OR into the frame's mask slot; lanes track the bounding box.
// k_apply then replays the masks of the batch in frame and class order (any update matrix, any grid).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, SMAP_STREAM_MINB)
k_stream_soa(const __grid_constant__ StreamParams F, const __grid_constant__ GridParams gp, FrameBox* __restrict__ box) {
    __shared__ uint32_t s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_tab_r[256], s_tab_g[256];
    __shared__ int s_box[4];

    const FrameParams& fp = F.fp;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* const queue = s_queue[warp];

    build_color_tables(gp, s_tab_r, s_tab_g);
    if (threadIdx.x == 0) box_reset(s_box);
    __syncthreads();

    const double* const pd = reinterpret_cast<const double*>(F.pts);
    const uint8_t* const f_image = F.image;
    uint32_t* const f_mask = F.mask;
    const int64_t f_n = F.n, f_ld = F.ld;

    uint32_t qn = 0;       // entries on this warp's stack (warp-uniform)
    int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1;

    // ---- one batch of <= 32 stacked survivors, one per lane
    auto drain_batch = [&](uint32_t count) {
        const uint32_t first = qn - count;
        qn = first;
        if ((uint32_t)lane >= count) return;
        const int64_t k = (int64_t)queue[first + lane];
        const double x = __ldg(pd + k), y = __ldg(pd + f_ld + k), z = __ldg(pd + 2 * f_ld + k);
        const bool coords_ok = fmax(fmax(fabs(x), fabs(y)), fabs(z)) < kCoordBound;
        // the boost test compares the float64 intensity with 2 and 14; rounding to float32 could move a
        // value across them, so map the double onto a float on the same side (NaN: neither)
        const double itd = __ldg(pd + 3 * f_ld + k);
        const float it = (itd < 2.0) ? 0.0f : ((itd > 14.0) ? 15.0f : 8.0f);
        int pix = fast_project(fp, x, y, z, coords_ok);
        if (pix == kAsk) pix = exact_project_slow(&fp, x, y, z);
        if (pix < 0) return;
        const size_t off = 3u * ((size_t)(pix >> 16) * (size_t)fp.img_w + (size_t)(pix & 0xffff));
        const uint32_t r = __ldg(f_image + off), g = __ldg(f_image + off + 1);
        const uint32_t bits = class_bits_lut(gp, s_tab_r, s_tab_g, (uint8_t)r, (uint8_t)g, it);
        if (!bits) return;
        int cx = 0, cy = 0;
        const int on = fast_cell(gp, x, y, cx, cy);
        if (on == kAsk) {
            const long long c2 = exact_cell_slow(&gp, x, y);
            if (c2 < 0) return;
            cx = (int)(c2 >> 32); cy = (int)(c2 & 0xffffffffll);
        } else if (on == kDrop) {
            return;
        }
        const uint32_t cell = (uint32_t)cx * (uint32_t)gp.mw + (uint32_t)cy;
        atomicOr(f_mask + cell, bits);   // result unused: RED.OR, nothing waits for it
        bx0 = min(bx0, cx); bx1 = max(bx1, cx); by0 = min(by0, cy); by1 = max(by1, cy);
    };

    const int64_t stride = (int64_t)gridDim.x * kBlockRoundPts;
    for (int64_t rbase = (int64_t)blockIdx.x * kBlockRoundPts + (int64_t)warp * kRoundPts; rbase < f_n; rbase += stride) {
#pragma unroll 2
        for (int j = 0; j < kRound; ++j) {
            const int64_t k = rbase + j * 32 + lane;
            bool pass = false;
            if (k < f_n) {
                // a coordinate that is not float32-representable only moves by one rounding: covered by kCullSlack
                pass = precull_pass(fp, (float)__ldg(pd + k), (float)__ldg(pd + f_ld + k), (float)__ldg(pd + 2 * f_ld + k));
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (pass) queue[qn + __popc(ballot & lt_mask)] = (uint32_t)k;
            qn += __popc(ballot);
        }
        __syncwarp();
        while (qn >= 32u) {
            drain_batch(32u);
            __syncwarp();
        }
    }
    while (qn) {
        drain_batch(qn < 32u ? qn : 32u);
        __syncwarp();
    }
    // fold the lanes' bounding boxes: warp -> block (shared atomics) -> frame (4 global atomics per block)
    if (__any_sync(0xffffffffu, bx1 >= bx0)) {
        const int a = __reduce_min_sync(0xffffffffu, bx0), b = __reduce_max_sync(0xffffffffu, bx1);
        const int c = __reduce_min_sync(0xffffffffu, by0), d = __reduce_max_sync(0xffffffffu, by1);
        if (lane == 0) {
            atomicMin(&s_box[0], a); atomicMax(&s_box[1], b);
            atomicMin(&s_box[2], c); atomicMax(&s_box[3], d);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_box[1] >= s_box[0]) {
        atomicMin(&box->x0, s_box[0]); atomicMax(&box->x1, s_box[1]);
        atomicMin(&box->y0, s_box[2]); atomicMax(&box->y1, s_box[3]);
    }
}

// Parity kernel for update_map (src/mapping_replay.py:261-277,:288-290): the scatter half from an already
// projected cloud (4, M) float64 + its (3, M) RGB labels, exact arithmetic throughout.
__global__ void __launch_bounds__(kThreads)
k_update_scatter(const double* __restrict__ pcd, int64_t ld, const uint8_t* __restrict__ label, int64_t ldl,
                 int64_t m, const __grid_constant__ GridParams gp, uint32_t* __restrict__ mask,
                 FrameBox* __restrict__ box) {
    __shared__ int s_box[4];
    if (threadIdx.x == 0) box_reset(s_box);
    __syncthreads();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t bits = class_bits(gp, label[k], label[ldl + k], pcd[3 * ld + k]);
        int cx, cy;
        if (bits && cell_xy(gp, pcd[k], pcd[ld + k], cx, cy)) {
            atomicOr(mask + (uint32_t)cx * (uint32_t)gp.mw + (uint32_t)cy, bits);
            atomicMin(&s_box[0], cx); atomicMax(&s_box[1], cx);
            atomicMin(&s_box[2], cy); atomicMax(&s_box[3], cy);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_box[1] >= s_box[0]) {
        atomicMin(&box->x0, s_box[0]); atomicMax(&box->x1, s_box[1]);
        atomicMin(&box->y0, s_box[2]); atomicMax(&box->y1, s_box[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// K3b: ordered apply for up to kMaxBatch frame slots in one pass.
// The union of the frames' bounding boxes is walked two horizontally adjacent cells per thread, so that a
// thread reads the two mask words of a slot with ONE 8-byte load and only from the slots whose row span covers
// the cell (a slot is all zero outside its own frame's box).  For each cell touched by any frame the C-element
// grid row is loaded once, the frames are replayed IN ORDER -- for each frame the classes in ascending order,
// "row += CM[:, i]" and the lane boost, exactly the statements at src/mapping_replay.py:281 and :294 -- the row
// is stored once and the words are zeroed.  Bit-exact for any update matrix; the grid traffic is shared by
// all frames of the batch.  NJ = ceil(C / 8).
// ------------------------------------------------------------------------------------------------
struct ApplyParams {
    int n_frames;
    int pad;
    uint32_t* mask[kMaxBatch];
};

template <int V> struct MaskVec;
template <> struct MaskVec<4> { typedef uint4 type; };
template <> struct MaskVec<2> { typedef uint2 type; };
template <> struct MaskVec<1> { typedef uint32_t type; };

template <int V> __device__ __forceinline__ void mask_vec_load(const uint32_t* p, uint32_t (&w)[V]);
template <> __device__ __forceinline__ void mask_vec_load<4>(const uint32_t* p, uint32_t (&w)[4]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
template <> __device__ __forceinline__ void mask_vec_load<2>(const uint32_t* p, uint32_t (&w)[2]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    w[0] = v.x; w[1] = v.y;
}
template <> __device__ __forceinline__ void mask_vec_load<1>(const uint32_t* p, uint32_t (&w)[1]) { w[0] = *p; }

// Before a slot set is (re)used by a chunk: its boxes empty, its touched-cell counter zero.
__global__ void k_reset_slot_state(FrameBox* __restrict__ boxes, unsigned long long* __restrict__ touched_total) {
    if (threadIdx.x < kMaxBatch) box_reset(&boxes[threadIdx.x].x0);
    if (threadIdx.x == 0) *touched_total = 0ull;
}

template <int NJ>
__global__ void __launch_bounds__(kThreads)
k_apply(double* __restrict__ map, const __grid_constant__ ApplyParams ap, FrameBox* __restrict__ boxes,
        FrameBox* __restrict__ next_boxes, unsigned long long* __restrict__ touched_total,
        unsigned long long* __restrict__ next_touched_total, FrameBox* __restrict__ ubox, const double* __restrict__ cm,
        int c, int lane_cls, int mw) {
    constexpr int V = 2;
    extern __shared__ double s_cm[];  // C x C, transposed: s_cm[i * c + j] = cm[j * c + i] (column i contiguous)
    __shared__ FrameBox s_boxes[kMaxBatch];
    __shared__ unsigned int s_count;
    for (int e = threadIdx.x; e < c * c; e += blockDim.x) s_cm[(e % c) * c + e / c] = cm[e];
    if (threadIdx.x < kMaxBatch) {
        FrameBox b;
        box_reset(&b.x0);
        if (threadIdx.x < ap.n_frames) b = boxes[threadIdx.x];
        s_boxes[threadIdx.x] = b;
    }
    if (threadIdx.x == 0) s_count = 0;
    // (next_boxes == nullptr: the host resets a slot set's boxes itself before it reuses the set, k_reset_slot_state)
    if (next_boxes && blockIdx.x == 0 && threadIdx.x < kMaxBatch) box_reset(&next_boxes[threadIdx.x].x0);
    if (next_touched_total && blockIdx.x == 0 && threadIdx.x == 0) *next_touched_total = 0ull;
    __syncthreads();
    int x0 = 0x7fffffff, x1 = -1, y0 = 0x7fffffff, y1 = -1;
    for (int f = 0; f < ap.n_frames; ++f) {
        if (s_boxes[f].x1 < s_boxes[f].x0) continue;
        x0 = min(x0, s_boxes[f].x0); x1 = max(x1, s_boxes[f].x1);
        y0 = min(y0, s_boxes[f].y0); y1 = max(y1, s_boxes[f].y1);
    }
    if (x1 < x0) return;   // block-uniform: nothing was touched
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // the handle's union window (what a multi-GPU exchange has to move)
        atomicMin(&ubox->x0, x0); atomicMax(&ubox->x1, x1);
        atomicMin(&ubox->y0, y0); atomicMax(&ubox->y1, y1);
    }
    // a thread owns V = 2 horizontally adjacent cells whose linear index is even, so that the two words of a
    // slot come with one 8-byte load; columns outside [y0, y1] that such a pair drags in are ignored
    const uint32_t gcols = (uint32_t)(y1 - y0 + 1) / V + 2u;          // pairs per row, generous
    const uint64_t total = (uint64_t)(x1 - x0 + 1) * gcols;
    const uint32_t boost = 1u << c;
    const uint32_t lane_bit = lane_cls >= 0 ? (1u << lane_cls) : 0u;
    unsigned int mine = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t rr = (uint32_t)(t / gcols), gc = (uint32_t)(t - (uint64_t)rr * gcols);
        const int cx = x0 + (int)rr;
        const uint32_t row0 = (uint32_t)cx * (uint32_t)mw;
        const uint32_t cell0 = ((row0 + (uint32_t)y0) & ~1u) + gc * V;   // even linear index
        // valid cells of the pair: inside this row's [y0, y1]
        bool valid[V];
        bool anyvalid = false;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t cell = cell0 + v;
            valid[v] = cell >= row0 + (uint32_t)y0 && cell <= row0 + (uint32_t)y1;
            anyvalid |= valid[v];
        }
        if (!anyvalid) continue;
        // all mask words of the pair, every frame whose box covers it: up to 16 independent loads in flight
        uint32_t w[kMaxBatch][V];
        uint32_t any = 0;
#pragma unroll
        for (int f = 0; f < kMaxBatch; ++f) {
#pragma unroll
            for (int v = 0; v < V; ++v) w[f][v] = 0;
            if (f < ap.n_frames && cx >= s_boxes[f].x0 && cx <= s_boxes[f].x1) mask_vec_load<V>(ap.mask[f] + cell0, w[f]);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (!valid[v]) w[f][v] = 0;
                any |= w[f][v];
            }
        }
        if (!any) continue;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            uint32_t anyv = 0;
#pragma unroll
            for (int f = 0; f < kMaxBatch; ++f) anyv |= w[f][v];
            if (!anyv) continue;
            const uint32_t cell = cell0 + v;
            SMAP_BOUNDS(cell / (uint32_t)mw <= (uint32_t)x1 && cell % (uint32_t)mw >= (uint32_t)y0 && cell % (uint32_t)mw <= (uint32_t)y1, 201);
            double* row = map + (size_t)cell * c;
            double acc[8 * NJ];
#pragma unroll
            for (int j = 0; j < 8 * NJ; ++j) acc[j] = (j < c) ? row[j] : 0.0;
            // replay the frames in order on the row held in registers
#pragma unroll
            for (int f = 0; f < kMaxBatch; ++f) {
                const uint32_t wf = w[f][v];
                if (!wf) continue;
                ap.mask[f][cell] = 0u;
                ++mine;
#pragma unroll 1
                for (int i = 0; i < c; ++i) {
                    if (!((wf >> i) & 1u)) continue;
                    const double* col = s_cm + i * c;
#pragma unroll
                    for (int j = 0; j < 8 * NJ; ++j)
                        if (j < c) acc[j] = __dadd_rn(acc[j], col[j]);
                    if (i == lane_cls && (wf & boost)) {
                        // a bit test per (compile-time) j: acc[] must never be indexed dynamically
#pragma unroll
                        for (int j = 0; j < 8 * NJ; ++j)
                            if ((lane_bit >> j) & 1u) acc[j] = __dadd_rn(acc[j], 2.0);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8 * NJ; ++j)
                if (j < c) row[j] = acc[j];
        }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(touched_total, (unsigned long long)s_count);
}

// ------------------------------------------------------------------------------------------------
// Parity path of project_pcd (src/mapping_replay.py:223-244): flags + pixel indices, then an
// order-preserving compaction (block counts -> scan -> scatter).
// ------------------------------------------------------------------------------------------------
constexpr int kCompactPts = 4;                         // points per thread
constexpr int kCompactTile = kThreads * kCompactPts;   // thread t owns points [t*4, t*4+4) of the tile

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads)
k_project_flags(const void* __restrict__ pts, int64_t n, int64_t ld, const __grid_constant__ FrameParams fp,
                uint8_t* __restrict__ keep, int32_t* __restrict__ iu_out, int32_t* __restrict__ iv_out,
                uint32_t* __restrict__ block_count) {
    const int64_t base = (int64_t)blockIdx.x * kCompactTile + (int64_t)threadIdx.x * kCompactPts;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kCompactPts; ++j) {
        const int64_t k = base + j;
        if (k < n) {
            double x, y, z, it;
            load_point<LAYOUT>(pts, ld, k, x, y, z, it);
            int iu, iv;
            const bool ok = project_point(fp, x, y, z, iu, iv);
            keep[k] = ok ? 1 : 0;
            iu_out[k] = iu;
            iv_out[k] = iv;
            mine += ok;
        }
    }
    __shared__ int s_sum[kThreads / 32];
    int w = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < kThreads / 32; ++i) s += s_sum[i];
        block_count[blockIdx.x] = (uint32_t)s;
    }
}

// exclusive scan of the block counts by one block; writes the grand total to total_out
__global__ void __launch_bounds__(1024)
k_scan_blocks(const uint32_t* __restrict__ block_count, int64_t* __restrict__ block_offset, int64_t nblocks,
              int64_t* __restrict__ total_out) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t start = 0; start < nblocks; start += blockDim.x) {
        const int64_t i = start + threadIdx.x;
        const int64_t v = i < nblocks ? (int64_t)block_count[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += up;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t ws = s_warp[threadIdx.x];
            int64_t wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += up;
            }
            s_warp[threadIdx.x] = wi - ws;  // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int64_t carry = s_carry;
        if (i < nblocks) block_offset[i] = carry + s_warp[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[threadIdx.x >> 5] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads)
k_compact(const void* __restrict__ pts, int64_t n, int64_t ld, const uint8_t* __restrict__ image, int img_w,
          const uint8_t* __restrict__ keep, const int32_t* __restrict__ iu_in, const int32_t* __restrict__ iv_in,
          const int64_t* __restrict__ block_offset, double* __restrict__ out_pcd, uint8_t* __restrict__ out_label,
          int32_t* __restrict__ out_uv, int64_t out_ld) {
    const int64_t base = (int64_t)blockIdx.x * kCompactTile + (int64_t)threadIdx.x * kCompactPts;
    bool k_ok[kCompactPts];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kCompactPts; ++j) {
        const int64_t k = base + j;
        k_ok[j] = (k < n) && keep[k];
        mine += k_ok[j];
    }
    // exclusive prefix of `mine` over the block (thread order == point order)
    __shared__ int s_warp[kThreads / 32];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += up;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int warp_base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_base += s_warp[w];
    int64_t dst = block_offset[blockIdx.x] + warp_base + incl - mine;
#pragma unroll
    for (int j = 0; j < kCompactPts; ++j) {
        if (!k_ok[j]) continue;
        const int64_t k = base + j;
        double x, y, z, it;
        load_point<LAYOUT>(pts, ld, k, x, y, z, it);
        out_pcd[dst] = x;
        out_pcd[out_ld + dst] = y;
        out_pcd[2 * out_ld + dst] = z;
        out_pcd[3 * out_ld + dst] = it;
        const int iu = iu_in[k], iv = iv_in[k];
        const uint8_t* px = image + 3 * ((int64_t)iv * img_w + iu);
        out_label[dst] = px[0];
        out_label[out_ld + dst] = px[1];
        out_label[2 * out_ld + dst] = px[2];
        if (out_uv) {
            out_uv[dst] = iu;
            out_uv[out_ld + dst] = iv;
        }
        ++dst;
    }
}

// ------------------------------------------------------------------------------------------------
// Rendering.  K4 + K5 (apply_filter + render_bev_map, fused) live in smap_render.cuh; below: the class-axis
// reduction shared with K6 and K6 itself.
// ------------------------------------------------------------------------------------------------
// Streams the C values of one cell in ascending class order through `value(ch)` and returns the first
// argmax (np.argmax semantics: a NaN wins and stops the scan) and the class-axis sum in numpy's order
// (add.reduce over a contiguous axis = 0 + pairwise sum: fewer than 8 addends left to right, otherwise
// eight running partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail).
template <typename F>
__device__ __forceinline__ void argmax_and_npsum(int c, F value, int& best, double& total) {
    double mp = 0.0;
    best = 0;
    bool stop = false;
    auto track = [&](int ch, double v) {
        if (ch == 0) {
            mp = v;
            stop = (v != v);
        } else if (!stop && !(v <= mp)) {
            mp = v;
            best = ch;
            stop = (v != v);
        }
    };
    double res = 0.0;
    if (c < 8) {
        for (int i = 0; i < c; ++i) {
            const double v = value(i);
            track(i, v);
            res = __dadd_rn(res, v);
        }
    } else {
        double r[8];
        const int main = c - (c % 8);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            r[jj] = value(jj);
            track(jj, r[jj]);
        }
        for (int i = 8; i < main; i += 8) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const double v = value(i + jj);
                track(i + jj, v);
                r[jj] = __dadd_rn(r[jj], v);
            }
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (int i = main; i < c; ++i) {
            const double v = value(i);
            track(i, v);
            res = __dadd_rn(res, v);
        }
    }
    total = res;
}

// K6: render_bev_map_with_thresholds (src/renderer.py:131-172)
struct ThresholdParams {
    int32_t priority[32];
    double thresholds[32];
};

__global__ void __launch_bounds__(kThreads)
k_render_thresholds(const double* __restrict__ map, int64_t cells, int c, const __grid_constant__ RenderColors colors,
                    const __grid_constant__ ThresholdParams tp, uint8_t* __restrict__ rgb) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const double* a = map + cell * c;
    int best;
    double s;
    argmax_and_npsum(c, [&](int ch) { return a[ch]; }, best, s);
    uint8_t r = 0, g = 0, b = 0;
    if (s != 0.0) {  // known_region; a NaN sum counts as known, as in numpy
        for (int i = 0; i < c; ++i) {
            const int p = tp.priority[i];
            if (__ddiv_rn(a[p], s) >= tp.thresholds[i]) {
                r = colors.rgb[3 * p]; g = colors.rgb[3 * p + 1]; b = colors.rgb[3 * p + 2];
            }
        }
    }
    uint8_t* o = rgb + 3 * cell;
    o[0] = r; o[1] = g; o[2] = b;
}

// ------------------------------------------------------------------------------------------------
// K7: evaluation counts -- convert_labels (test/test_semantic_mapping.py:6-18) fused with the sums of Test.iou
// (:127-161) for the rendered map against the ground-truth label map (the slice the reference takes at :123-124).
//   counts[0..2]  pixels where ground truth and map both hold class 1 (road), 2 (crosswalk), 3 (lane)
//   counts[3..5]  ground-truth pixels of the class          counts[6..8]  map pixels of the class
//   counts[9]     known ground truth (g > 0)                 counts[10]    known and mapped (g > 0 and m > 0)
//   counts[11]    map equals ground truth where it is known
// Integer sums: exact in any order; the host turns them into IoU / accuracy / missing rate as the reference does.
struct EvalParams {
    int mh, mw;            // rendered map (rows, columns)
    int truth_ld;          // row stride of the ground-truth label map (uint8)
    int shift_r, shift_c;  // the map's (0, 0) sits at ground truth (shift_r, shift_c)   (Test.shift_w, Test.shift_h)
    int mask_ld;           // row stride of the validity mask, 0: no mask (convert_labels(gmap) with mask=None)
};

__global__ void __launch_bounds__(kThreads)
k_eval_counts(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ truth, const uint8_t* __restrict__ mask,
              const __grid_constant__ EvalParams p, unsigned long long* __restrict__ counts) {
    uint32_t n[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) n[i] = 0u;
    const int64_t cells = (int64_t)p.mh * p.mw;
    for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < cells; cell += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(cell / p.mw), c = (int)(cell - (int64_t)r * p.mw);
        const uint8_t* px = rgb + 3 * cell;
        const uint32_t R = px[0], G = px[1], B = px[2];
        uint32_t m = 0u;   // convert_labels: all three channels must match
        if (R == 128u && G == 64u && B == 128u) m = 1u;         // road
        else if (R == 140u && G == 140u && B == 200u) m = 2u;   // crosswalk
        else if (R == 255u && G == 255u && B == 255u) m = 3u;   // lane
        else if (R == 244u && G == 35u && B == 232u) m = 4u;    // sidewalk
        else if (R == 107u && G == 142u && B == 35u) m = 5u;    // vegetation
        if (p.mask_ld > 0 && mask[(int64_t)r * p.mask_ld + c] == 0u) m = 0u;
        const uint32_t g = truth[(int64_t)(r + p.shift_r) * p.truth_ld + (c + p.shift_c)];
#pragma unroll
        for (uint32_t k = 1; k <= 3; ++k) {
            n[k - 1] += (g == k && m == k);
            n[k + 2] += (g == k);
            n[k + 5] += (m == k);
        }
        n[9] += (g > 0u);
        n[10] += (g > 0u && m > 0u);
        n[11] += (g > 0u && g == m);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, n[i]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(counts + i, (unsigned long long)w);
    }
}

}  // namespace smap
