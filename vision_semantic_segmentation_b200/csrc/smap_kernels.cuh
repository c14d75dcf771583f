// Kernels of the semantic-mapping path (sm_100a).  See DESIGN.md for the data layout and the
// roofline of each kernel.  None of this is GEMM-shaped: the work is HBM-bound streaming of the cloud,
// a sector-granular gather from the label image and a scatter of per-cell class masks, so the design
// rules that matter are coalescing (float4 / contiguous rows), keeping atomics off single hot
// addresses, and grid sizes that fill 148 SMs.
#pragma once
#include "smap_device.cuh"

namespace smap {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// Epoch-tagged cell masks.
// Cell-mask word:  [ tag | boost bit (bit C) | C class bits ].  The tag is the serial number of the frame
// that last wrote the word; a word carrying an older tag is stale and reads as empty, so masks are never
// cleared.  Tags only grow within a mask slot, hence atomicMax installs the new tag (dropping the stale
// bits) without a compare-and-swap loop, and the following atomicOr returns what THIS frame had already
// put there: one L2 round trip on the critical path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tagged_or(uint32_t* __restrict__ word, uint32_t tagword, uint32_t bits, uint32_t low) {
    atomicMax(word, tagword);
    return atomicOr(word, bits) & low;
}

// Bounding box (in cells) of everything a frame touched + number of touched cells; written by the ordered
// (two-kernel) update so that k_apply sweeps only that window of the mask.
struct FrameBox {
    int x0, x1, y0, y1;   // inclusive; empty when x1 < x0
    uint32_t touched;     // filled by k_apply (statistics)
    uint32_t pad[3];
};

// One frame of a batched launch.
struct BatchFrame {
    FrameParams fp;
    const void* pts;
    const uint8_t* image;
    uint32_t* mask;       // this frame's mask slot (MH*MW words)
    int64_t n;
    int64_t ld;
    uint32_t tagword;     // frame tag << tag_shift
    uint32_t unit_begin;  // first work unit of this frame in the launch
};

struct BatchParams {
    int n_frames;
    uint32_t n_units;
    BatchFrame f[kMaxBatch];
};

// tuning knobs (overridable at compile time for the variant sweeps recorded in profiles/)
#ifndef SMAP_STREAM_ROUND
#define SMAP_STREAM_ROUND 8     // 32-point chunks a warp loads back to back (LDG.128 in flight per lane)
#endif
#ifndef SMAP_STREAM_ROUNDS
#define SMAP_STREAM_ROUNDS 2    // rounds per work unit and warp
#endif
#ifndef SMAP_STREAM_MINB
#define SMAP_STREAM_MINB 3      // resident blocks per SM the register allocation aims for
#endif
constexpr int kWarps = kThreads / 32;
constexpr int kRound = SMAP_STREAM_ROUND;
constexpr int kRounds = SMAP_STREAM_ROUNDS;
constexpr int kWarpUnitPts = 32 * kRound * kRounds;   // points of a unit owned by one warp
constexpr int kUnitPts = kWarps * kWarpUnitPts;       // points per work unit (per block iteration)
constexpr int kQueueCap = 32 * kRound + 32;           // per-warp survivor stack: < 32 left over + one round

template <int LAYOUT> struct QueueEntry { typedef float4 type; };
template <> struct QueueEntry<1> { typedef uint32_t type; };

// ------------------------------------------------------------------------------------------------
// K1+K2+K3 fused, persistent and warp-autonomous: the whole per-frame rule of SURVEY.md section 9 in one
// kernel, up to kMaxBatch frames per launch.
//
// A block walks work units (kUnitPts consecutive points of one frame) round-robin; inside a unit every
// warp owns a contiguous slice and never synchronises with the other warps:
//   stream   kRound coalesced LDG.128 per lane in flight, float32 conservative cull (precull_pass), survivors
//            (~35 %) pushed on the warp's private stack in shared memory (ballot + popc, no atomics);
//   drain    whenever >= 32 survivors are stacked, pop 32 - one per lane, all lanes busy: exact double
//            projection (src/mapping_replay.py:223-240), label gather (:244), class bits from the shared colour
//            tables (:276,:288-290), cell (:261-268), tagged OR into the frame's mask slot (the per-frame
//            (cell, class) de-duplication of :281/:294);
//   MODE 1   (count update, CM = identity) every NEWLY set class bit adds 1.0 to map[cell, class] and a newly
//            set boost bit adds 2.0 to map[cell, lane] with a float64 atomic: integer-valued sums, exact in any
//            order;
//   MODE 0   (ordered update) only the masks are written, plus the bounding box of the touched cells;
//            k_apply then adds the matrix columns in class order.
// Stacks survive unit boundaries and are flushed (partial warps) only when the block moves to another frame.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(kThreads, SMAP_STREAM_MINB)
k_stream(const __grid_constant__ BatchParams bp, const __grid_constant__ GridParams gp, double* __restrict__ map,
         FrameBox* __restrict__ box) {
    typedef typename QueueEntry<LAYOUT>::type Entry;
    __shared__ __align__(16) Entry s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_tab_r[256], s_tab_g[256];
    __shared__ FrameParams s_fp;
    __shared__ int s_box[4];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    Entry* const queue = s_queue[warp];
    const uint32_t low = (1u << gp.tag_shift) - 1u;

    build_color_tables(gp, s_tab_r, s_tab_g);
    if (MODE == 0 && threadIdx.x == 0) {
        s_box[0] = 0x7fffffff; s_box[1] = -1; s_box[2] = 0x7fffffff; s_box[3] = -1;
    }

    int cur = -1;          // frame whose constants are in s_fp
    uint32_t qn = 0;       // entries on this warp's stack (warp-uniform)
    int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1;

    // frame-dependent values a lane keeps in registers
    const void* f_pts = nullptr;
    const uint8_t* f_image = nullptr;
    uint32_t* f_mask = nullptr;
    int64_t f_n = 0, f_ld = 0;
    uint32_t f_tagword = 0;
    bool f_words = false;   // label image can be read with aligned 32-bit loads

    // ---- one batch of <= 32 stacked survivors, one per lane
    auto drain_batch = [&](uint32_t count) {
        const uint32_t first = qn - count;
        qn = first;
        if ((uint32_t)lane >= count) return;
        double x, y, z;
        float it;
        if (LAYOUT == 0) {
            const float4 w = *reinterpret_cast<const float4*>(&queue[first + lane]);
            x = (double)w.x; y = (double)w.y; z = (double)w.z; it = w.w;
        } else {
            const double* pd = reinterpret_cast<const double*>(f_pts);
            const int64_t k = (int64_t)*reinterpret_cast<const uint32_t*>(&queue[first + lane]);
            x = __ldg(pd + k); y = __ldg(pd + f_ld + k); z = __ldg(pd + 2 * f_ld + k);
            // the boost test compares the float64 intensity with 2 and 14; rounding to float32 could move a
            // value across them, so map the double onto a float on the same side (NaN: neither)
            const double itd = __ldg(pd + 3 * f_ld + k);
            it = (itd < 2.0) ? 0.0f : ((itd > 14.0) ? 15.0f : 8.0f);
        }
        int iu, iv;
        if (!project_point(s_fp, x, y, z, iu, iv)) return;
        const uint32_t off = 3u * ((uint32_t)iv * (uint32_t)s_fp.img_w + (uint32_t)iu);
        uint32_t r, g;
        if (f_words) {
            // R and G sit in one aligned 32-bit word unless R is its last byte
            const uintptr_t addr = reinterpret_cast<uintptr_t>(f_image) + off;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(addr & 3u) * 8u;
            const uint32_t w0 = __ldg(wp);
            r = (w0 >> sh) & 0xffu;
            g = (sh == 24u) ? (__ldg(wp + 1) & 0xffu) : ((w0 >> (sh + 8u)) & 0xffu);
        } else {
            r = __ldg(f_image + off);
            g = __ldg(f_image + off + 1);
        }
        const uint32_t bits = class_bits_lut(gp, s_tab_r, s_tab_g, (uint8_t)r, (uint8_t)g, it);
        if (!bits) return;
        int cx, cy;
        if (!cell_xy(gp, x, y, cx, cy)) return;
        const uint32_t cell = (uint32_t)cx * (uint32_t)gp.mw + (uint32_t)cy;
        const uint32_t prev = tagged_or(f_mask + cell, f_tagword, bits, low);
        if (MODE == 0) {
            bx0 = min(bx0, cx); bx1 = max(bx1, cx); by0 = min(by0, cy); by1 = max(by1, cy);
        } else {
            uint32_t fresh = bits & ~prev;
            double* row = map + (size_t)cell * gp.c;
            if (fresh >> gp.c) {  // boost bit newly set: +2 on the lane class (src/mapping_replay.py:294)
                atomicAdd(row + gp.lane, 2.0);
                fresh &= (1u << gp.c) - 1u;
            }
            while (fresh) {
                const int i = __ffs(fresh) - 1;
                fresh &= fresh - 1u;
                atomicAdd(row + i, 1.0);
            }
        }
    };

    auto flush_box = [&]() {  // MODE 0: fold the lanes' boxes into the block's
        if (MODE != 0) return;
        if (bx1 >= bx0) {
            atomicMin(&s_box[0], bx0); atomicMax(&s_box[1], bx1);
            atomicMin(&s_box[2], by0); atomicMax(&s_box[3], by1);
        }
    };

    for (uint32_t unit = blockIdx.x; unit < bp.n_units; unit += gridDim.x) {
        int fi = cur < 0 ? 0 : cur;
        while (fi + 1 < bp.n_frames && unit >= bp.f[fi + 1].unit_begin) ++fi;
        if (fi != cur) {  // block-uniform: units are frame-major and every warp sees the same sequence
            while (qn) {  // the old frame's leftovers still need the old constants
                drain_batch(qn < 32u ? qn : 32u);
                __syncwarp();
            }
            __syncthreads();
            const BatchFrame& F = bp.f[fi];
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&F.fp);
            uint32_t* dst = reinterpret_cast<uint32_t*>(&s_fp);
            for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x) dst[i] = src[i];
            f_pts = F.pts; f_image = F.image; f_mask = F.mask; f_n = F.n; f_ld = F.ld; f_tagword = F.tagword;
            f_words = ((reinterpret_cast<uintptr_t>(F.image) & 3u) == 0u) &&
                      (((int64_t)F.fp.img_w * F.fp.img_h * 3) % 4 == 0);
            cur = fi;
            __syncthreads();
        }
        CullConsts kc;
#pragma unroll
        for (int i = 0; i < 16; ++i) kc.m[i] = s_fp.Mf[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            kc.ea[i] = s_fp.Ea[i];
            kc.eb[i] = s_fp.Eb[i];
        }
        kc.range_hi = s_fp.range_hi;
        kc.wf = s_fp.img_wf;
        kc.hf = s_fp.img_hf;

        const int64_t wbase = (int64_t)(unit - bp.f[fi].unit_begin) * kUnitPts + (int64_t)warp * kWarpUnitPts;
#pragma unroll 1
        for (int rd = 0; rd < kRounds; ++rd) {
            const int64_t rbase = wbase + (int64_t)rd * (32 * kRound);
            if (rbase >= f_n) break;
            if (LAYOUT == 0) {
                const float4* p4 = reinterpret_cast<const float4*>(f_pts);
                float4 p[kRound];
#pragma unroll
                for (int c = 0; c < kRound; ++c) {
                    const int64_t k = rbase + c * 32 + lane;
                    p[c] = (k < f_n) ? __ldcs(p4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int c = 0; c < kRound; ++c) {
                    const int64_t k = rbase + c * 32 + lane;
                    const bool pass = (k < f_n) & precull_pass(kc, p[c].x, p[c].y, p[c].z);
                    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                    if (pass) *reinterpret_cast<float4*>(&queue[qn + __popc(ballot & lt_mask)]) = p[c];
                    qn += __popc(ballot);
                }
            } else {
                const double* pd = reinterpret_cast<const double*>(f_pts);
#pragma unroll 2
                for (int c = 0; c < kRound; ++c) {
                    const int64_t k = rbase + c * 32 + lane;
                    bool pass = false;
                    if (k < f_n)
                        pass = precull_pass(kc, (float)__ldg(pd + k), (float)__ldg(pd + f_ld + k), (float)__ldg(pd + 2 * f_ld + k));
                    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                    if (pass) *reinterpret_cast<uint32_t*>(&queue[qn + __popc(ballot & lt_mask)]) = (uint32_t)k;
                    qn += __popc(ballot);
                }
            }
            __syncwarp();
            while (qn >= 32u) {
                drain_batch(32u);
                __syncwarp();
            }
        }
    }
    while (qn) {
        drain_batch(qn < 32u ? qn : 32u);
        __syncwarp();
    }
    if (MODE == 0) {
        flush_box();
        __syncthreads();
        if (threadIdx.x == 0 && s_box[1] >= s_box[0]) {
            atomicMin(&box->x0, s_box[0]); atomicMax(&box->x1, s_box[1]);
            atomicMin(&box->y0, s_box[2]); atomicMax(&box->y1, s_box[3]);
        }
    }
}

// Parity kernel for update_map (src/mapping_replay.py:261-277,:288-290): the scatter half of the ordered update from
// an already projected cloud (4, M) float64 + its (3, M) RGB labels.
__global__ void __launch_bounds__(kThreads)
k_update_scatter(const double* __restrict__ pcd, int64_t ld, const uint8_t* __restrict__ label, int64_t ldl,
                 int64_t m, const __grid_constant__ GridParams gp, uint32_t* __restrict__ mask, uint32_t tagword,
                 FrameBox* __restrict__ box) {
    __shared__ int s_box[4];
    if (threadIdx.x == 0) {
        s_box[0] = 0x7fffffff; s_box[1] = -1; s_box[2] = 0x7fffffff; s_box[3] = -1;
    }
    __syncthreads();
    const uint32_t low = (1u << gp.tag_shift) - 1u;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t bits = class_bits(gp, label[k], label[ldl + k], pcd[3 * ld + k]);
        int cx, cy;
        if (bits && cell_xy(gp, pcd[k], pcd[ld + k], cx, cy)) {
            tagged_or(mask + (uint32_t)cx * (uint32_t)gp.mw + (uint32_t)cy, tagword, bits, low);
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
// K3b (ordered update): sweep the frame's bounding box of the mask slot; every cell whose word carries this
// frame's tag gets the matrix columns of its classes added in ascending class order, then the lane boost:
// the "+=" statements at src/mapping_replay.py:281 and :294 bit for bit, for any update matrix.
// The box rows are contiguous in memory (axis 1), so the 4-byte mask reads are coalesced; the window of a
// 100 m frustum at 0.1 m is ~1.2 M words = 5 MB, cheaper than maintaining a list of touched cells with
// contended atomics.  `box` is this frame's window, `next_box` (other half of the double buffer) is reset.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_apply(double* __restrict__ map, const uint32_t* __restrict__ mask, uint32_t tagword, int tag_shift,
        FrameBox* __restrict__ box, FrameBox* __restrict__ next_box, const double* __restrict__ cm, int c, int lane_cls, int mw) {
    extern __shared__ double s_cm[];  // C x C, transposed: s_cm[i * c + j] = cm[j * c + i] (column i contiguous)
    __shared__ uint32_t s_count;
    for (int e = threadIdx.x; e < c * c; e += blockDim.x) s_cm[(e % c) * c + e / c] = cm[e];
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        next_box->x0 = 0x7fffffff; next_box->x1 = -1; next_box->y0 = 0x7fffffff; next_box->y1 = -1;
        next_box->touched = 0;
    }
    const int x0 = box->x0, x1 = box->x1, y0 = box->y0, y1 = box->y1;
    if (x1 < x0) return;
    const uint32_t ncols = (uint32_t)(y1 - y0 + 1);
    const uint64_t total = (uint64_t)(x1 - x0 + 1) * ncols;
    const uint32_t boost = 1u << c;
    const uint32_t tag = tagword >> tag_shift;
    uint32_t mine = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t rr = (uint32_t)(t / ncols), cc = (uint32_t)(t - (uint64_t)rr * ncols);
        const uint32_t cell = (uint32_t)(x0 + (int)rr) * (uint32_t)mw + (uint32_t)(y0 + (int)cc);
        const uint32_t word = __ldcg(mask + cell);
        if ((word >> tag_shift) != tag) continue;
        ++mine;
        double* row = map + (size_t)cell * c;
        for (int j0 = 0; j0 < c; j0 += 8) {
            double acc[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] = (j0 + jj < c) ? row[j0 + jj] : 0.0;
            for (int i = 0; i < c; ++i) {
                if (!((word >> i) & 1u)) continue;
                const double* col = s_cm + i * c + j0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    if (j0 + jj < c) acc[jj] = __dadd_rn(acc[jj], col[jj]);
                if (i == lane_cls && (word & boost) && i >= j0 && i < j0 + 8) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        if (j0 + jj == i) acc[jj] = __dadd_rn(acc[jj], 2.0);
                }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
                if (j0 + jj < c) row[j0 + jj] = acc[jj];
        }
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(&box->touched, s_count);
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
// Rendering.  K4 (3x3 box, BORDER_REFLECT_101, acc = acc + kf*p row-major from 0; src/renderer.py:175-189),
// K5 (first-argmax colour, zero-sum -> black; src/renderer.py:32-59) fused; a tile of the grid with
// a one-cell halo is staged in shared memory with coalesced row-segment loads.
// ------------------------------------------------------------------------------------------------
constexpr int kTileX = 32;
constexpr int kTileY = 8;

struct RenderColors {
    uint8_t rgb[32 * 3];
};

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

template <bool FILTER>
__global__ void __launch_bounds__(kTileX * kTileY)
k_render(const double* __restrict__ map, int mh, int mw, int c, const __grid_constant__ RenderColors colors,
         uint8_t* __restrict__ rgb, double* __restrict__ filtered) {
    extern __shared__ double s_tile[];  // (kTileY+2) x (kTileX+2) x C   (FILTER only)
    const int x0 = blockIdx.x * kTileX, y0 = blockIdx.y * kTileY;
    const int tx = threadIdx.x % kTileX, ty = threadIdx.x / kTileX;
    const int rowlen = (kTileX + 2) * c;
    if (FILTER) {
        for (int r = 0; r < kTileY + 2; ++r) {
            const int yy = y0 - 1 + r;
            if (yy > mh) break;  // rows past the bottom halo are never read
            const int ys = reflect101(yy, mh);
            for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
                const int col = x0 - 1 + e / c;
                if (col > mw) break;
                const int xs = reflect101(col, mw);
                s_tile[r * rowlen + e] = map[((size_t)ys * mw + xs) * c + (e % c)];
            }
        }
        __syncthreads();
    }
    const int x = x0 + tx, y = y0 + ty;
    if (x >= mw || y >= mh) return;
    const size_t cell = (size_t)y * mw + x;
    const double kf = (double)(1.0f / 9.0f);
    auto value = [&](int ch) -> double {
        if constexpr (!FILTER) return map[cell * c + ch];
        double acc = 0.0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
                acc = __dadd_rn(acc, __dmul_rn(kf, s_tile[(ty + dy) * rowlen + (tx + dx) * c + ch]));
        if (filtered) filtered[cell * c + ch] = acc;
        return acc;
    };
    int best;
    double total;
    argmax_and_npsum(c, value, best, total);
    if (rgb) {
        uint8_t* o = rgb + 3 * cell;
        if (total == 0.0) {
            o[0] = 0; o[1] = 0; o[2] = 0;
        } else {
            o[0] = colors.rgb[3 * best]; o[1] = colors.rgb[3 * best + 1]; o[2] = colors.rgb[3 * best + 2];
        }
    }
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

}  // namespace smap
