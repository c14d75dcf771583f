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
// Block-aggregated append of first-touched cells to the frame's touched list.
// One shared counter per block, one global atomic per block (a single hot global address would
// otherwise serialise ~1e5 atomics per frame in L2).
// ------------------------------------------------------------------------------------------------
template <int CAP>
struct TouchList {
    uint32_t cells[CAP];
    uint32_t n;
    uint32_t base;
};

template <int CAP>
__device__ __forceinline__ void touch_push(TouchList<CAP>& tl, bool first, uint32_t cell) {
    const unsigned ballot = __ballot_sync(0xffffffffu, first);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(ballot) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&tl.n, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (first) tl.cells[base + __popc(ballot & ((1u << lane) - 1u))] = cell;
    }
}

template <int CAP>
__device__ __forceinline__ void touch_flush(TouchList<CAP>& tl, uint32_t* __restrict__ touched,
                                            uint32_t* __restrict__ counter) {
    __syncthreads();
    if (threadIdx.x == 0 && tl.n) tl.base = atomicAdd(counter, tl.n);
    __syncthreads();
    const uint32_t n = tl.n, base = tl.base;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) touched[base + i] = tl.cells[i];
}

// ------------------------------------------------------------------------------------------------
// Epoch-tagged cell masks.
// OR `bits` into the word of `cell` for the frame whose tag is `tagword` (tag already shifted into the
// high bits) and return the class/boost bits the word held FOR THIS FRAME before the call.
// A word carrying an older tag is stale and reads as empty; tags only grow within a mask slot, so
// atomicMax installs the new tag (clearing the old bits) without a read-modify-write loop.
// A plain L2 load first: a point whose bits are already present costs no atomic at all.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tagged_or(uint32_t* __restrict__ word, uint32_t tagword, uint32_t bits, int shift) {
    const uint32_t low = (1u << shift) - 1u;
    uint32_t cur = __ldcg(word);
    if ((cur >> shift) != (tagword >> shift)) {
        atomicMax(word, tagword);
    } else if ((cur & bits) == bits) {
        return cur & low;
    }
    return atomicOr(word, bits) & low;
}

// One frame of a batched launch.
struct BatchFrame {
    FrameParams fp;
    const void* pts;
    const uint8_t* image;
    uint32_t* mask;       // this frame's mask slot (MH*MW words)
    int64_t n;
    int64_t ld;
    uint32_t tagword;     // frame tag << tag_shift
    uint32_t block_begin; // first block of this frame in the launch
};

struct BatchParams {
    int n_frames;
    int pad;
    BatchFrame f[kMaxBatch];
};

constexpr int kFusePts = 8;                      // points per thread in the pre-cull phase
constexpr int kFuseTile = kThreads * kFusePts;   // 2048 points per block

// ------------------------------------------------------------------------------------------------
// K1+K2+K3 fused: the whole per-frame rule of SURVEY.md section 9 in one kernel, several frames per launch.
//   phase 1  every thread streams kFusePts points (coalesced float4 / double rows), runs the float32
//            conservative cull and appends the survivors (~35 %) to a shared-memory list;
//   phase 2  the block walks the dense survivor list: exact double projection (src/mapping_replay.py:223-240),
//            label gather (:244), class bits from the shared colour tables (:276,:288-290), cell (:261-268),
//            tagged OR into the frame's mask (the per-frame (cell, class) de-duplication of :281/:294);
//   MODE 0   (deterministic) the thread that first touches a cell records it for k_apply;
//   MODE 1   (count update, CM = identity) every NEWLY set class bit adds 1.0 to map[cell, class] and a newly
//            set boost bit adds 2.0 to map[cell, lane] with a float64 atomic: integer-valued sums, exact in any order.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(kThreads, 3)
k_fuse(const __grid_constant__ BatchParams bp, const __grid_constant__ GridParams gp, double* __restrict__ map,
       uint32_t* __restrict__ touched, uint32_t* __restrict__ counter) {
    __shared__ uint32_t s_tab_r[256], s_tab_g[256];
    __shared__ float4 s_surv[LAYOUT == 0 ? kFuseTile : 1];     // surviving points (float4 layout)
    __shared__ uint32_t s_idx[LAYOUT == 0 ? 1 : kFuseTile];    // or their index in the tile (float64 layout)
    __shared__ uint32_t s_nsurv;
    __shared__ TouchList<MODE == 0 ? kFuseTile : 1> tl;

    int fi = 0;
    while (fi + 1 < bp.n_frames && blockIdx.x >= bp.f[fi + 1].block_begin) ++fi;
    const BatchFrame& F = bp.f[fi];
    const FrameParams& fp = F.fp;

    build_color_tables(gp, s_tab_r, s_tab_g);
    if (threadIdx.x == 0) {
        s_nsurv = 0;
        tl.n = 0;
    }
    __syncthreads();

    // ---- phase 1: stream + float32 pre-cull + compaction
    const int64_t base = (int64_t)(blockIdx.x - F.block_begin) * kFuseTile;
    const int lane = threadIdx.x & 31;
    if (LAYOUT == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(F.pts);
        float4 p[kFusePts];
#pragma unroll
        for (int j = 0; j < kFusePts; ++j) {
            const int64_t k = base + j * kThreads + threadIdx.x;
            p[j] = (k < F.n) ? __ldcs(p4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < kFusePts; ++j) {
            const int64_t k = base + j * kThreads + threadIdx.x;
            const bool pass = (k < F.n) && precull_pass(fp, p[j].x, p[j].y, p[j].z);
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (ballot) {
                const int leader = __ffs(ballot) - 1;
                uint32_t slot = 0;
                if (lane == leader) slot = atomicAdd(&s_nsurv, (uint32_t)__popc(ballot));
                slot = __shfl_sync(0xffffffffu, slot, leader);
                if (pass) s_surv[slot + __popc(ballot & ((1u << lane) - 1u))] = p[j];
            }
        }
    } else {
        const double* pd = reinterpret_cast<const double*>(F.pts);
#pragma unroll 2
        for (int j = 0; j < kFusePts; ++j) {
            const int64_t k = base + j * kThreads + threadIdx.x;
            bool pass = false;
            if (k < F.n) pass = precull_pass(fp, (float)__ldg(pd + k), (float)__ldg(pd + F.ld + k), (float)__ldg(pd + 2 * F.ld + k));
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (ballot) {
                const int leader = __ffs(ballot) - 1;
                uint32_t slot = 0;
                if (lane == leader) slot = atomicAdd(&s_nsurv, (uint32_t)__popc(ballot));
                slot = __shfl_sync(0xffffffffu, slot, leader);
                if (pass) s_idx[slot + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)(j * kThreads + threadIdx.x);
            }
        }
    }
    __syncthreads();

    // ---- phase 2: exact path on the dense survivor list
    const uint32_t nsurv = s_nsurv;
    for (uint32_t s0 = 0; s0 < nsurv; s0 += kThreads) {
        const uint32_t s = s0 + threadIdx.x;
        bool first = false;
        uint32_t cell = 0;
        if (s < nsurv) {
            double x, y, z;
            float it;
            if (LAYOUT == 0) {
                const float4 q = s_surv[s];
                x = (double)q.x; y = (double)q.y; z = (double)q.z; it = q.w;
            } else {
                const double* pd = reinterpret_cast<const double*>(F.pts);
                const int64_t k = base + s_idx[s];
                x = __ldg(pd + k); y = __ldg(pd + F.ld + k); z = __ldg(pd + 2 * F.ld + k);
                // the boost test compares the float64 intensity with 2 and 14; rounding to float32 could move a
                // value across them, so clamp the float onto the same side as the double
                const double itd = __ldg(pd + 3 * F.ld + k);
                it = (itd < 2.0) ? 0.0f : ((itd > 14.0) ? 15.0f : 8.0f);
                if (itd != itd) it = 8.0f;  // NaN: neither < 2 nor > 14
            }
            int iu, iv;
            if (project_point(fp, x, y, z, iu, iv)) {
                const uint8_t* px = F.image + 3 * ((int64_t)iv * fp.img_w + iu);
                const uint8_t r = __ldg(px), g = __ldg(px + 1);
                const uint32_t bits = class_bits_lut(gp, s_tab_r, s_tab_g, r, g, it);
                if (bits && cell_of(gp, x, y, cell)) {
                    const uint32_t prev = tagged_or(F.mask + cell, F.tagword, bits, gp.tag_shift);
                    if (MODE == 0) {
                        first = (prev == 0u);
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
                }
            }
        }
        if (MODE == 0) touch_push(tl, first, cell);
    }
    if (MODE == 0) touch_flush(tl, touched, counter);
}

// Parity kernel for update_map (src/mapping_replay.py:261-277,:288-290): same scatter as k_fuse MODE 0, but from
// an already projected cloud (4, M) float64 + its (3, M) RGB labels.
template <int PTS>
__global__ void __launch_bounds__(kThreads)
k_update_scatter(const double* __restrict__ pcd, int64_t ld, const uint8_t* __restrict__ label, int64_t ldl,
                 int64_t m, const __grid_constant__ GridParams gp, uint32_t* __restrict__ mask, uint32_t tagword,
                 uint32_t* __restrict__ touched, uint32_t* __restrict__ counter) {
    __shared__ TouchList<kThreads * PTS> tl;
    if (threadIdx.x == 0) tl.n = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (kThreads * PTS);
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const int64_t k = base + j * kThreads + threadIdx.x;
        bool first = false;
        uint32_t cell = 0;
        if (k < m) {
            const uint32_t bits = class_bits(gp, label[k], label[ldl + k], pcd[3 * ld + k]);
            if (bits && cell_of(gp, pcd[k], pcd[ld + k], cell))
                first = (tagged_or(mask + cell, tagword, bits, gp.tag_shift) == 0u);
        }
        touch_push(tl, first, cell);
    }
    touch_flush(tl, touched, counter);
}

// ------------------------------------------------------------------------------------------------
// K3b (deterministic path): apply the frame's cell masks to the grid, classes in ascending order.
// Replaces the "+=" statements at src/mapping_replay.py:281 and :294 bit for bit, any update matrix.
// One thread per touched cell: touched[] is read coalesced, then mask word and the C-element grid row are
// independent random accesses, so ~all touched cells of a frame are in flight at once (the first version
// walked 8-lane groups through a serial load chain and took 100 us for 137 k cells).
// `counter` is this frame's touched count; `next_counter` (other half of the double buffer) is zeroed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_apply(double* __restrict__ map, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ touched,
        const uint32_t* __restrict__ counter, uint32_t* __restrict__ next_counter,
        const double* __restrict__ cm, int c, int lane_cls) {
    extern __shared__ double s_cm[];  // C x C, transposed: s_cm[i * c + j] = cm[j * c + i] (column i contiguous)
    for (int e = threadIdx.x; e < c * c; e += blockDim.x) s_cm[(e % c) * c + e / c] = cm[e];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_counter = 0u;
    const uint32_t count = *counter;
    const uint32_t boost = 1u << c;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint32_t cell = touched[t];
        const uint32_t bits = __ldcg(mask + cell);
        double* row = map + (size_t)cell * c;
        for (int j0 = 0; j0 < c; j0 += 8) {
            double acc[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] = (j0 + jj < c) ? row[j0 + jj] : 0.0;
            for (int i = 0; i < c; ++i) {
                if (!((bits >> i) & 1u)) continue;
                const double* col = s_cm + i * c + j0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    if (j0 + jj < c) acc[jj] = __dadd_rn(acc[jj], col[jj]);
                if (i == lane_cls && (bits & boost) && i >= j0 && i < j0 + 8) {
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
