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
// K1+K2+K3a fused (deterministic path): project -> cull -> gather label -> class bits -> cell ->
// atomicOr into the frame's cell mask; the thread that turns a mask from 0 to non-zero records the cell.
// Replaces src/mapping_replay.py:223-244 and :261-277,:288-290 for one frame.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT, int PTS>
__global__ void __launch_bounds__(kThreads)
k_integrate(const void* __restrict__ pts, int64_t n, int64_t ld, const uint8_t* __restrict__ image,
            const __grid_constant__ FrameParams fp, const __grid_constant__ GridParams gp,
            uint32_t* __restrict__ mask, uint32_t* __restrict__ touched, uint32_t* __restrict__ counter) {
    __shared__ TouchList<kThreads * PTS> tl;
    if (threadIdx.x == 0) tl.n = 0;
    __syncthreads();

    const int64_t base = (int64_t)blockIdx.x * (kThreads * PTS);
    double x[PTS], y[PTS], z[PTS], it[PTS];
    bool live[PTS];
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const int64_t k = base + j * kThreads + threadIdx.x;
        live[j] = k < n;
        if (live[j]) load_point<LAYOUT>(pts, ld, k, x[j], y[j], z[j], it[j]);
    }
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        bool first = false;
        uint32_t cell = 0;
        if (live[j]) {
            int iu, iv;
            if (project_point(fp, x[j], y[j], z[j], iu, iv)) {
                const uint8_t* px = image + 3 * ((int64_t)iv * fp.img_w + iu);
                const uint8_t r = __ldg(px), g = __ldg(px + 1);
                const uint32_t bits = class_bits(gp, r, g, it[j]);
                if (bits && cell_of(gp, x[j], y[j], cell)) {
                    const uint32_t old = atomicOr(mask + cell, bits);
                    first = (old == 0u);
                }
            }
        }
        touch_push(tl, first, cell);
    }
    touch_flush(tl, touched, counter);
}

// Parity kernel for update_map (src/mapping_replay.py:261-277,:288-290): same scatter, but from an already
// projected cloud (4, M) float64 + its (3, M) RGB labels.
template <int PTS>
__global__ void __launch_bounds__(kThreads)
k_update_scatter(const double* __restrict__ pcd, int64_t ld, const uint8_t* __restrict__ label, int64_t ldl,
                 int64_t m, const __grid_constant__ GridParams gp, uint32_t* __restrict__ mask,
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
            if (bits && cell_of(gp, pcd[k], pcd[ld + k], cell)) first = (atomicOr(mask + cell, bits) == 0u);
        }
        touch_push(tl, first, cell);
    }
    touch_flush(tl, touched, counter);
}

// ------------------------------------------------------------------------------------------------
// K3b: apply the frame's cell masks to the grid, classes in ascending order, then clear the masks.
// Replaces the "+=" statements at src/mapping_replay.py:281 and :294.
// A group of G lanes (G = 8, 16 or 32 >= C) owns one touched cell; lane j owns class j, so the
// 8*C-byte row of the grid is read and written coalesced.
// `counter` points at this frame's touched count; `next_counter` (the other half of the double
// buffer) is zeroed for the next frame.
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kThreads)
k_apply(double* __restrict__ map, uint32_t* __restrict__ mask, const uint32_t* __restrict__ touched,
        const uint32_t* __restrict__ counter, uint32_t* __restrict__ next_counter,
        const double* __restrict__ cm, int c, int lane_cls) {
    extern __shared__ double s_cm[];  // C x C
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) s_cm[i] = cm[i];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_counter = 0u;

    const uint32_t count = *counter;
    const int j = threadIdx.x % G;
    const uint32_t groups_per_block = kThreads / G;
    const uint32_t stride = gridDim.x * groups_per_block;
    // every lane of a warp runs the same number of iterations (count is uniform, groups differ only in t)
    const uint32_t iters = (count + stride - 1) / stride;
    uint32_t t = blockIdx.x * groups_per_block + threadIdx.x / G;
    for (uint32_t it = 0; it < iters; ++it, t += stride) {
        const bool active = t < count;
        uint32_t cell = 0, bits = 0;
        if (active) {
            cell = touched[t];
            bits = mask[cell];
        }
        __syncwarp();  // all lanes of the group have read the mask before it is cleared
        if (active) {
            if (j == 0) mask[cell] = 0u;
            if (j < c) {
                double* p = map + (size_t)cell * c + j;
                double acc = *p;
                for (int i = 0; i < c; ++i) {
                    if ((bits >> i) & 1u) {
                        acc = __dadd_rn(acc, s_cm[j * c + i]);
                        if (i == lane_cls && j == i && (bits & kBoostBit)) acc = __dadd_rn(acc, 2.0);
                    }
                }
                *p = acc;
            }
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
