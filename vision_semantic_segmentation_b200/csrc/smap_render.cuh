// Rendering kernels (sm_100a): K4 apply_filter (cv2.filter2D 3x3 box, BORDER_REFLECT_101, src/renderer.py:175-189) and
// K5 render_bev_map (first-argmax colour, zero-sum cells black, src/renderer.py:32-59), fused.
//
// Exact forms reproduced: every tap is kf * p with kf = (double)(float)(1/9), un-fused,
// accumulated from 0 in row-major tap order; np.argmax lets the first maximum -- and the first NaN -- win; np.sum over
// the class axis is sequential below 8 addends and eight running partial sums above.
//
// Bound: HBM.  Algorithmic bytes MH MW C 8 (grid read once) + MH MW 3 (image), + MH MW C 8 when the filtered grid is
// wanted.  What round 1's kernel lost (11 % of the HBM peak) was instruction issue in the staging loop (a division and
// a remainder per element, scalar border logic for every element) and nine shared-memory loads per output; here
//   * a tile row is ONE contiguous run of the grid (C-contiguous cells), copied by one warp with coalesced 8-byte
//     loads and no index arithmetic (cells are padded to an odd number of doubles in shared memory only when C is
//     even, so that the 32 lanes of a warp -- 32 adjacent cells -- hit 32 distinct bank pairs); the at most two
//     reflected halo cells of a row are fetched separately; kf * p is formed once per element while staging;
//   * a thread owns a vertical strip of R cells and slides a 3x3 register window down it, one class at a time:
//     (R + 2) * 3 shared-memory loads per R outputs instead of 9 R;
//   * the colours leave through shared memory as aligned 32-bit stores.
#pragma once
#include "smap_device.cuh"

namespace smap {

#ifndef SMAP_RENDER_STRIPS
#define SMAP_RENDER_STRIPS 8
#endif
#ifndef SMAP_RENDER_R_SMALL
#define SMAP_RENDER_R_SMALL 4      // rows per strip when C < 8
#endif
#ifndef SMAP_RENDER_MINB
#define SMAP_RENDER_MINB 4         // resident blocks per SM the register allocation of the C < 8 kernels must allow
#endif
constexpr int kRX = 32;        // tile columns = lanes of a warp
constexpr int kRStrips = SMAP_RENDER_STRIPS;    // strips (= warps) per block
constexpr int kRThreads = kRX * kRStrips;
constexpr int kRSmall = SMAP_RENDER_R_SMALL;

struct RenderColors {
    uint8_t rgb[32 * 3];
};

// np.argmax / np.sum state of one cell while its class values stream by in ascending class order.
// NPS = 1: fewer than 8 classes (sequential sum); NPS = 8: numpy's eight running sums.
template <int NPS>
struct CellAcc {
    double mp;        // running maximum
    double r[NPS];    // running sum(s)
    int best;
    bool stop;        // a NaN has won: np.argmax stops looking
    __device__ __forceinline__ void track(int ch, double v) {
        if (ch == 0) {
            mp = v; best = 0; stop = (v != v);
        } else if (!stop && !(v <= mp)) {
            mp = v; best = ch; stop = (v != v);
        }
    }
};

__host__ __device__ constexpr int render_tile_rows(int r) { return kRStrips * r; }
// shared memory: the staged tile (with a one-cell halo when filtering) + the colour tile (+ one word of slack)
inline size_t render_smem_bytes(bool filter, int r, int cs) {
    const int h = filter ? 1 : 0;
    return sizeof(double) * (size_t)(render_tile_rows(r) + 2 * h) * (kRX + 2 * h) * cs + (size_t)render_tile_rows(r) * kRX * 3 + 8;
}

// Occupancy: the filtered C < 8 kernel wants 96 registers (two blocks = 16 warps per SM, 130 us for the 2000 x 2000 x 5
// grid); capped at 64 (four blocks, 8 bytes of spill) it takes 92 us.  The C >= 8 kernels keep their registers: with
// the same cap they spill inside the class loop (290 -> 383 us at C = 19).  gpurun_out/r2f_render_variants.log
template <bool FILTER, int R, int NPS>
__global__ void __launch_bounds__(kRThreads, NPS == 1 ? SMAP_RENDER_MINB : 1)
k_render(const double* __restrict__ map, int mh, int mw, int c, int cs, uint32_t div_c, const __grid_constant__ RenderColors colors,
         uint8_t* __restrict__ rgb, double* __restrict__ filtered) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int H = FILTER ? 1 : 0;
    constexpr int TY = kRStrips * R;
    double* const tile = reinterpret_cast<double*>(s_raw);
    const int pitch = (kRX + 2 * H) * cs;   // doubles per staged row
    uint8_t* const s_rgb = s_raw + sizeof(double) * (size_t)(TY + 2 * H) * pitch;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * kRX, y0 = blockIdx.y * TY;
    const double kf = (double)(1.0f / 9.0f);

    // ---- stage: warp w copies tile rows w, w + 8, ...
    const int gxa = max(x0 - H, 0), gxb = min(x0 + kRX - 1 + H, mw - 1);   // in-map columns of the tile (with halo)
    const int run = (gxb - gxa + 1) * c;                                    // contiguous doubles of one row
    const int dst0 = (gxa - (x0 - H)) * cs;
    for (int r = warp; r < TY + 2 * H; r += kRStrips) {
        const int gy = y0 - H + r;
        if (gy > mh - 1 + H) break;   // rows below the bottom halo are never read
        const int ys = reflect101(gy, mh);
        const double* src = map + ((size_t)ys * mw + gxa) * c;
        double* dst = tile + (size_t)r * pitch + dst0;
        if (cs == c) {
            for (int e = lane; e < run; e += 32) dst[e] = FILTER ? __dmul_rn(kf, __ldg(src + e)) : __ldg(src + e);
        } else {
            for (int e = lane; e < run; e += 32) {
                const int q = (int)__umulhi((uint32_t)e, div_c);   // e / c (exact for e < 2^16)
                dst[e + q] = FILTER ? __dmul_rn(kf, __ldg(src + e)) : __ldg(src + e);   // q * cs + e - q * c, cs = c + 1
            }
        }
        if (FILTER) {
            // the reflected halo cells (BORDER_REFLECT_101): column -1 -> 1, column mw -> mw - 2 (0 when mw == 1)
            if (x0 == 0 && lane < c)
                tile[(size_t)r * pitch + lane] = __dmul_rn(kf, __ldg(map + ((size_t)ys * mw + reflect101(-1, mw)) * c + lane));
            if (x0 + kRX >= mw && lane < c)
                tile[(size_t)r * pitch + (mw - (x0 - H)) * cs + lane] =
                    __dmul_rn(kf, __ldg(map + ((size_t)ys * mw + reflect101(mw, mw)) * c + lane));
        }
    }
    __syncthreads();

    // ---- compute: lane = column, warp = strip of R rows; classes stream through a sliding 3x3 register window
    const int x = x0 + lane;
    const int ys0 = y0 + warp * R;
    if (x < mw && ys0 < mh) {
        CellAcc<NPS> acc[R];
        double res[R];
        const double* base = tile + (size_t)(warp * R) * pitch + lane * cs;
        auto one_class = [&](int ch, auto&& consume) {
            const double* p = base + ch;
            if (FILTER) {
                double a0 = p[0], a1 = p[cs], a2 = p[2 * cs];
                double b0 = p[pitch], b1 = p[pitch + cs], b2 = p[pitch + 2 * cs];
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    if (ys0 + i < mh) {
                        const double* q = p + (size_t)(i + 2) * pitch;
                        const double c0 = q[0], c1 = q[cs], c2 = q[2 * cs];
                        double v = __dadd_rn(0.0, a0);
                        v = __dadd_rn(v, a1); v = __dadd_rn(v, a2);
                        v = __dadd_rn(v, b0); v = __dadd_rn(v, b1); v = __dadd_rn(v, b2);
                        v = __dadd_rn(v, c0); v = __dadd_rn(v, c1); v = __dadd_rn(v, c2);
                        if (filtered) filtered[((size_t)(ys0 + i) * mw + x) * c + ch] = v;
                        consume(i, v);
                        a0 = b0; a1 = b1; a2 = b2;
                        b0 = c0; b1 = c1; b2 = c2;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (ys0 + i < mh) consume(i, p[(size_t)i * pitch]);
            }
        };
        if (NPS == 1) {
            for (int ch = 0; ch < c; ++ch)
                one_class(ch, [&](int i, double v) {
                    acc[i].track(ch, v);
                    acc[i].r[0] = (ch == 0) ? __dadd_rn(0.0, v) : __dadd_rn(acc[i].r[0], v);
                });
#pragma unroll
            for (int i = 0; i < R; ++i) res[i] = acc[i].r[0];
        } else {
            const int main = c - (c % 8);
            for (int c8 = 0; c8 < main; c8 += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    one_class(c8 + j, [&](int i, double v) {
                        acc[i].track(c8 + j, v);
                        acc[i].r[j % NPS] = (c8 == 0) ? v : __dadd_rn(acc[i].r[j % NPS], v);
                    });
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                res[i] = __dadd_rn(__dadd_rn(__dadd_rn(acc[i].r[0], acc[i].r[1 % NPS]), __dadd_rn(acc[i].r[2 % NPS], acc[i].r[3 % NPS])),
                                   __dadd_rn(__dadd_rn(acc[i].r[4 % NPS], acc[i].r[5 % NPS]), __dadd_rn(acc[i].r[6 % NPS], acc[i].r[7 % NPS])));
            for (int ch = main; ch < c; ++ch)
                one_class(ch, [&](int i, double v) {
                    acc[i].track(ch, v);
                    res[i] = __dadd_rn(res[i], v);
                });
        }
        if (rgb) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (ys0 + i < mh) {
                    uint8_t* o = s_rgb + ((warp * R + i) * kRX + lane) * 3;
                    const bool black = res[i] == 0.0;
                    o[0] = black ? 0 : colors.rgb[3 * acc[i].best];
                    o[1] = black ? 0 : colors.rgb[3 * acc[i].best + 1];
                    o[2] = black ? 0 : colors.rgb[3 * acc[i].best + 2];
                }
            }
        }
    }
    if (!rgb) return;
    __syncthreads();
    // ---- colours out: warp w writes tile rows w, w + 8, ...: bytes up to the first 4-byte boundary, words, bytes
    const int nb = min(kRX, mw - x0) * 3;
    for (int r = warp; r < TY; r += kRStrips) {
        const int y = y0 + r;
        if (y >= mh) break;
        uint8_t* g = rgb + ((size_t)y * mw + x0) * 3;
        const uint8_t* s = s_rgb + r * kRX * 3;   // 4-byte aligned: kRX * 3 = 96
        int head = (int)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(g) & 3u)) & 3u);
        head = min(head, nb);
        if (lane < head) g[lane] = s[lane];
        const int nwords = (nb - head) >> 2;
        if (lane < nwords) {
            const int off = head + 4 * lane;
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(s) + (off >> 2);
            const uint32_t word = __funnelshift_r(sw[0], sw[1], (off & 3) * 8);   // sw[1]: at most the slack word
            *reinterpret_cast<uint32_t*>(g + off) = word;
        }
        const int t0 = head + 4 * nwords;
        if (lane < nb - t0) g[t0 + lane] = s[t0 + lane];
    }
}


// ------------------------------------------------------------------------------------------------
// k_render_bulk: the same computation with the tile staged by the TMA engine (round 2).
//
// k_render above reaches 29 % (filter + render) / 40 % (render only) of the HBM peak on the 2000 x 2000 x 5 grid: its
// staging loop keeps at most a few 256-byte requests per warp in flight, and only the warps that happen to be staging
// have any, so an SM never holds the ~44 KB in flight that 6.5 TB/s over 148 SMs at ~1 us of loaded latency asks for.
// Here a tile row -- one contiguous run of the C-contiguous grid -- is ONE bulk copy (cp.async.bulk.shared::cluster.global,
// completion counted in bytes on an mbarrier): the lanes of warp 0 issue the TY + 2 rows of the tile back to back, the
// whole 26 - 98 KB tile is in flight at once and no register or issue slot is spent on it.  Shared memory holds the
// rows exactly as they lie in HBM (no padding), so this path takes the grids whose cells are an ODD number of doubles
// (5 and 19 classes: 32 adjacent cells then fall on 32 distinct bank pairs) with an even number of columns and a
// 16-byte aligned base (every row run then starts and ends on a 16-byte boundary); everything else goes through
// k_render.  kf * p is formed when a value enters the 3 x 3 register window (three products per output instead of
// one per staged element: FP64 issue is not what bounds the kernel on B200).
//
// Slots of a staged row: slot s = column x0 - OFF + s, OFF = 2 when filtering (the halo column x0 - 1 is slot 1; the
// run starts one column further left so that it starts on an even column), 0 otherwise.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr int render_bulk_slots(bool filter) { return filter ? kRX + 4 : kRX; }
inline size_t render_bulk_smem_bytes(bool filter, int r, int c) {
    const int h = filter ? 1 : 0;
    return sizeof(double) * (size_t)(render_tile_rows(r) + 2 * h) * render_bulk_slots(filter) * c + (size_t)render_tile_rows(r) * kRX * 3 + 8;
}

// CT: the number of classes as a compile-time constant (5 and 19: the reference's two configurations), 0 = run-time.
// With a run-time class count two thirds of the kernel's instructions were index arithmetic (ncu, r2i: 53.5 M warp
// instructions for 20 M outputs, 21 % IMAD, 74 % issue-active); with CT every shared-memory address is an immediate.
// WF: the filtered grid is written as well (a template parameter, not a pointer test inside the unrolled loops).
template <bool FILTER, bool WF, int R, int NPS, int CT>
__global__ void __launch_bounds__(kRThreads, NPS == 1 ? SMAP_RENDER_MINB : 1)
k_render_bulk(const double* __restrict__ map, int mh, int mw, int c_arg, const __grid_constant__ RenderColors colors,
              uint8_t* __restrict__ rgb, double* __restrict__ filtered) {
    const int c = CT ? CT : c_arg;
    extern __shared__ __align__(128) unsigned char s_bulk[];
    __shared__ __align__(8) unsigned long long s_bar;
    constexpr int H = FILTER ? 1 : 0;
    constexpr int OFF = FILTER ? 2 : 0;
    constexpr int SL = render_bulk_slots(FILTER);
    constexpr int TY = kRStrips * R;
    double* const tile = reinterpret_cast<double*>(s_bulk);
    const int pitch = SL * c;   // doubles per staged row (even: SL is)
    uint8_t* const s_rgb = s_bulk + sizeof(double) * (size_t)(TY + 2 * H) * pitch;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * kRX, y0 = blockIdx.y * TY;
    const double kf = (double)(1.0f / 9.0f);

    // ---- stage: one bulk copy per tile row, all issued by warp 0
    const int xa = max(x0 - OFF, 0), xb = min(x0 - OFF + SL, mw);   // columns of the run (even .. even)
    const uint32_t row_bytes = (uint32_t)(xb - xa) * (uint32_t)c * 8u;
    int rows = TY + 2 * H;
    if (y0 - H + rows - 1 > mh - 1 + H) rows = mh + H - (y0 - H);   // rows below the bottom halo are never read
    const uint32_t bar = smem_addr32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)rows) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
            const int ys = reflect101(y0 - H + r, mh);
            const double* src = map + ((size_t)ys * mw + xa) * c;
            SMAP_BOUNDS(ys >= 0 && ys < mh && xa >= 0 && xb <= mw && xb > xa && (row_bytes & 15u) == 0u &&
                        ((reinterpret_cast<uintptr_t>(src)) & 15u) == 0u && r < TY + 2 * H &&
                        (xa - (x0 - OFF)) >= 0 && (xa - (x0 - OFF)) + (xb - xa) <= SL, 301);
            const uint32_t dst = smem_addr32(tile + (size_t)r * pitch + (xa - (x0 - OFF)) * c);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
        }
    }
    __syncthreads();   // the barrier is initialised for every waiter
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    }
    if (FILTER && (x0 == 0 || x0 + kRX >= mw)) {
        // the reflected halo columns (BORDER_REFLECT_101): column -1 = column 1, column mw = column mw - 2
        for (int i = threadIdx.x; i < rows * c; i += kRThreads) {
            const int r = i / c, k = i - r * c;
            double* row = tile + (size_t)r * pitch;
            if (x0 == 0) row[1 * c + k] = row[(reflect101(-1, mw) + OFF) * c + k];
            if (x0 + kRX >= mw) row[(mw - x0 + OFF) * c + k] = row[(reflect101(mw, mw) - x0 + OFF) * c + k];
        }
        __syncthreads();
    }

    // ---- compute: lane = column, warp = strip of R rows; classes stream through a sliding 3x3 register window.
    // Straight-line code: every lane of a strip that has at least one row inside the map computes all R rows (rows and
    // columns beyond the map read staged bytes nobody wrote -- inside the tile's allocation -- and are dropped at the
    // stores), the argmax / sum state is updated with selects.  With the bottom-row, filtered-pointer and argmax
    // branches inside the unrolled loops half of the executed instructions were moves and branches (ncu, r2k).
    const int x = x0 + lane;
    const int ys0 = y0 + warp * R;
    if (ys0 < mh) {
        double mp[R], rs[R][NPS], res[R];
        int best[R];
#pragma unroll
        for (int i = 0; i < R; ++i) { mp[i] = 0.0; best[i] = 0; }
        const bool col_ok = x < mw;
        const double* base = tile + (warp * R) * pitch + (lane + OFF - H) * c;
        // np.argmax: the first maximum wins, and so does the first NaN (mp != mp from then on)
        auto track = [&](int i, int ch, double v) {
            const bool take = (ch == 0) | (!(v <= mp[i]) & (mp[i] == mp[i]));
            mp[i] = take ? v : mp[i];
            best[i] = take ? ch : best[i];
        };
        auto one_class = [&](int ch, auto&& consume) {
            const double* p = base + ch;
            if (FILTER) {
                double a0 = __dmul_rn(kf, p[0]), a1 = __dmul_rn(kf, p[c]), a2 = __dmul_rn(kf, p[2 * c]);
                double b0 = __dmul_rn(kf, p[pitch]), b1 = __dmul_rn(kf, p[pitch + c]), b2 = __dmul_rn(kf, p[pitch + 2 * c]);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const double* q = p + (i + 2) * pitch;
                    const double c0 = __dmul_rn(kf, q[0]), c1 = __dmul_rn(kf, q[c]), c2 = __dmul_rn(kf, q[2 * c]);
                    double v = __dadd_rn(0.0, a0);
                    v = __dadd_rn(v, a1); v = __dadd_rn(v, a2);
                    v = __dadd_rn(v, b0); v = __dadd_rn(v, b1); v = __dadd_rn(v, b2);
                    v = __dadd_rn(v, c0); v = __dadd_rn(v, c1); v = __dadd_rn(v, c2);
                    if (WF) {
                        if (col_ok && ys0 + i < mh) filtered[((size_t)(ys0 + i) * mw + x) * c + ch] = v;
                    }
                    consume(i, v);
                    a0 = b0; a1 = b1; a2 = b2;
                    b0 = c0; b1 = c1; b2 = c2;
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i) consume(i, p[i * pitch]);
            }
        };
        if (NPS == 1) {
#pragma unroll(CT ? CT : 1)
            for (int ch = 0; ch < c; ++ch)
                one_class(ch, [&](int i, double v) {
                    track(i, ch, v);
                    rs[i][0] = __dadd_rn(ch == 0 ? 0.0 : rs[i][0], v);
                });
#pragma unroll
            for (int i = 0; i < R; ++i) res[i] = rs[i][0];
        } else {
            const int main = c - (c % 8);
            for (int c8 = 0; c8 < main; c8 += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    one_class(c8 + j, [&](int i, double v) {
                        track(i, c8 + j, v);
                        rs[i][j % NPS] = (c8 == 0) ? v : __dadd_rn(rs[i][j % NPS], v);
                    });
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                res[i] = __dadd_rn(__dadd_rn(__dadd_rn(rs[i][0], rs[i][1 % NPS]), __dadd_rn(rs[i][2 % NPS], rs[i][3 % NPS])),
                                   __dadd_rn(__dadd_rn(rs[i][4 % NPS], rs[i][5 % NPS]), __dadd_rn(rs[i][6 % NPS], rs[i][7 % NPS])));
#pragma unroll(CT ? (CT % 8 ? CT % 8 : 1) : 1)
            for (int ch = main; ch < c; ++ch)
                one_class(ch, [&](int i, double v) {
                    track(i, ch, v);
                    res[i] = __dadd_rn(res[i], v);
                });
        }
        if (rgb) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                uint8_t* o = s_rgb + ((warp * R + i) * kRX + lane) * 3;   // rows / columns beyond the map: never copied out
                const bool black = res[i] == 0.0;
                o[0] = black ? 0 : colors.rgb[3 * best[i]];
                o[1] = black ? 0 : colors.rgb[3 * best[i] + 1];
                o[2] = black ? 0 : colors.rgb[3 * best[i] + 2];
            }
        }
    }
    if (!rgb) return;
    __syncwarp();
    // ---- colours out: every warp writes the R rows of its own strip (no block barrier: a warp that is done leaves):
    // bytes up to the first 4-byte boundary, words, bytes
    const int nb = min(kRX, mw - x0) * 3;
    for (int i = 0; i < R; ++i) {
        const int r = warp * R + i;
        const int y = y0 + r;
        if (y >= mh) break;
        uint8_t* g = rgb + ((size_t)y * mw + x0) * 3;
        const uint8_t* s = s_rgb + r * kRX * 3;
        int head = (int)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(g) & 3u)) & 3u);
        head = min(head, nb);
        if (lane < head) g[lane] = s[lane];
        const int nwords = (nb - head) >> 2;
        if (lane < nwords) {
            const int off = head + 4 * lane;
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(s) + (off >> 2);
            const uint32_t word = __funnelshift_r(sw[0], sw[1], (off & 3) * 8);
            *reinterpret_cast<uint32_t*>(g + off) = word;
        }
        const int t0 = head + 4 * nwords;
        if (lane < nb - t0) g[t0 + lane] = s[t0 + lane];
    }
}

}  // namespace smap
