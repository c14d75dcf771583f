// Device-side arithmetic shared by every kernel of the semantic-mapping path.
//
// The per-point rule is the specification in SURVEY.md section 9, i.e. the reference's
// project_pcd + update_map (src/mapping_replay.py:214-301) restated per point:
// IEEE double throughout, the two matrix products as explicit fused chains (what
// np.matmul -> OpenBLAS dgemm computes), un-fused add/sub/div elsewhere.  All double
// operations go through the _rn intrinsics so that nvcc can neither contract nor
// reassociate them, whatever -fmad says.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smap {

// -DSMAP_DEBUG_BOUNDS (dev build, tools/bounds_check.sh): every index the kernels form for a stack push, a label load, a
// mask / tag / grid update or a staged tile is checked against its array; violations are counted (never trapped) and
// read back with smap_debug_bounds.  compute-sanitizer is not available on the GPU pool; this is the stand-in.
#ifdef SMAP_DEBUG_BOUNDS
__device__ unsigned long long g_bounds[2];   // violations, code of the first one
#define SMAP_BOUNDS(cond, code)                                                   \
    do {                                                                           \
        if (!(cond)) {                                                             \
            if (atomicAdd(&g_bounds[0], 1ull) == 0ull) g_bounds[1] = (code);       \
        }                                                                          \
    } while (0)
#else
#define SMAP_BOUNDS(cond, code) do { } while (0)
#endif

constexpr int kMaxBatch = 16;  // frames per launch (one cell-mask slot each)

// Per-frame projection constants (kernel parameter -> copied to shared memory once per block and frame).
struct FrameParams {
    double T[16];      // world -> velodyne, row-major (src/mapping_replay.py:225-226)
    double P[12];      // camera projection (src/camera.py:28)
    double range_max;  // cfg.MAPPING.PCD.RANGE_MAX
    // ---- certified fast projection (see fast_project): M = P * T composed in double on the host
    double M[12];      // rows q0, q1, q2
    double e3x4;       // 4 * e[2]: the depth must exceed this for the fast path
    double cgu, cgv;   // guard slopes: (4/3) (e_row + Umax e_depth)
    double c0;         // guard offset: Umax * 2^-44
    double img_wd, img_hd;
    double img_wd1, img_hd1;   // W + 1, H + 1
    // ---- float32 pre-cull (see precull_pass): row 0 = T row 0 (velodyne x), rows 1..3 = rows of P*T
    float Mf[16];
    float Ea[4];       // kCullSlack * max(|m_r0|, |m_r1|, |m_r2|)   error bound of row r = Ea[r] * (|x|+|y|+|z|) + Eb[r]
    float Eb[4];       // kCullSlack * |m_r3|
    float range_hi;    // range_max * (1 + kCullSlack)
    float img_wf, img_hf;
    int has_T;         // 0: cloud already in the velodyne frame
    int img_w, img_h;  // image.shape[1], image.shape[0]
    int pad;
};

constexpr float kCullSlack = 1e-6f;
constexpr double kCoordBound = 1048576.0;  // fast-path error bounds assume |x|, |y|, |z| < 2^20 m; beyond: exact path

// Grid + class constants of a mapper handle.
struct GridParams {
    double off_x, off_y;  // pcd origin w.r.t. map origin (src/mapping_replay.py:261)
    double bx0, by0;      // cfg.MAPPING.BOUNDARY[0][0], [1][0]
    double res;           // cfg.MAPPING.RESOLUTION
    double rinv;          // fl(1 / res) for the certified fast cell index
    double mh_d1, mw_d1;  // MH + 1, MW + 1
    int mh, mw, c;
    int lane;             // class index named "lane", or -1
    int use_intensity;
    int pad;
    uint8_t col_r[32];    // cfg.LABEL_COLORS[:, 0]
    uint8_t col_g[32];    // cfg.LABEL_COLORS[:, 1]   (blue is never compared, src/mapping_replay.py:276)
};

__device__ __forceinline__ double dot4(const double* __restrict__ r, double a, double b, double c, double w) {
    return __fma_rn(r[3], w, __fma_rn(r[2], c, __fma_rn(r[1], b, __dmul_rn(r[0], a))));
}

// float64 -> int32 the way numpy does it, followed by the reference's 0 <= i < n test:
// truncation toward zero (so (-1, 0) lands in 0), NaN / inf / overflow rejected.
__device__ __forceinline__ bool trunc_in_range(double g, int n, int& out) {
    bool ok = (g > -1.0) && (g < (double)n);
    out = ok ? __double2int_rz(g) : 0;
    return ok;
}

// src/mapping_replay.py:223-240 for one point.  Returns true when the point survives the range and
// frustum culls; (iu, iv) is its pixel.
__device__ __forceinline__ bool project_point(const FrameParams& f, double x, double y, double z, int& iu, int& iv) {
    double vx, vy, vz, vw;
    if (f.has_T) {
        vx = dot4(f.T + 0, x, y, z, 1.0);
        vy = dot4(f.T + 4, x, y, z, 1.0);
        vz = dot4(f.T + 8, x, y, z, 1.0);
        vw = dot4(f.T + 12, x, y, z, 1.0);
    } else {
        vx = x; vy = y; vz = z; vw = 1.0;
    }
    const double q0 = dot4(f.P + 0, vx, vy, vz, vw);
    const double q1 = dot4(f.P + 4, vx, vy, vz, vw);
    const double q2 = dot4(f.P + 8, vx, vy, vz, vw);
    const double u = __ddiv_rn(q0, q2);
    const double v = __ddiv_rn(q1, q2);
    const bool in_front = (0.0 < vx) && (vx < f.range_max);
    const bool a = trunc_in_range(u, f.img_w, iu);
    const bool b = trunc_in_range(v, f.img_h, iv);
    return in_front && a && b;
}

// src/mapping_replay.py:261-268 for one point: map cell of world (x, y).
__device__ __forceinline__ bool cell_of(const GridParams& g, double x, double y, uint32_t& cell) {
    const double gx = __ddiv_rn(__dsub_rn(__dadd_rn(x, g.off_x), g.bx0), g.res);
    const double gy = __ddiv_rn(__dsub_rn(__dadd_rn(y, g.off_y), g.by0), g.res);
    int cx, cy;
    const bool a = trunc_in_range(gx, g.mh, cx);
    const bool b = trunc_in_range(gy, g.mw, cy);
    cell = (uint32_t)cx * (uint32_t)g.mw + (uint32_t)cy;
    return a && b;
}

__device__ __forceinline__ bool cell_xy(const GridParams& g, double x, double y, int& cx, int& cy) {
    const double gx = __ddiv_rn(__dsub_rn(__dadd_rn(x, g.off_x), g.bx0), g.res);
    const double gy = __ddiv_rn(__dsub_rn(__dadd_rn(y, g.off_y), g.by0), g.res);
    const bool a = trunc_in_range(gx, g.mh, cx);
    const bool b = trunc_in_range(gy, g.mw, cy);
    return a && b;
}

// src/mapping_replay.py:276 and :288-290: classes whose (R, G) equal the pixel's, plus the boost flag.
__device__ __forceinline__ uint32_t class_bits(const GridParams& g, uint8_t r, uint8_t gch, double intensity) {
    uint32_t bits = 0;
    for (int i = 0; i < g.c; ++i) bits |= (uint32_t)((r == g.col_r[i]) & (gch == g.col_g[i])) << i;
    if (g.use_intensity && g.lane >= 0 && ((bits >> g.lane) & 1u) && (intensity < 2.0 || intensity > 14.0))
        bits |= 1u << g.c;
    return bits;
}

// Same classes from two 256-entry tables (built once per block in shared memory):
// tab_r[v] has bit i set iff LABEL_COLORS[i].R == v, tab_g likewise for G.  Zero the tables, then one thread per
// class ORs its bit in (shared atomics): a handful of instructions per thread instead of a C-iteration loop for each
// of the 256 entries.  Contains one block barrier; the caller adds the one that publishes the tables.
__device__ __forceinline__ void build_color_tables(const GridParams& g, uint32_t* tab_r, uint32_t* tab_g) {
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        tab_r[v] = 0u;
        tab_g[v] = 0u;
    }
    __syncthreads();
    if ((int)threadIdx.x < g.c) {
        atomicOr(&tab_r[g.col_r[threadIdx.x]], 1u << threadIdx.x);
        atomicOr(&tab_g[g.col_g[threadIdx.x]], 1u << threadIdx.x);
    }
}

__device__ __forceinline__ uint32_t class_bits_lut(const GridParams& g, const uint32_t* tab_r, const uint32_t* tab_g,
                                                   uint8_t r, uint8_t gch, float intensity) {
    uint32_t bits = tab_r[r] & tab_g[gch];
    // float32 intensity compared as a double in the reference; 2 and 14 are exact in both, so the tests agree
    if (g.use_intensity && g.lane >= 0 && ((bits >> g.lane) & 1u) && (intensity < 2.0f || intensity > 14.0f))
        bits |= 1u << g.c;
    return bits;
}

// Conservative float32 cull.  Returns false only when the exact double-precision rule (project_point) is
// CERTAIN to reject the point, so that the expensive path runs on ~1/3 of the cloud.
// Row r of Mf evaluated in float differs from the exact value by at most
//     e_r = Ea[r] * (|x|+|y|+|z|) + Eb[r]   >=   kCullSlack * sum_i |m_ri| |x_i|
// where kCullSlack = 1e-6 covers the float rounding of the inputs (6e-8), of the composed matrix (6e-8) and
// of the four-term fused sum (< 2.4e-7) with a 2x margin for the comparisons below; the double chain's own
// rounding and the final division's (1e-16) disappear in it.  Branch-free: one predicate per point.
__device__ __forceinline__ bool precull_pass(const FrameParams& k, float x, float y, float z) {
    const float s = fabsf(x) + fabsf(y) + fabsf(z);
    float v[4], e[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        v[r] = fmaf(k.Mf[4 * r], x, fmaf(k.Mf[4 * r + 1], y, fmaf(k.Mf[4 * r + 2], z, k.Mf[4 * r + 3])));
        e[r] = fmaf(k.Ea[r], s, k.Eb[r]);
    }
    // anything not comfortably finite in float goes to the exact path (which rejects NaN / inf itself)
    const bool finite = (e[0] + e[1] + e[2] + e[3]) < 1e30f;
    const bool range_ok = (v[0] + e[0] > 0.0f) & (v[0] - e[0] < k.range_hi);
    // depth certainly positive: -q2 < q0 < W q2 and -q2 < q1 < H q2 must be possible; otherwise decided exactly
    const bool depth_pos = (v[3] - e[3]) > 0.0f;
    const float q2hi = (v[3] + e[3]) * (1.0f + kCullSlack);
    const bool u_ok = (v[1] + e[1] > -q2hi) & (v[1] - e[1] < k.img_wf * q2hi);
    const bool v_ok = (v[2] + e[2] > -q2hi) & (v[2] - e[2] < k.img_hf * q2hi);
    return !finite | (range_ok & (!depth_pos | (u_ok & v_ok)));
}

// ------------------------------------------------------------------------------------------------
// Certified fast path ("filtered predicate"): the integer results the reference produces -- keep / drop,
// pixel (iu, iv), cell (cx, cy) -- are floors of real quantities; a cheaper evaluation with a rigorous error
// bound yields the same integers unless the quantity sits within the bound of an integer.  Only then is the
// reference's own rounding chain (project_point / cell_xy) evaluated.  Results are bit-identical by construction.
//
// Projection: q~ = M p with M = P T composed on the host (12 FMAs instead of 28).  Both q~ and the reference
// chain q^ approximate the exact product within 8.1 u sum_j (|P||T|)_rj |x_j|, so |q~ - q^| <= e_r with
// e_r = 64 u ((|P||T|)_r,xyz 3 B + (|P||T|)_r,w) for |x|,|y|,|z| < B (u = 2^-53; 4x safety factor).
// With depth |q~2| > 4 e_3:  |q~0/q~2 - q^0/q^2| <= (4/3)(e_1 + |u| e_3) / |q~2|.  The reciprocal (MUFU seed + two
// Newton steps) is good to 2^-45, the reference's own division to 2^-53.  Guard:
//      g = (cg * r + c0),  cg = (4/3)(e_row + Umax e_3),  c0 = Umax 2^-44,  Umax = max(W, H) + 2.
// The range test uses the velodyne x of the reference chain itself (same four operations), so it is exact.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double a) {
    // MUFU.RCP64H seed (only the high word is used: good to at least 2^-8), three Newton steps square the
    // error each time: 2^-8 -> 2^-16 -> 2^-32 -> 2^-64, i.e. limited by the roundings (a few 2^-53) < 2^-45
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    r = __fma_rn(r, __fma_rn(-a, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-a, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-a, r, 1.0), r);
    return r;
}

// floor of t when t is farther than `guard` from every integer; returns false when it is not.
// Valid for |t| < 2^31 (the callers check the range first).
__device__ __forceinline__ bool certified_floor(double t, double guard, int& k) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double rn = __dadd_rn(__dadd_rn(t, kMagic), -kMagic);
    const double d = __dadd_rn(t, -rn);
    k = __double2int_rn(rn) - (d < 0.0 ? 1 : 0);
    return fabs(d) > guard;
}

// Decisions are returned packed in one int: >= 0: kept, value = (row << 16 | col) style payload built by the
// caller; kDrop: dropped; kAsk: undecided, evaluate the reference chain.
constexpr int kDrop = -1;
constexpr int kAsk = -2;

// Fast version of project_point.  Returns (iv << 16 | iu) >= 0 when kept
// (images up to 32767 x 65535), kDrop, or kAsk.
__device__ __forceinline__ int fast_project(const FrameParams& f, double x, double y, double z, bool coords_ok) {
    const double vx = f.has_T ? dot4(f.T, x, y, z, 1.0) : x;   // the reference's own value
    if (!((0.0 < vx) && (vx < f.range_max))) return kDrop;      // src/mapping_replay.py:235, exact
    const double q0 = __fma_rn(f.M[0], x, __fma_rn(f.M[1], y, __fma_rn(f.M[2], z, f.M[3])));
    const double q1 = __fma_rn(f.M[4], x, __fma_rn(f.M[5], y, __fma_rn(f.M[6], z, f.M[7])));
    const double q2 = __fma_rn(f.M[8], x, __fma_rn(f.M[9], y, __fma_rn(f.M[10], z, f.M[11])));
    // the bounds hold for either sign of the depth (the reference has no depth test: a point behind the image
    // plane but in front of the LiDAR is projected like any other), as long as |q2| clears its own error
    if (!coords_ok || !(fabs(q2) > f.e3x4)) return kAsk;
    const double r = fast_rcp(q2);
    const double tu = __dmul_rn(q0, r), tv = __dmul_rn(q1, r);
    const double gu = __fma_rn(fabs(r), f.cgu, f.c0), gv = __fma_rn(fabs(r), f.cgv, f.c0);
    if (!(gu < 0.25 && gv < 0.25)) return kAsk;
    // farther than 1 outside the image: dropped for certain (the guards are < 1/4)
    if (!(tu > -2.0 && tu < f.img_wd1 && tv > -2.0 && tv < f.img_hd1)) return kDrop;
    int ku, kv;
    const bool su = certified_floor(tu, gu, ku), sv = certified_floor(tv, gv, kv);
    if (!(su && sv)) return kAsk;
    if (ku < -1 || ku >= f.img_w || kv < -1 || kv >= f.img_h) return kDrop;
    return (max(kv, 0) << 16) | max(ku, 0);   // (-1, 0) truncates to 0
}

// Fast version of cell_xy: n = (x + off) - b0 exactly as the reference, then n * fl(1/res) instead of n / res
// (they differ by at most 4 u |n / res| < 2^-19 for |n / res| < 2^31).  Returns the cell index, kDrop or kAsk.
__device__ __forceinline__ int fast_cell(const GridParams& g, double x, double y, int& cx, int& cy) {
    const double nx = __dsub_rn(__dadd_rn(x, g.off_x), g.bx0);
    const double ny = __dsub_rn(__dadd_rn(y, g.off_y), g.by0);
    const double tx = __dmul_rn(nx, g.rinv), ty = __dmul_rn(ny, g.rinv);
    if (!(tx > -2.0 && tx < g.mh_d1 && ty > -2.0 && ty < g.mw_d1))
        return (tx == tx && ty == ty) ? kDrop : kAsk;   // clearly off the grid; NaN: let the exact path decide
    const double kGuard = 1.9073486328125e-06;          // 2^-19
    const bool sx = certified_floor(tx, kGuard, cx), sy = certified_floor(ty, kGuard, cy);
    if (!(sx && sy)) return kAsk;
    if (cx < -1 || cx >= g.mh || cy < -1 || cy >= g.mw) return kDrop;
    cx = max(cx, 0);
    cy = max(cy, 0);
    return 0;
}

template <int LAYOUT>
__device__ __forceinline__ void load_point(const void* __restrict__ pts, int64_t ld, int64_t k, double& x, double& y,
                                           double& z, double& it) {
    if (LAYOUT == 0) {  // SMAP_PTS_F32X4
        const float4 p = __ldcs(reinterpret_cast<const float4*>(pts) + k);
        x = (double)p.x; y = (double)p.y; z = (double)p.z; it = (double)p.w;
    } else {            // SMAP_PTS_F64_SOA
        const double* p = reinterpret_cast<const double*>(pts);
        x = __ldcs(p + k); y = __ldcs(p + ld + k); z = __ldcs(p + 2 * ld + k); it = __ldcs(p + 3 * ld + k);
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

}  // namespace smap
