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

// Cell-mask word (32 bit):  [ tag : 31-C bits | boost : 1 bit (bit C) | class bits : C bits ].
// The tag is the serial number of the frame that last wrote the word, so a mask never has to be cleared:
// a word whose tag is not the current frame's reads as empty.
constexpr int kMaxBatch = 16;  // frames per launch of the batched kernel (one mask slot each)

// Per-frame projection constants (passed by value as a kernel parameter -> constant bank).
struct FrameParams {
    double T[16];      // world -> velodyne, row-major (src/mapping_replay.py:225-226)
    double P[12];      // camera projection (src/camera.py:28)
    double range_max;  // cfg.MAPPING.PCD.RANGE_MAX
    // float32 pre-cull (see precull_pass): row 0 = T row 0 (velodyne x), rows 1..3 = rows of P*T (q0, q1, q2)
    float Mf[16];
    float Ea[4];       // kSlack * max(|m_r0|, |m_r1|, |m_r2|)   error bound of row r = Ea[r] * (|x|+|y|+|z|) + Eb[r]
    float Eb[4];       // kSlack * |m_r3|
    float range_hi;    // range_max * (1 + kSlack)
    float img_wf, img_hf;
    int has_T;         // 0: cloud already in the velodyne frame
    int img_w, img_h;  // image.shape[1], image.shape[0]
    int pad;
};

constexpr float kCullSlack = 1e-6f;

// Grid + class constants of a mapper handle.
struct GridParams {
    double off_x, off_y;  // pcd origin w.r.t. map origin (src/mapping_replay.py:261)
    double bx0, by0;      // cfg.MAPPING.BOUNDARY[0][0], [1][0]
    double res;           // cfg.MAPPING.RESOLUTION
    int mh, mw, c;
    int lane;             // class index named "lane", or -1
    int use_intensity;
    int tag_shift;        // C + 1: first bit of the frame tag in a cell-mask word
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
// tab_r[v] has bit i set iff LABEL_COLORS[i].R == v, tab_g likewise for G.
__device__ __forceinline__ void build_color_tables(const GridParams& g, uint32_t* tab_r, uint32_t* tab_g) {
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        uint32_t br = 0, bg = 0;
        for (int i = 0; i < g.c; ++i) {
            br |= (uint32_t)(g.col_r[i] == v) << i;
            bg |= (uint32_t)(g.col_g[i] == v) << i;
        }
        tab_r[v] = br;
        tab_g[v] = bg;
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
struct CullConsts {
    float m[16], ea[4], eb[4], range_hi, wf, hf;
};

__device__ __forceinline__ bool precull_pass(const CullConsts& k, float x, float y, float z) {
    const float s = fabsf(x) + fabsf(y) + fabsf(z);
    float v[4], e[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        v[r] = fmaf(k.m[4 * r], x, fmaf(k.m[4 * r + 1], y, fmaf(k.m[4 * r + 2], z, k.m[4 * r + 3])));
        e[r] = fmaf(k.ea[r], s, k.eb[r]);
    }
    // anything not comfortably finite in float goes to the exact path (which rejects NaN / inf itself)
    const bool finite = (e[0] + e[1] + e[2] + e[3]) < 1e30f;
    const bool range_ok = (v[0] + e[0] > 0.0f) & (v[0] - e[0] < k.range_hi);
    // depth certainly positive: -q2 < q0 < W q2 and -q2 < q1 < H q2 must be possible; otherwise decided exactly
    const bool depth_pos = (v[3] - e[3]) > 0.0f;
    const float q2hi = (v[3] + e[3]) * (1.0f + kCullSlack);
    const bool u_ok = (v[1] + e[1] > -q2hi) & (v[1] - e[1] < k.wf * q2hi);
    const bool v_ok = (v[2] + e[2] > -q2hi) & (v[2] - e[2] < k.hf * q2hi);
    return !finite | (range_ok & (!depth_pos | (u_ok & v_ok)));
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
