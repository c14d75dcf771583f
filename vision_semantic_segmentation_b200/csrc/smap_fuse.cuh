// k_fuse: the fused project -> cull -> label lookup -> cell -> update kernel for float4 clouds (sm_100a).
//
// Same per-point rule as k_stream_soa (SURVEY.md section 9 = src/mapping_replay.py:214-301), but every point is first
// decided in FLOAT32 with a rigorous error bound ("filtered predicate"), because the kernel is bound by
// instruction issue, not by HBM: the integers the reference produces -- keep / drop, pixel (iu, iv), cell
// (cx, cy) -- are floors of real quantities, and a float32 evaluation in coordinates re-centred on the vehicle
// yields the same integers unless a quantity lies within its error bound of an integer.  Those points (2.7 %)
// are set aside on a second per-warp stack and decided by the float64 certified path of smap_device.cuh
// (fast_project / fast_cell, themselves backed by the reference's own rounding chain), 32 at a time.
// The results are therefore bit-identical to the reference by construction; the float32 arithmetic only
// decides how a point is ROUTED.
//
// Error analysis (u = 2^-24; all constants are composed in double on the host, see fill_fast32 in smap.cu).
//   re-centring   xl = fl(x - cx) = D (1 + d), |d| <= u  (D = x - cx exactly; x and cx are float32 values)
//   row r         q~ = fma(a0, xl, fma(a1, yl, fma(a2, zl, b))) with a_j = fl32(A_j), b = fl32(beta); against the
//                 exact Q = sum A_j D_j + beta:  |q~ - Q| <= 5.1 u (sum |A_j| |D_j| + |beta|)
//                 -> E_r(rho) = 6 u (amax_r rho + |beta_r|) + e_ref_r,   rho = |xl| + |yl| + |zl|
//                 (e_ref_r: what the reference's own float64 chain and the host's composition may differ from Q)
//   quotient      us = q0' r~ (kept unrounded inside two FMAs), r~ = MUFU.RCP(q2~) (<= 2 ulp), q0' = q0 - q2 / 2, so
//                 that us ~ u - 1/2 and RN(us) = floor(u) away from the integers.  With |q2~| > 4 E_2:
//                 |us - (u_ref - 1/2)| <= (4/3)(E_0' + Umax E_2) |r~| + Umax 2^-21  =: g
//                 certified  <=>  |us - RN(us)| < 1/2 - g   (implies |q2~| > 4 E_2, see smap.cu)
//   range         |d~ - vx_ref| <= E_d(rho);  certified inside  <=>  |d~ - R/2| < R/2 - E_d
//   cell          ts = fma(xl, fl32(1/res), f0) ~ gx_ref - I0 - 1/2,  |ts - ...| <= 3.5 u rho / res + c
//
// Variants measured and NOT kept (numbers in profiles/r1*_ncu_summary.md, profiles/r2_sweep*.log; the code is in the
// history up to commit cf06422): a per-warp TMA stage ring for the cloud, one persistent launch per batch, launches
// with blockIdx.y = frame, register-carried software pipelines of the label loads / atomics, L2 / L1 prefetch hints
// for the label image, the cloud and the records' label bytes; round 2 (profiles/r2_rejected_variants.md): the cull as
// scalar FFMAs with constant-bank operands, the cull's constants read from shared memory (LDS.128), third grids on three
// or six streams, the running box in uniform registers for the count update.
// ------------------------------------------------------------------------------------------------
#pragma once
#include "smap_kernels.cuh"

namespace smap {

// Per-frame constants of the float32 path.  float2 members are operand pairs of the packed FFMA2 / FADD2
// instructions (two rows per instruction); as kernel parameters they live in the constant bank.
struct Fast32 {
    // ---- conservative cull, world coordinates: rows {velodyne x, q2} and {q0, q1}, columns x, y, z, 1
    float2 c_dc[4];
    float2 c_ab[4];
    float2 c_wh;           // {W, H}
    float c_bw;            // a coordinate beyond this magnitude: not culled here (decided downstream)
    float c_rh, c_rthr;    // R / 2,  R / 2 + E_d
    float c_depth;         // q2 above this: depth certainly positive
    float c_lo_u, c_hi_u, c_lo_v, c_hi_v;   // pass iff  a + c > c_lo_u,  W c - a > c_hi_u, ... (all <= 0)
    // ---- certified float32 decisions, coordinates re-centred on the vehicle
    float2 n_ctr_xy;       // {-cx, -cy}
    float n_ctr_z;         // -cz
    float coord_l;         // rho must stay below this (<= 0 switches the float32 path off)
    float2 d_uv[4];        // rows {q0 - q2 / 2, q1 - q2 / 2}
    float2 d_cd[4];        // rows {q2, velodyne x}
    float g_k1, g_k0, g_hc;        // 1/2 - g = fma(-(fma(g_k1, rho, g_k0)), |r|, g_hc)
    float r_h, r_kd, r_thr0;       // |d - r_h| < fma(-r_kd, rho, r_thr0)
    float2 mid_uv, half_uv;        // |RN - mid| <= half  <=>  floor in [-1, W - 1] (resp. H - 1)
    float2 cell_f0;                // fractional parts of the centre's cell coordinate, minus 1/2
    float cell_rf;                 // fl32(1 / res)
    float cell_kc, cell_hg0;       // 1/2 - g_cell = fma(-cell_kc, rho, cell_hg0)
    float2 mid_c, half_c;          // floor + I0 in [-1, MH - 1] (resp. MW - 1)
    float2 clamp_c;                // magic - I0: lower clamp of the magic-shifted cell coordinate (cell >= 0)
    uint32_t pix_k;                // pixel = bits(tv) * W + bits(tu) + pix_k    (mod 2^32)
    uint32_t cell_k;               // cell  = bits(tx) * MW + bits(ty) + cell_k  (mod 2^32)
    uint32_t tag;                  // count update: this frame's tag (strictly increasing per tag plane)
    uint32_t pad;
};

constexpr float kMagic32 = 12582912.0f;   // 1.5 * 2^23: adding it rounds |t| < 2^22 to the nearest integer

// One frame = one launch: every per-frame constant below is a kernel parameter at a fixed offset of the constant bank.
struct FuseFrame {
    FrameParams fp;        // float64 constants (deferred points only)
    Fast32 fk;
    const float4* pts;
    const uint8_t* image;
    uint32_t* mask;        // MODE 0 / 2: this frame's mask slot
    int64_t n;
    int32_t per_warp;      // points per warp: ceil(n / (gridDim.x * kFWarps)), computed by the host
    int32_t img64;         // label image readable with aligned 8-byte loads (base aligned, size a multiple of 8)
    // FMT 1 (class-id plane at the network's resolution, smap.h SMAP_IMG_CLASS_IDS): camera pixel (u, v) reads
    // ids[nn(v) * src_w + nn(u)], nn = cv2.resize's INTER_NEAREST index map (min(floor(x * fl(1 / (W / w))), w - 1),
    // src/vision_semantic_segmentation_node.py:109-110).  The host tabulates the map in double as OpenCV does and hands it
    // over as a multiply-shift when one reproduces the whole table (checked exhaustively: at most 65535 entries), as
    // the table itself (u entries, then v entries, W and H of them) otherwise.
    const uint16_t* nn_tab;
    uint32_t nn_mx, nn_my; // nn(u) = (u * nn_mx) >> nn_sx when nn_tab == nullptr
    uint32_t nn_sx, nn_sy;
    int32_t src_w;
    int32_t pad;
};

// Index of the label byte(s) of camera pixel (iu, iv): the pixel itself for an RGB image, the nearest-neighbour source
// pixel of the class-id plane otherwise.
template <int FMT>
__device__ __forceinline__ uint32_t label_index(const FuseFrame& F, uint32_t iu, uint32_t iv) {
    if (FMT == 0) return iv * (uint32_t)F.fp.img_w + iu;
    uint32_t sx, sy;
    if (F.nn_tab) {
        // decide32 also evaluates lanes whose point it then rejects or defers: their (iu, iv) may be anything
        iu = min(iu, (uint32_t)F.fp.img_w - 1u);
        iv = min(iv, (uint32_t)F.fp.img_h - 1u);
        sx = __ldg(F.nn_tab + iu);
        sy = __ldg(F.nn_tab + (uint32_t)F.fp.img_w + iv);
    } else {
        sx = (iu * F.nn_mx) >> F.nn_sx;
        sy = (iv * F.nn_my) >> F.nn_sy;
    }
    return sy * (uint32_t)F.src_w + sx;
}

// Kernel parameter of k_fuse.
struct FuseLaunch {
    FuseFrame f;
    uint32_t* tags;           // MODE 1: this launch's tag plane, cells * (C + 1) uint32
    const uint32_t* id_lut;   // FMT 1: class bits of the 256 class ids (palette x cfg.LABEL_COLORS, folded by the host)
};

#ifndef SMAP_FUSE_ROUND
#define SMAP_FUSE_ROUND 2
#endif
#ifndef SMAP_FUSE_THREADS
#define SMAP_FUSE_THREADS 256   // threads per k_fuse block
#endif
#ifndef SMAP_FUSE_MINB
#define SMAP_FUSE_MINB (1024 / SMAP_FUSE_THREADS)   // resident blocks per SM: 32 warps, 64 registers per thread
#endif
constexpr int kFThreads = SMAP_FUSE_THREADS;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFRound = SMAP_FUSE_ROUND;
constexpr int kFRoundPts = 32 * kFRound;
constexpr int kFBlockRoundPts = kFWarps * kFRoundPts;
constexpr int kFQueueCap = 32 * kFRound + 32;      // survivor stack: < 32 left over + one round
constexpr int kFDeferCap = 64;
constexpr uint32_t kNone = 0xffffffffu;

#ifdef SMAP_FUSE_STATS   // diagnostic builds only (tools/fuse_stats.py): how the points were routed
__device__ unsigned long long g_fuse_stats[4];   // survivors of the cull, deferred to float64, float32-accepted, updates
#endif

// The conservative cull: false only when the reference rule is CERTAIN to drop the point.
__device__ __forceinline__ bool cull32(const Fast32& k, float x, float y, float z) {
    const float2 xx = make_float2(x, x), yy = make_float2(y, y), zz = make_float2(z, z);
    const float2 dc = __ffma2_rn(k.c_dc[0], xx, __ffma2_rn(k.c_dc[1], yy, __ffma2_rn(k.c_dc[2], zz, k.c_dc[3])));
    const float2 ab = __ffma2_rn(k.c_ab[0], xx, __ffma2_rn(k.c_ab[1], yy, __ffma2_rn(k.c_ab[2], zz, k.c_ab[3])));
    const float2 cc = make_float2(dc.y, dc.y);
    const float2 lo = __fadd2_rn(ab, cc);                                       // a + c, b + c
    const float2 hi = __ffma2_rn(k.c_wh, cc, make_float2(-ab.x, -ab.y));        // W c - a, H c - b
    bool p = (lo.x > k.c_lo_u) & (hi.x > k.c_hi_u) & (lo.y > k.c_lo_v) & (hi.y > k.c_hi_v);
    p |= !(dc.y > k.c_depth);                                                   // depth not certainly positive
    p &= fabsf(dc.x - k.c_rh) < k.c_rthr;                                       // NaN: dropped, as the reference does
    p |= fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > k.c_bw;
    return p;
}

// A deferred point: the float64 certified path (fast_project / fast_cell, exact chain behind them).  Returns
// {pixel index, cell index}; cell == kNone: dropped.  About 3 % of the cull's survivors come here.
// Inlined on purpose: `fp` and `gp` are kernel parameters at fixed offsets, so the float64 instructions take their
// constants straight from the constant bank; behind a call they were ~40 dependent generic loads per point.  Only
// the exact chains (a few points per 10 000) stay out of line.
template <int FMT>
__device__ __forceinline__ uint2 fuse_decide64(const FuseFrame& F, const GridParams& gp, float4 w) {
    const FrameParams& fp = F.fp;
    const double x = (double)w.x, y = (double)w.y, z = (double)w.z;
    const bool coords_ok = fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fabsf(w.z)) < (float)kCoordBound;
    int pix = fast_project(fp, x, y, z, coords_ok);
    if (pix == kAsk) pix = exact_project_slow(&fp, x, y, z);
    if (pix < 0) return make_uint2(0u, kNone);
    int cx = 0, cy = 0;
    const int on = fast_cell(gp, x, y, cx, cy);
    if (on == kAsk) {
        const long long c2 = exact_cell_slow(&gp, x, y);
        if (c2 < 0) return make_uint2(0u, kNone);
        cx = (int)(c2 >> 32); cy = (int)(c2 & 0xffffffffll);
    } else if (on == kDrop) {
        return make_uint2(0u, kNone);
    }
    return make_uint2(label_index<FMT>(F, (uint32_t)(pix & 0xffff), (uint32_t)(pix >> 16)),
                      (uint32_t)cx * (uint32_t)gp.mw + (uint32_t)cy);
}

#ifndef SMAP_FUSE_GATHER
#define SMAP_FUSE_GATHER 2
#endif
#ifndef SMAP_FUSE_UPDATE
#define SMAP_FUSE_UPDATE 1
#endif
constexpr int kFGather = SMAP_FUSE_GATHER;         // records per lane in one label-lookup pass
constexpr int kFUpdate = SMAP_FUSE_UPDATE;         // updates per lane in one tag + add pass (MODE 1)
constexpr int kFRecCap = 32 * kFGather + 64;       // record stack: < 32 * kFGather left over + 32 (float32) + 32 (float64)
constexpr int kFUpdCap = 32 * kFUpdate + 32 * kFGather;   // update stack: < 32 * kFUpdate left over + one lookup pass
// dynamic shared memory of k_fuse, per warp: the survivor stack, the deferred stack, the record stack, the update
// stack; then the two colour tables of the block.  Keeping this small matters beyond occupancy: the unified L1 /
// shared memory is carved in steps, and a block size that pushes the SM from the 196 KB to the 228 KB carve-out
// (28 KB of L1 left) cost 20 % in round 1.
__host__ __device__ constexpr int fuse_warp_smem(int mode) {
    return (kFQueueCap + kFDeferCap) * 16 + kFRecCap * 8 + (mode == 1 ? kFUpdCap * 8 : 0);
}
__host__ __device__ constexpr int fuse_block_smem(int mode) { return kFWarps * fuse_warp_smem(mode) + 2 * 256 * 4; }

// ------------------------------------------------------------------------------------------------
// MODE 0: masks only (RED.OR into the frame's slot, bounding box) -- k_apply replays the frames in order.
// MODE 2: count update through the masks (matrix == np.eye(C), grid of integer-valued counts): ATOM.OR returns what the
//         frame had already put in the cell, every NEWLY set class bit adds 1.0 to map[cell, class] (a newly set boost
//         bit 2.0 to map[cell, lane]) with a float64 RED; k_clear_masks zeroes the touched windows afterwards.  One
//         word per cell whatever the number of classes: used when C + 1 > 8, where MODE 1's per-class tags would
//         double the scattered traffic (C = 19: 45 us -> see DESIGN.md).
// MODE 1: count update (matrix == np.eye(C), grid of integer-valued counts): one uint32 tag per (cell, class) and one
//         per (cell, boost); ATOM.MAX with the frame's tag returns an older tag exactly once per frame, and that lane
//         adds 1.0 (boost: 2.0 on the lane class; src/mapping_replay.py:281,294) with a float64 RED.  Sums of small
//         integers are exact in any order.  Nothing to clear, no second kernel.  Launches that may overlap (different
//         internal streams) use different tag planes.
//
// Every warp is autonomous (no block barrier after the prologue).  A warp owns a contiguous, equally sized slice of
// the cloud, which it walks in rounds of kFRoundPts points:
//   cloud     the next round is prefetched into registers (LDG.128, streaming) while the current one is processed;
//   cull      conservative float32 test (cull32), survivors (~36 %) pushed on the warp's stack (ballot + popc);
//   decide    whenever >= 32 survivors are stacked, pop 32 -- one per lane, all lanes busy: float32 decisions; the
//             undecided points go to the deferred stack (decided in float64, 32 at a time), the accepted ones to
//             the record stack as {pixel, cell | intensity flag};
//   lookup    whenever >= 32 * kFGather records are stacked: all their label bytes are requested before the first is
//             used, class bits from the shared tables.  MODE 0 / 2 update the mask word right here.  MODE 1 pushes
//             the records whose pixel is a MAPPED class (5 of the 19 classes by default: a quarter of them) on the
//             update stack -- a third compaction, so that
//   update    (MODE 1) the tag atomics and the float64 REDs run with all 32 lanes busy instead of a quarter of them
//             (the divergent region cost the same ~75 instructions per pass whatever the number of active lanes),
//             and their round trip to L2 is paid once per 32 * kFUpdate updates, decoupled from the label loads'.
// ------------------------------------------------------------------------------------------------
//
// FMT 0: RGB label image, class bits = tabR[R] & tabG[G].  FMT 1: class-id plane (1 byte per network pixel), class
// bits = id_lut[id]; the only other difference is the label index (label_index<FMT>).
//
// `boxes`: MODE 0 / 2 the frame's bounding box (consumed and reset by k_apply / k_clear_masks, which also fold it into
// the handle's union window); MODE 1 the union window itself (never reset by a kernel: smap_clear does).
template <int MODE, int FMT = 0>
__global__ void __launch_bounds__(kFThreads, SMAP_FUSE_MINB)
k_fuse(const __grid_constant__ FuseLaunch B, const __grid_constant__ GridParams gp, FrameBox* __restrict__ boxes,
       double* __restrict__ map) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ int s_box[4];

    const FuseFrame& F = B.f;
    const Fast32& fk = B.f.fk;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int kFWarpSmem = fuse_warp_smem(MODE);
    unsigned char* const wbase = s_dyn + (size_t)warp * kFWarpSmem;
    float4* const queue = reinterpret_cast<float4*>(wbase);
    float4* const defer = queue + kFQueueCap;
    uint2* const recs = reinterpret_cast<uint2*>(defer + kFDeferCap);
    uint2* const upds = recs + kFRecCap;   // MODE 1 only
    uint32_t* const s_tab_r = reinterpret_cast<uint32_t*>(s_dyn + (size_t)kFWarps * kFWarpSmem);
    uint32_t* const s_tab_g = s_tab_r + 256;

    const int64_t gw = (int64_t)blockIdx.x * kFWarps + warp;   // this warp's index in the grid
    // this warp's slice of the cloud: [gw * per_warp, ...) clipped
    int w_pts;
    {
        const int64_t left = F.n - gw * F.per_warp;
        w_pts = left <= 0 ? 0 : (left < F.per_warp ? (int)left : F.per_warp);
    }
    // the first round is on its way while the block builds its colour tables
    const float4* gp_pts = F.pts + gw * F.per_warp + lane;
    float4 buf[kFRound];
#pragma unroll
    for (int j = 0; j < kFRound; ++j)
        buf[j] = (j * 32 + lane < w_pts) ? __ldcs(gp_pts + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);

    if (FMT == 0) {
        build_color_tables(gp, s_tab_r, s_tab_g);
    } else {
        for (int i = threadIdx.x; i < 256; i += kFThreads) s_tab_r[i] = __ldg(B.id_lut + i);
    }
    if (threadIdx.x == 0) box_reset(s_box);
    __syncthreads();

    const uint32_t c1 = (uint32_t)gp.c + 1u;
    const uint32_t lane_bit = (gp.use_intensity && gp.lane >= 0) ? (1u << gp.lane) : 0u;

    uint32_t qn = 0, dn = 0, rn = 0, un = 0;   // entries on the survivor / deferred / record / update stacks (warp-uniform)
    // bounding box of what this warp touched: magic-shifted floats (monotone in the cell coordinates)
    float fbx0 = 3.0e38f, fbx1 = -3.0e38f, fby0 = 3.0e38f, fby1 = -3.0e38f;
    // (the uniform-register box in MODE 1 / 2 as well: measured, no gain -- 13.01 vs 12.89 us, profiles/r2s_*)
    constexpr bool kBoxInRegs = MODE != 0;
    int ubx0 = 0x7fffffff, ubx1 = -1, uby0 = 0x7fffffff, uby1 = -1;   // MODE 0: the warp's box, warp-uniform

    // a decided point -> record stack: {pixel index, cell index | intensity flag << 31}
    auto push_record = [&](bool have, uint32_t pix, uint32_t cell, float it) {
        const unsigned ballot = __ballot_sync(0xffffffffu, have);
        if (have) {
            const uint32_t extreme = (it < 2.0f || it > 14.0f) ? 0x80000000u : 0u;   // src/mapping_replay.py:290
            SMAP_BOUNDS(rn + __popc(ballot & lt_mask) < (uint32_t)kFRecCap, 101);
            recs[rn + __popc(ballot & lt_mask)] = make_uint2(pix, cell | extreme);
        }
        rn += __popc(ballot);
    };

    // float32 decisions for up to 32 stacked survivors, one per lane
    auto decide32 = [&](uint32_t count) {
        const uint32_t first = qn - count;
        qn = first;
        bool defer_me = false, have = false;
        uint32_t pix = 0, cell = 0;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 tcb = make_float2(0.f, 0.f);   // the magic-shifted cell coordinates (MODE 0: for the box below)
        if ((uint32_t)lane < count) {
            w = queue[first + lane];
            const float2 lxy = __fadd2_rn(make_float2(w.x, w.y), fk.n_ctr_xy);
            const float lz = w.z + fk.n_ctr_z;
            const float rho = fabsf(lxy.x) + fabsf(lxy.y) + fabsf(lz);
            const float2 xx = make_float2(lxy.x, lxy.x), yy = make_float2(lxy.y, lxy.y), zz = make_float2(lz, lz);
            const float2 uv = __ffma2_rn(fk.d_uv[0], xx, __ffma2_rn(fk.d_uv[1], yy, __ffma2_rn(fk.d_uv[2], zz, fk.d_uv[3])));
            const float2 cd = __ffma2_rn(fk.d_cd[0], xx, __ffma2_rn(fk.d_cd[1], yy, __ffma2_rn(fk.d_cd[2], zz, fk.d_cd[3])));
            float rc;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(cd.x));
            // us = uv * rc is never rounded on its own: both uses are fused (one rounding each), which the error
            // bound covers either way.  (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
            // -fmad=false, so the fusion is spelled out instead of being left to the compiler.)
            const float2 rc2 = make_float2(rc, rc);
            const float hg = fmaf(-fmaf(fk.g_k1, rho, fk.g_k0), fabsf(rc), fk.g_hc);
            const float2 mg = make_float2(kMagic32, kMagic32), nmg = make_float2(-kMagic32, -kMagic32);
            float2 tp = __ffma2_rn(uv, rc2, mg);                               // RN(us) + magic
            const float2 rp = __fadd2_rn(tp, nmg);                             // RN(us), exact
            const float2 dp = __ffma2_rn(uv, rc2, make_float2(-rp.x, -rp.y));  // us - RN(us)
            const float2 op = __fadd2_rn(rp, make_float2(-fk.mid_uv.x, -fk.mid_uv.y));
            // cell
            const float2 ts = __ffma2_rn(lxy, make_float2(fk.cell_rf, fk.cell_rf), fk.cell_f0);
            const float hgc = fmaf(-fk.cell_kc, rho, fk.cell_hg0);
            float2 tc = __fadd2_rn(ts, mg);
            const float2 rcl = __fadd2_rn(tc, nmg);
            const float2 dcl = __fadd2_rn(ts, make_float2(-rcl.x, -rcl.y));
            const float2 oc = __fadd2_rn(rcl, make_float2(-fk.mid_c.x, -fk.mid_c.y));
            // every comparison is written so that a NaN anywhere means "not certified"
            bool cert = rho < fk.coord_l;
            cert &= fabsf(cd.y - fk.r_h) < fmaf(-fk.r_kd, rho, fk.r_thr0);
            cert &= (fabsf(dp.x) < hg) & (fabsf(dp.y) < hg);
            cert &= (fabsf(dcl.x) < hgc) & (fabsf(dcl.y) < hgc);
            const bool inside = (fabsf(op.x) <= fk.half_uv.x) & (fabsf(op.y) <= fk.half_uv.y) &
                                (fabsf(oc.x) <= fk.half_c.x) & (fabsf(oc.y) <= fk.half_c.y);
            defer_me = !cert;
            have = cert & inside;
            tp.x = fmaxf(tp.x, kMagic32); tp.y = fmaxf(tp.y, kMagic32);              // floor -1 -> pixel 0
            tc.x = fmaxf(tc.x, fk.clamp_c.x); tc.y = fmaxf(tc.y, fk.clamp_c.y);      // floor -1 -> cell 0
            if (FMT == 0) {
                pix = __float_as_uint(tp.y) * (uint32_t)F.fp.img_w + __float_as_uint(tp.x) + fk.pix_k;
            } else {
                constexpr uint32_t kMagicBits = 0x4B400000u;   // bits of kMagic32: RN(u - 1/2) = bits(tp) - kMagicBits
                pix = label_index<1>(F, __float_as_uint(tp.x) - kMagicBits, __float_as_uint(tp.y) - kMagicBits);
            }
            cell = __float_as_uint(tc.x) * (uint32_t)gp.mw + __float_as_uint(tc.y) + fk.cell_k;
            tcb = tc;
            if (kBoxInRegs && have) {
                fbx0 = fminf(fbx0, tc.x); fbx1 = fmaxf(fbx1, tc.x);
                fby0 = fminf(fby0, tc.y); fby1 = fmaxf(fby1, tc.y);
            }
        }
        if (!kBoxInRegs) {
            // MODE 0: the four running extremes do not fit next to the prefetched round (ptxas spilled a float4 of the
            // prefetch at the top of every round and waited for it right there): the batch's box is reduced over the
            // warp right here and kept in four warp-uniform integers (uniform registers)
            const uint32_t ox = __float_as_uint(fk.clamp_c.x), oy = __float_as_uint(fk.clamp_c.y);
            const int vx = (int)(__float_as_uint(tcb.x) - ox), vy = (int)(__float_as_uint(tcb.y) - oy);
            ubx0 = min(ubx0, __reduce_min_sync(0xffffffffu, have ? vx : 0x7fffffff));
            ubx1 = max(ubx1, __reduce_max_sync(0xffffffffu, have ? vx : -1));
            uby0 = min(uby0, __reduce_min_sync(0xffffffffu, have ? vy : 0x7fffffff));
            uby1 = max(uby1, __reduce_max_sync(0xffffffffu, have ? vy : -1));
        }
        push_record(have, pix, cell, w.w);
        const unsigned dballot = __ballot_sync(0xffffffffu, defer_me);
#ifdef SMAP_FUSE_STATS
        {
            const unsigned aballot = __ballot_sync(0xffffffffu, have);
            if (lane == 0) {
                atomicAdd(&g_fuse_stats[0], (unsigned long long)count);
                atomicAdd(&g_fuse_stats[1], (unsigned long long)__popc(dballot));
                atomicAdd(&g_fuse_stats[2], (unsigned long long)__popc(aballot));
            }
        }
#endif
        if (dballot) {
            SMAP_BOUNDS(!defer_me || dn + __popc(dballot & lt_mask) < (uint32_t)kFDeferCap, 102);
            if (defer_me) defer[dn + __popc(dballot & lt_mask)] = w;
            dn += __popc(dballot);
        }
    };

    // float64 decisions for up to 32 deferred points
    auto decide64 = [&](uint32_t count) {
        const uint32_t first = dn - count;
        dn = first;
        bool have = false;
        uint2 pc = make_uint2(0u, kNone);
        float it = 0.f;
        if ((uint32_t)lane < count) {
            const float4 w = defer[first + lane];
            it = w.w;
            pc = fuse_decide64<FMT>(F, gp, w);
            have = pc.y != kNone;
            if (have) {   // rare: straight into the block's box
                const int cx = (int)(pc.y / (uint32_t)gp.mw), cy = (int)(pc.y - (uint32_t)cx * (uint32_t)gp.mw);
                atomicMin(&s_box[0], cx); atomicMax(&s_box[1], cx);
                atomicMin(&s_box[2], cy); atomicMax(&s_box[3], cy);
            }
        }
        push_record(have, pc.x, pc.y, it);
    };

    // MODE 1: tag atomics + float64 REDs for up to 32 * kFUpdate stacked updates {cell | boost << 31, class bits}
    auto update = [&](uint32_t count) {
        const uint32_t first = un - count;
        un = first;
        uint32_t elem[kFUpdate], old0[kFUpdate], old1[kFUpdate];
        const uint32_t t = fk.tag;
#pragma unroll
        for (int k = 0; k < kFUpdate; ++k) {
            const uint32_t i = (uint32_t)(k * 32 + lane);
            elem[k] = 0u; old0[k] = t; old1[k] = t;
            if (i < count) {
                const uint2 u = upds[first + i];
                const uint32_t cell = u.x & 0x7fffffffu, bits = u.y;
                SMAP_BOUNDS(cell < (uint32_t)gp.mh * (uint32_t)gp.mw && bits < (1u << gp.c), 111);
                const bool boost = (bits & lane_bit) && (u.x >> 31);
                // element indices fit 32 bits (checked by the host)
                uint32_t* trow = B.tags + (size_t)(cell * c1);
                if (bits & (bits - 1u)) {
                    // several classes share this pixel's (R, G): rare, done in place
                    double* row = map + cell * (uint32_t)gp.c;
                    uint32_t b = bits;
                    while (b) {
                        const int c = __ffs(b) - 1;
                        b &= b - 1u;
                        if (atomicMax(trow + c, t) != t) atomicAdd(row + c, 1.0);
                    }
                    if (boost && atomicMax(trow + gp.c, t) != t) atomicAdd(row + gp.lane, 2.0);
                } else {
                    const uint32_t cls = (uint32_t)__ffs(bits) - 1u;
                    old0[k] = atomicMax(trow + cls, t);
                    if (boost) old1[k] = atomicMax(trow + gp.c, t);
                    elem[k] = cell * (uint32_t)gp.c + cls;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kFUpdate; ++k) {
            if (old0[k] != t) atomicAdd(map + elem[k], 1.0);
            if (old1[k] != t) atomicAdd(map + elem[k], 2.0);
        }
    };

    // label lookup for up to 32 * kFGather records, kFGather per lane; MODE 0 / 2: mask update as well
    auto lookup = [&](uint32_t count) {
        const uint32_t first = rn - count;
        rn = first;
        uint32_t cellf[kFGather], lr[kFGather], lg[kFGather];
#pragma unroll
        for (int k = 0; k < kFGather; ++k) {
            const uint32_t i = (uint32_t)(k * 32 + lane);
            cellf[k] = kNone;
            lr[k] = 0; lg[k] = 0;
            if (i < count) {
                const uint2 rec = recs[first + i];
                cellf[k] = rec.y;
                const uint32_t pix = rec.x;
                const uint8_t* img = F.image;
                if (FMT == 1) {
                    SMAP_BOUNDS(pix < (uint32_t)F.src_w * 65536u, 121);   // (the plane's height is not a kernel constant)
                    lr[k] = __ldg(img + pix);   // the class id
                } else if (F.img64) {
                    // R and G from ONE aligned 8-byte load (a second one only when R is the last byte of its 8: one
                    // pixel in eight): the L1 sees one sector request per point instead of two
                    const uint32_t a = pix * 3u;   // < 3 * 2^28
                    SMAP_BOUNDS(pix < (uint32_t)F.fp.img_w * (uint32_t)F.fp.img_h, 122);
                    const uint2 wv = __ldg(reinterpret_cast<const uint2*>(img + (a & ~7u)));
                    const uint32_t sh = (a & 7u) * 8u;
                    const uint64_t v = (((uint64_t)wv.y << 32) | wv.x) >> sh;
                    lr[k] = (uint32_t)v & 0xffu;
                    lg[k] = (sh == 56u) ? __ldg(img + a + 1) : (((uint32_t)v >> 8) & 0xffu);
                } else {
                    const size_t a = (size_t)pix * 3u;
                    lr[k] = __ldg(img + a);
                    lg[k] = __ldg(img + a + 1);
                }
            }
        }
        uint32_t bits[kFGather];
#pragma unroll
        for (int k = 0; k < kFGather; ++k) {
            if (FMT == 0) bits[k] = (cellf[k] != kNone) ? (s_tab_r[lr[k]] & s_tab_g[lg[k]]) : 0u;
            else bits[k] = (cellf[k] != kNone) ? s_tab_r[lr[k]] : 0u;
            if constexpr (MODE == 0) {
                // ordered update: fire-and-forget RED.OR into the frame's slot, nothing comes back
                if (!bits[k]) continue;
                const bool boost = (bits[k] & lane_bit) && (cellf[k] >> 31);
                SMAP_BOUNDS((cellf[k] & 0x7fffffffu) < (uint32_t)gp.mh * (uint32_t)gp.mw, 112);
                atomicOr(F.mask + (cellf[k] & 0x7fffffffu), boost ? (bits[k] | (1u << gp.c)) : bits[k]);
            }
        }
        if constexpr (MODE == 0) {
        } else if constexpr (MODE == 1) {
            // third compaction: only the records of mapped classes go on
#pragma unroll
            for (int k = 0; k < kFGather; ++k) {
                const unsigned ballot = __ballot_sync(0xffffffffu, bits[k] != 0u);
                SMAP_BOUNDS(!bits[k] || un + __popc(ballot & lt_mask) < (uint32_t)kFUpdCap, 103);
                if (bits[k]) upds[un + __popc(ballot & lt_mask)] = make_uint2(cellf[k], bits[k]);
                un += __popc(ballot);
            }
        } else {
            uint32_t want[kFGather], old[kFGather];
#pragma unroll
            for (int k = 0; k < kFGather; ++k) {
                want[k] = 0u; old[k] = 0u;
                if (!bits[k]) continue;
                const uint32_t cell = cellf[k] & 0x7fffffffu;
                const bool boost = (bits[k] & lane_bit) && (cellf[k] >> 31);
                want[k] = boost ? (bits[k] | (1u << gp.c)) : bits[k];          // the bits this point wants set
                SMAP_BOUNDS(cell < (uint32_t)gp.mh * (uint32_t)gp.mw, 113);
                old[k] = atomicOr(F.mask + cell, want[k]);                     // what the frame had set before
                cellf[k] = cell;
            }
            {
#pragma unroll
                for (int k = 0; k < kFGather; ++k) {
                    uint32_t fresh = want[k] & ~old[k];   // inactive slots: want == 0
                    if (!fresh) continue;
                    double* row = map + cellf[k] * (uint32_t)gp.c;   // element indices fit 32 bits (checked by the host)
                    if (fresh >> gp.c) {   // boost bit newly set: +2 on the lane class (src/mapping_replay.py:294)
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
    };

    // SMAP_ABL (dev builds only, results are WRONG): the stages after the cull switched off one by one, to see what each
    // costs in the real launch shape -- 1: no tag / grid update, 2: no label lookup either, 3: no decisions either
    // (stream + cull + compaction only).  profiles/r2g_ablation.md
    auto drain = [&](uint32_t count) {
#if defined(SMAP_ABL) && SMAP_ABL >= 3
        qn -= count;
        if (queue[qn + lane].x == 1.2345e-30f) fbx0 = 0.f;
        return;
#endif
        decide32(count);
        __syncwarp();
        if (dn >= 32u) {
            decide64(32u);
            __syncwarp();
        }
#if defined(SMAP_ABL) && SMAP_ABL >= 2
        if (rn >= 32u * kFGather) {
            rn -= 32u * kFGather;
            if (recs[rn + lane].x == 0xfffffff0u) fbx0 = 0.f;
        }
        return;
#endif
        if (rn >= 32u * kFGather) {
            lookup(32u * kFGather);
            __syncwarp();
            if (MODE == 1) {
                while (un >= 32u * kFUpdate) {
#if defined(SMAP_ABL) && SMAP_ABL >= 1
                    un -= 32u * kFUpdate;
                    if (upds[un + lane].x == 0xfffffff0u) fbx0 = 0.f;
#else
                    update(32u * kFUpdate);
#endif
                    __syncwarp();
                }
            }
        }
    };

    // ---- the cloud, in rounds; the next round is prefetched into registers while the current one is processed
    const int n_rounds = (w_pts + kFRoundPts - 1) / kFRoundPts;
    for (int r = 0; r < n_rounds; ++r) {
        const int left = w_pts - r * kFRoundPts;
        const int pts = left < kFRoundPts ? left : kFRoundPts;
        float4 nxt[kFRound];
#pragma unroll
        for (int j = 0; j < kFRound; ++j)
            nxt[j] = (kFRoundPts + j * 32 + lane < left) ? __ldcs(gp_pts + (r + 1) * kFRoundPts + j * 32)
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kFRound; ++j) {
            const float4 w = buf[j];
            const bool pass = cull32(fk, w.x, w.y, w.z) & (j * 32 + lane < pts);
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            SMAP_BOUNDS(!pass || qn + __popc(ballot & lt_mask) < (uint32_t)kFQueueCap, 104);
            if (pass) queue[qn + __popc(ballot & lt_mask)] = w;
            qn += __popc(ballot);
        }
        __syncwarp();
        while (qn >= 32u) drain(32u);
#pragma unroll
        for (int j = 0; j < kFRound; ++j) buf[j] = nxt[j];
    }
    // ---- flush: the survivors left over, the deferred points, the records, the updates
    if (qn) drain(qn);
    if (dn) {
        decide64(dn);
        __syncwarp();
    }
    while (rn) {
        lookup(rn < 32u * kFGather ? rn : 32u * kFGather);
        __syncwarp();
    }
    if (MODE == 1) {
        while (un) {
            update(un < 32u * kFUpdate ? un : 32u * kFUpdate);
            __syncwarp();
        }
    }

    // ---- bounding box: lane -> warp -> block (shared atomics) -> global
    if (!kBoxInRegs) {
        if (lane == 0 && ubx1 >= ubx0) {
            atomicMin(&s_box[0], ubx0); atomicMax(&s_box[1], ubx1);
            atomicMin(&s_box[2], uby0); atomicMax(&s_box[3], uby1);
        }
    } else if (__any_sync(0xffffffffu, fbx1 >= fbx0)) {
        // magic-shifted float -> cell coordinate: bits - (bits(magic) - I0); clamp_c = magic - I0
        const int ox = (int)__float_as_uint(fk.clamp_c.x), oy = (int)__float_as_uint(fk.clamp_c.y);
        const bool any = fbx1 >= fbx0;
        const int a = __reduce_min_sync(0xffffffffu, any ? (int)__float_as_uint(fbx0) - ox : 0x7fffffff);
        const int b = __reduce_max_sync(0xffffffffu, any ? (int)__float_as_uint(fbx1) - ox : -1);
        const int c = __reduce_min_sync(0xffffffffu, any ? (int)__float_as_uint(fby0) - oy : 0x7fffffff);
        const int d = __reduce_max_sync(0xffffffffu, any ? (int)__float_as_uint(fby1) - oy : -1);
        if (lane == 0) {
            atomicMin(&s_box[0], a); atomicMax(&s_box[1], b);
            atomicMin(&s_box[2], c); atomicMax(&s_box[3], d);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_box[1] >= s_box[0]) {
        atomicMin(&boxes->x0, s_box[0]); atomicMax(&boxes->x1, s_box[1]);
        atomicMin(&boxes->y0, s_box[2]); atomicMax(&boxes->y1, s_box[3]);
    }
}

// After a MODE 2 batch: zero every frame slot inside its frame's bounding box, fold the boxes into the handle's union
// window, reset the boxes of the next batch.  blockIdx.y = slot; the blocks of a slot stride over the rows of its box.
__global__ void __launch_bounds__(kThreads)
k_clear_masks(const __grid_constant__ ApplyParams ap, const FrameBox* __restrict__ boxes, FrameBox* __restrict__ next_boxes,
              unsigned long long* __restrict__ next_touched_total, FrameBox* __restrict__ ubox, int mw) {
    const int f = blockIdx.y;
    if (next_boxes && blockIdx.x == 0 && threadIdx.x == 0) {
        box_reset(&next_boxes[f].x0);
        if (f == 0) *next_touched_total = 0ull;
    }
    if (f >= ap.n_frames) return;
    const FrameBox b = boxes[f];
    if (b.x1 < b.x0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicMin(&ubox->x0, b.x0); atomicMax(&ubox->x1, b.x1);
        atomicMin(&ubox->y0, b.y0); atomicMax(&ubox->y1, b.y1);
    }
    uint32_t* const m = ap.mask[f];
    for (int x = b.x0 + (int)blockIdx.x; x <= b.x1; x += (int)gridDim.x) {
        uint32_t* row = m + (size_t)x * mw;
        for (int y = b.y0 + (int)threadIdx.x; y <= b.y1; y += kThreads) row[y] = 0u;
    }
}

}  // namespace smap
