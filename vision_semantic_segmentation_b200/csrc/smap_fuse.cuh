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
// ------------------------------------------------------------------------------------------------
#pragma once
#include "smap_kernels.cuh"

namespace smap {

// Per-frame constants of the float32 path.  float2 members are operand pairs of the packed FFMA2 / FADD2
// instructions (two rows per instruction); as kernel parameters they live in the constant bank and reach the
// packed instructions through uniform registers.
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

// One frame of a batch.
struct FuseFrame {
    FrameParams fp;        // float64 constants (deferred points only)
    Fast32 fk;
    const float4* pts;
    const uint8_t* image;
    uint32_t* mask;        // MODE 0: this frame's mask slot
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
    uint32_t pf_bytes;     // SMAP_FUSE_PF_IMAGE: size of the label image at pf_image (0: nothing to prefetch)
    const uint8_t* pf_image;   // SMAP_FUSE_PF_IMAGE: label image of a LATER frame of the batch, pulled into L2 by this launch
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

// Kernel parameter of k_fuse<MODE, NF>: the NF frames a launch walks.
//   NF == 1          one launch per frame (on alternating internal streams, so that the ramp-up and tail of one
//                    launch overlap the next one's body); every per-frame constant then sits at a fixed offset of the
//                    constant bank.  This is what the library uses (13.4 us / frame on the benchmark workload; 17.2
//                    at the time of the comparison below).
//   NF == kMaxBatch  ONE persistent launch walks all the frames of a batch: frame f + 1's cloud is already in flight
//                    while frame f's last survivors are decided, records and deferred points carry their frame
//                    index, nothing is flushed between frames.  Kept as a measured alternative (SMAP_FUSE_PERSISTENT):
//                    18.4 - 19.0 us / frame -- the run-time frame index turns every constant operand into an indexed LDC,
//                    which costs more than the per-launch ramp it saves.  (A third variant, one launch per batch with
//                    frame = blockIdx.y, measured 21.4 us / frame.)
template <int NF>
struct FuseBatchT {
    FuseFrame f[NF];
    uint32_t* tags;        // MODE 1: (cells * (C + 1), tag_planes) uint32; frame f of the launch uses plane f
    const uint32_t* id_lut;   // FMT 1: class bits of the 256 class ids (palette x cfg.LABEL_COLORS, folded by the host)
    int32_t n_frames;
    int32_t tag_planes;    // plane stride (>= n_frames)
};

constexpr int kFidShift = 28;   // a record's pixel index carries the frame index above bit 28

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
#ifndef SMAP_FUSE_GROUP
#define SMAP_FUSE_GROUP SMAP_FUSE_ROUND   // chunks culled back to back (unrolled) before the survivor stack is looked at
#endif
constexpr int kFGroup = SMAP_FUSE_GROUP;
static_assert(kFRound % kFGroup == 0, "a round is a whole number of groups");
#if !defined(SMAP_FUSE_TMA) || !SMAP_FUSE_TMA
static_assert(kFGroup == kFRound, "the register-prefetch path culls a whole round before it looks at the stack");
#endif
constexpr int kFQueueCap = 32 * kFGroup + 32;      // survivor stack: < 32 left over + one group
constexpr int kFDeferCap = 64;
constexpr uint32_t kNone = 0xffffffffu;

#ifdef SMAP_FUSE_STATS   // diagnostic builds only (tools/fuse_stats.py): how the points were routed
__device__ unsigned long long g_fuse_stats[4];   // survivors of the cull, deferred to float64, float32-accepted, unused
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
// Inlined on purpose: with one frame per launch `fp` and `gp` are kernel parameters at fixed offsets, so the float64
// instructions take their constants straight from the constant bank; behind a call they were ~40 dependent generic
// loads per point.  Only the exact chains (a few points per 10 000) stay out of line.
template <int FMT>
#ifdef SMAP_FUSE_DECIDE64_CALL
__device__ __noinline__
#else
__device__ __forceinline__
#endif
uint2 fuse_decide64(const FuseFrame& F, const GridParams& gp, float4 w) {
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

// ---- TMA (bulk async copy) + mbarrier plumbing of the per-warp cloud pipeline
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// (the bulk copies themselves -- cp.async.bulk global -> shared, 16-byte granular, completion counted in bytes on
// the stage's mbarrier, evict-first in L2 because the cloud is read once -- are issued inline in k_fuse)

// How the cloud reaches the cull.  Default: the next round is prefetched into registers with plain LDG.128 while the
// current one is processed.  -DSMAP_FUSE_TMA=1: a private ring of SMAP_FUSE_STAGES shared-memory stages per warp filled
// by TMA bulk copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first), one LDS.128 per point.  Both keep one
// round per warp in flight; the TMA ring costs 52 instructions of round set-up per 64 points (elect, expect_tx,
// descriptor operands through uniform registers, try_wait) against 25 for the register prefetch, and its stages push
// the block over a shared-memory carve-out step: 15.3 vs 14.5 us / frame (profiles/r1_sweep27.log).
#ifndef SMAP_FUSE_TMA
#define SMAP_FUSE_TMA 0
#endif
#ifndef SMAP_FUSE_STAGES
#define SMAP_FUSE_STAGES 2
#endif
constexpr int kFStages = SMAP_FUSE_TMA ? SMAP_FUSE_STAGES : 0;
// Experiments prepared from the stall samples of profiles/r1k_stall_breakdown.md (both off by default, not yet measured):
//   SMAP_FUSE_PF_IMAGE=1   every launch pulls the label image of the frame that will run after the one running beside
//                          it into L2 (one prefetch.global.L2 per lane at kernel start): the label gather then waits
//                          for L2 instead of DRAM, and the image costs sequential lines instead of scattered sectors
//   SMAP_FUSE_PF_CLOUD=D   lanes 0..7 prefetch the 1 KB of round r + D into L2 (D >= 2; round r + 1 is already on its
//                          way into registers): cull-only rounds are shorter than the DRAM latency
//   SMAP_FUSE_PF_LABEL=1|2 a record's label bytes are prefetched (1: into L2, 2: into L1) when the float32 decision
//                          pushes the record, one to two passes before the gather loads them (+1 instruction per
//                          batch of 32 survivors)
#ifndef SMAP_FUSE_PF_LABEL
#define SMAP_FUSE_PF_LABEL 0
#endif
#ifndef SMAP_FUSE_PF_IMAGE
#define SMAP_FUSE_PF_IMAGE 0
#endif
#ifndef SMAP_FUSE_PF_CLOUD
#define SMAP_FUSE_PF_CLOUD 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
#ifndef SMAP_FUSE_GATHER
#define SMAP_FUSE_GATHER 2
#endif
constexpr int kFGather = SMAP_FUSE_GATHER;         // records per lane in one label-lookup + update pass
constexpr int kFRecCap = 32 * kFGather + 64;       // record stack: < 32 * kFGather left over + 32 (float32) + 32 (float64)
// dynamic shared memory of k_fuse, per warp: kFStages cloud stages, the survivor stack, the deferred stack, the
// record stack, the deferred points' frame indices, the stages' mbarriers; then the two colour tables of the block
// (the deferred points' frame indices are only needed when a launch walks several frames.)  Keeping this small
// matters beyond occupancy: the unified L1 / shared memory is carved in steps, and a block size that pushes the SM
// from the 196 KB to the 228 KB carve-out (28 KB of L1 left) costs 20 % on the label gather.
__host__ __device__ constexpr int fuse_warp_smem(int nf) {
    return (kFStages * kFRoundPts + kFQueueCap + kFDeferCap) * 16 + kFRecCap * 8 + (nf > 1 ? kFDeferCap : 0) +
           (kFStages * 8 + 15) / 16 * 16;
}
__host__ __device__ constexpr int fuse_block_smem(int nf) { return kFWarps * fuse_warp_smem(nf) + 2 * 256 * 4; }

// ------------------------------------------------------------------------------------------------
// MODE 0: masks only (RED.OR into the frame's slot, bounding box) -- k_apply replays the frames in order.
// MODE 2: count update through the masks (matrix == np.eye(C), grid of integer-valued counts): ATOM.OR returns what the
//         frame had already put in the cell, every NEWLY set class bit adds 1.0 to map[cell, class] (a newly set boost
//         bit 2.0 to map[cell, lane]) with a float64 RED; k_clear_masks zeroes the touched windows afterwards.  One
//         word per cell whatever the number of classes: used when C + 1 > 8, where MODE 1's per-class tags would
//         double the scattered traffic (C = 19: 45 us -> see DESIGN.md).
// MODE 1: count update (matrix == np.eye(C), grid of integer-valued counts): one uint32 tag per (cell, class, frame
//         of the batch) and one per (cell, boost, frame); ATOM.MAX with the frame's tag returns an older tag exactly
//         once per frame, and that lane adds 1.0 (boost: 2.0 on the lane class; src/mapping_replay.py:281,294) with a
//         float64 RED.  Sums of small integers are exact in any order.  Nothing to clear, no second kernel.  The
//         frames of a batch use different tag planes (interleaved: the planes of one element share a sector), so
//         warps may be at different frames without any synchronisation.
//
// Every warp is autonomous (no block barrier after the prologue).  Per frame a warp owns a contiguous, equally sized
// slice of the cloud, which it walks in rounds of kFRoundPts points:
//   cloud     the next round is prefetched into registers (LDG.128, streaming) while the current one is processed;
//             -DSMAP_FUSE_TMA=1 selects a per-warp ring of shared-memory stages filled by TMA bulk copies instead
//             (measured: more round set-up instructions than it saves, see above);
//   cull      conservative float32 test (cull32), survivors (~36 %) pushed on the warp's stack (ballot + popc);
//   decide    whenever >= 32 survivors are stacked, pop 32 -- one per lane, all lanes busy: float32 decisions; the
//             undecided points go to the deferred stack (decided in float64, 32 at a time), the accepted ones to
//             the record stack as {pixel | frame, cell | intensity flag};
//   gather    whenever >= 32 * kFGather records are stacked: label lookup + update in straight-line code, all the
//             label bytes requested before the first is used, all the tag atomics issued before the first result
//             is used -- the two memory round trips are paid once per 32 * kFGather records.  (A software pipeline
//             that carried loads and atomics across loop iterations was tried first: ptxas puts every carried
//             operation on one scoreboard and waits for it at the loop head, which serialised everything.)
// ------------------------------------------------------------------------------------------------
//
// FMT 0: RGB label image, class bits = tabR[R] & tabG[G].  FMT 1: class-id plane (1 byte per network pixel), class
// bits = id_lut[id]; the only other difference is the label index (label_index<FMT>).
template <int MODE, int NF, int FMT = 0>
__global__ void __launch_bounds__(kFThreads, SMAP_FUSE_MINB)
k_fuse(const __grid_constant__ FuseBatchT<NF> B, const __grid_constant__ GridParams gp, FrameBox* __restrict__ boxes,
       double* __restrict__ map) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ int s_box[NF][4];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int kFWarpSmem = fuse_warp_smem(NF);
    unsigned char* const wbase = s_dyn + (size_t)warp * kFWarpSmem;
    float4* const stages = reinterpret_cast<float4*>(wbase);
    float4* const queue = stages + kFStages * kFRoundPts;
    float4* const defer = queue + kFQueueCap;
    uint2* const recs = reinterpret_cast<uint2*>(defer + kFDeferCap);
    uint8_t* const defer_f = reinterpret_cast<uint8_t*>(recs + kFRecCap);   // NF > 1 only
    uint64_t* const bars = reinterpret_cast<uint64_t*>(defer_f + (NF > 1 ? kFDeferCap : 0));
    (void)bars;
    uint32_t* const s_tab_r = reinterpret_cast<uint32_t*>(s_dyn + (size_t)kFWarps * kFWarpSmem);
    uint32_t* const s_tab_g = s_tab_r + 256;

    const int nf = (NF == 1) ? 1 : B.n_frames;   // NF == 1: every B.f[f] below is B.f[0], a fixed constant-bank offset
    const int64_t gw = (int64_t)blockIdx.x * kFWarps + warp;   // this warp's index in the grid
    // this warp's slice of frame f: [gw * per_warp, ...) clipped to the cloud
    auto slice_pts = [&](int f) -> int {
        const int64_t left = B.f[f].n - gw * B.f[f].per_warp;
        return left <= 0 ? 0 : (left < B.f[f].per_warp ? (int)left : B.f[f].per_warp);
    };

#if SMAP_FUSE_TMA
    // ---- TMA producer state (meaningful in lane 0): frame, source and points left of the next round to issue
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const uint32_t stg0 = smem_u32(stages), bar0 = smem_u32(bars);
    int i_f = -1, i_left = 0;
    const float4* i_src = nullptr;
    uint32_t i_st = 0;
    auto issue_round = [&]() {   // no-op once the batch is exhausted; issues exactly the rounds the consumer walks
        while (i_left <= 0 && i_f + 1 < nf) {
            ++i_f;
            i_left = slice_pts(i_f);
            i_src = B.f[i_f].pts + gw * B.f[i_f].per_warp;
        }
        if (i_left <= 0) return;
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(i_left < kFRoundPts ? i_left : kFRoundPts) * 16u;
            const uint32_t bar = bar0 + i_st * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(stg0 + i_st * (uint32_t)(kFRoundPts * 16)), "l"(i_src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
        }
        i_src += kFRoundPts;
        i_left -= kFRoundPts;
        i_st = (i_st + 1u) % (uint32_t)kFStages;
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kFStages; ++s) mbar_init(bars + s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kFStages - 1; ++r) issue_round();   // one more is issued at the top of every round
    // the first rounds are on their way while the block builds its colour tables
#endif
    if (FMT == 0) {
        build_color_tables(gp, s_tab_r, s_tab_g);
    } else {
        for (int i = threadIdx.x; i < 256; i += kFThreads) s_tab_r[i] = __ldg(B.id_lut + i);
    }
    if (threadIdx.x < NF) box_reset(s_box[threadIdx.x]);
    __syncthreads();

#if SMAP_FUSE_PF_IMAGE
    if (NF == 1 && B.f[0].pf_bytes) {
        const uint32_t lines = (B.f[0].pf_bytes + 127u) >> 7;
        const uint32_t lanes = gridDim.x * (uint32_t)kFThreads;
        for (uint32_t l = (uint32_t)gw * 32u + (uint32_t)lane; l < lines; l += lanes)
            prefetch_l2(B.f[0].pf_image + (size_t)l * 128u);
    }
#endif
    const uint32_t c1 = (uint32_t)gp.c + 1u;
    const uint32_t lane_bit = (gp.use_intensity && gp.lane >= 0) ? (1u << gp.lane) : 0u;

    uint32_t qn = 0, dn = 0, rn = 0;   // entries on the survivor / deferred / record stacks (warp-uniform)
    // MODE 0 / 2 bounding box of the current frame: magic-shifted floats (monotone in the cell coordinates)
    float fbx0 = 3.0e38f, fbx1 = -3.0e38f, fby0 = 3.0e38f, fby1 = -3.0e38f;

    // a decided point -> record stack: {pixel index | frame << 28, cell index | intensity flag << 31}
    auto push_record = [&](bool have, uint32_t pix_f, uint32_t cell, float it) {
        const unsigned ballot = __ballot_sync(0xffffffffu, have);
        if (have) {
            const uint32_t extreme = (it < 2.0f || it > 14.0f) ? 0x80000000u : 0u;   // src/mapping_replay.py:290
            recs[rn + __popc(ballot & lt_mask)] = make_uint2(pix_f, cell | extreme);
        }
        rn += __popc(ballot);
    };

    // float32 decisions for up to 32 stacked survivors of frame f, one per lane
    auto decide32 = [&](int f, uint32_t count) {
        const Fast32& fk = B.f[f].fk;
        const uint32_t first = qn - count;
        qn = first;
        bool defer_me = false, have = false;
        uint32_t pix = 0, cell = 0;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef SMAP_ABL_NO_DRAIN   // ablation builds (profiles/): streaming + cull + stack traffic alone
        if ((uint32_t)lane < count) {
            w = queue[first + lane];
            if (w.x == 1234.5f) atomicOr(B.f[f].mask, (uint32_t)w.z);
        }
        count = 0;
#endif
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
                pix = __float_as_uint(tp.y) * (uint32_t)B.f[f].fp.img_w + __float_as_uint(tp.x) + fk.pix_k;
            } else {
                constexpr uint32_t kMagicBits = 0x4B400000u;   // bits of kMagic32: RN(u - 1/2) = bits(tp) - kMagicBits
                pix = label_index<1>(B.f[f], __float_as_uint(tp.x) - kMagicBits, __float_as_uint(tp.y) - kMagicBits);
            }
            cell = __float_as_uint(tc.x) * (uint32_t)gp.mw + __float_as_uint(tc.y) + fk.cell_k;
            if (MODE != 1 && have) {
                fbx0 = fminf(fbx0, tc.x); fbx1 = fmaxf(fbx1, tc.x);
                fby0 = fminf(fby0, tc.y); fby1 = fmaxf(fby1, tc.y);
            }
        }
#if SMAP_FUSE_PF_LABEL
        if (have) {
            const uint8_t* lp = B.f[f].image + (FMT == 1 ? (size_t)pix : (size_t)pix * 3u);
            if (SMAP_FUSE_PF_LABEL == 2) prefetch_l1(lp); else prefetch_l2(lp);
        }
#endif
        push_record(have, pix | ((uint32_t)f << kFidShift), cell, w.w);
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
#ifndef SMAP_ABL_NO_DEFER   // ablation builds: deferred points are dropped (wrong results, timing only)
        if (dballot) {
            if (defer_me) {
                const uint32_t slot = dn + __popc(dballot & lt_mask);
                defer[slot] = w;
                if (NF > 1) defer_f[slot] = (uint8_t)f;
            }
            dn += __popc(dballot);
        }
#endif
    };

    // float64 decisions for up to 32 deferred points (possibly of different frames)
    auto decide64 = [&](uint32_t count) {
        const uint32_t first = dn - count;
        dn = first;
        bool have = false;
        uint2 pc = make_uint2(0u, kNone);
        float it = 0.f;
        uint32_t fid = 0;
        if ((uint32_t)lane < count) {
            const float4 w = defer[first + lane];
            fid = (NF == 1) ? 0u : defer_f[first + lane];
            it = w.w;
            pc = fuse_decide64<FMT>(B.f[fid], gp, w);
            have = pc.y != kNone;
            if (MODE != 1 && have) {   // rare: straight into the block's box of that frame
                const int cx = (int)(pc.y / (uint32_t)gp.mw), cy = (int)(pc.y - (uint32_t)cx * (uint32_t)gp.mw);
                atomicMin(&s_box[fid][0], cx); atomicMax(&s_box[fid][1], cx);
                atomicMin(&s_box[fid][2], cy); atomicMax(&s_box[fid][3], cy);
            }
        }
        push_record(have, pc.x | (fid << kFidShift), pc.y, it);
    };

    // label lookup + update for up to 32 * kFGather records, kFGather per lane
    auto gather = [&](uint32_t count) {
        const uint32_t first = rn - count;
        rn = first;
        uint32_t cellf[kFGather], lr[kFGather], lg[kFGather], fid[kFGather];
#pragma unroll
        for (int k = 0; k < kFGather; ++k) {
            const uint32_t i = (uint32_t)(k * 32 + lane);
            cellf[k] = kNone;
            lr[k] = 0; lg[k] = 0; fid[k] = 0;
            if (i < count) {
                const uint2 rec = recs[first + i];
                cellf[k] = rec.y;
                fid[k] = (NF == 1) ? 0u : (rec.x >> kFidShift);
                const uint32_t pix = rec.x & ((1u << kFidShift) - 1u);
#ifdef SMAP_ABL_NO_GATHER
                lr[k] = (pix & 1u) ? 128u : 255u; lg[k] = (pix & 1u) ? 64u : 255u;
#else
                const uint8_t* img = B.f[fid[k]].image;
                const size_t a = (size_t)pix * 3u;
                if (FMT == 1) {
                    lr[k] = __ldg(img + pix);   // the class id
                } else
#ifdef SMAP_FUSE_LABEL8
                {
                lr[k] = __ldg(img + a);
                lg[k] = __ldg(img + a + 1);
                }
#else
                if (B.f[fid[k]].img64) {
                    // R and G from ONE aligned 8-byte load (a second one only when R is the last byte of its 8: one
                    // pixel in eight): the L1 sees one sector request per point instead of two
                    const uint2 wv = __ldg(reinterpret_cast<const uint2*>(img + (a & ~(size_t)7)));
                    const uint32_t sh = ((uint32_t)a & 7u) * 8u;
                    const uint64_t v = (((uint64_t)wv.y << 32) | wv.x) >> sh;
                    lr[k] = (uint32_t)v & 0xffu;
                    lg[k] = (sh == 56u) ? __ldg(img + a + 1) : (((uint32_t)v >> 8) & 0xffu);
                } else {
                    lr[k] = __ldg(img + a);
                    lg[k] = __ldg(img + a + 1);
                }
#endif
#endif
            }
        }
        uint32_t bits[kFGather], old0[kFGather], old1[kFGather], tag[kFGather];
#pragma unroll
        for (int k = 0; k < kFGather; ++k) {
            if (FMT == 0) bits[k] = (cellf[k] != kNone) ? (s_tab_r[lr[k]] & s_tab_g[lg[k]]) : 0u;
            else bits[k] = (cellf[k] != kNone) ? s_tab_r[lr[k]] : 0u;
            tag[k] = 0u; old0[k] = 0u; old1[k] = 0u;
#ifdef SMAP_ABL_NO_SCATTER
            if (bits[k] && cellf[k] == 0x7ffffff0u) atomicOr(B.f[0].mask, bits[k]);
            bits[k] = 0;
#endif
            if (!bits[k]) continue;
            const uint32_t cell = cellf[k] & 0x7fffffffu;
            const bool boost = (bits[k] & lane_bit) && (cellf[k] >> 31);
            if (MODE == 0) {
                atomicOr(B.f[fid[k]].mask + cell, boost ? (bits[k] | (1u << gp.c)) : bits[k]);   // result unused: RED.OR
                bits[k] = 0;
            } else if (MODE == 2) {
                tag[k] = boost ? (bits[k] | (1u << gp.c)) : bits[k];          // the bits this point wants set
                old0[k] = atomicOr(B.f[fid[k]].mask + cell, tag[k]);          // what the frame had set before
                bits[k] = cell;                                               // from here on: the cell
            } else {
                const uint32_t t = B.f[fid[k]].fk.tag;
                // element indices fit 32 bits (checked by the host); the planes of one element are adjacent
                uint32_t* trow = B.tags + ((size_t)(cell * c1) * (uint32_t)B.tag_planes + fid[k]);
                const size_t tstride = (size_t)(uint32_t)B.tag_planes;
                if (bits[k] & (bits[k] - 1u)) {
                    // several classes share this pixel's (R, G): rare, done in place
                    double* row = map + cell * (uint32_t)gp.c;
                    uint32_t b = bits[k];
                    while (b) {
                        const int i = __ffs(b) - 1;
                        b &= b - 1u;
                        if (atomicMax(trow + i * tstride, t) != t) atomicAdd(row + i, 1.0);
                    }
                    if (boost && atomicMax(trow + gp.c * tstride, t) != t) atomicAdd(row + gp.lane, 2.0);
                    bits[k] = 0;
                } else {
                    const uint32_t cls = (uint32_t)__ffs(bits[k]) - 1u;
                    tag[k] = t; old0[k] = t; old1[k] = t;
                    old0[k] = atomicMax(trow + cls * tstride, t);
                    if (boost) old1[k] = atomicMax(trow + gp.c * tstride, t);
                    bits[k] = cell * (uint32_t)gp.c + cls;   // from here on: the grid element
                }
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < kFGather; ++k) {
                if (old0[k] != tag[k]) atomicAdd(map + bits[k], 1.0);
                if (old1[k] != tag[k]) atomicAdd(map + bits[k], 2.0);
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < kFGather; ++k) {
                uint32_t fresh = tag[k] & ~old0[k];   // inactive slots: tag == 0
                if (!fresh) continue;
                double* row = map + bits[k] * (uint32_t)gp.c;   // element indices fit 32 bits (checked by the host)
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
    };

    auto drain = [&](int f, uint32_t count) {
        decide32(f, count);
        __syncwarp();
        if (dn >= 32u) {
            decide64(32u);
            __syncwarp();
        }
        if (rn >= 32u * kFGather) {
            gather(32u * kFGather);
            __syncwarp();
        }
    };

#if SMAP_FUSE_TMA
    uint32_t rr = 0;   // rounds consumed so far (all frames): stage = rr % kFStages, parity from rr / kFStages
#endif
    for (int f = 0; f < nf; ++f) {
        const Fast32& fk = B.f[f].fk;
        const int w_pts = slice_pts(f);
        const int n_rounds = (w_pts + kFRoundPts - 1) / kFRoundPts;
#if !SMAP_FUSE_TMA
        // the next round is prefetched into registers (LDG.128, streaming) while the current one is processed
        {
            const float4* gp_pts = B.f[f].pts + gw * B.f[f].per_warp + lane;
            float4 buf[kFRound];
#pragma unroll
            for (int j = 0; j < kFRound; ++j)
                buf[j] = (j * 32 + lane < w_pts) ? __ldcs(gp_pts + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < n_rounds; ++r) {
                const int left = w_pts - r * kFRoundPts;
                const int pts = left < kFRoundPts ? left : kFRoundPts;
                float4 nxt[kFRound];
#pragma unroll
                for (int j = 0; j < kFRound; ++j)
                    nxt[j] = (kFRoundPts + j * 32 + lane < left) ? __ldcs(gp_pts + (r + 1) * kFRoundPts + j * 32)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#if SMAP_FUSE_PF_CLOUD
                // only lines that start inside this warp's slice (a prefetch is a hint, but it is kept in bounds anyway)
                if (lane < kFRoundPts * 16 / 128 && (r + SMAP_FUSE_PF_CLOUD) * kFRoundPts + lane * 8 < w_pts)
                    prefetch_l2(reinterpret_cast<const char*>(gp_pts - lane) +
                                (size_t)(r + SMAP_FUSE_PF_CLOUD) * (kFRoundPts * 16) + lane * 128);
#endif
#pragma unroll
                for (int j = 0; j < kFRound; ++j) {
                    const float4 w = buf[j];
                    const bool pass = cull32(fk, w.x, w.y, w.z) & (j * 32 + lane < pts);
                    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                    if (pass) queue[qn + __popc(ballot & lt_mask)] = w;
                    qn += __popc(ballot);
                }
                __syncwarp();
                while (qn >= 32u) drain(f, 32u);
#pragma unroll
                for (int j = 0; j < kFRound; ++j) buf[j] = nxt[j];
            }
        }
#else
        for (int r = 0; r < n_rounds; ++r, ++rr) {
            // keep kFStages - 1 rounds in flight: the next one goes into the stage that the previous round used
            // (every lane has consumed its reads of that stage, and the warp has re-converged since)
            issue_round();
            const uint32_t st = rr % (uint32_t)kFStages;
            const uint32_t parity = (rr / (uint32_t)kFStages) & 1u;
            while (!mbar_try_wait(bars + st, parity)) {}
            const int left = w_pts - r * kFRoundPts;
            const int pts = left < kFRoundPts ? left : kFRoundPts;
            const float4* sp = stages + st * kFRoundPts + lane;
#pragma unroll 1
            for (int g = 0; g < kFRound; g += kFGroup) {
#pragma unroll
                for (int j = 0; j < kFGroup; ++j) {
                    const float4 w = sp[(g + j) * 32];
                    const bool pass = cull32(fk, w.x, w.y, w.z) & ((g + j) * 32 + lane < pts);
                    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                    if (pass) queue[qn + __popc(ballot & lt_mask)] = w;
                    qn += __popc(ballot);
                }
                __syncwarp();
                while (qn >= 32u) drain(f, 32u);
            }
        }
#endif
        // end of the frame for this warp: the survivors left over are decided with this frame's constants (records
        // and deferred points carry their frame and stay stacked)
        if (qn) drain(f, qn);
        if (MODE != 1) {
            // fold this frame's bounding box: lane -> warp -> block (shared atomics); flushed to the frame's box at the end
            if (__any_sync(0xffffffffu, fbx1 >= fbx0)) {
                // magic-shifted float -> cell coordinate: bits - (bits(magic) - I0); clamp_c = magic - I0
                const int ox = (int)__float_as_uint(fk.clamp_c.x), oy = (int)__float_as_uint(fk.clamp_c.y);
                const bool any = fbx1 >= fbx0;
                const int a = __reduce_min_sync(0xffffffffu, any ? (int)__float_as_uint(fbx0) - ox : 0x7fffffff);
                const int b = __reduce_max_sync(0xffffffffu, any ? (int)__float_as_uint(fbx1) - ox : -1);
                const int c = __reduce_min_sync(0xffffffffu, any ? (int)__float_as_uint(fby0) - oy : 0x7fffffff);
                const int d = __reduce_max_sync(0xffffffffu, any ? (int)__float_as_uint(fby1) - oy : -1);
                if (lane == 0) {
                    atomicMin(&s_box[f][0], a); atomicMax(&s_box[f][1], b);
                    atomicMin(&s_box[f][2], c); atomicMax(&s_box[f][3], d);
                }
                fbx0 = 3.0e38f; fbx1 = -3.0e38f; fby0 = 3.0e38f; fby1 = -3.0e38f;
            }
        }
    }
    // flush: the deferred points, then the records
    if (dn) {
        decide64(dn);
        __syncwarp();
    }
    while (rn) {
        gather(rn < 32u * kFGather ? rn : 32u * kFGather);
        __syncwarp();
    }

    if (MODE != 1) {
        __syncthreads();
        if ((int)threadIdx.x < nf && s_box[threadIdx.x][1] >= s_box[threadIdx.x][0]) {
            FrameBox* box = boxes + threadIdx.x;
            atomicMin(&box->x0, s_box[threadIdx.x][0]); atomicMax(&box->x1, s_box[threadIdx.x][1]);
            atomicMin(&box->y0, s_box[threadIdx.x][2]); atomicMax(&box->y1, s_box[threadIdx.x][3]);
        }
    }
}

// After a MODE 2 batch: zero every frame slot inside its frame's bounding box, reset the boxes of the next batch.
// blockIdx.y = slot; the blocks of a slot stride over the rows of its box.
__global__ void __launch_bounds__(kThreads)
k_clear_masks(const __grid_constant__ ApplyParams ap, const FrameBox* __restrict__ boxes, FrameBox* __restrict__ next_boxes,
              unsigned long long* __restrict__ next_touched_total, int mw) {
    const int f = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        box_reset(&next_boxes[f].x0);
        if (f == 0) *next_touched_total = 0ull;
    }
    if (f >= ap.n_frames) return;
    const FrameBox b = boxes[f];
    if (b.x1 < b.x0) return;
    uint32_t* const m = ap.mask[f];
    for (int x = b.x0 + (int)blockIdx.x; x <= b.x1; x += (int)gridDim.x) {
        uint32_t* row = m + (size_t)x * mw;
        for (int y = b.y0 + (int)threadIdx.x; y <= b.y1; y += kThreads) row[y] = 0u;
    }
}

}  // namespace smap
