// Planar projection of the label image onto the map plane: cv2.warpPerspective(image, h, (MW, MH)) as
// generate_homography calls it (src/homography.py:53-55, from update_map_planar, src/mapping.py:465-466) -- 8-bit image,
// INTER_LINEAR, BORDER_CONSTANT 0.  OpenCV's arithmetic (third-party, restated in oracle/warp_port.py and held to cv2
// itself there): destination blocks of bw0 columns, source coordinates X = cvRound((X0 + M0 x1) * (32 / W)) in fixed point
// with 5 fractional bits (X0 = M0 bx + M1 y + M2 formed per block row in double, un-fused), integer part saturated to
// int16, four taps weighted (32 - ax)(32 - ay) 32 ..., taps outside the source are 0, result (acc + 2^14) >> 15.
//
// Bound: HBM on the destination side (MH MW cn bytes written once) plus a gather from the source image through L2: one
// thread per destination pixel, consecutive threads = consecutive columns, so the taps of a warp walk a line of the source.
#pragma once
#include "smap_device.cuh"

namespace smap {

struct WarpParams {
    double m[9];      // inv(h), cv::invert's closed formula (host)
    int src_h, src_w, cn;
    int dst_h, dst_w, bw0;
};

// std::max((double)INT_MIN, std::min((double)INT_MAX, v)) with the C++ comparison semantics (NaN -> INT_MAX)
__device__ __forceinline__ double clamp_like_std(double v) {
    const double t = (v < 2147483647.0) ? v : 2147483647.0;
    return (-2147483648.0 < t) ? t : -2147483648.0;
}

template <int CN>
__global__ void __launch_bounds__(256)
k_warp_perspective(const __grid_constant__ WarpParams p, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= p.dst_w) return;
    const int bxi = (x / p.bw0) * p.bw0;
    const double bx = (double)bxi, x1 = (double)(x - bxi), yd = (double)y;
    // X0 = M[0] * bx + M[1] * y + M[2]: products rounded, added left to right
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(p.m[0], bx), __dmul_rn(p.m[1], yd)), p.m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(p.m[3], bx), __dmul_rn(p.m[4], yd)), p.m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(p.m[6], bx), __dmul_rn(p.m[7], yd)), p.m[8]);
    double W = __dadd_rn(W0, __dmul_rn(p.m[6], x1));
    W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
    const double fX = clamp_like_std(__dmul_rn(__dadd_rn(X0, __dmul_rn(p.m[0], x1)), W));
    const double fY = clamp_like_std(__dmul_rn(__dadd_rn(Y0, __dmul_rn(p.m[3], x1)), W));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);   // cvRound: to nearest, ties to even
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    const bool y0ok = sy >= 0 && sy < p.src_h, y1ok = sy + 1 >= 0 && sy + 1 < p.src_h;
    const bool x0ok = sx >= 0 && sx < p.src_w, x1ok = sx + 1 >= 0 && sx + 1 < p.src_w;
    const uint8_t* r0 = src + ((size_t)(y0ok ? sy : 0) * p.src_w) * CN;
    const uint8_t* r1 = src + ((size_t)(y1ok ? sy + 1 : 0) * p.src_w) * CN;
    const int c0 = (x0ok ? sx : 0) * CN, c1 = (x1ok ? sx + 1 : 0) * CN;
    SMAP_BOUNDS((!y0ok || (sy >= 0 && sy < p.src_h)) && (!y1ok || sy + 1 < p.src_h) && (!x0ok || (sx >= 0 && sx < p.src_w)) &&
                (!x1ok || sx + 1 < p.src_w), 501);
    uint8_t* o = dst + ((size_t)y * p.dst_w + x) * CN;
#pragma unroll
    for (int k = 0; k < CN; ++k) {
        const int v00 = (y0ok && x0ok) ? (int)__ldg(r0 + c0 + k) : 0;
        const int v01 = (y0ok && x1ok) ? (int)__ldg(r0 + c1 + k) : 0;
        const int v10 = (y1ok && x0ok) ? (int)__ldg(r1 + c0 + k) : 0;
        const int v11 = (y1ok && x1ok) ? (int)__ldg(r1 + c1 + k) : 0;
        const int acc = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
        o[k] = (uint8_t)((acc + (1 << 14)) >> 15);
    }
}

}  // namespace smap
