// Connected components of one class of a label image and the row extremes of a component: the per-pixel work of
// generate_convex_hull (src/semantic_convex_hull.py:17-91).
//
//   mask      img == index_care_about                                                    (:37-38)
//   erode     cv2.erode(mask, ones(3, 3)): a pixel survives when its whole 3 x 3 neighbourhood inside the image is set
//             (OpenCV's default border value for erosion never removes a pixel)          (:45)
//   label     8-connected components (skimage.measure.label(..., connectivity=2))       (:51): union-find over the
//             pixel grid with atomicMin (every pixel is united with its W, NW, N, NE neighbours; the root of a
//             component is its raster-first pixel, which is also what orders scikit-image's label numbers)
//   area      pixels per component (Counter(...).most_common, :59), accumulated at the root
//   extremes  per image row the smallest and the largest column of a component, its raster-first pixel left out (the
//             reference drops the first point, :70): the convex hull of these <= 2 H points is the hull of the component
//
// All integer work, HBM-bound: the label plane (4 B per pixel) is read a handful of times.
#pragma once
#include "smap_device.cuh"

namespace smap {

__global__ void __launch_bounds__(256)
k_hull_erode(const uint8_t* __restrict__ img, int h, int w, int index, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    bool keep = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= w) continue;
            keep &= (int)__ldg(img + (size_t)yy * w + xx) == index;
        }
    }
    out[(size_t)y * w + x] = keep ? 1 : 0;
}

// (loads through L2 only: other SMs are linking roots with atomics while this one reads; a stale entry would still name
// an ancestor and the atomicMin in ccl_unite re-checks, but there is no reason to read one)
__device__ __forceinline__ int ccl_find(const int* lab, int a) {
    int p = __ldcg(lab + a);
    while (p != a) { a = p; p = __ldcg(lab + a); }
    return a;
}

// lock-free union by smallest index: the root of a set is its smallest pixel index
__device__ __forceinline__ void ccl_unite(int* __restrict__ lab, int a, int b) {
    bool done;
    do {
        a = ccl_find(lab, a);
        b = ccl_find(lab, b);
        if (a < b) {
            const int old = atomicMin(lab + b, a);
            done = old == b;
            b = old;
        } else if (b < a) {
            const int old = atomicMin(lab + a, b);
            done = old == a;
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

__global__ void __launch_bounds__(256)
k_ccl_init(const uint8_t* __restrict__ mask, int64_t n, int* __restrict__ lab, int* __restrict__ area) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    lab[p] = mask[p] ? (int)p : -1;
    area[p] = 0;
}

__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t* __restrict__ mask, int h, int w, int* __restrict__ lab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int p = y * w + x;
    SMAP_BOUNDS(p >= 0 && (int64_t)p < (int64_t)h * w, 401);
    if (!mask[p]) return;
    if (x > 0 && mask[p - 1]) ccl_unite(lab, p, p - 1);
    if (y > 0) {
        const int q = p - w;
        if (mask[q]) ccl_unite(lab, p, q);
        if (x > 0 && mask[q - 1]) ccl_unite(lab, p, q - 1);
        if (x + 1 < w && mask[q + 1]) ccl_unite(lab, p, q + 1);
    }
}

__global__ void __launch_bounds__(256)
k_ccl_flatten(int64_t n, int* __restrict__ lab, int* __restrict__ area) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = p < n && lab[p >= n ? 0 : p] >= 0;
    int root = -1;
    if (in) {
        root = ccl_find(lab, (int)p);
        lab[p] = root;   // a pixel's entry only ever moves towards its root: concurrent finds stay correct
    }
    // the lanes of a warp are consecutive pixels, mostly of one component: one atomic per distinct root in the warp
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (in && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(area + root, __popc(peers));
}

// rowmin / rowmax: h entries, initialised to INT_MAX / -1 by the host wrapper
__global__ void __launch_bounds__(256)
k_hull_row_extremes(const int* __restrict__ lab, int h, int w, int root, int* __restrict__ rowmin, int* __restrict__ rowmax) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int p = y * w + x;
    const bool mine = x < w && lab[x < w ? p : 0] == root && p != root;   // the raster-first pixel is dropped (:70)
    const int lo = __reduce_min_sync(0xffffffffu, mine ? x : 0x7fffffff);
    const int hi = __reduce_max_sync(0xffffffffu, mine ? x : -1);
    if ((threadIdx.x & 31) == 0 && hi >= 0) {
        atomicMin(rowmin + y, lo);
        atomicMax(rowmax + y, hi);
    }
}

__global__ void k_fill_i32(int* __restrict__ a, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace smap
