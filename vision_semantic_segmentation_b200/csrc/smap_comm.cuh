// Multi-GPU exchange of the BEV grid: frames are sharded over the ranks, every rank integrates its block
// of frames into its own full-size grid, and the grids are SUMMED once -- update_map only ever adds frame-determined
// constants (src/mapping_replay.py:281,294).  This file holds what smap_allreduce / smap_reduce_scatter_rows (smap.cu)
// are made of:
//   * NCCL, resolved at run time from the libnccl already loaded in the process (torch's) or found by the loader --
//     the library itself has no link-time dependency on it, and a single-GPU user never loads it;
//   * k_comm_prep       the few integers the ranks agree on before the exchange (union window, value bound);
//   * k_pack_window     the touched window of the float64 grid -> a dense buffer: two uint16 counts per 32-bit word
//                       (exact while the global sum stays below 2^16), one uint32 per count, or the doubles themselves;
//   * k_unpack_window   the reduced buffer -> the grid (the window is overwritten with the global sum).
// A count grid holds small non-negative integers, so the packed exchange is EXACT and moves 4x (uint16) or 2x (uint32)
// fewer bytes than the float64 grid, and only the rows / columns some rank touched since the last clear.
// Both kernels are HBM-bound streams: blockIdx.y = window row, threads stride over the row's contiguous run.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: every function is resolved with dlsym (see NcclApi)

#include "smap_kernels.cuh"

namespace smap {

struct NcclApi {
    void* lib = nullptr;
    bool tried = false;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;

    // 1. SMAP_NCCL_LIB (explicit path), 2. the libnccl.so.2 already mapped into the process (RTLD_NOLOAD: torch's
    // bundled copy when the caller is a torch program), 3. whatever the loader finds.
    bool load() {
        if (tried) return lib != nullptr;
        tried = true;
        const char* env = getenv("SMAP_NCCL_LIB");
        if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
#define SMAP_NCCL_SYM(name) *reinterpret_cast<void**>(&name) = dlsym(lib, "nccl" #name)
        SMAP_NCCL_SYM(GetErrorString); SMAP_NCCL_SYM(GetUniqueId); SMAP_NCCL_SYM(CommInitRank);
        SMAP_NCCL_SYM(CommDestroy); SMAP_NCCL_SYM(CommCount); SMAP_NCCL_SYM(CommUserRank);
        SMAP_NCCL_SYM(AllReduce); SMAP_NCCL_SYM(ReduceScatter); SMAP_NCCL_SYM(AllGather);
#undef SMAP_NCCL_SYM
        if (!GetErrorString || !GetUniqueId || !CommInitRank || !CommDestroy || !CommCount || !CommUserRank ||
            !AllReduce || !ReduceScatter || !AllGather) {
            lib = nullptr;
            return false;
        }
        return true;
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

// What the ranks agree on before an exchange: element-wise MAX over the ranks of
//   [0] -x0  [1] x1  [2] -y0  [3] y1     union window of the touched cells (empty: x1 < x0)
//   [4] value bound: no grid element exceeds it (3 per integrated frame: 1 per class observation + 2 lane boost)
//   [5] 1 when the grid may hold non-integer values (log-likelihood update, caller-written grid): exchange as float64
constexpr int kCommWords = 8;

__global__ void k_comm_prep(const FrameBox* __restrict__ ubox, int full, int mh, int mw, int value_bound, int not_integer,
                            int* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    FrameBox b = *ubox;
    if (full) { b.x0 = 0; b.x1 = mh - 1; b.y0 = 0; b.y1 = mw - 1; }
    out[0] = -b.x0; out[1] = b.x1; out[2] = -b.y0; out[3] = b.y1;
    out[4] = value_bound; out[5] = not_integer; out[6] = 0; out[7] = 0;
}

__global__ void k_box_set(FrameBox* __restrict__ box, int x0, int x1, int y0, int y1) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { box->x0 = x0; box->x1 = x1; box->y0 = y0; box->y1 = y1; }
}

enum { kPackU16 = 0, kPackU32 = 1, kPackF64 = 2 };

// Integer-valued double in [0, 2^32) <-> uint32 with one DADD (2^52 puts the integer into the low mantissa bits)
// instead of the FP64 conversion instructions.
__device__ __forceinline__ uint32_t count_to_u32(double x) {
    return (uint32_t)(unsigned long long)__double_as_longlong(__dadd_rn(x, 4503599627370496.0));
}
__device__ __forceinline__ double u32_to_count(uint32_t v) {
    return __dadd_rn(__longlong_as_double(0x4330000000000000ll | (long long)v), -4503599627370496.0);
}

// Grid rows [x0, x0 + gridDim.y) x columns [y0, y0 + cols) of the (MH, MW, C) grid -> dense rows of `run` = cols * C
// elements, each padded to `run_words` 32-bit words (kPackU16: 2 elements per word, kPackU32: 1; kPackF64: run doubles).
// Rows outside the touched rows [wx0, wx1] (untouched on every rank, or the padding rows of the row-tiled exchange
// beyond the grid) are written as zeros without reading the grid.
// CLEAR: the window is zeroed behind the read (streaming exchange: the packed grid is a buffer of local increments
// that starts over after every exchange).  The zero that is stored is derived from the loaded value (AND with
// `zmask`, a kernel argument that is always 0, so the compiler cannot fold it): a store of a CONSTANT to an address
// whose load is still in flight stalls the memory pipeline -- measured on B200 at 0.29 ms for this kernel against
// 0.028 ms without the stores and 0.023 ms without the loads (profiles/r2c_exchange_phases.md); a store that waits for
// its data in the register scoreboard, as any read-modify-write does, costs nothing extra.
__device__ __forceinline__ void clear_behind(double* p, double loaded, long long zmask) {
    __stcs(p, __longlong_as_double(__double_as_longlong(loaded) & zmask));
}

template <int PACK, bool CLEAR = false>
__global__ void __launch_bounds__(kThreads)
k_pack_window(double* __restrict__ map, int wx0, int wx1, int mw, int c, int x0, int y0, int run, int run_words,
              void* __restrict__ out, long long zmask) {
    const int row = x0 + (int)blockIdx.y;
    const bool real = row >= wx0 && row <= wx1;
    double* src = map + ((size_t)row * mw + y0) * c;
    if (PACK == kPackF64) {
        double* dst = reinterpret_cast<double*>(out) + (size_t)blockIdx.y * run;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < run; e += gridDim.x * blockDim.x) {
            double x = 0.0;
            if (real) {
                x = __ldcs(src + e);
                if (CLEAR) clear_behind(src + e, x, zmask);
            }
            dst[e] = x;
        }
    } else if (PACK == kPackU32) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(out) + (size_t)blockIdx.y * run_words;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < run; e += gridDim.x * blockDim.x) {
            uint32_t v = 0u;
            if (real) {
                const double x = __ldcs(src + e);
                v = count_to_u32(x);
                if (CLEAR) clear_behind(src + e, x, zmask);
            }
            dst[e] = v;
        }
    } else {
        // one element per thread and iteration: coalesced 8-byte loads, 2-byte stores (a row of the buffer holds
        // 2 * run_words uint16, the last one padding when run is odd)
        uint16_t* dst = reinterpret_cast<uint16_t*>(out) + (size_t)blockIdx.y * run_words * 2;
        const int n16 = 2 * run_words;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += gridDim.x * blockDim.x) {
            uint32_t v = 0u;
            if (real && e < run) {
                const double x = __ldcs(src + e);
                v = count_to_u32(x);
                if (CLEAR) clear_behind(src + e, x, zmask);
            }
            dst[e] = (uint16_t)v;
        }
    }
}

// box <- box UNION [x0, x1] x [y0, y1]
__global__ void k_box_fold(FrameBox* __restrict__ box, int x0, int x1, int y0, int y1) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && x1 >= x0 && y1 >= y0) {
        box->x0 = min(box->x0, x0); box->x1 = max(box->x1, x1);
        box->y0 = min(box->y0, y0); box->y1 = max(box->y1, y1);
    }
}

// The reduced buffer -> the grid: dst rows [dst_x0, dst_x0 + gridDim.y) of a (*, dst_mw, C) float64 array receive the
// buffer rows [src_row0, ...), columns [y0, y0 + cols).  Used with dst = the handle's grid (all-reduce) and with
// dst = the caller's row tile (reduce-scatter; dst_x0 = tile row of the first received row).
// ADD: the window receives grid + buffer instead of the buffer (streaming exchange: the buffer holds the ranks' summed
// increments since the previous exchange).
template <int PACK, bool ADD = false>
__global__ void __launch_bounds__(kThreads)
k_unpack_window(double* __restrict__ dst, int dst_mw, int c, int dst_x0, int y0, int run, int run_words,
                const void* __restrict__ in, int src_row0) {
    const size_t srow = (size_t)src_row0 + blockIdx.y;
    double* out = dst + ((size_t)(dst_x0 + (int)blockIdx.y) * dst_mw + y0) * c;
    if (PACK == kPackF64) {
        const double* src = reinterpret_cast<const double*>(in) + srow * run;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < run; e += gridDim.x * blockDim.x)
            out[e] = ADD ? __dadd_rn(out[e], src[e]) : src[e];
    } else if (PACK == kPackU32) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + srow * run_words;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < run; e += gridDim.x * blockDim.x)
            out[e] = ADD ? __dadd_rn(out[e], u32_to_count(src[e])) : u32_to_count(src[e]);
    } else {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(in) + srow * run_words * 2;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < run; e += gridDim.x * blockDim.x) {
            const double v = u32_to_count((uint32_t)src[e]);
            out[e] = ADD ? __dadd_rn(out[e], v) : v;
        }
    }
}

}  // namespace smap
