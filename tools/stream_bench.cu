// Dev microbenchmark (not part of the library): how fast can the cloud be streamed + culled + compacted on B200, and
// with which staging?  The stream + cull stage of k_fuse alone, in the launch shape of the library (256 threads,
// 4 blocks per SM, every warp owns a contiguous slice walked in rounds), with the staging as the variable:
//   0  register prefetch one round ahead (what k_fuse does), rounds of 64 points
//   1  the same with rounds of 128 points
//   2  cp.async (LDGSTS.128) ring in shared memory, DEPTH rounds of 64 points ahead
//   3  variant 0 + cp.async.bulk.prefetch.L2 of the warp's slice, 4 KB blocks, 8 KB ahead
//   4  read only (no cull): the ceiling of the access pattern
//   5  per-warp 1-D TMA bulk copies (cp.async.bulk + mbarrier) into a ring, DEPTH stages of 1 KB
// Build + run on the GPU box:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/stream_bench tools/stream_bench.cu
//   gpurun_out/stream_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e__)); exit(1); } } while (0)

struct Cull {
    float2 c_dc[4], c_ab[4], c_wh;
    float c_bw, c_rh, c_rthr, c_depth, c_lo_u, c_hi_u, c_lo_v, c_hi_v;
};

__device__ __forceinline__ bool cull32(const Cull& k, float x, float y, float z) {
    const float2 xx = make_float2(x, x), yy = make_float2(y, y), zz = make_float2(z, z);
    const float2 dc = __ffma2_rn(k.c_dc[0], xx, __ffma2_rn(k.c_dc[1], yy, __ffma2_rn(k.c_dc[2], zz, k.c_dc[3])));
    const float2 ab = __ffma2_rn(k.c_ab[0], xx, __ffma2_rn(k.c_ab[1], yy, __ffma2_rn(k.c_ab[2], zz, k.c_ab[3])));
    const float2 cc = make_float2(dc.y, dc.y);
    const float2 lo = __fadd2_rn(ab, cc);
    const float2 hi = __ffma2_rn(k.c_wh, cc, make_float2(-ab.x, -ab.y));
    bool p = (lo.x > k.c_lo_u) & (hi.x > k.c_hi_u) & (lo.y > k.c_lo_v) & (hi.y > k.c_hi_v);
    p |= !(dc.y > k.c_depth);
    p &= fabsf(dc.x - k.c_rh) < k.c_rthr;
    p |= fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > k.c_bw;
    return p;
}

constexpr int kThreads = 256, kWarps = 8;
constexpr int kQueueCap = 32 * 4 + 32;
constexpr int kPadSmem = 4352 - kQueueCap * 16 > 0 ? 4352 - kQueueCap * 16 : 0;   // the library's per-warp footprint

__device__ unsigned long long g_sink;

template <int ROUND>
__device__ __forceinline__ void cull_round(const Cull& ck, const float4* buf, int pts, int lane, uint32_t lt_mask, float4* queue,
                                           uint32_t& qn, uint32_t& acc) {
#pragma unroll
    for (int j = 0; j < ROUND; ++j) {
        const float4 w = buf[j];
        const bool pass = cull32(ck, w.x, w.y, w.z) & (j * 32 + lane < pts);
        const unsigned ballot = __ballot_sync(0xffffffffu, pass);
        if (pass) queue[qn + __popc(ballot & lt_mask)] = w;
        qn += __popc(ballot);
    }
    __syncwarp();
    while (qn >= 32u) {   // stand-in for the drain: consume 32 survivors
        qn -= 32u;
        const float4 w = queue[qn + lane];
        acc += __float_as_uint(w.x) ^ __float_as_uint(w.w);
        __syncwarp();
    }
}

// ---- variants 0 / 1 / 3 / 4: register prefetch
template <int ROUND, int VAR>
__global__ void __launch_bounds__(kThreads, 4)
k_reg(const __grid_constant__ Cull ck, const float4* __restrict__ pts, int64_t n, int per_warp) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    float4* const queue = reinterpret_cast<float4*>(s_dyn + (size_t)warp * (kQueueCap * 16 + kPadSmem));
    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
    int w_pts;
    {
        const int64_t left = n - gw * per_warp;
        w_pts = left <= 0 ? 0 : (left < per_warp ? (int)left : per_warp);
    }
    constexpr int RP = 32 * ROUND;
    const float4* gp = pts + gw * per_warp + lane;
    float4 buf[ROUND];
#pragma unroll
    for (int j = 0; j < ROUND; ++j) buf[j] = (j * 32 + lane < w_pts) ? __ldcs(gp + j * 32) : make_float4(0, 0, 0, 0);
    uint32_t qn = 0, acc = 0;
    const int n_rounds = (w_pts + RP - 1) / RP;
    for (int r = 0; r < n_rounds; ++r) {
        const int left = w_pts - r * RP;
        const int p = left < RP ? left : RP;
        if (VAR == 3) {
            // every 4 KB (4 rounds of 64): prefetch the block 8 KB ahead into L2 with one bulk instruction
            if ((r & 3) == 0 && lane == 0) {
                const int64_t ahead = (int64_t)(r + 8) * RP;
                if (ahead + 256 <= w_pts) {
                    const float4* a = pts + gw * per_warp + ahead;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(4096) : "memory");
                }
            }
        }
        float4 nxt[ROUND];
#pragma unroll
        for (int j = 0; j < ROUND; ++j)
            nxt[j] = (RP + j * 32 + lane < left) ? __ldcs(gp + (r + 1) * RP + j * 32) : make_float4(0, 0, 0, 0);
        if (VAR == 4) {
#pragma unroll
            for (int j = 0; j < ROUND; ++j) acc += __float_as_uint(buf[j].x) ^ __float_as_uint(buf[j].y) ^ __float_as_uint(buf[j].z) ^ __float_as_uint(buf[j].w);
        } else {
            cull_round<ROUND>(ck, buf, p, lane, lt_mask, queue, qn, acc);
        }
#pragma unroll
        for (int j = 0; j < ROUND; ++j) buf[j] = nxt[j];
    }
    if (acc == 0x12345678u) atomicAdd(&g_sink, 1ull);
}

// ---- variant 2: cp.async ring per warp, DEPTH rounds of 64 points in flight
template <int DEPTH>
__global__ void __launch_bounds__(kThreads, 4)
k_cpasync(const __grid_constant__ Cull ck, const float4* __restrict__ pts, int64_t n, int per_warp) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    constexpr int ROUND = 2, RP = 64;
    constexpr int kWarpBytes = kQueueCap * 16 + kPadSmem + DEPTH * RP * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned char* wb = s_dyn + (size_t)warp * kWarpBytes;
    float4* const queue = reinterpret_cast<float4*>(wb);
    float4* const ring = reinterpret_cast<float4*>(wb + kQueueCap * 16 + kPadSmem);
    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
    int w_pts;
    {
        const int64_t left = n - gw * per_warp;
        w_pts = left <= 0 ? 0 : (left < per_warp ? (int)left : per_warp);
    }
    const float4* gp = pts + gw * per_warp + lane;
    const int n_rounds = (w_pts + RP - 1) / RP;
    auto issue = [&](int r) {   // round r -> slot r % DEPTH (one commit group per round, empty when past the end)
        if (r < n_rounds) {
            float4* slot = ring + (r % DEPTH) * RP;
#pragma unroll
            for (int j = 0; j < ROUND; ++j) {
                const int i = r * RP + j * 32 + lane;
                if (i < w_pts) {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot + j * 32 + lane);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gp + r * RP + j * 32) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) issue(d);
    uint32_t qn = 0, acc = 0;
    for (int r = 0; r < n_rounds; ++r) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        const int left = w_pts - r * RP;
        const int p = left < RP ? left : RP;
        float4 buf[ROUND];
        const float4* slot = ring + (r % DEPTH) * RP;
#pragma unroll
        for (int j = 0; j < ROUND; ++j) buf[j] = (j * 32 + lane < p) ? slot[j * 32 + lane] : make_float4(0, 0, 0, 0);
        cull_round<ROUND>(ck, buf, p, lane, lt_mask, queue, qn, acc);
        issue(r + DEPTH);   // the slot just consumed
    }
    if (acc == 0x12345678u) atomicAdd(&g_sink, 1ull);
}

// ---- variant 5: per-warp TMA bulk copies into a ring of 1 KB stages
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            (uint32_t)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}
template <int DEPTH>
__global__ void __launch_bounds__(kThreads, 4)
k_tma(const __grid_constant__ Cull ck, const float4* __restrict__ pts, int64_t n, int per_warp) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ uint64_t s_bar[kWarps][DEPTH];
    constexpr int ROUND = 2, RP = 64;
    constexpr int kWarpBytes = kQueueCap * 16 + kPadSmem + DEPTH * RP * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned char* wb = s_dyn + (size_t)warp * kWarpBytes;
    float4* const queue = reinterpret_cast<float4*>(wb);
    float4* const ring = reinterpret_cast<float4*>(wb + kQueueCap * 16 + kPadSmem);
    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
    int w_pts;
    {
        const int64_t left = n - gw * per_warp;
        w_pts = left <= 0 ? 0 : (left < per_warp ? (int)left : per_warp);
    }
    const float4* base = pts + gw * per_warp;
    const int n_rounds = (w_pts + RP - 1) / RP;
    if (lane == 0) {
        for (int d = 0; d < DEPTH; ++d) mbar_init(&s_bar[warp][d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int r) {
        if (r < n_rounds && lane == 0) {
            const int left = w_pts - r * RP;
            const uint32_t bytes = (uint32_t)(left < RP ? left : RP) * 16u;
            uint64_t* bar = &s_bar[warp][r % DEPTH];
            mbar_expect(bar, bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(ring + (r % DEPTH) * RP)),
                         "l"(base + (int64_t)r * RP), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                         : "memory");
        }
    };
    for (int d = 0; d < DEPTH; ++d) issue(d);
    uint32_t qn = 0, acc = 0;
    for (int r = 0; r < n_rounds; ++r) {
        mbar_wait(&s_bar[warp][r % DEPTH], (uint32_t)((r / DEPTH) & 1));
        const int left = w_pts - r * RP;
        const int p = left < RP ? left : RP;
        float4 buf[ROUND];
        const float4* slot = ring + (r % DEPTH) * RP;
#pragma unroll
        for (int j = 0; j < ROUND; ++j) buf[j] = (j * 32 + lane < p) ? slot[j * 32 + lane] : make_float4(0, 0, 0, 0);
        cull_round<ROUND>(ck, buf, p, lane, lt_mask, queue, qn, acc);
        __syncwarp();
        issue(r + DEPTH);
    }
    if (acc == 0x12345678u) atomicAdd(&g_sink, 1ull);
}

__global__ void k_fill(float4* p, int64_t n, uint32_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = seed + (uint32_t)i * 2654435761u;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (float)(s >> 8) * (1.0f / 16777216.0f); };
    p[i] = make_float4(-5.f + 125.f * rnd(), -60.f + 120.f * rnd(), -2.4f + rnd(), 30.f * rnd());
}

template <typename F>
float time_it(F&& launch, int frames, cudaStream_t* st, int n_streams) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cudaEvent_t fork, join[4];
    CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, 0));
        CK(cudaEventRecord(fork, 0));
        for (int i = 0; i < n_streams; ++i) CK(cudaStreamWaitEvent(st[i], fork, 0));
        for (int f = 0; f < frames; ++f) launch(f, st[f % n_streams]);
        for (int i = 0; i < n_streams; ++i) { CK(cudaEventRecord(join[i], st[i])); CK(cudaStreamWaitEvent(0, join[i], 0)); }
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / frames * 1000.f;
}

int main() {
    const int64_t n = 2000000;
    const int ring = 16, frames = 256;
    float4* pts;
    CK(cudaMalloc(&pts, sizeof(float4) * n * ring));
    for (int f = 0; f < ring; ++f) k_fill<<<(unsigned)((n + 255) / 256), 256>>>(pts + f * n, n, 1234u + f);
    CK(cudaDeviceSynchronize());
    // a camera looking along +x from the origin: ~36 % of the cloud inside range + frustum
    Cull ck{};
    const float fx = 1500.f, cx = 960.f, cy = 720.f;
    // rows: d = x, c (depth) = x; a = fx * (-y) + cx * x, b = fx * (-z) + cy * x
    ck.c_dc[0] = make_float2(1.f, 1.f); ck.c_dc[1] = make_float2(0.f, 0.f); ck.c_dc[2] = make_float2(0.f, 0.f); ck.c_dc[3] = make_float2(0.f, 0.f);
    ck.c_ab[0] = make_float2(cx, cy); ck.c_ab[1] = make_float2(-fx, 0.f); ck.c_ab[2] = make_float2(0.f, -fx); ck.c_ab[3] = make_float2(0.f, 0.f);
    ck.c_wh = make_float2(1920.f - 1.f, 1440.f - 1.f);   // W c - a > hi  with a shifted by c (lo = a + c)
    ck.c_bw = 1.0e6f; ck.c_rh = 50.f; ck.c_rthr = 50.01f; ck.c_depth = 0.01f;
    ck.c_lo_u = -0.01f; ck.c_hi_u = -0.01f; ck.c_lo_v = -0.01f; ck.c_hi_v = -0.01f;

    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaStream_t st[4];
    for (int i = 0; i < 4; ++i) CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    const int base_smem = kWarps * (kQueueCap * 16 + kPadSmem) + 2048;

    auto report = [&](const char* name, float us) {
        printf("%-44s %7.2f us/frame  %7.1f GB/s (32 MB cloud)\n", name, us, 32.0e6 / us / 1e3);
        fflush(stdout);
    };
    for (int shape = 0; shape < 2; ++shape) {
        const int div = shape == 0 ? 1 : 2;
        const int ns = shape == 0 ? 1 : 4;
        const int64_t gx = (int64_t)sms * 4 / div;
        const int per_warp = (int)((n + gx * kWarps - 1) / (gx * kWarps));
        printf("---- %s: grid %lld, %d points per warp\n", shape == 0 ? "full grid, one stream" : "half grids on four streams", (long long)gx, per_warp);
#define RUN(NAME, KERNEL, SMEM)                                                                                   \
    {                                                                                                              \
        CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                       \
        report(NAME, time_it([&](int f, cudaStream_t s) { KERNEL<<<(unsigned)gx, kThreads, SMEM, s>>>(ck, pts + (int64_t)(f % ring) * n, n, per_warp); }, frames, st, ns)); \
        CK(cudaGetLastError());                                                                                    \
    }
        RUN("4 read only, rounds of 64", (k_reg<2, 4>), base_smem);
        RUN("4 read only, rounds of 128", (k_reg<4, 4>), base_smem);
        RUN("0 register prefetch, rounds of 64", (k_reg<2, 0>), base_smem);
        RUN("1 register prefetch, rounds of 128", (k_reg<4, 0>), base_smem);
        RUN("3 register prefetch + bulk L2 prefetch", (k_reg<2, 3>), base_smem);
        RUN("2 cp.async ring, depth 2", (k_cpasync<2>), base_smem + kWarps * 2 * 1024);
        RUN("2 cp.async ring, depth 3", (k_cpasync<3>), base_smem + kWarps * 3 * 1024);
        RUN("2 cp.async ring, depth 4", (k_cpasync<4>), base_smem + kWarps * 4 * 1024);
        RUN("5 TMA bulk ring, depth 2", (k_tma<2>), base_smem + kWarps * 2 * 1024);
        RUN("5 TMA bulk ring, depth 4", (k_tma<4>), base_smem + kWarps * 4 * 1024);
    }
    return 0;
}
