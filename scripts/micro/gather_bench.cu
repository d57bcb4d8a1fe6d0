// Micro-benchmark: how fast can one B200 gather random 128-byte rows?
//   A: LDG.128, 8 lanes per row, U rows in flight per lane group (register staged)
//   B: cp.async.bulk 128-byte row copies into shared memory (mbarrier completion), then LDS
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) gatherA(const float* __restrict__ tab, const int* __restrict__ idx, const int n_idx,
                                               float* __restrict__ out) {
    const int lane = threadIdx.x & 31, c = lane & 7, g = lane >> 3;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = warp * 4 * U; base < n_idx; base += n_warps * 4 * U) {
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldg(idx + min(base + u * 4 + g, n_idx - 1));
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r[u] * 32 + 4 * c));
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

// A32: one row per warp instruction (LDG.32, 32 lanes per row): is the per-warp limit in instructions or in lines?
template <int U>
__global__ void __launch_bounds__(256) gatherA32(const float* __restrict__ tab, const int* __restrict__ idx, const int n_idx,
                                                 float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int base = warp * U; base < n_idx; base += n_warps * U) {
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldg(idx + min(base + u, n_idx - 1));
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(tab + (size_t)r[u] * 32 + lane);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 12345.678f) out[0] = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p);
// D: like the node kernel's gather: one CTA of NT threads per SM, per thread U cp.async then wait_all then consume (no overlap)
template <int U, bool LDGV>
__global__ void gatherD(const float* __restrict__ tab, const int* __restrict__ idx, const int n_idx, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, c = lane & 7, g = lane >> 3;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t ring = s32(smem) + threadIdx.x * 16;
    const float* ring_g = reinterpret_cast<const float*>(smem) + threadIdx.x * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = warp * 4 * U; base < n_idx; base += n_warps * 4 * U) {
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldg(idx + min(base + u * 4 + g, n_idx - 1));
        if (LDGV) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r[u] * 32 + 4 * c));
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + u * blockDim.x * 16), "l"(tab + (size_t)r[u] * 32 + 4 * c) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 v = *reinterpret_cast<const float4*>(ring_g + u * blockDim.x * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

// C: cp.async (LDGSTS) 16 bytes per lane, 8 lanes per row, U row-instructions per stage, 2 stages per warp
template <int U>
__global__ void __launch_bounds__(256) gatherC(const float* __restrict__ tab, const int* __restrict__ idx, const int n_idx,
                                               float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, c = lane & 7, g = lane >> 3, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned char* buf = smem + (size_t)wib * 2 * U * 512;      // [2 stages][U][4 rows][128 B]
    const int n_iter = (n_idx - warp * 4 * U + n_warps * 4 * U - 1) / (n_warps * 4 * U);
    auto issue = [&](int it) {
        const int base = warp * 4 * U + it * n_warps * 4 * U;
        unsigned char* st = buf + (it & 1) * U * 512;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = __ldg(idx + min(base + u * 4 + g, n_idx - 1));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(st + u * 512 + g * 128 + c * 16)), "l"(tab + (size_t)r * 32 + 4 * c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_iter > 0) issue(0);
    for (int it = 0; it < n_iter; ++it) {
        if (it + 1 < n_iter) { issue(it + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const unsigned char* st = buf + (it & 1) * U * 512;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float4 v = *reinterpret_cast<const float4*>(st + u * 512 + g * 128 + c * 16);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
    }
    if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// B: one warp = one pipeline.  STAGES stages of 32 rows (4 KB each).  Each lane issues one 128-byte bulk copy per stage.
template <int STAGES>
__global__ void __launch_bounds__(256) gatherB(const float* __restrict__ tab, const int* __restrict__ idx, const int n_idx,
                                               float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned char* buf = smem + (size_t)wib * STAGES * 4096;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(blockDim.x >> 5) * STAGES * 4096) + wib * STAGES;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bars + s)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int n_iter = (n_idx - warp * 32 + n_warps * 32 - 1) / (n_warps * 32);
    auto issue = [&](int it) {
        const int s = it % STAGES;
        const int j = warp * 32 + it * n_warps * 32 + lane;
        const int r = __ldg(idx + min(j, n_idx - 1));
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + s)), "r"(32 * 128) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(buf + s * 4096 + lane * 128)), "l"(tab + (size_t)r * 32), "r"(128), "r"(s32(bars + s)) : "memory");
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < STAGES - 1 && it < n_iter; ++it) issue(it);
    for (int it = 0; it < n_iter; ++it) {
        if (it + STAGES - 1 < n_iter) issue(it + STAGES - 1);
        const int s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bars + s)), "r"(parity) : "memory");
        // consume: lane reads 4 float4 of "its" rows (row-major stage: 32 rows x 128 B); 8 lanes per row, 4 rows per pass, 8 passes
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const float4 v = *reinterpret_cast<const float4*>(buf + s * 4096 + (p * 4 + (lane >> 3)) * 128 + (lane & 7) * 16);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
    }
    if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

int main(int argc, char** argv) {
    const int n_idx = 8 << 20;                                   // 8M row gathers = 1 GB
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (size_t mb : {16, 96}) {
        const size_t rows = mb * 1024 * 1024 / 128;
        float* tab; int* idx; float* out;
        CK(cudaMalloc(&tab, rows * 128)); CK(cudaMemset(tab, 0, rows * 128));
        CK(cudaMalloc(&idx, n_idx * sizeof(int))); CK(cudaMalloc(&out, 16));
        std::vector<int> h(n_idx);
        uint64_t x = 88172645463325252ull;
        for (int i = 0; i < n_idx; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (int)(x % rows); }
        CK(cudaMemcpy(idx, h.data(), n_idx * sizeof(int), cudaMemcpyHostToDevice));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        auto time = [&](const char* name, auto launch) {
            launch(); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a)); for (int i = 0; i < 3; ++i) launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3;
            printf("table %4zu MB  %-28s %8.1f us  %7.1f GB/s  (%5.2f B/clk/SM @1.9GHz)\n", mb, name, ms * 1e3, n_idx * 128.0 / ms / 1e6,
                   n_idx * 128.0 / (ms * 1e-3) / sms / 1.9e9);
            CK(cudaGetLastError());
        };
        for (int occ : {2, 4}) {
            char nm[64];
            snprintf(nm, 64, "LDG U=4 ctas/SM=%d", occ);  time(nm, [&] { gatherA<4><<<sms * occ, 256>>>(tab, idx, n_idx, out); });
            snprintf(nm, 64, "LDG U=8 ctas/SM=%d", occ);  time(nm, [&] { gatherA<8><<<sms * occ, 256>>>(tab, idx, n_idx, out); });
            snprintf(nm, 64, "LDG U=16 ctas/SM=%d", occ); time(nm, [&] { gatherA<16><<<sms * occ, 256>>>(tab, idx, n_idx, out); });
        }
        {
            CK(cudaFuncSetAttribute(gatherD<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 1024 * 16));
            CK(cudaFuncSetAttribute(gatherD<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 16));
            time("D cp.async U=6 512thr 1cta/SM", [&] { gatherD<6, false><<<sms, 512, 6 * 512 * 16>>>(tab, idx, n_idx, out); });
            time("D cp.async U=8 512thr 1cta/SM", [&] { gatherD<8, false><<<sms, 512, 8 * 512 * 16>>>(tab, idx, n_idx, out); });
            time("D cp.async U=8 1024thr 1cta/SM", [&] { gatherD<8, false><<<sms, 1024, 8 * 1024 * 16>>>(tab, idx, n_idx, out); });
            time("D LDG U=6 512thr 1cta/SM", [&] { gatherD<6, true><<<sms, 512, 0>>>(tab, idx, n_idx, out); });
            time("D LDG U=8 1024thr 1cta/SM", [&] { gatherD<8, true><<<sms, 1024, 0>>>(tab, idx, n_idx, out); });
        }
        for (int occ : {2, 4}) {
            char nm[64];
            snprintf(nm, 64, "LDG.32 U=8 ctas/SM=%d", occ);  time(nm, [&] { gatherA32<8><<<sms * occ, 256>>>(tab, idx, n_idx, out); });
            snprintf(nm, 64, "LDG.32 U=32 ctas/SM=%d", occ); time(nm, [&] { gatherA32<32><<<sms * occ, 256>>>(tab, idx, n_idx, out); });
            CK(cudaFuncSetAttribute(gatherC<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8 * 512));
            CK(cudaFuncSetAttribute(gatherC<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 16 * 512));
            snprintf(nm, 64, "LDGSTS U=8 ctas/SM=%d", occ);  time(nm, [&] { gatherC<8><<<sms * occ, 256, 8 * 2 * 8 * 512>>>(tab, idx, n_idx, out); });
            if (occ <= 1 || 8 * 2 * 16 * 512 * occ <= 220 * 1024) {
                snprintf(nm, 64, "LDGSTS U=16 ctas/SM=%d", occ); time(nm, [&] { gatherC<16><<<sms * occ, 256, 8 * 2 * 16 * 512>>>(tab, idx, n_idx, out); });
            }
        }
        for (int occ : {1}) {
            char nm[64];
            const size_t sm4 = 8 * 4 * 4096 + 8 * 4 * 8, sm8 = 8 * 8 * 4096 + 8 * 8 * 8, sm2 = 8 * 2 * 4096 + 8 * 2 * 8;
            CK(cudaFuncSetAttribute(gatherB<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            CK(cudaFuncSetAttribute(gatherB<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
            snprintf(nm, 64, "BULK stages=2 ctas/SM=%d", occ); time(nm, [&] { gatherB<2><<<sms * occ, 256, sm2>>>(tab, idx, n_idx, out); });
            snprintf(nm, 64, "BULK stages=4 ctas/SM=%d", occ); time(nm, [&] { gatherB<4><<<sms * occ, 256, sm4>>>(tab, idx, n_idx, out); });
            if (occ == 1) {
                CK(cudaFuncSetAttribute(gatherB<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 6 * 4096 + 8 * 6 * 8)));
                snprintf(nm, 64, "BULK stages=6 ctas/SM=%d", occ); time(nm, [&] { gatherB<6><<<sms * occ, 256, 8 * 6 * 4096 + 8 * 6 * 8>>>(tab, idx, n_idx, out); });
            }
        }
        CK(cudaFree(tab)); CK(cudaFree(idx)); CK(cudaFree(out));
    }
    return 0;
}
