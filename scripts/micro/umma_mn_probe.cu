// Probe: does tcgen05.mma kind::tf32 read MN-major operands out of tiles staged in the K-major canonical (no swizzle)
// layout the forward kernels use?  Tile X [128 nodes][32 cols] and Y [128 nodes][40 cols], rows = nodes.  Wanted:
//     D[m][n] = sum_node X[node][m] * Y[node][n]        (a weight-gradient GEMM: K = nodes, A = X^T, B = Y^T)
// i.e. A with M = cols of X (MN-major: M contiguous in memory), B with N = cols of Y (MN-major).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gnn_fpga_b200/csrc -I../../include -o umma_mn_probe umma_mn_probe.cu
// usage: umma_mn_probe [variant]   variant bit 0: swap the LBO / SBO fields of the MN-major descriptors
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gnnseg_tc.cuh"
using namespace gnnseg;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TM = 128, CX = 32, CY = 40;
constexpr int SBO_X = (CX / 4) * 128, SBO_Y = (CY / 4) * 128;      // K-major staging: 8-row group stride
constexpr int X_BYTES = (TM / 8) * SBO_X, Y_BYTES = (TM / 8) * SBO_Y;

__host__ __device__ constexpr uint32_t idesc_tf32_mn(const int M, const int N) {
    return idesc_tf32(M, N) | (1u << 15) | (1u << 16);              // a_major = b_major = MN
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, const int variant) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // stage both tiles exactly like the forward's loader: canon_off(row = node, k = col, SBO)
    for (int i = tid; i < TM * CX; i += 128) { const int r = i / CX, c = i % CX; *reinterpret_cast<float*>(smem + canon_off(r, c, SBO_X)) = X[i]; }
    for (int i = tid; i < TM * CY; i += 128) { const int r = i / CY, c = i % CY; *reinterpret_cast<float*>(smem + X_BYTES + 16384 + canon_off(r, c, SBO_Y)) = Y[i]; }
    if (tid == 0) { mbar_init(smem_u32(&mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t sa = smem_u32(smem), sx = sa, sy = sa + X_BYTES + 16384;
        constexpr uint32_t ID = idesc_tf32_mn(128, CY);
        if (variant & 2) {                                          // sanity: the K-major form the forward uses, D[node][n] = sum_c X[node][c] Y[n][c]
            constexpr uint32_t IDK = idesc_tf32(128, CY);
            for (int kq = 0; kq < CX / 8; ++kq)
                umma_ss(tmem, smem_desc(sx + kq * 256, 128, SBO_X), smem_desc(sy + kq * 256, 128, SBO_Y), IDK, kq > 0);
        } else if (variant & 4) {
            // A MN-major (M = cols of X, K = nodes), B K-major: N = rows of X (the first 40 nodes), K = ... not the same K:
            // use B = the X tile itself read K-major over its first 8-node groups as [N = 40 nodes][K = 8 cols] per step --
            // only a liveness test: is anything non-zero produced when ONLY a_major is set?
            constexpr uint32_t IDA = idesc_tf32(128, CY) | (1u << 15);
            for (int kg = 0; kg < 4; ++kg)
                umma_ss(tmem, smem_desc(sx + kg * SBO_X, (variant & 1) ? 128 : SBO_X, (variant & 1) ? SBO_X : 128),
                        smem_desc(sy + kg * 256, 128, SBO_Y), IDA, kg > 0);
        } else if (variant & 8) {
            constexpr uint32_t IDB = idesc_tf32(128, CY) | (1u << 16);
            for (int kg = 0; kg < 4; ++kg)
                umma_ss(tmem, smem_desc(sx + kg * 256, 128, SBO_X),
                        smem_desc(sy + kg * SBO_Y, (variant & 1) ? 128 : SBO_Y, (variant & 1) ? SBO_Y : 128), IDB, kg > 0);
        } else
        for (int kg = 0; kg < TM / 8; ++kg) {                       // one MMA per 8 nodes (K = 8)
            // MN-major: SBO = stride between groups of 4 along M / N (128 B here), LBO = stride between groups of 8 along K
            uint64_t a, b;
            if (variant & 1) {
                a = smem_desc(sx + kg * SBO_X, 128, SBO_X);
                b = smem_desc(sy + kg * SBO_Y, 128, SBO_Y);
            } else {
                a = smem_desc(sx + kg * SBO_X, SBO_X, 128);
                b = smem_desc(sy + kg * SBO_Y, SBO_Y, 128);
            }
            umma_ss(tmem, a, b, ID, kg > 0);
        }
        umma_commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), 0);
    tc_fence_after();
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 48; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + c0, v);
        for (int i = 0; i < 16; ++i)
            if (c0 + i < CY) D[(size_t)tid * CY + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float> X(TM * CX), Y(TM * CY), D(TM * CY), R(CX * CY, 0.f);
    for (int i = 0; i < TM * CX; ++i) X[i] = (float)((i * 7 + 3) % 5 - 2);          // small integers: exact in tf32
    for (int i = 0; i < TM * CY; ++i) Y[i] = (float)((i * 5 + 1) % 7 - 3);
    for (int n = 0; n < TM; ++n)
        for (int m = 0; m < CX; ++m)
            for (int k = 0; k < CY; ++k) R[m * CY + k] += X[n * CX + m] * Y[n * CY + k];
    float *dX, *dY, *dD;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dY, Y.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice));
    const int smem = X_BYTES + 16384 + Y_BYTES + 16384;             // slack behind each tile: M = 128 reads past the 32 columns
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(dX, dY, dD, variant);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (variant & 2) {
        double err = 0;
        for (int n = 0; n < TM; ++n)
            for (int k = 0; k < CY; ++k) {
                double r = 0;
                for (int c = 0; c < CX; ++c) r += (double)X[n * CX + c] * Y[k * CY + c];
                err = fmax(err, fabs(r - D[n * CY + k]));
            }
        printf("K-major sanity: max |D - ref| = %g   D[0][0..3] = %g %g %g %g\n", err, D[0], D[1], D[2], D[3]);
        return 0;
    }
    if (variant & 12) {
        int nz = 0; for (float v : D) nz += v != 0.f;
        printf("variant %d (liveness): %d of %d outputs non-zero; D[0][0..7] =", variant, nz, (int)D.size());
        for (int k = 0; k < 8; ++k) printf(" %g", D[k]);
        printf("\n");
        return 0;
    }
    double err = 0;
    int bad = 0;
    for (int m = 0; m < CX; ++m)
        for (int k = 0; k < CY; ++k) {
            const double e = fabs((double)D[m * CY + k] - R[m * CY + k]);
            if (e > err) err = e;
            bad += e > 1e-3;
        }
    printf("variant %d: max |D - ref| over the 32 x 40 valid block = %g, %d of %d wrong\n", variant, err, bad, CX * CY);
    printf("D[0][0..7]   :"); for (int k = 0; k < 8; ++k) printf(" %g", D[k]); printf("\nref[0][0..7] :"); for (int k = 0; k < 8; ++k) printf(" %g", R[k]);
    printf("\nD[1][0..7]   :"); for (int k = 0; k < 8; ++k) printf(" %g", D[CY + k]); printf("\nref[1][0..7] :"); for (int k = 0; k < 8; ++k) printf(" %g", R[CY + k]);
    printf("\nD[5][0..7]   :"); for (int k = 0; k < 8; ++k) printf(" %g", D[5 * CY + k]); printf("\nref[5][0..7] :"); for (int k = 0; k < 8; ++k) printf(" %g", R[5 * CY + k]);
    printf("\n");
    return 0;
}
