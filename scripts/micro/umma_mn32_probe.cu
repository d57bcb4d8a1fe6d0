// Probe 2: tcgen05.mma kind::tf32 with MN-major operands in the ONE shared-memory layout CUTLASS allows for them
// (cutlass/gemm/collective/builders/sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem
// layout"): layout type 1 = SWIZZLE_128B_BASE32B, canonical form ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in elements, i.e. per
// K index (= node) a 128-byte row of 32 consecutive M / N elements, four such rows per atom, the 32-byte chunks of a row
// XOR-ed with the row index mod 4 (Swizzle<2,5,2> on the byte address), blocks of 32 M / N elements LBO apart, groups of four
// K rows SBO apart.  umma_mn_probe.cu (SWIZZLE_NONE) got zeros; this is the weight-gradient GEMM of dense_bwd:
//     D[m][n] = sum_node X[node][m] * Y[node][n]        X [128 nodes][32], Y [128 nodes][40], rows = nodes (node-major tiles)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gnn_fpga_b200/csrc -I../../include -o umma_mn32_probe umma_mn32_probe.cu
// usage: umma_mn32_probe [variant]   bit 0: swap LBO / SBO in the descriptors; bit 1: M = 64 (prints which TMEM lane holds which row of D)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gnnseg_tc.cuh"
using namespace gnnseg;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TM = 128, CX = 32, CY = 40;
constexpr int BLOCK_BYTES = TM * 128;                       // one block of 32 columns: 128 nodes x 128 bytes
constexpr int A_BYTES = 4 * BLOCK_BYTES, B_BYTES = 2 * BLOCK_BYTES;

// byte offset of (node, column c) inside a tile of 32-column blocks
__host__ __device__ inline int mn32_off(const int node, const int c) {
    const int mb = c >> 5, cc = c & 31;
    return mb * BLOCK_BYTES + (node >> 2) * 512 + (node & 3) * 128 + (((cc >> 3) ^ (node & 3)) << 5) + ((cc & 7) << 2);
}
__device__ __forceinline__ uint64_t smem_desc_mn32(const uint32_t addr, const uint32_t lbo, const uint32_t sbo) {
    return smem_desc(addr, lbo, sbo) | ((uint64_t)1 << 61);          // layout type 1: SWIZZLE_128B_BASE32B
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, const int variant) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle atoms want their own alignment
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < TM * CX; i += 128) { const int r = i / CX, c = i % CX; *reinterpret_cast<float*>(smem + mn32_off(r, c)) = X[i]; }
    for (int i = tid; i < TM * CY; i += 128) { const int r = i / CY, c = i % CY; *reinterpret_cast<float*>(smem + A_BYTES + mn32_off(r, c)) = Y[i]; }
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
        const uint32_t sa = smem_u32(smem), sb = sa + A_BYTES;
        const uint32_t ID = idesc_tf32((variant & 2) ? 64 : 128, CY) | (1u << 15) | (1u << 16);      // a_major = b_major = MN
        for (int kg = 0; kg < TM / 8; ++kg) {                                        // K = 8 nodes per instruction: two atoms of four rows
            uint64_t a, b;
            if (variant & 1) {
                a = smem_desc_mn32(sa + kg * 1024, 512, BLOCK_BYTES);
                b = smem_desc_mn32(sb + kg * 1024, 512, BLOCK_BYTES);
            } else {
                a = smem_desc_mn32(sa + kg * 1024, BLOCK_BYTES, 512);
                b = smem_desc_mn32(sb + kg * 1024, BLOCK_BYTES, 512);
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
    CK(cudaMemset(dD, 0, D.size() * 4));
    const int smem = A_BYTES + B_BYTES + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<<<1, 128, smem>>>(dX, dY, dD, variant);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (variant & 2) {          // X block 1 of A holds zeros, so rows 32..63 of D are zero rows: look for rows 0..31
        for (int l = 0; l < TM; ++l) {
            int hit = -1, nzl = 0;
            for (int k = 0; k < CY; ++k) nzl += D[l * CY + k] != 0.f;
            for (int m = 0; m < CX && hit < 0; ++m) {
                bool same = true;
                for (int k = 0; k < CY; ++k) same &= D[l * CY + k] == R[m * CY + k];
                if (same) hit = m;
            }
            if (hit >= 0 || nzl) printf("lane %3d: row %d (%d non-zero)\n", l, hit, nzl);
        }
        return 0;
    }
    double err = 0;
    int bad = 0, nz = 0;
    for (int m = 0; m < CX; ++m)
        for (int k = 0; k < CY; ++k) {
            const double e = fabs((double)D[m * CY + k] - R[m * CY + k]);
            if (e > err) err = e;
            bad += e > 1e-3;
            nz += D[m * CY + k] != 0.f;
        }
    int nz_rest = 0;
    for (int m = CX; m < TM; ++m) for (int k = 0; k < CY; ++k) nz_rest += D[m * CY + k] != 0.f;
    printf("variant %d: max |D - ref| over the 32 x 40 valid block = %g, %d of %d wrong, %d non-zero; rows 32..127 (zero blocks): %d non-zero\n",
           variant, err, bad, CX * CY, nz, nz_rest);
    printf("D[0][0..7]   :"); for (int k = 0; k < 8; ++k) printf(" %g", D[k]); printf("\nref[0][0..7] :"); for (int k = 0; k < 8; ++k) printf(" %g", R[k]);
    printf("\nD[1][0..7]   :"); for (int k = 0; k < 8; ++k) printf(" %g", D[CY + k]); printf("\nref[1][0..7] :"); for (int k = 0; k < 8; ++k) printf(" %g", R[CY + k]);
    printf("\nD[9][32..39] :"); for (int k = 32; k < 40; ++k) printf(" %g", D[9 * CY + k]); printf("\nref[9][32..39]:"); for (int k = 32; k < 40; ++k) printf(" %g", R[9 * CY + k]);
    printf("\n");
    return 0;
}
