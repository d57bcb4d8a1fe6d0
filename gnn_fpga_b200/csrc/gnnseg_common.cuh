// Shared definitions for the gnnseg kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include "gnnseg.h"

namespace gnnseg {

// GnnsegGraph with writable members: what the graph-building kernels fill in
struct GnnsegGraphMut {
    int32_t* src; int32_t* dst;
    int32_t* in_ptr; int32_t* in_eid; int32_t* in_nbr; int32_t* in_pos;
    int32_t* out_ptr; int32_t* out_eid; int32_t* out_nbr; int32_t* out_pos;
};

// Where the columns [Ps|Pd|Qi|Qo|Qs] of a projection GEMM go (tcgen05 kernels, gnnseg_node_tc.cu):
//   mode 0  [Ps|Pd] -> rows of p (2H floats), [Qi|Qo|Qs] -> rows of q (3H floats)        the step-by-step path
//   mode 1  state rows of p (5H floats): [SPs|SPd|Qi|Qo|Qs], the edge projections stored as exponentials
//           2^(log2e v)                                                                   gnnseg_fused.cu
// n_cols: how many of the 5H columns are stored (2H when only the edge projections have a reader).
struct ProjOut {
    float* p;
    float* q;
    int mode;
    int n_cols;
    int* range_flag;     // nullable: raised when an exponent had to be clamped (modes 1, 2)
};

// Dataflow ("projection first", exact algebra of gnn/model.py:69-81,113-125):
//   every node carries, besides its features HX[n] = [H[n] | X[n]], five H-wide projections
//     Ps[n] = W1[:, 0:D].HX[n] + b1      Pd[n] = W1[:, D:2D].HX[n]            (edge step inputs)
//     Qi[n] = W3[:, 0:D].HX[n]           Qo[n] = W3[:, D:2D].HX[n]
//     Qs[n] = W3[:, 2D:3D].HX[n] + b3                                          (node step inputs)
//   edge step:  e_j  = sigmoid(W2 . tanh(Ps[src_j] + Pd[dst_j]) + b2)
//   node step:  h1[n] = tanh(Qs[n] + sum_{dst_j=n} e_j Qi[src_j] + sum_{src_j=n} e_j Qo[dst_j])
//               H'[n] = tanh(W4 . h1[n] + b4),  then the five projections of [H'[n] | X[n]].
//   (W3.[mi; mo; HX] = sum e.(W3a.HX[src]) + sum e.(W3b.HX[dst]) + W3c.HX: the segment sums move
//   behind the first linear layer, so a gathered row is one aligned H-float line.)
// State arrays: X4 (n,4) = X zero padded; P (n,2H) = [Ps|Pd]; Q (n,3H) = [Qi|Qo|Qs].
//
// Offsets (in floats) of the packed weight blob.  Matrices are stored transposed ([k][out]),
// and the D-wide input [hidden | X] is padded to D4 = h+4 rows (zero weights on the padding).
// Column map of the reference weights: SURVEY.md §3.1 / gnn/model.py:73,120,146.
template <int H>
struct Blob {
    static constexpr int D4  = H + 4;
    static constexpr int WIN = 0;                    // [4][H]      input_network.0.weight^T
    static constexpr int BIN = WIN + 4 * H;          // [H]
    static constexpr int WP  = BIN + H;              // [D4][5H]    [W1a | W1b | W3a | W3b | W3c]^T
    static constexpr int BP  = WP + D4 * 5 * H;      // [5H]        [b1 | 0 | 0 | 0 | b3]
    static constexpr int W2  = BP + 5 * H;           // [H]         edge layer 2
    static constexpr int B2  = W2 + H;               // [4]         b2, 0, 0, 0
    static constexpr int W4  = B2 + 4;               // [H][H]      node layer 2 ^T
    static constexpr int B4  = W4 + H * H;           // [H]
    // operand images for the tcgen05 kernels: W4 and WP as [out][k] matrices, split into tf32
    // hi / lo parts, in the canonical K-major core-matrix layout (8 rows x 16 bytes per core
    // matrix, core matrices contiguous along K) so that a CTA copies them to shared memory verbatim
    static constexpr int D4P = (D4 + 7) / 8 * 8;
    static constexpr int TC_W4H = B4 + H;            // [H][H]
    static constexpr int TC_W4L = TC_W4H + H * H;
    static constexpr int TC_WPH = TC_W4L + H * H;    // [5H][D4P]
    static constexpr int TC_WPL = TC_WPH + 5 * H * D4P;
    // the fused inference path (gnnseg_fused.cu): the edge network's second layer in the form the
    // reciprocal formulation wants
    static constexpr int W2N = TC_WPL + 5 * H * D4P;      // [H]  -2 w2
    static constexpr int Z0  = W2N + H;                   // [4]  b2 + sum(w2), 0, 0, 0
    static constexpr int SB1 = Z0 + 4;                    // [H]  2^(log2e b1): SPs of an absent start node
    static constexpr int TOTAL = SB1 + H;
};

__host__ __device__ inline int blob_total(int h) {
    const int d4 = h + 4, d4p = (d4 + 7) / 8 * 8;
    return 4 * h + h + d4 * 5 * h + 5 * h + h + 4 + h * h + h + 2 * h * h + 2 * 5 * h * d4p + 2 * h + 4;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 lds4(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
// L2 eviction policies: rows that are gathered several times per step (P, Q) are loaded
// evict_last; arrays that are written once and read by the NEXT kernel are stored evict_first
// so that they do not push the gathered rows out of the L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg4_hint(const float* p, const uint64_t policy) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st4_hint(float* p, const float4 v, const uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
// Programmatic dependent launch: a forward kernel lets its successor start (get resident, run
// its prologue: weights to shared memory, TMEM allocation) while it is still running; the
// successor calls pdl_wait() before it touches anything the predecessor writes, which returns when
// the predecessor grid has completed and its writes are visible.  Every kernel of the forward is
// a single resident wave, so a waiting successor never keeps a predecessor CTA from being scheduled.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-serialisation attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), const int grid, const int block, const size_t smem,
                              cudaStream_t st, const bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Immutable facts about the device and the kernels, looked up once per device instead of on every
// launch (the host-side launch path is on the critical path of the pipelined end-to-end call).
// Idempotent writes of the same values: safe without a lock.
constexpr int MAX_DEVICES = 64;
inline int cached_sm_count() {
    static std::atomic<int> cache[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return -1;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) return -1;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// opt the kernel in to `bytes` of dynamic shared memory (once per device)
template <auto Kern>
inline bool ensure_dynamic_smem(const int bytes) {
    static std::atomic<int> done[MAX_DEVICES];                        // largest size opted in so far
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return false;
    if (done[dev].load(std::memory_order_acquire) < bytes) {
        if (cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
        done[dev].store(bytes, std::memory_order_release);
    }
    return true;
}
// resident CTAs per SM of the kernel at this launch shape.  One cached value per (kernel, device), tagged with
// the launch shape it was computed for: another shape recomputes instead of reusing it.  The slot is one atomic
// 64-bit word (shape tag | occupancy), so concurrent host threads read either nothing or a complete entry.
template <auto Kern>
inline int cached_occupancy(const int threads, const size_t smem) {
    static std::atomic<unsigned long long> cache[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return -1;
    const unsigned long long tag = ((unsigned long long)(threads & 0xFFFF) << 48) | ((unsigned long long)(smem & 0xFFFFFFFFull) << 16);
    const unsigned long long v = cache[dev].load(std::memory_order_acquire);
    if ((v & ~0xFFFFull) == tag && (v & 0xFFFF) != 0) return (int)(v & 0xFFFF);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, Kern, threads, smem) != cudaSuccess || occ < 1) return -1;
    cache[dev].store(tag | (unsigned long long)(occ & 0xFFFF), std::memory_order_release);
    return occ;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) with the hardware ex2 / rcp approximations: 2 MUFU + 3 FMA-pipe
// instructions, branch free (tanhf is ~20 instructions over two divergent paths).  The formula holds
// for both signs (x -> -inf: exp -> 0, result -1; x -> +inf: exp -> inf, rcp -> 0, result +1; NaN
// propagates), so no |x| / copysign is needed.  Absolute error <= ~2e-7 over the whole range; the
// relative error near 0 is larger than tanhf's, which the MLPs do not see: every use feeds a dot
// product.  Measured end to end: edge scores agree with the reference to ~3e-7 relative (gate 1e-5).
__device__ __forceinline__ float tanh_fast(const float x) {
    float e, r;
    const float a = x * 2.885390081777927f;                   // 2 x log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}
// The same with one Newton step on the reciprocal (+2 FMA): the rcp.approx error (1 ulp of 2/(e+1), up to
// 2.4e-7 absolute near -1) goes away, what is left is ex2.approx's 2^-22 relative error on e, at most
// 1.2e-7 absolute on the result.  Used where a tanh feeds the recurrence directly (the node network's
// two activations, the input network's): n*h evaluations per step, not E*h.
__device__ __forceinline__ float tanh_node(const float x) {
#ifdef GNNSEG_NODE_TANH_FAST
    return tanh_fast(x);
#else
    float e, r;
    const float a = fminf(x * 2.885390081777927f, 126.f);      // e stays finite: the Newton step would turn inf * 0 into NaN
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));       // (fminf keeps a NaN a NaN only by accident: see below)
    const float d = e + 1.f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(r, fmaf(-d, r, 1.f), r);
    const float t = fmaf(-2.f, r, 1.f);
    return x != x ? x : t;                                      // fminf(NaN, 126) = 126: put the NaN back
#endif
}
// the edge network's hidden activation (E*h evaluations per step): the 5-instruction form
__device__ __forceinline__ float tanh_edge(const float x) {
#ifdef GNNSEG_EDGE_TANH_ACC
    return tanh_node(x);
#else
    return tanh_fast(x);
#endif
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

// Row stride (floats) for a [rows][k] shared tile read as float4 by up to 8 consecutive
// rows at once: stride/4 odd makes those 8 float4 land on distinct banks.
__host__ __device__ constexpr int tile_stride(int k) { return ((k / 4) & 1) ? k : k + 4; }

}  // namespace gnnseg
