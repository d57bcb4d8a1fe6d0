// Node step on the 5th-generation tensor cores (tcgen05 + TMEM) for hidden_dim = 32.
//
// Same mathematics as node_kernel in gnnseg_forward.cu (gnn/model.py:113-125,154 of the
// reference) and the same producer side (staged CSR slices, branch-free batched row gather, no
// atomics, ascending-slot summation order).  What changes is the MLP: the three per-node GEMMs
//     h1 = tanh([mi|mo|self] . W3^T + b3)      128 x 112(108) x 32
//     H' = tanh(h1 . W4^T + b4)                128 x  32      x 32
//     P' = [H'|X] . W1^T (+ b1)                128 x  40(36)  x 64
// run as tcgen05.mma kind::tf32 with fp32 accumulators in tensor memory.  fp32 accuracy is kept
// with the 3xTF32 split: x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and
// a.b ~= lo_a.hi_b + hi_a.lo_b + hi_a.hi_b, all three accumulated into the same TMEM tile.
//
// Data flow per 128-node tile (one CTA per SM, persistent):
//   stagers (4 warps)     CSR slices of the tile after next -> (neighbour, weight) pairs in smem
//   gather  (16 warps)    4 lanes per node: row gather -> A1_hi / A1_lo in shared memory,
//                         canonical K-major core-matrix layout
//   thread 0              GEMM1 (SS: A1 from smem, W3 from smem) -> D1 in TMEM, commit -> mbarrier
//   epilogue (4 warps)    D1 -> regs -> +b3, tanh, split -> A2_hi / A2_lo in TMEM
//   thread 0              GEMM2 (TS: A2 from TMEM, W4 from smem) -> D2
//   epilogue              D2 -> +b4, tanh -> HX' to global; split, with X and zero pad -> A3 in TMEM
//   thread 0              GEMM3 (TS) -> D3
//   epilogue              D3 -> +b1 -> P' to global
// A1 is released to the producers as soon as GEMM1 has completed, so the gather of tile t+1
// overlaps everything after GEMM1 of tile t.
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {

template <int H>
struct TcCfg {
    static constexpr int TM   = 128;                 // nodes per tile = UMMA M
    static constexpr int EW   = 4;                   // MLP warps (issuer + epilogue): warp w owns TMEM lanes 32w..32w+31
    static constexpr int SW   = 4;                   // stager warps: CSR slices of the NEXT tile -> shared memory
    static constexpr int GW   = 16;                  // gather warps: 4 lanes per node, the whole tile at once
    static constexpr int ET   = EW * 32;
    static constexpr int ST   = SW * 32;
    static constexpr int GT   = GW * 32;
    static constexpr int NT   = ET + ST + GT;
    static constexpr int G    = 4;                   // lanes per node in the gather
    static_assert(GT / G == TM, "one gather group per node of the tile");
    static_assert(H == 32, "the gather lane map (2 float4 + 1 scalar per lane) is written for H = 32");
    static constexpr int D4   = H + 4;
    static constexpr int K1   = 3 * D4;
    static constexpr int K1P  = (K1 + 7) / 8 * 8;    // K of GEMM1, padded to the tf32 k-step
    static constexpr int D4P  = (D4 + 7) / 8 * 8;    // K of GEMM3
    // canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous 128 bytes;
    // LBO = distance between core matrices along K, SBO = distance between 8-row groups
    static constexpr int LBO     = 128;
    static constexpr int SBO_K1  = (K1P / 4) * LBO;
    static constexpr int SBO_H   = (H / 4) * LBO;
    static constexpr int SBO_D4  = (D4P / 4) * LBO;
    static constexpr int A1_BYTES = (TM / 8) * SBO_K1;       // one of hi / lo
    static constexpr int W3_BYTES = (H / 8) * SBO_K1;
    static constexpr int W4_BYTES = (H / 8) * SBO_H;
    static constexpr int W1_BYTES = (2 * H / 8) * SBO_D4;
    static constexpr int CAP  = 1536;                // staged CSR slots per direction per tile
    // shared memory map (bytes)
    static constexpr int O_A1H = 0;
    static constexpr int O_A1L = O_A1H + A1_BYTES;
    static constexpr int O_W3H = O_A1L + A1_BYTES;
    static constexpr int O_W3L = O_W3H + W3_BYTES;
    static constexpr int O_W4H = O_W3L + W3_BYTES;
    static constexpr int O_W4L = O_W4H + W4_BYTES;
    static constexpr int O_W1H = O_W4L + W4_BYTES;
    static constexpr int O_W1L = O_W1H + W1_BYTES;
    static constexpr int O_BIAS = O_W1L + W1_BYTES;          // b3, b4, b1
    static constexpr int STAGE_BYTES = 2 * CAP * 8 + 2 * (TM + 4) * 4;   // [2][CAP] int2 + [2][TM+4] int
    static constexpr int O_STAGE = O_BIAS + 3 * H * 4;       // two staging buffers (double buffered)
    static constexpr int O_MBAR = O_STAGE + 2 * STAGE_BYTES; // mbarrier (8 B) + tmem base (4 B)
    static constexpr int SMEM_BYTES = O_MBAR + 16;
    // tensor memory columns (fp32 cells, 128 lanes)
    static constexpr int C_D1  = 0;
    static constexpr int C_A2H = C_D1 + H;
    static constexpr int C_A2L = C_A2H + H;
    static constexpr int C_D2  = C_A2L + H;
    static constexpr int C_A3H = C_D2 + H;
    static constexpr int C_A3L = C_A3H + D4P;
    static constexpr int C_D3  = C_A3L + D4P;
    static constexpr int C_END = C_D3 + 2 * H;
    static constexpr int TMEM_COLS = C_END <= 32 ? 32 : C_END <= 64 ? 64 : C_END <= 128 ? 128 : C_END <= 256 ? 256 : 512;
    static_assert(C_END <= 512, "tensor memory");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

// ---- small PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_bar_sync(const int id, const int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void tc_bar_arrive(const int id, const int n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t tf32_rna(const float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split3(const float x, float& hi, float& lo) {
    hi = __uint_as_float(tf32_rna(x));
    lo = __uint_as_float(tf32_rna(x - hi));
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(const uint32_t addr, const uint32_t lbo, const uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t idesc_tf32(const int M, const int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(const uint32_t d_tmem, const uint64_t a_desc, const uint64_t b_desc,
                                        const uint32_t idesc, const uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts(const uint32_t d_tmem, const uint32_t a_tmem, const uint64_t b_desc,
                                        const uint32_t idesc, const uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(const uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(const uint32_t mbar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t mbar, const uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n"
        ::"r"(mbar), "r"(parity) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane <-> registers
__device__ __forceinline__ void tmem_ld16(const uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(const uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
          "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
          "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st8(const uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
          "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of element (row, k) in a canonical K-major no-swizzle operand
__device__ __forceinline__ int canon_off(const int row, const int k, const int sbo) {
    return (row >> 3) * sbo + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4;
}

// One CSR row of the gather (same contract as csr_row_sum in gnnseg_forward.cu, staged form):
// lane c of the node's 4-lane group owns float4 chunks c and c+4 of the hidden part and X[c].
// Slots go in batches of U: every row load of the batch is issued before the first FMA (no
// branch in between: an absent neighbour loads row 0 and is dropped by predication), then the
// FMAs run in ascending slot order.
template <int H>
__device__ __forceinline__ void tc_row_sum(const int2* __restrict__ pairs, const float* __restrict__ HX,
                                           const int beg, const int end, const int c, float4& acc0,
                                           float4& acc1, float& acc_x) {
    constexpr int D4 = H + 4, U = 4;
    for (int s0 = beg; s0 < end; s0 += U) {
        float w[U], vx[U];
        bool ok[U];
        float4 v0[U], v1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int2 pr = pairs[min(s0 + u, end - 1)];
            w[u] = __int_as_float(pr.y);
            ok[u] = (s0 + u < end) && pr.x >= 0;     // pr.x < 0: half edge, gathers the zero row
            const float* row = HX + (size_t)max(pr.x, 0) * D4;
            v0[u] = ldg4(row + 4 * c);
            v1[u] = ldg4(row + 16 + 4 * c);
            vx[u] = __ldg(row + H + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) {
                fma4(acc0, w[u], v0[u]);
                fma4(acc1, w[u], v1[u]);
                acc_x = fmaf(w[u], vx[u], acc_x);
            }
    }
}
// Path for a tile whose CSR slice does not fit the staging buffer: same batching, but the
// (neighbour, weight) pairs come straight from global memory (one more dependent load level).
template <int H>
__device__ __forceinline__ void tc_row_sum_direct(const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr,
                                                  const float* __restrict__ e, const float* __restrict__ HX,
                                                  const int beg, const int end, const int c, float4& acc0,
                                                  float4& acc1, float& acc_x) {
    constexpr int D4 = H + 4, U = 4;
    for (int s0 = beg; s0 < end; s0 += U) {
        float w[U], vx[U];
        bool ok[U];
        float4 v0[U], v1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = min(s0 + u, end - 1);
            const int nb = __ldg(nbr + s);
            w[u] = __ldg(e + __ldg(eid + s));
            ok[u] = (s0 + u < end) && nb >= 0;
            const float* row = HX + (size_t)max(nb, 0) * D4;
            v0[u] = ldg4(row + 4 * c);
            v1[u] = ldg4(row + 16 + 4 * c);
            vx[u] = __ldg(row + H + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) {
                fma4(acc0, w[u], v0[u]);
                fma4(acc1, w[u], v1[u]);
                acc_x = fmaf(w[u], vx[u], acc_x);
            }
    }
}
// split the lane's share of a row part and store it into A1_hi / A1_lo (canonical layout);
// part = 0 mi, 1 mo, 2 self
template <int H>
__device__ __forceinline__ void tc_store_part(unsigned char* __restrict__ a1h, unsigned char* __restrict__ a1l,
                                              const int sbo, const int ln, const int part, const int c,
                                              const float4 h0, const float4 h1, const float x) {
    constexpr int D4 = H + 4;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4 h = half ? h1 : h0;
        float4 hh, hl;
        split3(h.x, hh.x, hl.x); split3(h.y, hh.y, hl.y); split3(h.z, hh.z, hl.z); split3(h.w, hh.w, hl.w);
        const int off = canon_off(ln, part * D4 + 16 * half + 4 * c, sbo);
        *reinterpret_cast<float4*>(a1h + off) = hh;
        *reinterpret_cast<float4*>(a1l + off) = hl;
    }
    float xh, xl;
    split3(x, xh, xl);
    const int ox = canon_off(ln, part * D4 + H + c, sbo);
    *reinterpret_cast<float*>(a1h + ox) = xh;
    *reinterpret_cast<float*>(a1l + ox) = xl;
}

template <int H>
__global__ void __launch_bounds__(TcCfg<H>::NT, 1)
node_kernel_tc(const float* __restrict__ blob, const GnnsegGraph g, const float* __restrict__ HX_in,
               const float* __restrict__ e, const int n_tiles, float* __restrict__ HX_out,
               float* __restrict__ P_out) {
    using C = TcCfg<H>;
    using B = Blob<H>;
    constexpr int TM = C::TM, NT = C::NT, ET = C::ET, ST = C::ST, GT = C::GT, D4 = C::D4, K1 = C::K1;
    // named barriers: A1 full / empty (gather <-> MLP), MLP-internal, staging full / empty x2
    // (stager <-> gather), stager-internal
    constexpr int BAR_FULL = 1, BAR_EMPTY = 2, BAR_EPI = 3, BAR_SFULL = 4, BAR_SEMPTY = 6, BAR_STG = 8;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* a1h = smem + C::O_A1H;
    unsigned char* a1l = smem + C::O_A1L;
    float* sB3 = reinterpret_cast<float*>(smem + C::O_BIAS);
    float* sB4 = sB3 + H;
    float* sB1 = sB4 + H;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::O_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_MBAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_nodes = g.n_nodes;

    // ---- prologue: weights (hi/lo, canonical layout), biases, zero padding, TMEM, mbarrier ----
    for (int i = tid; i < H * C::K1P; i += NT) {          // W3: [H outputs][K1P]
        const int j = i / C::K1P, k = i % C::K1P;
        float hi, lo;
        split3(k < K1 ? __ldg(blob + B::W3 + k * H + j) : 0.f, hi, lo);
        const int off = canon_off(j, k, C::SBO_K1);
        *reinterpret_cast<float*>(smem + C::O_W3H + off) = hi;
        *reinterpret_cast<float*>(smem + C::O_W3L + off) = lo;
    }
    for (int i = tid; i < H * H; i += NT) {               // W4: [H][H]
        const int j = i / H, k = i % H;
        float hi, lo;
        split3(__ldg(blob + B::W4 + k * H + j), hi, lo);
        const int off = canon_off(j, k, C::SBO_H);
        *reinterpret_cast<float*>(smem + C::O_W4H + off) = hi;
        *reinterpret_cast<float*>(smem + C::O_W4L + off) = lo;
    }
    for (int i = tid; i < 2 * H * C::D4P; i += NT) {      // W1: [2H][D4P]
        const int j = i / C::D4P, k = i % C::D4P;
        float hi, lo;
        split3(k < D4 ? __ldg(blob + B::W1 + k * 2 * H + j) : 0.f, hi, lo);
        const int off = canon_off(j, k, C::SBO_D4);
        *reinterpret_cast<float*>(smem + C::O_W1H + off) = hi;
        *reinterpret_cast<float*>(smem + C::O_W1L + off) = lo;
    }
    for (int i = tid; i < H; i += NT) {
        sB3[i] = __ldg(blob + B::B3 + i);
        sB4[i] = __ldg(blob + B::B4 + i);
        sB1[i] = __ldg(blob + B::B1 + i);
    }
    for (int i = tid; i < TM * (C::K1P - K1); i += NT) {   // K padding of A1 stays zero for ever
        const int off = canon_off(i / (C::K1P - K1), K1 + i % (C::K1P - K1), C::SBO_K1);
        *reinterpret_cast<float*>(a1h + off) = 0.f;
        *reinterpret_cast<float*>(a1l + off) = 0.f;
    }
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();            // weights were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (tid >= ET + ST) {
        // ================================ gather warps ==================================
        constexpr int CAP = C::CAP;
        const int gt = tid - ET - ST;
        const int ln = gt >> 2, c = gt & 3;                   // node of the tile, lane in its group
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int sb = it & 1;
            const int2* sPair = reinterpret_cast<const int2*>(smem + C::O_STAGE + sb * C::STAGE_BYTES);
            const int* sPtr = reinterpret_cast<const int*>(smem + C::O_STAGE + sb * C::STAGE_BYTES + 2 * CAP * 8);
            const int n = tile * TM + ln;
            const bool live = n < n_nodes;
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            float sx = 0.f;
            if (live) {                                        // own row first: needs no staging
                const float* row = HX_in + (size_t)n * D4;
                s0 = ldg4(row + 4 * c);
                s1 = ldg4(row + 16 + 4 * c);
                sx = __ldg(row + H + c);
            }
            tc_bar_sync(BAR_SFULL + sb, ST + GT);              // this tile's CSR slices are staged
            const int ib = sPtr[0], ic = sPtr[TM] - ib;
            const int ob = sPtr[TM + 4], oc = sPtr[TM + 4 + TM] - ob;
            const bool staged = ic <= CAP && oc <= CAP;       // CTA-uniform
            const int i0 = sPtr[ln], i1 = live ? sPtr[ln + 1] : i0;
            const int o0 = sPtr[TM + 4 + ln], o1 = live ? sPtr[TM + 4 + ln + 1] : o0;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 i_0 = zero, i_1 = zero, o_0 = zero, o_1 = zero;
            float i_x = 0.f, o_x = 0.f;
            if (staged) {
                tc_row_sum<H>(sPair, HX_in, i0 - ib, i1 - ib, c, i_0, i_1, i_x);
                tc_row_sum<H>(sPair + CAP, HX_in, o0 - ob, o1 - ob, c, o_0, o_1, o_x);
            } else {
                tc_row_sum_direct<H>(g.in_eid, g.in_nbr, e, HX_in, i0, i1, c, i_0, i_1, i_x);
                tc_row_sum_direct<H>(g.out_eid, g.out_nbr, e, HX_in, o0, o1, c, o_0, o_1, o_x);
            }
            if (tile + 2 * (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_SEMPTY + sb, ST + GT);   // staging buffer free
            if (it >= 1) tc_bar_sync(BAR_EMPTY, ET + GT);      // GEMM1 of the previous tile has read A1
            tc_store_part<H>(a1h, a1l, C::SBO_K1, ln, 0, c, i_0, i_1, i_x);
            tc_store_part<H>(a1h, a1l, C::SBO_K1, ln, 1, c, o_0, o_1, o_x);
            tc_store_part<H>(a1h, a1l, C::SBO_K1, ln, 2, c, s0, s1, sx);
            fence_async_smem();                                // generic-proxy writes -> tensor core reads
            tc_bar_arrive(BAR_FULL, ET + GT);
        }
    } else if (tid >= ET) {
        // ================================ stager warps ==================================
        // Run ahead of the gather: row pointers, then (neighbour, edge weight) pairs of the tile's
        // two CSR slices, fetched with coalesced loads into one of two staging buffers.
        constexpr int CAP = C::CAP;
        const int stid = tid - ET;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int sb = it & 1;
            int2* sPair = reinterpret_cast<int2*>(smem + C::O_STAGE + sb * C::STAGE_BYTES);
            int* sPtr = reinterpret_cast<int*>(smem + C::O_STAGE + sb * C::STAGE_BYTES + 2 * CAP * 8);
            const int node0 = tile * TM;
            if (it >= 2) tc_bar_sync(BAR_SEMPTY + sb, ST + GT);   // the gather is done with this buffer
            for (int i = stid; i <= TM; i += ST) {
                const int n = min(node0 + i, n_nodes);
                sPtr[i] = __ldg(g.in_ptr + n);
                sPtr[TM + 4 + i] = __ldg(g.out_ptr + n);
            }
            tc_bar_sync(BAR_STG, ST);
            const int ib = sPtr[0], ic = sPtr[TM] - ib;
            const int ob = sPtr[TM + 4], oc = sPtr[TM + 4 + TM] - ob;
            if (ic <= CAP && oc <= CAP) {
                for (int s = stid; s < ic; s += ST)
                    sPair[s] = make_int2(__ldg(g.in_nbr + ib + s), __float_as_int(__ldg(e + __ldg(g.in_eid + ib + s))));
                for (int s = stid; s < oc; s += ST)
                    sPair[CAP + s] = make_int2(__ldg(g.out_nbr + ob + s), __float_as_int(__ldg(e + __ldg(g.out_eid + ob + s))));
            }
            tc_bar_arrive(BAR_SFULL + sb, ST + GT);
        }
    } else {
        // ================================ MLP (issuer + epilogue) ======================
        const uint32_t mb = smem_u32(mbar);
        const int row = warp * 32 + lane;                     // node within the tile = TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        constexpr uint32_t ID1 = idesc_tf32(TM, H), ID3 = idesc_tf32(TM, 2 * H);
        const uint32_t sa = smem_u32(smem);
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int node0 = tile * TM;
            const int n = node0 + row;
            const bool live = n < n_nodes;
            // ---- GEMM1: D1 = A1 . W3^T  (A and B from shared memory) ----------------------
            tc_bar_sync(BAR_FULL, ET + GT);
            if (tid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int kq = 0; kq < C::K1P / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;      // 8 tf32 = two 16-byte core columns
                    const uint64_t ah = smem_desc(sa + C::O_A1H + ko, C::LBO, C::SBO_K1);
                    const uint64_t al = smem_desc(sa + C::O_A1L + ko, C::LBO, C::SBO_K1);
                    const uint64_t bh = smem_desc(sa + C::O_W3H + ko, C::LBO, C::SBO_K1);
                    const uint64_t bl = smem_desc(sa + C::O_W3L + ko, C::LBO, C::SBO_K1);
                    umma_ss(tmem + C::C_D1, al, bh, ID1, kq > 0);
                    umma_ss(tmem + C::C_D1, ah, bl, ID1, 1);
                    umma_ss(tmem + C::C_D1, ah, bh, ID1, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            if (tile + (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_EMPTY, ET + GT);   // A1 may be refilled
            // ---- epilogue 1: h1 = tanh(D1 + b3) -> A2 (hi, lo) in TMEM ---------------------
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 16) {
                float v[16], hi[16], lo[16];
                tmem_ld16(lane_base + C::C_D1 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) split3(tanhf(v[i] + sB3[c0 + i]), hi[i], lo[i]);
                tmem_st16(lane_base + C::C_A2H + c0, hi);
                tmem_st16(lane_base + C::C_A2L + c0, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);
            // ---- GEMM2: D2 = A2 . W4^T  (A from tensor memory) ----------------------------
            if (tid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_W4H + ko, C::LBO, C::SBO_H);
                    const uint64_t bl = smem_desc(sa + C::O_W4L + ko, C::LBO, C::SBO_H);
                    umma_ts(tmem + C::C_D2, tmem + C::C_A2L + 8 * kq, bh, ID1, kq > 0);
                    umma_ts(tmem + C::C_D2, tmem + C::C_A2H + 8 * kq, bl, ID1, 1);
                    umma_ts(tmem + C::C_D2, tmem + C::C_A2H + 8 * kq, bh, ID1, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            // ---- epilogue 2: H' = tanh(D2 + b4) -> global HX', and [H'|X|0] -> A3 in TMEM ---
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 16) {
                float v[16], hi[16], lo[16];
                tmem_ld16(lane_base + C::C_D2 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = tanhf(v[i] + sB4[c0 + i]);
                    split3(v[i], hi[i], lo[i]);
                }
                if (live) {
                    float* dst = HX_out + (size_t)n * D4 + c0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) st4(dst + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
                tmem_st16(lane_base + C::C_A3H + c0, hi);
                tmem_st16(lane_base + C::C_A3L + c0, lo);
            }
            {                                                     // columns H..H+7: X and the zero K padding
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    x = ldg4(HX_in + (size_t)n * D4 + H);
                    st4(HX_out + (size_t)n * D4 + H, x);
                }
                float xh[8], xl[8];
                split3(x.x, xh[0], xl[0]); split3(x.y, xh[1], xl[1]);
                split3(x.z, xh[2], xl[2]); split3(x.w, xh[3], xl[3]);
#pragma unroll
                for (int i = 4; i < 8; ++i) xh[i] = xl[i] = 0.f;
                tmem_st8(lane_base + C::C_A3H + H, xh);
                tmem_st8(lane_base + C::C_A3L + H, xl);
            }
            tmem_st_wait();
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);
            // ---- GEMM3: D3 = A3 . W1^T ----------------------------------------------------
            if (tid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int kq = 0; kq < C::D4P / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_W1H + ko, C::LBO, C::SBO_D4);
                    const uint64_t bl = smem_desc(sa + C::O_W1L + ko, C::LBO, C::SBO_D4);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            // ---- epilogue 3: P' = D3 (+ b1 on the source half) -> global ---------------------
#pragma unroll
            for (int c0 = 0; c0 < 2 * H; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + C::C_D3 + c0, v);
                if (c0 < H) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += sB1[c0 + i];
                }
                if (live) {
                    float* dst = P_out + (size_t)n * 2 * H + c0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) st4(dst + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
            }
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);     // TMEM tiles are rewritten by the next tile
        }
    }
    // ---- teardown --------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

int launch_node_tc32(const float* blob, const GnnsegGraph* g, const float* HX_in, const float* e,
                     float* HX_out, float* P_out, cudaStream_t st) {
    using C = TcCfg<32>;
    if (g->n_nodes == 0) return GNNSEG_OK;
    const int n_tiles = (g->n_nodes + C::TM - 1) / C::TM;
    if (cudaFuncSetAttribute(node_kernel_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
        return GNNSEG_ECUDA;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
        return GNNSEG_ENODEVICE;
    const int grid = n_tiles < sms ? n_tiles : sms;
    node_kernel_tc<32><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, *g, HX_in, e, n_tiles, HX_out, P_out);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
