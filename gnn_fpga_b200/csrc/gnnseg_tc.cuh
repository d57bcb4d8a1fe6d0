// PTX wrappers shared by the tcgen05 kernels (gnnseg_node_tc.cu, gnnseg_mlp_pipe.cu): shared-memory and
// instruction descriptors, tcgen05.mma / commit / ld / st, mbarriers, cp.async, the canonical K-major operand
// layout, and where the columns of a projection GEMM go.
#pragma once
#include <cuda.h>
#include "gnnseg_common.cuh"

namespace gnnseg {

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
// x = hi + lo for the 3xTF32 products: hi = x rounded to tf32 (cvt.rna), lo = (x - hi) rounded to tf32.  cvt runs on the
// 16-lane XU pipe that the tanh / ex2 of the same warps need (measured: the conversions were a third of the epilogue's
// critical path), so the second rounding is done on the integer pipe: add half an ulp of tf32 to the magnitude, clear the
// 13 low bits -- round to nearest, ties away, what cvt.rna does.  (x - hi is exact and finite for finite x; a NaN or an
// infinity shows in hi already.)
__device__ __forceinline__ void split3(const float x, float& hi, float& lo) {
    hi = __uint_as_float(tf32_rna(x));
    lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA store of one box (shared -> global) described by a tensor map; bulk async group per issuing thread
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const uint32_t smem_addr, const int x, const int y,
                                             const uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_addr), "r"(x), "r"(y), "l"(policy) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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

// 16-byte asynchronous copy global -> shared (LDGSTS): the weight images of a CTA's prologue are all
// in flight at once instead of one dependent load/store pair after the other (the prologue was 11 % of
// the MLP kernel's warp time: 296 CTAs each pull the same 58 KB through L2)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// byte offset of element (row, k) in a canonical K-major no-swizzle operand
__device__ __forceinline__ int canon_off(const int row, const int k, const int sbo) {
    return (row >> 3) * sbo + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4;
}

// Where the columns of the projection GEMM go (see ProjOut in gnnseg_common.cuh).
template <int H>
__device__ __forceinline__ float* proj_ptr(const ProjOut& o, const int node, const int col, int& ld) {
    if (o.mode == 1) { ld = 5 * H; return o.p + (size_t)node * 5 * H + col; }
    if (col < 2 * H) { ld = 2 * H; return o.p + (size_t)node * 2 * H + col; }
    ld = 3 * H;
    return o.q + (size_t)node * 3 * H + (col - 2 * H);
}
template <int H>
__device__ __forceinline__ bool proj_is_exp(const ProjOut& o, const int col) {
    return o.mode == 1 && col < 2 * H;
}
// A warp's swizzled 32 x 32 store tile (16-byte chunk j of row r at chunk j ^ (r & 7)) -> 32 rows of global memory, 128
// bytes each: every store instruction writes four full lines.  `p0` = column c0 of the warp's first row, LD = row stride
// in floats (compile time: the eight row offsets fold into the instructions), rows_valid = how many of the 32 rows exist.
// All eight shared-memory loads are issued before the first store: written as load / store pairs the compiler reuses ONE
// register quad and every pair waits for the one before (measured: 2.3 k cycles per 4 KB tile and warp).
template <int LD>
__device__ __forceinline__ void store_tile_rows(const float* sOut, float* p0, const int rows_valid, const int lane, const uint64_t policy) {
    const int rq = lane >> 3, jj = lane & 7;
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rq;
        t[i] = lds4(sOut + r * 32 + ((jj ^ (r & 7)) << 2));
    }
    float* p = p0 + (size_t)rq * LD + 4 * jj;
    if (rows_valid >= 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) st4_hint(p + (size_t)(4 * i) * LD, t[i], policy);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (4 * i + rq < rows_valid) st4_hint(p + (size_t)(4 * i) * LD, t[i], policy);
    }
}
// the same, row stride picked from where the columns go (ProjOut): state rows, or P' / Q'
template <int H>
__device__ __forceinline__ void store_tile_proj(const ProjOut& o, const float* sOut, const int node_w0, const int c0, const int rows_valid,
                                                const int lane, const uint64_t policy) {
    int ld;
    float* p0 = proj_ptr<H>(o, node_w0, c0, ld);
    if (o.mode == 1) store_tile_rows<5 * H>(sOut, p0, rows_valid, lane, policy);
    else if (c0 < 2 * H) store_tile_rows<2 * H>(sOut, p0, rows_valid, lane, policy);
    else store_tile_rows<3 * H>(sOut, p0, rows_valid, lane, policy);
}
// 32 consecutive fp32 columns of this thread's TMEM lane -> registers (one wait for both halves)
__device__ __forceinline__ void tmem_ld32(const uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// edge projection -> 2^(log2e v), exponent clamped to [-63, 63] (gnnseg_fused.cu); a clamp raises the flag
__device__ __forceinline__ void to_exponentials(float (&v)[16], int* __restrict__ flag) {
    bool clamped = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a = v[i] * 1.4426950408889634f;
        clamped |= !(fabsf(a) <= 63.f);                         // also a NaN
        float c, e;
        asm("min.NaN.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(63.f));        // the NaN-propagating forms: a NaN stays visible
        asm("max.NaN.f32 %0, %1, %2;" : "=f"(c) : "f"(c), "f"(-63.f));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(c));
        v[i] = e;
    }
    if (clamped && flag) atomicOr(flag, 1);
}

}  // namespace gnnseg
