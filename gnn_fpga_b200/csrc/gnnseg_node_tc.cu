// Node step on the 5th-generation tensor cores (tcgen05 + TMEM) for hidden_dim = 32.
//
// Same mathematics and dataflow as node_kernel in gnnseg_forward.cu (projection-first form of
// gnn/model.py:113-125,154, see gnnseg_common.cuh) and the same gather contract (staged CSR
// slices, branch-free batched row gather, no atomics, own term then in-edges then out-edges in
// ascending slot order).  What changes is the MLP after the gather: the two per-node GEMMs
//     H'      = tanh(h1 . W4^T + b4)                        128 x 32 x 32
//     [P'|Q'] = [H'|X] . [W1a|W1b|W3a|W3b|W3c]^T + bias     128 x 40(36) x 160
// run as tcgen05.mma kind::tf32 with fp32 accumulators in tensor memory.  fp32 accuracy is kept
// with the 3xTF32 split: x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and
// a.b ~= lo_a.hi_b + hi_a.lo_b + hi_a.hi_b, all three accumulated into the same TMEM tile.
//
// Data flow per 128-node tile (one CTA per SM, persistent):
//   stagers (4 warps)     CSR slices of the tile after next -> (neighbour, weight) pairs in smem
//   gather  (16 warps)    8 lanes per node, one aligned 128-byte Q row per edge and instruction:
//                         h1 = tanh(Qs + sum e Qi + sum e Qo) -> A_hi / A_lo in shared memory,
//                         canonical K-major core-matrix layout (double buffered)
//   thread 0              GEMM2 (SS: A from smem, W4 from smem) -> D2 in TMEM, commit -> mbarrier
//   epilogue (4 warps)    D2 -> regs -> +b4, tanh, split, X and zero pad appended -> A3 in TMEM
//   thread 0              GEMM3 (TS: A3 from TMEM, WP from smem) -> D3 (160 columns)
//   epilogue              D3 -> +bias -> P' and Q' to global
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "gnnseg_common.cuh"
#include "gnnseg_tc.cuh"

namespace gnnseg {

bool use_pdl(int n_slots);   // gnnseg_forward.cu

// gnnseg_mlp_pipe.cu: the warp-specialised form of the MLP and input kernels of this file (the default for hidden_dim 32
// and 64; GNNSEG_MLP_PIPE=0 picks the kernels below)
int launch_node_mlp_pipe(const float*, const float*, const float*, int, int, int, const ProjOut&, float*, bool, cudaStream_t);
int launch_input_pipe(const float*, const float*, int, int, int, float*, const ProjOut&, float*, cudaStream_t);
static bool mlp_pipe() {
    static const bool v = [] { const char* e = std::getenv("GNNSEG_MLP_PIPE"); return !(e && e[0] == '0'); }();
    return v;
}

// `make trace` builds ../libgnnseg_trace.so with -DGNNSEG_TRACE: node_mlp_kernel_tc<32> then records
// clock64 stamps per role of CTA 0 at the phase boundaries of every tile and %globaltimer at entry /
// exit of every CTA (scripts/mlp_trace.py reads them).  The shipped library has none of this.
#ifdef GNNSEG_TRACE
__device__ long long g_trace[2][16][12];
__device__ unsigned long long g_cta[512][2];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE_MLP(t, k) do { if (blockIdx.x == 0 && tid == 0 && (t) < 16) g_trace[0][(t)][(k)] = clock64(); } while (0)
#define TRACE_LDR(t, k) do { if (blockIdx.x == 0 && tid == ET && (t) < 16) g_trace[1][(t)][(k)] = clock64(); } while (0)
#define TRACE_CTA(k) do { if (tid == 0 && blockIdx.x < 512) g_cta[blockIdx.x][(k)] = gtimer(); } while (0)
#else
#define TRACE_MLP(t, k) do { } while (0)
#define TRACE_LDR(t, k) do { } while (0)
#define TRACE_CTA(k) do { } while (0)
#endif

template <int H>
struct TcCfg {
    static constexpr int TM   = 128;                 // nodes per tile = UMMA M
    static constexpr int EW   = 8;                   // MLP warps (issuer + epilogue): warp w owns TMEM lanes 32(w%4)..+31 and half of the columns
    static constexpr int SW   = 4;                   // stager warps
    static constexpr int GW   = 16;                  // gather warps
    static constexpr int ET   = EW * 32;
    static constexpr int ST   = SW * 32;
    static constexpr int GT   = GW * 32;
    static constexpr int NT   = ET + ST + GT;
    static constexpr int G    = H / 4;               // lanes per node in the gather: one float4 each
    static constexpr int NGRP = GT / G;              // nodes gathered concurrently
    static_assert(TM % NGRP == 0, "whole passes");
    static_assert(ST == TM, "one stager thread per row pointer");
    static constexpr int D4   = H + 4;
    static constexpr int D4P  = (D4 + 7) / 8 * 8;    // K of GEMM3, padded to the tf32 k-step
    static constexpr int NP   = 5 * H;               // projection width
    static_assert(H % 16 == 0 && NP % 16 == 0 && NP <= 256, "UMMA N constraints at M = 128");
    // canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous 128 bytes;
    // LBO = distance between core matrices along K, SBO = distance between 8-row groups
    static constexpr int LBO     = 128;
    static constexpr int SBO_H   = (H / 4) * LBO;
    static constexpr int SBO_D4  = (D4P / 4) * LBO;
    static constexpr int A_BYTES  = (TM / 8) * SBO_H;        // one of hi / lo, one buffer
    static constexpr int W4_BYTES = (H / 8) * SBO_H;
    static constexpr int WP_BYTES = (NP / 8) * SBO_D4;
    static constexpr int CAP  = 1536;                // staged CSR slots per direction per tile
    // shared memory map (bytes)
    static constexpr int O_A    = 0;                          // [2 buffers][hi, lo]
    static constexpr int O_W4H  = O_A + 4 * A_BYTES;
    static constexpr int O_W4L  = O_W4H + W4_BYTES;
    static constexpr int O_WPH  = O_W4L + W4_BYTES;
    static constexpr int O_WPL  = O_WPH + WP_BYTES;
    static constexpr int O_BIAS = O_WPL + WP_BYTES;           // b4 [H] (the projection bias travels inside the weight image)
    static constexpr int STAGE_BYTES = 2 * CAP * 8 + 2 * (TM + 4) * 4;   // [2][CAP] int2 + [2][TM+4] int
    static constexpr int O_STAGE = O_BIAS + 6 * H * 4;        // two staging buffers
    static constexpr int OUT_STRIDE = 36;                     // floats per staged output row (32 + pad)
    static constexpr int OUT_BYTES = 32 * OUT_STRIDE * 4;     // per epilogue warp: 32 rows x 32 columns
    static constexpr int O_OUT  = O_STAGE + 2 * STAGE_BYTES;  // [EW] output transposition buffers
    static constexpr int O_MBAR = O_OUT + EW * OUT_BYTES;     // mbarrier (8 B) + tmem base (4 B)
    static constexpr int SMEM_BYTES = O_MBAR + 16;
    // tensor memory columns (fp32 cells, 128 lanes)
    static constexpr int C_D2  = 0;
    static constexpr int C_A3H = C_D2 + H;
    static constexpr int C_A3L = C_A3H + D4P;
    static constexpr int C_D3  = C_A3L + D4P;
    static constexpr int C_END = C_D3 + NP;
    static constexpr int TMEM_COLS = C_END <= 32 ? 32 : C_END <= 64 ? 64 : C_END <= 128 ? 128 : C_END <= 256 ? 256 : 512;
    static_assert(C_END <= 512, "tensor memory");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};


// One CSR row of the gather, one float4 chunk of the aligned Q row per lane:
// acc += w_s * Qcol[nbr_s] for s = beg..end-1 in ascending order.  Slots go in batches of U: all
// row loads of the batch are issued before the first FMA (no branch in between: an absent slot or
// neighbour is predicated off).  STAGED: pairs come from shared memory.
template <bool STAGED>
__device__ __forceinline__ void tc_row_sum(const int2* __restrict__ pairs, const int32_t* __restrict__ nbr,
                                           const float* __restrict__ ew, const float* __restrict__ Qcol,
                                           const int row_floats, const int beg, const int end,
                                           const uint64_t keep, float4& acc) {
    constexpr int U = 8;
    for (int s0 = beg; s0 < end; s0 += U) {
        float w[U];
        bool ok[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = min(s0 + u, end - 1);
            int nb;
            if (STAGED) {
                const int2 pr = pairs[s];
                nb = pr.x;
                w[u] = __int_as_float(pr.y);
            } else {
                nb = __ldg(nbr + s);
                w[u] = __ldg(ew + s);
            }
            ok[u] = (s0 + u < end) && nb >= 0;       // nb < 0: half edge, gathers the zero row
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[u]) v[u] = ldg4_hint(Qcol + (size_t)nb * row_floats, keep);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) fma4(acc, w[u], v[u]);
    }
}

// Epilogue of the projection GEMM: D3 (this warp's 32 TMEM lanes = 32 consecutive nodes, NP
// columns; the bias is already in it) -> P' and Q' in global memory.  Each thread reads its own row from tensor
// memory; the 32x32 block is turned through a padded shared buffer so that every store
// instruction writes four full 128-byte lines (a row of P or Q is contiguous in global memory).
template <int H>
__device__ __forceinline__ void tc_store_projections(const uint32_t lane_base, const uint32_t col_d3,
                                                     float* __restrict__ sOut,
                                                     const int node_w0, const int n_nodes, const int lane,
                                                     const ProjOut out, const int chunk0 = 0,
                                                     const int chunk_step = 1) {
    constexpr int OS = TcCfg<H>::OUT_STRIDE;
    const uint64_t stream = l2_policy_evict_first();           // read next by another kernel, not by this one
    const int c_end = out.n_cols;
    // 32-column chunks; two warps that share a TMEM lane quarter take alternate chunks
#pragma unroll 1
    for (int c0 = 32 * chunk0; c0 < c_end; c0 += 32 * chunk_step) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float v[16];
            tmem_ld16(lane_base + col_d3 + c0 + 16 * hb, v);
            if (proj_is_exp<H>(out, c0)) to_exponentials(v, out.range_flag);
#pragma unroll
            for (int i = 0; i < 4; ++i)       // the bias is already in D3 (constant-1 column of A x bias row of WP)
                st4(sOut + lane * OS + 16 * hb + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        }
        __syncwarp();
        int ld;
        float* base = proj_ptr<H>(out, node_w0, c0, ld);
        float4 t[8];                                               // all loads before the first store (see store_tile_rows)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3), c4 = (lane & 7) * 4;
            t[i] = lds4(sOut + r * OS + c4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3), c4 = (lane & 7) * 4;
            if (node_w0 + r < n_nodes) st4_hint(base + (size_t)r * ld + c4, t[i], stream);
        }
        __syncwarp();
    }
}

template <int H>
__global__ void __launch_bounds__(TcCfg<H>::NT, 1)
node_kernel_tc(const float* __restrict__ blob, const GnnsegGraph g, const float* __restrict__ X4,
               const float* __restrict__ Q_in, const float* __restrict__ e_in, const float* __restrict__ e_out,
               const int n_tiles, float* __restrict__ P_out, float* __restrict__ Q_out, const int write_q,
               float* __restrict__ h1_save, float* __restrict__ H_save) {
    using C = TcCfg<H>;
    using B = Blob<H>;
    constexpr int TM = C::TM, NT = C::NT, ET = C::ET, ST = C::ST, GT = C::GT, D4 = C::D4, NP = C::NP;
    // named barriers: A full / empty per buffer (gather <-> MLP), MLP-internal, staging full /
    // empty per buffer (stager <-> gather), stager-internal
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_EPI = 5, BAR_SFULL = 6, BAR_SEMPTY = 8, BAR_STG = 10;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sB4 = reinterpret_cast<float*>(smem + C::O_BIAS);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::O_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_MBAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_nodes = g.n_nodes;

    // ---- prologue: weights (hi/lo, canonical layout), biases, TMEM, mbarrier ------------------
    // the four operand images (W4 hi, W4 lo, WP hi, WP lo) are contiguous in the blob and in smem
    static_assert(C::O_W4L == C::O_W4H + H * H * 4 && C::O_WPH == C::O_W4L + H * H * 4 &&
                  C::O_WPL == C::O_WPH + NP * C::D4P * 4, "image order");
    for (int i = tid * 4; i < 2 * H * H + 2 * NP * C::D4P; i += NT * 4)
        cp_async16(smem + C::O_W4H + i * 4, blob + B::TC_W4H + i);
    for (int i = tid; i < H; i += NT) sB4[i] = __ldg(blob + B::B4 + i);
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cp_async_wait_all();
    fence_async_smem();            // weights were written through the generic proxy
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                    // prologue done; Q_in / e_in / e_out come from the kernels before

    if (tid >= ET + ST) {
        // ================================ gather warps ==================================
        constexpr int CAP = C::CAP, G = C::G, NGRP = C::NGRP;
        const int gt = tid - ET - ST;
        const int grp = gt / G, c = gt % G;                   // node slot of the pass, float4 chunk
        const uint64_t keep = l2_policy_evict_last();          // gathered rows are re-read ~deg times
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int sb = it & 1;
            const int2* sPair = reinterpret_cast<const int2*>(smem + C::O_STAGE + sb * C::STAGE_BYTES);
            const int* sPtr = reinterpret_cast<const int*>(smem + C::O_STAGE + sb * C::STAGE_BYTES + 2 * CAP * 8);
            unsigned char* ah = smem + C::O_A + sb * 2 * C::A_BYTES;
            unsigned char* al = ah + C::A_BYTES;
            const int node0 = tile * TM;
            tc_bar_sync(BAR_SFULL + sb, ST + GT);              // this tile's CSR slices are staged
            if (it >= 2) tc_bar_sync(BAR_EMPTY + sb, ET + GT); // GEMM2 of tile it-2 has read this A buffer
            const int ib = sPtr[0], ic = sPtr[TM] - ib;
            const int ob = sPtr[TM + 4], oc = sPtr[TM + 4 + TM] - ob;
            const bool staged = ic <= CAP && oc <= CAP;       // CTA-uniform
#pragma unroll 1
            for (int ln = grp; ln < TM; ln += NGRP) {
                const int n = node0 + ln;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < n_nodes) {
                    acc = ldg4(Q_in + (size_t)n * 3 * H + 2 * H + 4 * c);            // Qs[n] (holds b3)
                    const int i0 = sPtr[ln], i1 = sPtr[ln + 1], o0 = sPtr[TM + 4 + ln], o1 = sPtr[TM + 4 + ln + 1];
                    if (staged) {
                        tc_row_sum<true>(sPair, nullptr, nullptr, Q_in + 4 * c, 3 * H, i0 - ib, i1 - ib, keep, acc);
                        tc_row_sum<true>(sPair + CAP, nullptr, nullptr, Q_in + H + 4 * c, 3 * H, o0 - ob, o1 - ob, keep, acc);
                    } else {
                        tc_row_sum<false>(nullptr, g.in_nbr, e_in, Q_in + 4 * c, 3 * H, i0, i1, keep, acc);
                        tc_row_sum<false>(nullptr, g.out_nbr, e_out, Q_in + H + 4 * c, 3 * H, o0, o1, keep, acc);
                    }
                    acc.x = tanh_node(acc.x); acc.y = tanh_node(acc.y); acc.z = tanh_node(acc.z); acc.w = tanh_node(acc.w);
                    if (h1_save) st4(h1_save + (size_t)n * H + 4 * c, acc);          // training: kept for the backward pass
                }
                float4 hh, hl;
                split3(acc.x, hh.x, hl.x); split3(acc.y, hh.y, hl.y); split3(acc.z, hh.z, hl.z); split3(acc.w, hh.w, hl.w);
                const int off = canon_off(ln, 4 * c, C::SBO_H);
                *reinterpret_cast<float4*>(ah + off) = hh;
                *reinterpret_cast<float4*>(al + off) = hl;
            }
            if (tile + 2 * (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_SEMPTY + sb, ST + GT);   // staging buffer free
            fence_async_smem();                                // generic-proxy writes -> tensor core reads
            tc_bar_arrive(BAR_FULL + sb, ET + GT);
        }
    } else if (tid >= ET) {
        // ================================ stager warps ==================================
        // Run ahead of the gather: row pointers, then (neighbour, edge weight) pairs of the tile's
        // two CSR slices, fetched with coalesced loads into one of two staging buffers.  The row
        // pointers of the NEXT tile are prefetched into registers while this tile's pairs are in
        // flight, and the pair loads are issued in batches of UNR per thread before any of them
        // is stored, so a tile costs about two memory latencies instead of one per slot.
        constexpr int CAP = C::CAP, UNR = 6;
        const int stid = tid - ET;
        int it = 0;
        // pointer prefetch registers: entry stid, and entry TM for thread 0
        int p_in = 0, p_out = 0, p_in_last = 0, p_out_last = 0;
        auto load_ptrs = [&](const int tile) {
            const int node0 = tile * TM;
            const int n = min(node0 + stid, n_nodes);
            p_in = __ldg(g.in_ptr + n);
            p_out = __ldg(g.out_ptr + n);
            if (stid == 0) {
                const int nl = min(node0 + TM, n_nodes);
                p_in_last = __ldg(g.in_ptr + nl);
                p_out_last = __ldg(g.out_ptr + nl);
            }
        };
        if ((int)blockIdx.x < n_tiles) load_ptrs(blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int sb = it & 1;
            int2* sPair = reinterpret_cast<int2*>(smem + C::O_STAGE + sb * C::STAGE_BYTES);
            int* sPtr = reinterpret_cast<int*>(smem + C::O_STAGE + sb * C::STAGE_BYTES + 2 * CAP * 8);
            if (it >= 2) tc_bar_sync(BAR_SEMPTY + sb, ST + GT);   // the gather is done with this buffer
            sPtr[stid] = p_in;
            sPtr[TM + 4 + stid] = p_out;
            if (stid == 0) { sPtr[TM] = p_in_last; sPtr[TM + 4 + TM] = p_out_last; }
            tc_bar_sync(BAR_STG, ST);
            const int ib = sPtr[0], ic = sPtr[TM] - ib;
            const int ob = sPtr[TM + 4], oc = sPtr[TM + 4 + TM] - ob;
            if (tile + (int)gridDim.x < n_tiles) load_ptrs(tile + gridDim.x);   // in flight during the pair loads
            if (ic <= CAP && oc <= CAP) {
#pragma unroll 1
                for (int base = 0; base < max(ic, oc); base += ST * UNR) {
                    int nb_i[UNR], nb_o[UNR];
                    float w_i[UNR], w_o[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int s = base + u * ST + stid;
                        if (s < ic) { nb_i[u] = __ldg(g.in_nbr + ib + s); w_i[u] = __ldg(e_in + ib + s); }
                        if (s < oc) { nb_o[u] = __ldg(g.out_nbr + ob + s); w_o[u] = __ldg(e_out + ob + s); }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int s = base + u * ST + stid;
                        if (s < ic) {
                            sPair[s] = make_int2(nb_i[u], __float_as_int(w_i[u]));
                            // pull the row the gather will want into the L2 one tile ahead
                            if (nb_i[u] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(Q_in + (size_t)nb_i[u] * 3 * H));
                        }
                        if (s < oc) {
                            sPair[CAP + s] = make_int2(nb_o[u], __float_as_int(w_o[u]));
                            if (nb_o[u] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(Q_in + (size_t)nb_o[u] * 3 * H + H));
                        }
                    }
                }
            }
            tc_bar_arrive(BAR_SFULL + sb, ST + GT);
        }
    } else {
        // ================================ MLP (issuer + epilogue) ======================
        const uint32_t mb = smem_u32(mbar);
        const int q = warp & 3, hf = warp >> 2;               // TMEM lane quarter, which half of the columns
        const int row = q * 32 + lane;                        // node within the tile = TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        constexpr uint32_t ID2 = idesc_tf32(TM, H), ID3 = idesc_tf32(TM, NP);
        const uint32_t sa = smem_u32(smem);
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int sb = it & 1;
            const int n = tile * TM + row;
            const bool live = n < n_nodes;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && hf == 0) x = ldg4(X4 + (size_t)n * 4); // early: independent of the gather
            // ---- GEMM2: D2 = h1 . W4^T  (A and B from shared memory) ----------------------
            tc_bar_sync(BAR_FULL + sb, ET + GT);
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = sa + C::O_A + sb * 2 * C::A_BYTES, a_lo = a_hi + C::A_BYTES;
#pragma unroll
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;      // 8 tf32 = two 16-byte core columns
                    const uint64_t ah = smem_desc(a_hi + ko, C::LBO, C::SBO_H);
                    const uint64_t al = smem_desc(a_lo + ko, C::LBO, C::SBO_H);
                    const uint64_t bh = smem_desc(sa + C::O_W4H + ko, C::LBO, C::SBO_H);
                    const uint64_t bl = smem_desc(sa + C::O_W4L + ko, C::LBO, C::SBO_H);
                    umma_ss(tmem + C::C_D2, al, bh, ID2, kq > 0);
                    umma_ss(tmem + C::C_D2, ah, bl, ID2, 1);
                    umma_ss(tmem + C::C_D2, ah, bh, ID2, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            if (tile + 2 * (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_EMPTY + sb, ET + GT);   // A buffer may be refilled
            // ---- epilogue 2: H' = tanh(D2 + b4); [H'|X|0] -> A3 (hi, lo) in TMEM -------------
            {
                const int c0 = hf * (H / 2);                       // this warp's 16 columns
                float v[16], hi[16], lo[16];
                tmem_ld16(lane_base + C::C_D2 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = tanh_node(v[i] + sB4[c0 + i]);
                    split3(v[i], hi[i], lo[i]);
                }
                if (H_save && live) {                                              // training: H' kept for the backward pass
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st4(H_save + (size_t)n * H + c0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
                tmem_st16(lane_base + C::C_A3H + c0, hi);
                tmem_st16(lane_base + C::C_A3L + c0, lo);
            }
            if (hf == 0) {                                        // columns H..H+7: X and the zero K padding
                float xh[8], xl[8];
                split3(x.x, xh[0], xl[0]); split3(x.y, xh[1], xl[1]);
                split3(x.z, xh[2], xl[2]); split3(x.w, xh[3], xl[3]);
#pragma unroll
                for (int i = 4; i < 8; ++i) xh[i] = xl[i] = 0.f;
                xh[4] = 1.f;                                      // times the bias row of the weight image
                tmem_st8(lane_base + C::C_A3H + H, xh);
                tmem_st8(lane_base + C::C_A3L + H, xl);
            }
            tmem_st_wait();
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);
            // ---- GEMM3: D3 = [H'|X] . WP^T  (A from tensor memory) --------------------------
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int kq = 0; kq < C::D4P / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_WPH + ko, C::LBO, C::SBO_D4);
                    const uint64_t bl = smem_desc(sa + C::O_WPL + ko, C::LBO, C::SBO_D4);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            // ---- epilogue 3: [P'|Q'] = D3 + bias -> global ------------------------------------
            tc_store_projections<H>(lane_base, C::C_D3, reinterpret_cast<float*>(smem + C::O_OUT + warp * C::OUT_BYTES),
                                    tile * TM + q * 32, n_nodes, lane, ProjOut{P_out, Q_out, 0, write_q ? NP : 2 * H, nullptr}, hf, 2);
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

// ------------------------------------------------------------------------------------------
// Node step, second half, on the tensor cores: the two per-node GEMMs on rows h1 that
// node_gather_kernel (gnnseg_forward.cu) left in global memory.  Same GEMMs, same 3xTF32 split and
// the same epilogues as node_kernel_tc, but no gather inside: the CTA needs 93 KB of shared memory
// and 256 tensor-memory columns, so TWO CTAs share an SM and the serial chain of one tile
// (barrier -> GEMM2 -> epilogue -> GEMM3 -> epilogue -> stores, ~9.6 k cycles, mostly latency) overlaps
// the other CTA's.  The store transposition buffers alias the A tile (free once GEMM2 has read it).
// ------------------------------------------------------------------------------------------
template <int H>
struct TcMlpCfg {
    using N = TcCfg<H>;
    static constexpr int TM = N::TM, EW = 8, LW = 4, ET = EW * 32, LT = LW * 32, NT = ET + LT;
    static constexpr int O_A    = 0;                            // hi, lo (one buffer); reused as [EW] 32x32 swizzled store tiles
    static constexpr int O_W4H  = O_A + 2 * N::A_BYTES;
    static constexpr int O_W4L  = O_W4H + N::W4_BYTES;
    static constexpr int O_WPH  = O_W4L + N::W4_BYTES;
    static constexpr int O_WPL  = O_WPH + N::WP_BYTES;
    static constexpr int O_BIAS = O_WPL + N::WP_BYTES;          // b4 [H] (the projection bias travels inside the weight image)
    static constexpr int O_MBAR = O_BIAS + 6 * H * 4;
    static constexpr int SMEM_BYTES = O_MBAR + 16;
    static_assert(EW * 32 * 32 * 4 <= 2 * N::A_BYTES, "store tiles fit in the A buffer");
    static constexpr int C_A3H = 0, C_A3L = N::D4P, C_D3 = 2 * N::D4P, C_D2 = C_D3;   // D2 is consumed before GEMM3 writes D3
    static constexpr int TMEM_COLS = 256;
    static_assert(C_D3 + N::NP <= TMEM_COLS, "tensor memory");
    static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");
};

// TMA = true: P', Q' leave through TMA stores (cp.async.bulk.tensor) of the swizzled 32 x 32 tiles: the
// hardware's 128-byte swizzle is the tile's own (16-byte chunk j of row r at j ^ (r & 7)), rows past
// n_nodes are clipped by the tensor map, and the MLP warps no longer read the tile back.
template <int H, bool TMA>
__global__ void __launch_bounds__(TcMlpCfg<H>::NT, 2)
node_mlp_kernel_tc(const float* __restrict__ blob, const float* __restrict__ X4, const float* h1,
                   const int ld_h1, const int n_nodes, const int n_tiles, const ProjOut out,
                   float* __restrict__ H_save,
                   const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                   const __grid_constant__ CUtensorMap tmQ16) {
    // h1 may alias the output rows (row n of h1 inside row n of P' / the state row): no __restrict__, no read-only loads on those
    float* const P_out = out.p;
    float* const Q_out = out.q;
    const int write_q = out.n_cols > 2 * H;
    using C = TcMlpCfg<H>;
    using N = TcCfg<H>;
    using B = Blob<H>;
    constexpr int TM = C::TM, NT = C::NT, ET = C::ET, LT = C::LT, NP = N::NP;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 2, BAR_EPI = 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the store tiles (offset 0, 4 KB apart) must sit on 1024-byte boundaries for the 128-byte swizzle
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* sB4 = reinterpret_cast<float*>(smem + C::O_BIAS);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::O_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_MBAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TRACE_MLP(15, 0);                                  // kernel entry
    TRACE_CTA(0);

    static_assert(C::O_W4L == C::O_W4H + H * H * 4 && C::O_WPH == C::O_W4L + H * H * 4 &&
                  C::O_WPL == C::O_WPH + NP * N::D4P * 4, "image order");
    for (int i = tid * 4; i < 2 * H * H; i += NT * 4)                     // W4 hi, lo
        cp_async16(smem + C::O_W4H + i * 4, blob + B::TC_W4H + i);
    for (int i = tid * 4; i < 2 * NP * N::D4P; i += NT * 4)               // projection images hi, lo (either row order)
        cp_async16(smem + C::O_WPH + i * 4, blob + B::TC_WPH + i);
    for (int i = tid; i < H; i += NT) sB4[i] = __ldg(blob + B::B4 + i);
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    TRACE_MLP(15, 1);                                  // copies issued, barrier + TMEM set up
    cp_async_wait_all();
    TRACE_MLP(15, 2);                                  // weight images have arrived
    fence_async_smem();
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                    // h1 comes from the gather kernel before
    TRACE_MLP(15, 3);                                  // prologue done

    if (tid >= ET) {
        // ================================ loader warps ==================================
        // h1 rows of the NEXT tile are fetched into registers while the MLP warps work on this one;
        // after the MLP warps have released the buffer the rows are split into tf32 hi / lo parts and
        // written as the canonical K-major operand tile.  A quarter warp holds 8 consecutive rows of
        // ONE float4 chunk, i.e. one contiguous 128-byte row group of the canonical tile: the stores
        // are bank-conflict free (8 lanes on one row would hit the same banks 8 ways).
        const int lt = tid - ET, lw = lt >> 5, r7 = lt & 7, cq = (lt >> 3) & 3;
        float4 v[8];
        auto fetch = [&](const int tile) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = tile * TM + (lw * 4 + (j >> 1)) * 8 + r7;
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < n_nodes) v[j] = lds4(h1 + (size_t)n * ld_h1 + 4 * (cq + 4 * (j & 1)));     // plain 16-byte load
            }
        };
        if ((int)blockIdx.x < n_tiles) fetch(blockIdx.x);
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            TRACE_LDR(it, 0);
            if (it >= 1) tc_bar_sync(BAR_EMPTY, ET + LT);       // the previous tile's stores have left the buffer
            TRACE_LDR(it, 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 hh, hl;
                split3(v[j].x, hh.x, hl.x); split3(v[j].y, hh.y, hl.y); split3(v[j].z, hh.z, hl.z); split3(v[j].w, hh.w, hl.w);
                const int off = canon_off((lw * 4 + (j >> 1)) * 8 + r7, 4 * (cq + 4 * (j & 1)), N::SBO_H);
                *reinterpret_cast<float4*>(smem + C::O_A + off) = hh;
                *reinterpret_cast<float4*>(smem + C::O_A + N::A_BYTES + off) = hl;
            }
            fence_async_smem();                                // generic-proxy writes -> tensor core reads
            tc_bar_arrive(BAR_FULL, ET + LT);
            TRACE_LDR(it, 2);
            if (tile + (int)gridDim.x < n_tiles) fetch(tile + gridDim.x);
            TRACE_LDR(it, 3);
        }
    } else {
        // ================================ MLP (issuer + epilogue) ======================
        const uint32_t mb = smem_u32(mbar);
        const int q = warp & 3, hf = warp >> 2;               // TMEM lane quarter, which half of the columns
        const int row = q * 32 + lane;                        // node within the tile = TMEM lane
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        constexpr uint32_t ID2 = idesc_tf32(TM, H), ID3 = idesc_tf32(TM, NP);
        const uint32_t sa = smem_u32(smem);
        float* sOut = reinterpret_cast<float*>(smem + C::O_A + warp * 4096);   // this warp's 32 x 32 store tile
        const uint64_t stream = l2_policy_evict_first();       // P', Q' are read next by another kernel
        uint32_t phase = 0;
        [[maybe_unused]] int itm = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++itm) {
            TRACE_MLP(itm, 0);
            const int n = tile * TM + row;
            const bool live = n < n_nodes;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && hf == 0) x = ldg4(X4 + (size_t)n * 4);
            // ---- GEMM2: D2 = h1 . W4^T ----------------------------------------------------
            tc_bar_sync(BAR_FULL, ET + LT);
            TRACE_MLP(itm, 1);
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = sa + C::O_A, a_lo = a_hi + N::A_BYTES;
#pragma unroll
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * N::LBO;
                    const uint64_t ah = smem_desc(a_hi + ko, N::LBO, N::SBO_H);
                    const uint64_t al = smem_desc(a_lo + ko, N::LBO, N::SBO_H);
                    const uint64_t bh = smem_desc(sa + C::O_W4H + ko, N::LBO, N::SBO_H);
                    const uint64_t bl = smem_desc(sa + C::O_W4L + ko, N::LBO, N::SBO_H);
                    umma_ss(tmem + C::C_D2, al, bh, ID2, kq > 0);
                    umma_ss(tmem + C::C_D2, ah, bl, ID2, 1);
                    umma_ss(tmem + C::C_D2, ah, bh, ID2, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            TRACE_MLP(itm, 2);
            // ---- epilogue 2: H' = tanh(D2 + b4); [H'|X|0] -> A3 (hi, lo) in TMEM -------------
            {
                const int c0 = hf * (H / 2);
                float v[16], hi[16], lo[16];
                tmem_ld16(lane_base + C::C_D2 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = tanh_node(v[i] + sB4[c0 + i]);
                    split3(v[i], hi[i], lo[i]);
                }
                if (H_save && live) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st4(H_save + (size_t)n * H + c0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
                tmem_st16(lane_base + C::C_A3H + c0, hi);
                tmem_st16(lane_base + C::C_A3L + c0, lo);
            }
            if (hf == 0) {
                float xh[8], xl[8];
                split3(x.x, xh[0], xl[0]); split3(x.y, xh[1], xl[1]);
                split3(x.z, xh[2], xl[2]); split3(x.w, xh[3], xl[3]);
#pragma unroll
                for (int i = 4; i < 8; ++i) xh[i] = xl[i] = 0.f;
                xh[4] = 1.f;                                      // times the bias row of the weight image
                tmem_st8(lane_base + C::C_A3H + H, xh);
                tmem_st8(lane_base + C::C_A3L + H, xl);
            }
            tmem_st_wait();
            tc_fence_before();
            TRACE_MLP(itm, 3);
            tc_bar_sync(BAR_EPI, ET);                          // A3 complete, every thread has read its D2 columns
            TRACE_MLP(itm, 4);
            // ---- GEMM3: D3 = [H'|X] . WP^T  (A from tensor memory) --------------------------
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int kq = 0; kq < N::D4P / 8; ++kq) {
                    const uint32_t ko = kq * 2 * N::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_WPH + ko, N::LBO, N::SBO_D4);
                    const uint64_t bl = smem_desc(sa + C::O_WPL + ko, N::LBO, N::SBO_D4);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                    umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
                }
                umma_commit(mb);
            }
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            TRACE_MLP(itm, 5);
            // ---- epilogue 3: [P'|Q'] = D3 + bias -> global, through a swizzled 32 x 32 tile per warp
            //      (16-byte chunk j of row r lives at chunk j ^ (r & 7)): conflict free both ways, every
            //      store instruction writes full 128-byte lines.  The two warps of a TMEM lane quarter
            //      share the 5 chunks of 32 columns as 2 + 2 + half of the last one each ----------------
            {
                const int node_w0 = tile * TM + q * 32;
                auto stage16 = [&](const int col, const int jb) {          // 16 columns of D3 -> chunks jb..jb+3 of the tile
                    float v[16];
                    tmem_ld16(lane_base + C::C_D3 + col, v);
                    if (proj_is_exp<H>(out, col)) to_exponentials(v, out.range_flag);
#pragma unroll
                    for (int i = 0; i < 4; ++i)       // the bias is already in D3 (constant-1 column of A x bias row of WP)
                        st4(sOut + lane * 32 + (((jb + i) ^ (lane & 7)) << 2),
                            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                };
                auto out_ptr = [&](const int col, int& ld) {               // column `col` of the projections for the warp's first node
                    return proj_ptr<H>(out, node_w0, col, ld);
                };
                const int n_full = write_q ? 2 : 1;                          // 32-column chunks hf, hf + 2 (P' only: chunk hf)
                auto chunk_of = [&](const int t) { return hf + 2 * t; };
                if constexpr (TMA) {
                    const uint32_t s_tile = smem_u32(sOut);
#pragma unroll 1
                    for (int t = 0; t < n_full; ++t) {
                        const int c0 = 32 * (hf + 2 * t);
                        float v0[16], v1[16];
                        tmem_ld16(lane_base + C::C_D3 + c0, v0);
                        tmem_ld16(lane_base + C::C_D3 + c0 + 16, v1);
                        if (lane == 0) tma_store_wait_read();                // the previous box has left the tile
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            st4(sOut + lane * 32 + ((i ^ (lane & 7)) << 2), make_float4(v0[4 * i], v0[4 * i + 1], v0[4 * i + 2], v0[4 * i + 3]));
                            st4(sOut + lane * 32 + (((4 + i) ^ (lane & 7)) << 2), make_float4(v1[4 * i], v1[4 * i + 1], v1[4 * i + 2], v1[4 * i + 3]));
                        }
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (c0 < 2 * H) tma_store_2d(&tmP, s_tile, c0, node_w0, stream);
                            else            tma_store_2d(&tmQ, s_tile, c0 - 2 * H, node_w0, stream);
                        }
                    }
                    if (write_q) {                                           // last chunk: 16 columns per warp, dense [32][16] tile
                        const int c0 = 128 + 16 * hf;
                        float v[16];
                        tmem_ld16(lane_base + C::C_D3 + c0, v);
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            st4(sOut + lane * 16 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) tma_store_2d(&tmQ16, s_tile, c0 - 2 * H, node_w0, stream);
                    }
                    if (lane == 0) tma_store_wait_read();                    // the tiles alias the A buffer of the next tile
                    __syncwarp();
                } else {
#pragma unroll 1
                    for (int t = 0; t < n_full; ++t) {
                        const int c0 = 32 * chunk_of(t);
                        stage16(c0, 0);
                        stage16(c0 + 16, 4);
                        __syncwarp();
                        store_tile_proj<H>(out, sOut, node_w0, c0, n_nodes - node_w0, lane, stream);
                        __syncwarp();
                    }
                    if (write_q) {                                               // last chunk: 16 columns per warp
                        const int c0 = 128 + 16 * hf;
                        stage16(c0, 0);
                        __syncwarp();
                        int ld;
                        float* base = out_ptr(c0, ld);
                        float4 t4[4];                                                // all loads before the first store
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = 8 * i + (lane >> 2), j = lane & 3;
                            t4[i] = lds4(sOut + r * 32 + ((j ^ (r & 7)) << 2));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = 8 * i + (lane >> 2), j = lane & 3;
                            if (node_w0 + r < n_nodes) st4_hint(base + (size_t)r * ld + 4 * j, t4[i], stream);
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            TRACE_MLP(itm, 6);
            tc_bar_sync(BAR_EPI, ET);     // TMEM tiles and the store tiles are rewritten by the next tile
            TRACE_MLP(itm, 7);
            if (tile + (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_EMPTY, ET + LT);
        }
        if (TMA && lane == 0) tma_store_wait_all();           // the boxes have reached global memory
    }
    TRACE_MLP(15, 4);                                  // this thread's tiles done
    tc_fence_before();
    __syncthreads();
    TRACE_MLP(15, 5);                                  // whole CTA done
    TRACE_CTA(1);
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// hidden_dim = 64: the same MLP on tcgen05.  What does not fit is the projection weight image: 320
// outputs x 72 (K padded) x {hi, lo} = 184 KB, and N = 320 exceeds one UMMA (N <= 256).  So GEMM3 runs as
// two halves of 160 outputs, and the half images (92 KB, hi + lo) are STREAMED through one shared
// buffer by the loader warps with cp.async: half 1 arrives while the MLP warps store the first 160
// columns (epilogue 3a), half 0 of the next tile while they store the last 160 (epilogue 3b).  One CTA
// per SM (190 KB of shared memory, 464 tensor-memory columns), 8 MLP + 8 loader warps.
// ------------------------------------------------------------------------------------------
struct Mlp64 {
    static constexpr int H = 64, TM = 128, D4 = 68, D4P = 72, NP = 320, NH = 160;
    static constexpr int EW = 8, LW = 8, ET = EW * 32, LT = LW * 32, NT = ET + LT;
    static constexpr int LBO = 128, SBO_H = (H / 4) * LBO, SBO_D4 = (D4P / 4) * LBO;
    static constexpr int A_BYTES = (TM / 8) * SBO_H;              // 32 KB, one of hi / lo
    static constexpr int W4_BYTES = (H / 8) * SBO_H;              // 16 KB
    static constexpr int WPH_BYTES = (NH / 8) * SBO_D4;           // 45 KB: one half of the outputs, one of hi / lo
    static constexpr int O_A = 0;                                 // hi, lo; the first 32 KB double as [EW] swizzled 32x32 store tiles
    static constexpr int O_W4H = O_A + 2 * A_BYTES;
    static constexpr int O_W4L = O_W4H + W4_BYTES;
    static constexpr int O_WP = O_W4L + W4_BYTES;                 // [hi half][lo half] of the half in flight
    static constexpr int O_BIAS = O_WP + 2 * WPH_BYTES;           // b4 [H] (the projection bias travels inside the weight image)
    static constexpr int O_MBAR = O_BIAS + 6 * H * 4;
    static constexpr int SMEM_BYTES = O_MBAR + 16;
    static constexpr int C_A3H = 0, C_A3L = D4P, C_D3 = 2 * D4P, C_D2 = C_D3, TMEM_COLS = 512;
    static_assert(C_D3 + NP <= TMEM_COLS && SMEM_BYTES <= 227 * 1024, "resources");
};

// INPUT = true is the input step (gnn/model.py:144-146): there is no GEMM2; the MLP warps compute
// H0 = tanh(Win.X + bin) (K = F <= 4) straight into the A3 operand in tensor memory, the loader warps
// only stream the weight halves.  X4 is then an output (X zero padded), Xraw the (n, F) input.
template <bool INPUT>
__global__ void __launch_bounds__(Mlp64::NT, 1)
node_mlp_kernel_tc64(const float* __restrict__ blob, float* __restrict__ X4, const float* h1, const int ld_h1,
                     const int n_nodes, const int n_tiles, const ProjOut out,
                     float* __restrict__ H_save, const float* __restrict__ Xraw, const int F) {
    using C = Mlp64;
    using B = Blob<64>;
    // with n_cols <= 160 only the first half of the projection image is needed (P' alone, or [SPs|SPd]): it
    // stays in shared memory for the whole launch
    const int write_q = out.n_cols > C::NH;
    constexpr int H = C::H, TM = C::TM, NT = C::NT, ET = C::ET, LT = C::LT;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 2, BAR_EPI = 3, BAR_WPF_A = 4, BAR_WPF_B = 5, BAR_WPE_A = 6, BAR_WPE_B = 7;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sB4 = reinterpret_cast<float*>(smem + C::O_BIAS);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::O_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_MBAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // W4 hi, lo (contiguous in the blob and in smem); the WP halves are streamed by the loader warps
    float* sWin = reinterpret_cast<float*>(smem + C::O_W4H);      // INPUT: Win^T [4][H] and bin [H] live where W4 would
    if (INPUT) {
        for (int i = tid; i < 5 * H; i += NT) sWin[i] = __ldg(blob + B::WIN + i);      // Win and bin are contiguous
    } else {
        for (int i = tid * 4; i < 2 * H * H; i += NT * 4) cp_async16(smem + C::O_W4H + i * 4, blob + B::TC_W4H + i);
        for (int i = tid; i < H; i += NT) sB4[i] = __ldg(blob + B::B4 + i);
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
    cp_async_wait_all();
    fence_async_smem();
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (tid >= ET) {
        // ================================ loader warps ==================================
        const int lt = tid - ET, lw = lt >> 5, r7 = lt & 7, cq = (lt >> 3) & 3;
        // half `hf` of the projection weight images -> the shared buffer (hi then lo), then make it
        // visible to the tensor core and tell the MLP warps
        auto load_wp_half = [&](const int hf, const int bar) {
            constexpr int HALF_FLOATS = C::WPH_BYTES / 4, IMG_FLOATS = C::NP * C::D4P;
            for (int i = lt * 4; i < HALF_FLOATS; i += LT * 4) {
                cp_async16(smem + C::O_WP + i * 4, blob + B::TC_WPH + hf * HALF_FLOATS + i);
                cp_async16(smem + C::O_WP + C::WPH_BYTES + i * 4, blob + B::TC_WPH + IMG_FLOATS + hf * HALF_FLOATS + i);
            }
            cp_async_wait_all();
            fence_async_smem();
            tc_bar_arrive(bar, ET + LT);
        };
        float4 v[8];
        auto fetch = [&](const int tile) {                  // 8 of the tile's 2048 float4: quarter warp = 8 rows of one chunk
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = tile * TM + (lw * 2 + (j >> 2)) * 8 + r7;
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < n_nodes) v[j] = lds4(h1 + (size_t)n * ld_h1 + 4 * (cq + 4 * (j & 3)));
            }
        };
        if ((int)blockIdx.x < n_tiles) {
            if (!INPUT) fetch(blockIdx.x);
            load_wp_half(0, BAR_WPF_A);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const bool more = tile + (int)gridDim.x < n_tiles;
            if (!INPUT) {
                if (it >= 1) tc_bar_sync(BAR_EMPTY, ET + LT);   // the previous tile's stores have left the buffer
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 hh, hl;
                    split3(v[j].x, hh.x, hl.x); split3(v[j].y, hh.y, hl.y); split3(v[j].z, hh.z, hl.z); split3(v[j].w, hh.w, hl.w);
                    const int off = canon_off((lw * 2 + (j >> 2)) * 8 + r7, 4 * (cq + 4 * (j & 3)), C::SBO_H);
                    *reinterpret_cast<float4*>(smem + C::O_A + off) = hh;
                    *reinterpret_cast<float4*>(smem + C::O_A + C::A_BYTES + off) = hl;
                }
                fence_async_smem();
                tc_bar_arrive(BAR_FULL, ET + LT);
                if (more) fetch(tile + gridDim.x);
            }
            if (write_q) {
                tc_bar_sync(BAR_WPE_A, ET + LT);                // GEMM3a has read half 0
                load_wp_half(1, BAR_WPF_B);
                tc_bar_sync(BAR_WPE_B, ET + LT);                // GEMM3b has read half 1
                if (more) load_wp_half(0, BAR_WPF_A);
            }
        }
    } else {
        // ================================ MLP (issuer + epilogue) ======================
        const uint32_t mb = smem_u32(mbar);
        const int q = warp & 3, hf = warp >> 2;               // TMEM lane quarter, which of its two warps
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        constexpr uint32_t ID2 = idesc_tf32(TM, H), ID3 = idesc_tf32(TM, C::NH);
        const uint32_t sa = smem_u32(smem);
        float* sOut = reinterpret_cast<float*>(smem + C::O_A + warp * 4096);
        const uint64_t stream = l2_policy_evict_first();
        uint32_t phase = 0;
        // 32 columns [c0, c0 + 32) of [P'|Q'] for this warp's 32 nodes: TMEM -> + bias -> swizzled tile -> full lines
        auto store_chunk = [&](const int c0, const int node_w0) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                float v[16];
                tmem_ld16(lane_base + C::C_D3 + c0 + 16 * hb, v);
                if (proj_is_exp<H>(out, c0)) to_exponentials(v, out.range_flag);
#pragma unroll
                for (int i = 0; i < 4; ++i)           // the bias is already in D3
                    st4(sOut + lane * 32 + (((4 * hb + i) ^ (lane & 7)) << 2),
                        make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
            __syncwarp();
            store_tile_proj<H>(out, sOut, node_w0, c0, n_nodes - node_w0, lane, stream);
            __syncwarp();
        };
        auto gemm3_half = [&](const int half) {               // D3[:, 160 half ...] = [H'|X] . WP_half^T
            tc_fence_after();
#pragma unroll
            for (int kq = 0; kq < C::D4P / 8; ++kq) {
                const uint32_t ko = kq * 2 * C::LBO;
                const uint64_t bh = smem_desc(sa + C::O_WP + ko, C::LBO, C::SBO_D4);
                const uint64_t bl = smem_desc(sa + C::O_WP + C::WPH_BYTES + ko, C::LBO, C::SBO_D4);
                const uint32_t d = tmem + C::C_D3 + half * C::NH;
                umma_ts(d, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                umma_ts(d, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                umma_ts(d, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
            }
            umma_commit(mb);
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int n = tile * TM + row;
            const bool live = n < n_nodes;
            const int node_w0 = tile * TM + q * 32;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (INPUT) {
                if (live) {
                    float xv[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int f = 0; f < F; ++f) xv[f] = __ldg(Xraw + (size_t)n * F + f);
                    x = make_float4(xv[0], xv[1], xv[2], xv[3]);
                    if (hf == 0) st4(X4 + (size_t)n * 4, x);
                }
            } else if (live && hf == 0) {
                x = ldg4(X4 + (size_t)n * 4);
            }
            // ---- GEMM2: D2 = h1 . W4^T ----------------------------------------------------
            if (!INPUT) tc_bar_sync(BAR_FULL, ET + LT);
            if (!INPUT && tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = sa + C::O_A, a_lo = a_hi + C::A_BYTES;
#pragma unroll
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t ah = smem_desc(a_hi + ko, C::LBO, C::SBO_H);
                    const uint64_t al = smem_desc(a_lo + ko, C::LBO, C::SBO_H);
                    const uint64_t bh = smem_desc(sa + C::O_W4H + ko, C::LBO, C::SBO_H);
                    const uint64_t bl = smem_desc(sa + C::O_W4L + ko, C::LBO, C::SBO_H);
                    umma_ss(tmem + C::C_D2, al, bh, ID2, kq > 0);
                    umma_ss(tmem + C::C_D2, ah, bl, ID2, 1);
                    umma_ss(tmem + C::C_D2, ah, bh, ID2, 1);
                }
                umma_commit(mb);
            }
            if (!INPUT) { mbar_wait(mb, phase); phase ^= 1; }
            tc_fence_after();
            // ---- epilogue 2: H' = tanh(D2 + b4) (INPUT: tanh(Win.X + bin)); [H'|X|0] -> A3 (hi, lo) in TMEM ----
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const int c0 = hf * 32 + part * 16;
                float v[16], hi[16], lo[16];
                if (!INPUT) tmem_ld16(lane_base + C::C_D2 + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (INPUT) {
                        float t = sWin[4 * H + c0 + i];
                        t = fmaf(x.x, sWin[0 * H + c0 + i], t);
                        t = fmaf(x.y, sWin[1 * H + c0 + i], t);
                        t = fmaf(x.z, sWin[2 * H + c0 + i], t);
                        t = fmaf(x.w, sWin[3 * H + c0 + i], t);
                        v[i] = tanh_node(t);
                    } else {
                        v[i] = tanh_node(v[i] + sB4[c0 + i]);
                    }
                    split3(v[i], hi[i], lo[i]);
                }
                if (H_save && live) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st4(H_save + (size_t)n * H + c0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
                // D2 aliases the first columns of D3, A3 does not: safe to write while other lanes still read D2
                tmem_st16(lane_base + C::C_A3H + c0, hi);
                tmem_st16(lane_base + C::C_A3L + c0, lo);
            }
            if (hf == 0) {
                float xh[8], xl[8];
                split3(x.x, xh[0], xl[0]); split3(x.y, xh[1], xl[1]);
                split3(x.z, xh[2], xl[2]); split3(x.w, xh[3], xl[3]);
#pragma unroll
                for (int i = 4; i < 8; ++i) xh[i] = xl[i] = 0.f;
                xh[4] = 1.f;                                      // times the bias row of the weight image
                tmem_st8(lane_base + C::C_A3H + H, xh);
                tmem_st8(lane_base + C::C_A3L + H, xl);
            }
            tmem_st_wait();
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);
            // ---- GEMM3a: outputs 0..159 = P' (128) and the first 32 columns of Q' -----------------
            if (write_q || tile == (int)blockIdx.x)            // half 0 of the weight images is in shared memory
                tc_bar_sync(BAR_WPF_A, ET + LT);               // (it stays there when no tile needs half 1)
            if (tid == 0) gemm3_half(0);
            mbar_wait(mb, phase); phase ^= 1;
            tc_fence_after();
            if (write_q) tc_bar_arrive(BAR_WPE_A, ET + LT);    // the loader may stream half 1 in
            // epilogue 3a (chunks 0..4 of 32 columns; as many as the caller wants stored)
            for (int ch = hf; ch < 5 && 32 * ch < out.n_cols; ch += 2) store_chunk(32 * ch, node_w0);
            if (write_q) {
                // ---- GEMM3b: outputs 160..319 ------------------------------------------------
                tc_bar_sync(BAR_WPF_B, ET + LT);
                if (tid == 0) gemm3_half(1);
                mbar_wait(mb, phase); phase ^= 1;
                tc_fence_after();
                tc_bar_arrive(BAR_WPE_B, ET + LT);             // ... and half 0 for the next tile
                for (int ch = 5 + (hf ^ 1); ch < 10; ch += 2) store_chunk(32 * ch, node_w0);
            }
            tc_fence_before();
            tc_bar_sync(BAR_EPI, ET);
            if (!INPUT && tile + (int)gridDim.x < n_tiles) tc_bar_arrive(BAR_EMPTY, ET + LT);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// input step on the tensor cores: H0 = tanh(Win.X + bin) per thread (K = F <= 4), [H0|X|0] split
// into tensor memory, one projection GEMM, same epilogue.  (gnn/model.py:144-146)
// ------------------------------------------------------------------------------------------
template <int H>
struct TcInCfg {
    using N = TcCfg<H>;
    static constexpr int NT = 256;                          // 8 warps: two per TMEM lane quarter, each half of the columns
    static constexpr int O_WPH  = 0;
    static constexpr int O_WPL  = O_WPH + N::WP_BYTES;
    static constexpr int O_BIAS = O_WPL + N::WP_BYTES;      // Win [4][H], bin [H]
    static constexpr int O_OUT  = O_BIAS + (4 * H + H + 5 * H) * 4;
    static constexpr int O_MBAR = O_OUT + 8 * N::OUT_BYTES;
    static constexpr int SMEM_BYTES = O_MBAR + 16;
    static constexpr int C_A3H = 0, C_A3L = N::D4P, C_D3 = 2 * N::D4P, C_END = C_D3 + 5 * H;
    static constexpr int TMEM_COLS = C_END <= 64 ? 64 : C_END <= 128 ? 128 : C_END <= 256 ? 256 : 512;
};

template <int H>
__global__ void __launch_bounds__(TcInCfg<H>::NT)
input_kernel_tc(const float* __restrict__ blob, const float* __restrict__ X, const int F, const int n_nodes,
                const int n_tiles, float* __restrict__ X4, const ProjOut out,
                float* __restrict__ H_save) {
    using C = TcInCfg<H>;
    using N = TcCfg<H>;
    using B = Blob<H>;
    constexpr int NT = C::NT, D4 = N::D4, NP = N::NP, TM = N::TM;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sWin = reinterpret_cast<float*>(smem + C::O_BIAS);
    float* sBin = sWin + 4 * H;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::O_MBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_MBAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid * 4; i < 2 * NP * N::D4P; i += NT * 4)      // WP hi, WP lo: contiguous images
        cp_async16(smem + C::O_WPH + i * 4, blob + B::TC_WPH + i);
    for (int i = tid; i < 5 * H; i += NT) sWin[i] = __ldg(blob + B::WIN + i);      // Win and bin are contiguous
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cp_async_wait_all();
    fence_async_smem();
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t mb = smem_u32(mbar);
    const int q = warp & 3, hf = warp >> 2;                   // TMEM lane quarter, which half of the columns
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    constexpr uint32_t ID3 = idesc_tf32(TM, NP);
    const uint32_t sa = smem_u32(smem);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n = tile * TM + q * 32 + lane;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < n_nodes) {
            for (int f = 0; f < F; ++f) x[f] = __ldg(X + (size_t)n * F + f);
            if (hf == 0) st4(X4 + (size_t)n * 4, make_float4(x[0], x[1], x[2], x[3]));
        }
        {
            const int c0 = hf * (H / 2);                           // this warp's 16 hidden columns
            float hi[16], lo[16], hv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v = sBin[c0 + i];
                v = fmaf(x[0], sWin[0 * H + c0 + i], v);
                v = fmaf(x[1], sWin[1 * H + c0 + i], v);
                v = fmaf(x[2], sWin[2 * H + c0 + i], v);
                v = fmaf(x[3], sWin[3 * H + c0 + i], v);
                hv[i] = tanh_node(v);
                split3(hv[i], hi[i], lo[i]);
            }
            if (H_save && n < n_nodes) {                                           // training: H0 kept for the backward pass
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st4(H_save + (size_t)n * H + c0 + 4 * i, make_float4(hv[4 * i], hv[4 * i + 1], hv[4 * i + 2], hv[4 * i + 3]));
            }
            tmem_st16(lane_base + C::C_A3H + c0, hi);
            tmem_st16(lane_base + C::C_A3L + c0, lo);
        }
        if (hf == 0) {
            float xh[8], xl[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) split3(x[i], xh[i], xl[i]);
#pragma unroll
            for (int i = 4; i < 8; ++i) xh[i] = xl[i] = 0.f;
            xh[4] = 1.f;                                          // times the bias row of the weight image
            tmem_st8(lane_base + C::C_A3H + H, xh);
            tmem_st8(lane_base + C::C_A3L + H, xl);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kq = 0; kq < N::D4P / 8; ++kq) {
                const uint32_t ko = kq * 2 * N::LBO;
                const uint64_t bh = smem_desc(sa + C::O_WPH + ko, N::LBO, N::SBO_D4);
                const uint64_t bl = smem_desc(sa + C::O_WPL + ko, N::LBO, N::SBO_D4);
                umma_ts(tmem + C::C_D3, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                umma_ts(tmem + C::C_D3, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
            }
            umma_commit(mb);
        }
        mbar_wait(mb, phase); phase ^= 1;
        tc_fence_after();
        tc_store_projections<H>(lane_base, C::C_D3, reinterpret_cast<float*>(smem + C::O_OUT + warp * N::OUT_BYTES),
                                tile * TM + q * 32, n_nodes, lane, out, hf, 2);
        tc_fence_before();
        __syncthreads();       // A3 / D3 are rewritten by the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

int launch_input_tc32_ex(const float* blob, const float* X, int n_nodes, int F, float* X4, const ProjOut& out, float* H_save,
                         cudaStream_t st);
int launch_input_tc32(const float* blob, const float* X, int n_nodes, int F, float* X4, float* P, float* Q,
                      float* H_save, cudaStream_t st) {
    return launch_input_tc32_ex(blob, X, n_nodes, F, X4, ProjOut{P, Q, 0, 160, nullptr}, H_save, st);
}
int launch_input_tc32_ex(const float* blob, const float* X, int n_nodes, int F, float* X4, const ProjOut& out, float* H_save,
                         cudaStream_t st) {
    using C = TcInCfg<32>;
    if (n_nodes == 0) return GNNSEG_OK;
    if (mlp_pipe()) return launch_input_pipe(blob, X, n_nodes, F, 32, X4, out, H_save, st);
    const int n_tiles = (n_nodes + TcCfg<32>::TM - 1) / TcCfg<32>::TM;
    if (!ensure_dynamic_smem<input_kernel_tc<32>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int cap = 2 * sms;     // 256 TMEM columns and ~95 KB of shared memory per CTA: two CTAs per SM
    const int grid = n_tiles < cap ? n_tiles : cap;
    input_kernel_tc<32><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, X, F, n_nodes, n_tiles, X4, out, H_save);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// fp32 matrix [rows][cols] (row major) seen as boxes of box_cols x 32 rows
static bool make_store_map(CUtensorMap* tm, float* base, int rows, int cols, int box_cols, bool swizzle128) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// `out` says where the projections go
int launch_node_mlp_tc32_ex(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, const ProjOut& out,
                            float* H_save, bool pdl, cudaStream_t st) {
    using C = TcMlpCfg<32>;
    if (n_nodes == 0) return GNNSEG_OK;
    if (mlp_pipe()) return launch_node_mlp_pipe(blob, X4, h1, ld_h1, n_nodes, 32, out, H_save, pdl, st);
    const int n_tiles = (n_nodes + C::TM - 1) / C::TM;
    constexpr int SMEM = C::SMEM_BYTES + 1024;                      // room to align the store tiles to 1024 bytes
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;        // two CTAs per SM
    static const bool tma_env = [] { const char* v = std::getenv("GNNSEG_MLP_STORE"); return v && v[0] == 't'; }();
    const bool tma = tma_env && out.mode == 0;
    CUtensorMap tmP, tmQ, tmQ16;
    memset(&tmP, 0, sizeof(tmP)); memset(&tmQ, 0, sizeof(tmQ)); memset(&tmQ16, 0, sizeof(tmQ16));
    if (tma) {
        const bool write_q = out.n_cols > 64;
        if (!make_store_map(&tmP, out.p, n_nodes, 64, 32, true)) return GNNSEG_ECUDA;
        if (write_q && (!make_store_map(&tmQ, out.q, n_nodes, 96, 32, true) || !make_store_map(&tmQ16, out.q, n_nodes, 96, 16, false)))
            return GNNSEG_ECUDA;
        if (!ensure_dynamic_smem<node_mlp_kernel_tc<32, true>>(SMEM)) return GNNSEG_ECUDA;
        if (launch_pdl(node_mlp_kernel_tc<32, true>, grid, C::NT, SMEM, st, pdl, blob, X4, h1, ld_h1, n_nodes, n_tiles, out,
                       H_save, tmP, tmQ, tmQ16) != cudaSuccess)
            return GNNSEG_ECUDA;
        return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
    }
    if (!ensure_dynamic_smem<node_mlp_kernel_tc<32, false>>(SMEM)) return GNNSEG_ECUDA;
    if (launch_pdl(node_mlp_kernel_tc<32, false>, grid, C::NT, SMEM, st, pdl, blob, X4, h1, ld_h1, n_nodes, n_tiles, out,
                   H_save, tmP, tmQ, tmQ16) != cudaSuccess)
        return GNNSEG_ECUDA;
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}
int launch_node_mlp_tc32(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, float* P_out,
                         float* Q_out, int write_q, float* H_save, bool pdl, cudaStream_t st) {
    return launch_node_mlp_tc32_ex(blob, X4, h1, ld_h1, n_nodes, ProjOut{P_out, Q_out, 0, write_q ? 160 : 64, nullptr}, H_save, pdl, st);
}

#ifdef GNNSEG_TRACE
extern "C" int gnnseg_debug_read_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 2 * 16 * 12) == cudaSuccess ? 0 : -4;
}
extern "C" int gnnseg_debug_read_cta(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_cta, sizeof(unsigned long long) * 512 * 2) == cudaSuccess ? 0 : -4;
}
#endif

int launch_node_mlp_tc64_ex(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, const ProjOut& out,
                            float* H_save, bool pdl, cudaStream_t st) {
    using C = Mlp64;
    if (n_nodes == 0) return GNNSEG_OK;
    if (mlp_pipe()) return launch_node_mlp_pipe(blob, X4, h1, ld_h1, n_nodes, 64, out, H_save, pdl, st);
    const int n_tiles = (n_nodes + C::TM - 1) / C::TM;
    if (!ensure_dynamic_smem<node_mlp_kernel_tc64<false>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int grid = n_tiles < sms ? n_tiles : sms;
    if (launch_pdl(node_mlp_kernel_tc64<false>, grid, C::NT, C::SMEM_BYTES, st, pdl, blob, const_cast<float*>(X4), h1, ld_h1, n_nodes,
                   n_tiles, out, H_save, (const float*)nullptr, 0) != cudaSuccess)
        return GNNSEG_ECUDA;
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}
int launch_node_mlp_tc64(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, float* P_out,
                         float* Q_out, int write_q, float* H_save, bool pdl, cudaStream_t st) {
    return launch_node_mlp_tc64_ex(blob, X4, h1, ld_h1, n_nodes, ProjOut{P_out, Q_out, 0, write_q ? 320 : 128, nullptr}, H_save, pdl, st);
}

int launch_input_tc64_ex(const float* blob, const float* X, int n_nodes, int F, float* X4, const ProjOut& out, float* H_save,
                         cudaStream_t st);
int launch_input_tc64(const float* blob, const float* X, int n_nodes, int F, float* X4, float* P, float* Q, float* H_save,
                      cudaStream_t st) {
    return launch_input_tc64_ex(blob, X, n_nodes, F, X4, ProjOut{P, Q, 0, 320, nullptr}, H_save, st);
}
int launch_input_tc64_ex(const float* blob, const float* X, int n_nodes, int F, float* X4, const ProjOut& out, float* H_save,
                         cudaStream_t st) {
    using C = Mlp64;
    if (n_nodes == 0) return GNNSEG_OK;
    if (mlp_pipe()) return launch_input_pipe(blob, X, n_nodes, F, 64, X4, out, H_save, st);
    const int n_tiles = (n_nodes + C::TM - 1) / C::TM;
    if (!ensure_dynamic_smem<node_mlp_kernel_tc64<true>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int grid = n_tiles < sms ? n_tiles : sms;
    // the first kernel of a forward: launched fully serialised (it reads the blob the pack kernels wrote)
    node_mlp_kernel_tc64<true><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, X4, nullptr, 0, n_nodes, n_tiles, out, H_save, X, F);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int launch_node_tc32(const float* blob, const GnnsegGraph* g, const float* X4, const float* Q_in, const float* e_in,
                     const float* e_out, float* P_out, float* Q_out, int write_q, float* h1_save, float* H_save,
                     cudaStream_t st) {
    using C = TcCfg<32>;
    if (g->n_nodes == 0) return GNNSEG_OK;
    const int n_tiles = (g->n_nodes + C::TM - 1) / C::TM;
    if (!ensure_dynamic_smem<node_kernel_tc<32>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int grid = n_tiles < sms ? n_tiles : sms;
    if (launch_pdl(node_kernel_tc<32>, grid, C::NT, C::SMEM_BYTES, st, use_pdl(g->n_slots), blob, *g, X4, Q_in, e_in, e_out, n_tiles, P_out,
                   Q_out, write_q, h1_save, H_save) != cudaSuccess)
        return GNNSEG_ECUDA;
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
