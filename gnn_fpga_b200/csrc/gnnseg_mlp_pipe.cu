// The per-node MLP of the node step (and the input step) as a warp-specialised pipeline, hidden_dim H = 32 and 64.
//
// Reference: NodeNetwork.forward's two Linear + Tanh layers after the segment sums, gnn/model.py:120-125, and
// the InputNet + first projections, gnn/model.py:144-146, in the projection-first form of gnnseg_common.cuh:
//     H'      = tanh(h1 . W4^T + b4)                                   128 x H x H           (GEMM2)
//     [P'|Q'] = [H'|X|1] . [W1a|W1b|W3a|W3b|W3c|bias]^T                128 x (H + 8) x 5H    (GEMM3, chunks of 160 columns)
// both as tcgen05.mma kind::tf32 with the 3xTF32 split (hi.hi + hi.lo + lo.hi, fp32 accumulators in tensor memory).
//
// Why: the round-1 kernels (gnnseg_node_tc.cu) run one tile's chain in sequence (GEMM2 -> epilogue -> GEMM3 -> stores;
// at H = 64: 24 k cycles per tile of which 10 k are the stores -- 1280 bytes per node leave the SM -- and 5 k the
// tensor pipe).  Here every stage has its own warps and its own mbarriers, so the chain of tile t + 1 runs under the
// stores of tile t, and the kernel sits at the SM's store rate:
//
//   loader warps (8)   h1 rows of the NEXT tile (registers) -> tf32 hi / lo -> canonical K-major A2 tile in shared memory
//   MMA warp (1)       one thread issues GEMM2 (SS), then per chunk GEMM3 (TS: A = [H'] hi / lo in tensor memory,
//                      plus one SS k-step for the [X|1] slab in shared memory); tcgen05.commit -> mbarriers
//   epilogue warps (8) D2 -> +b4, tanh, split -> A3 in tensor memory; X slab -> shared memory
//   store warps (8)    D3 chunk -> (2^(log2e v) for the edge projections) -> swizzled 32 x 32 tile -> full lines
//   weight warp (1)    H = 64: the projection weight image (184 KB hi + lo) does not fit next to A2 and W4: its two halves
//                      go through one 92 KB buffer, cp.async.bulk + mbarrier complete_tx, half c + 1 as soon as
//                      GEMM3 of chunk c has read half c.  H = 32 (one chunk, 51 KB): loaded once
//
// Tensor memory (512 columns allocated): A3 hi [0, H), A3 lo [H, 2H), D2 [2H, 3H), D3 slot 0 [3H, 3H + 160), slot 1
// [3H + 160, 3H + 320).  Chunk g of the launch (two per tile at H = 64 when Q' is wanted, else one) goes to slot g & 1.
// Every CTA takes ONE contiguous range of nodes (the stores bound the kernel, so the work is balanced by rows,
// not by whole tiles: 100 000 nodes on 148 SMs are 5.28 tiles each, not 6 for some and 5 for others).
// Launched as a programmatic dependent of the fused gather (gnnseg_forward.cu: use_pdl_mlp).
#include <cstdlib>
#include "gnnseg_tc.cuh"

namespace gnnseg {

// `make trace` (-DGNNSEG_TRACE): clock64 stamps of one lane per role of CTA 0 (scripts/pipe_trace.py reads them)
#ifdef GNNSEG_TRACE
__device__ long long g_ptrace[5][16][16];
__device__ unsigned long long g_pcta[160][2];
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PT_CTA(k) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_pcta[blockIdx.x][(k)] = gtimer_ns(); } while (0)
#define PT(role, t, k) do { if (blockIdx.x == 0 && lane == 0 && (t) < 16) g_ptrace[(role)][(t)][(k)] = clock64(); } while (0)
#else
#define PT(role, t, k) do { } while (0)
#define PT_CTA(k) do { } while (0)
#endif

template <int HD>
struct Pipe {
    static constexpr int H = HD, TM = 128, D4P = (H + 4 + 7) / 8 * 8, NP = 5 * H, NH = 160;      // NH: output columns per chunk
    static constexpr int MAXCH = NP / NH;                         // chunks per tile: 2 at hidden_dim 64 (weights streamed), 1 at 32
    static_assert(NP % NH == 0 && MAXCH <= 2, "chunking");
    static constexpr int W_EPI = 8, W_ST = 8, W_LD = 8;
    static constexpr int WARP_EPI0 = 0, WARP_ST0 = W_EPI, WARP_LD0 = WARP_ST0 + W_ST, WARP_MMA = WARP_LD0 + W_LD,
                         WARP_WP = WARP_MMA + 1, NT = (WARP_WP + 1) * 32;
    static_assert(WARP_EPI0 % 4 == 0 && WARP_ST0 % 4 == 0, "a warp reaches the TMEM lanes 32 (warp % 4) ..");
    static constexpr int LBO = 128, SBO_H = (H / 4) * LBO, SBO_D4 = (D4P / 4) * LBO, SBO_X = 2 * LBO;
    static constexpr int A_BYTES = (TM / 8) * SBO_H;              // 32 KB at hidden_dim 64, one of hi / lo
    static constexpr int W4_BYTES = (H / 8) * SBO_H;              // 16 KB
    static constexpr int WPH_BYTES = (NH / 8) * SBO_D4;           // 45 KB: the weights of one chunk, one of hi / lo
    static constexpr int XS_BYTES = (TM / 8) * SBO_X;             // 4 KB: the [X|1|0] k-step of A3, one of hi / lo
    static constexpr int O_A = 0;                                 // hi, lo  (INPUT: Win^T [4][H], bin [H] live here)
    static constexpr int O_W4 = O_A + 2 * A_BYTES;                // hi, lo
    static constexpr int O_WP = O_W4 + 2 * W4_BYTES;              // [hi half][lo half] of the half in flight
    static constexpr int O_ST = O_WP + 2 * WPH_BYTES;             // [W_ST] swizzled 32 x 32 store tiles
    static constexpr int O_XS = O_ST + W_ST * 4096;               // hi, lo
    static constexpr int O_BIAS = O_XS + 2 * XS_BYTES;            // b4 [H]
    static constexpr int O_BAR = O_BIAS + H * 4;
    enum { A2_FULL, A2_EMPTY, D2_FULL, A3_FULL, A3_EMPTY, D3_FULL0, D3_FULL1, D3_EMPTY0, D3_EMPTY1, WP_FULL, WP_EMPTY, N_BAR };
    static constexpr int SMEM_BYTES = O_BAR + N_BAR * 8 + 8;
    static constexpr int C_A3H = 0, C_A3L = H, C_D2 = 2 * H, C_D3 = 3 * H, TMEM_COLS = 512;
    static_assert(C_D3 + 2 * NH <= TMEM_COLS && SMEM_BYTES <= 227 * 1024, "resources");
};

__device__ __forceinline__ void mbar_arrive(const uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const uint32_t mbar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (the copy engine of the TMA unit, no tensor map), completion counted in bytes on `mbar`
__device__ __forceinline__ void bulk_copy_g2s(const uint32_t dst_smem, const void* src, const uint32_t bytes, const uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// INPUT = true is the input step: no GEMM2; the epilogue warps compute H0 = tanh(Win.X + bin) (K = F <= 4) straight
// into A3, X4 is then an output (X zero padded), Xraw the (n, F) input.
template <int HD, bool INPUT>
__global__ void __launch_bounds__(Pipe<HD>::NT, 1)
node_mlp_kernel_pipe(const float* __restrict__ blob, float* __restrict__ X4, const float* h1, const int ld_h1,
                       const int n_nodes, const int rows_per_cta, const ProjOut out,
                       float* __restrict__ H_save, const float* __restrict__ Xraw, const int F) {
    using C = Pipe<HD>;
    using B = Blob<HD>;
    constexpr int H = C::H, TM = C::TM, NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem[];
    float* sB4 = reinterpret_cast<float*>(smem + C::O_BIAS);
    float* sWin = reinterpret_cast<float*>(smem + C::O_A);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::O_BAR + C::N_BAR * 8);
    const uint32_t sa = smem_u32(smem);
    auto bar = [&](const int i) { return sa + C::O_BAR + 8 * i; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    PT_CTA(0);

    // this CTA's nodes, in tiles of 128 (the last one partial); chunks of 160 output columns per tile
    const int r0 = min((int)blockIdx.x * rows_per_cta, n_nodes), r1 = min(r0 + rows_per_cta, n_nodes);
    const int n_tiles = (r1 - r0 + TM - 1) / TM;
    const int CH = (C::MAXCH == 2 && out.n_cols > C::NH) ? 2 : 1;

    if (INPUT) {
        for (int i = tid; i < 5 * H; i += NT) sWin[i] = __ldg(blob + B::WIN + i);      // Win and bin are contiguous
    } else {
        for (int i = tid * 4; i < 2 * H * H; i += NT * 4) cp_async16(smem + C::O_W4 + i * 4, blob + B::TC_W4H + i);   // hi, lo
        for (int i = tid; i < H; i += NT) sB4[i] = __ldg(blob + B::B4 + i);
    }
    if (tid == 0) {
        mbar_init(bar(C::A2_FULL), C::W_LD * 32);
        mbar_init(bar(C::A2_EMPTY), 1);
        mbar_init(bar(C::D2_FULL), 1);
        mbar_init(bar(C::A3_FULL), C::W_EPI * 32);
        mbar_init(bar(C::A3_EMPTY), 1);
        mbar_init(bar(C::D3_FULL0), 1);
        mbar_init(bar(C::D3_FULL1), 1);
        mbar_init(bar(C::D3_EMPTY0), C::W_ST * 32);
        mbar_init(bar(C::D3_EMPTY1), C::W_ST * 32);
        mbar_init(bar(C::WP_FULL), 1);
        mbar_init(bar(C::WP_EMPTY), 1);
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
    // Launched as a programmatic dependent of the gather kernel: only the loader warps read what that kernel writes (h1)
    // and wait for it below; every other role hangs on the loader through the barriers, and the weight stream and the
    // X4 loads start under the gather's tail.

    if (warp == C::WARP_WP) {
        // ================================ weight warp =====================================
        if (lane == 0 && n_tiles > 0) {
            constexpr int HALF_FLOATS = C::WPH_BYTES / 4, IMG_FLOATS = C::NP * C::D4P;
            const int n_loads = CH == 2 ? 2 * n_tiles : 1;        // one chunk per tile: half 0 stays
            for (int l = 0; l < n_loads; ++l) {
                PT(4, l, 0);
                if (l > 0) mbar_wait(bar(C::WP_EMPTY), (l & 1) ^ 1);              // GEMM3 of chunk l - 1 has read the buffer
                PT(4, l, 1);
                const int half = l & 1;
                mbar_arrive_expect_tx(bar(C::WP_FULL), 2 * C::WPH_BYTES);
                bulk_copy_g2s(sa + C::O_WP, blob + B::TC_WPH + half * HALF_FLOATS, C::WPH_BYTES, bar(C::WP_FULL));
                bulk_copy_g2s(sa + C::O_WP + C::WPH_BYTES, blob + B::TC_WPH + IMG_FLOATS + half * HALF_FLOATS, C::WPH_BYTES,
                              bar(C::WP_FULL));
            }
        }
        __syncwarp();
    } else if (warp == C::WARP_MMA) {
        // ================================ MMA warp =========================================
        // Issue order: GEMM3 chunk 0 of tile t, GEMM2 of tile t + 1, GEMM3 chunk 1 of tile t.  The epilogue warps then turn
        // D2 of tile t + 1 into H' (registers) while the tensor pipe works on chunk 1, and only their write of A3 waits for it.
        constexpr uint32_t ID2 = idesc_tf32(TM, H), ID3 = idesc_tf32(TM, C::NH);
        const int total = CH * n_tiles;
        auto gemm2 = [&](const int it) {                          // D2 = h1 . W4^T of tile `it`
            mbar_wait(bar(C::A2_FULL), it & 1);
            PT(0, it, 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = sa + C::O_A, a_lo = a_hi + C::A_BYTES;
#pragma unroll
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t ah = smem_desc(a_hi + ko, C::LBO, C::SBO_H);
                    const uint64_t al = smem_desc(a_lo + ko, C::LBO, C::SBO_H);
                    const uint64_t bh = smem_desc(sa + C::O_W4 + ko, C::LBO, C::SBO_H);
                    const uint64_t bl = smem_desc(sa + C::O_W4 + C::W4_BYTES + ko, C::LBO, C::SBO_H);
                    umma_ss(tmem + C::C_D2, al, bh, ID2, kq > 0);
                    umma_ss(tmem + C::C_D2, ah, bl, ID2, 1);
                    umma_ss(tmem + C::C_D2, ah, bh, ID2, 1);
                }
                if (it + 1 < n_tiles) umma_commit(bar(C::A2_EMPTY));     // the loader may write the next tile
                umma_commit(bar(C::D2_FULL));
            }
            __syncwarp();
            PT(0, it, 2);
        };
        int g = 0;
        auto gemm3 = [&](const int it, const int c) {             // D3[slot] = [H'|X|1] . WP_half^T
            const bool more = it + 1 < n_tiles;
            const int slot = g & 1;
            mbar_wait(bar(C::D3_EMPTY0 + slot), ((g >> 1) & 1) ^ 1);              // the store warps have read this slot
            PT(0, it, 4 + 4 * c);
            if (CH == 2) mbar_wait(bar(C::WP_FULL), g & 1);
            else if (g == 0) mbar_wait(bar(C::WP_FULL), 0);
            PT(0, it, 5 + 4 * c);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d = tmem + C::C_D3 + slot * C::NH;
#pragma unroll
                for (int kq = 0; kq < H / 8; ++kq) {
                    const uint32_t ko = kq * 2 * C::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_WP + ko, C::LBO, C::SBO_D4);
                    const uint64_t bl = smem_desc(sa + C::O_WP + C::WPH_BYTES + ko, C::LBO, C::SBO_D4);
                    umma_ts(d, tmem + C::C_A3L + 8 * kq, bh, ID3, kq > 0);
                    umma_ts(d, tmem + C::C_A3H + 8 * kq, bl, ID3, 1);
                    umma_ts(d, tmem + C::C_A3H + 8 * kq, bh, ID3, 1);
                }
                {                                                 // the [X|1|0] k-step: A from shared memory
                    const uint32_t ko = (H / 8) * 2 * C::LBO;
                    const uint64_t bh = smem_desc(sa + C::O_WP + ko, C::LBO, C::SBO_D4);
                    const uint64_t bl = smem_desc(sa + C::O_WP + C::WPH_BYTES + ko, C::LBO, C::SBO_D4);
                    const uint64_t xh = smem_desc(sa + C::O_XS, C::LBO, C::SBO_X);
                    const uint64_t xl = smem_desc(sa + C::O_XS + C::XS_BYTES, C::LBO, C::SBO_X);
                    umma_ss(d, xl, bh, ID3, 1);
                    umma_ss(d, xh, bl, ID3, 1);
                    umma_ss(d, xh, bh, ID3, 1);
                }
                umma_commit(bar(C::D3_FULL0 + slot));
                if (CH == 2 && g + 1 < total) umma_commit(bar(C::WP_EMPTY));
                if (c + 1 == CH && more) umma_commit(bar(C::A3_EMPTY));
            }
            __syncwarp();
            PT(0, it, 6 + 4 * c);
            ++g;
        };
        if (!INPUT && n_tiles > 0) gemm2(0);
        for (int it = 0; it < n_tiles; ++it) {
            PT(0, it, 0);
            mbar_wait(bar(C::A3_FULL), it & 1);                    // [H'|X|1] of this tile is in place, D2 has been read
            PT(0, it, 3);
            if (CH == 1) {                                        // (no weight streaming to hide: GEMM2 of the next tile goes first)
                if (!INPUT && it + 1 < n_tiles) gemm2(it + 1);
                gemm3(it, 0);
            } else {
                gemm3(it, 0);
                if (!INPUT && it + 1 < n_tiles) gemm2(it + 1);
                gemm3(it, 1);
            }
        }
    } else if (warp >= C::WARP_LD0) {
        // ================================ loader warps =====================================
        // a quarter warp holds 8 consecutive rows of ONE float4 chunk, i.e. one contiguous 128-byte row group of
        // the canonical tile: the shared-memory stores are bank-conflict free
        // The rows of the NEXT tile are fetched into registers while this one is in flight (the loads queue behind the
        // store warps' traffic: 5 - 11 k cycles measured); only the split and the shared-memory stores wait for GEMM2.
        if (!INPUT) {
            const int lt = tid - C::WARP_LD0 * 32, lw = lt >> 5, r7 = lt & 7, cq = (lt >> 3) & 3;
            constexpr int CPQ = H / 16, J = 2 * CPQ;               // float4 chunk cq + 4 (j % CPQ) of row group 2 lw + j / CPQ
            static_assert(C::W_LD == 8, "row groups");
            float4 v[J];
            auto fetch = [&](const int it) {
                const int base = r0 + it * TM;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int n = base + (lw * 2 + j / CPQ) * 8 + r7;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n < r1) v[j] = lds4(h1 + (size_t)n * ld_h1 + 4 * (cq + 4 * (j % CPQ)));      // plain load: h1 may alias the output rows
                }
            };
            pdl_wait();                                            // h1 comes from the kernel before
            if (n_tiles > 0) fetch(0);
            for (int it = 0; it < n_tiles; ++it) {
                if (warp == C::WARP_LD0) PT(3, it, 0);
                if (it > 0) mbar_wait(bar(C::A2_EMPTY), (it & 1) ^ 1);            // GEMM2 of the previous tile has read A2
                if (warp == C::WARP_LD0) PT(3, it, 1);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float4 hh, hl;
                    split3(v[j].x, hh.x, hl.x); split3(v[j].y, hh.y, hl.y); split3(v[j].z, hh.z, hl.z); split3(v[j].w, hh.w, hl.w);
                    const int off = canon_off((lw * 2 + j / CPQ) * 8 + r7, 4 * (cq + 4 * (j % CPQ)), C::SBO_H);
                    *reinterpret_cast<float4*>(smem + C::O_A + off) = hh;
                    *reinterpret_cast<float4*>(smem + C::O_A + C::A_BYTES + off) = hl;
                }
                fence_async_smem();                                // generic-proxy writes -> tensor core reads
                mbar_arrive(bar(C::A2_FULL));
                if (warp == C::WARP_LD0) PT(3, it, 2);
                if (it + 1 < n_tiles) fetch(it + 1);
            }
        }
    } else if (warp >= C::WARP_ST0) {
        // ================================ store warps ======================================
        const int w = warp - C::WARP_ST0, q = w & 3, hf = w >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float* sOut = reinterpret_cast<float*>(smem + C::O_ST + w * 4096);
        const uint64_t stream = l2_policy_evict_first();
        int g = 0;
        for (int it = 0; it < n_tiles; ++it) {
            const int node_w0 = r0 + it * TM + q * 32;
            for (int c = 0; c < CH; ++c, ++g) {
                const int slot = g & 1;
                if (w == 0) PT(2, it, 3 * c);
                mbar_wait(bar(C::D3_FULL0 + slot), (g >> 1) & 1);
                if (w == 0) PT(2, it, 3 * c + 1);
                tc_fence_after();
                // the two warps of a lane quarter share the five 32-column pieces of a chunk, 3 + 2 alternating
                for (int j = (hf + g) & 1; j < 5; j += 2) {
                    const int c0 = C::NH * c + 32 * j;
                    if (c0 >= out.n_cols) break;
                    {
                        float v[32];
                        tmem_ld32(lane_base + C::C_D3 + slot * C::NH + 32 * j, v);
                        if (proj_is_exp<H>(out, c0)) {
                            to_exponentials(*reinterpret_cast<float (*)[16]>(&v[0]), out.range_flag);
                            to_exponentials(*reinterpret_cast<float (*)[16]>(&v[16]), out.range_flag);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)           // the bias is already in D3
                            st4(sOut + lane * 32 + ((i ^ (lane & 7)) << 2), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                    }
                    __syncwarp();
                    store_tile_proj<H>(out, sOut, node_w0, c0, r1 - node_w0, lane, stream);
                    __syncwarp();
                }
                tc_fence_before();
                mbar_arrive(bar(C::D3_EMPTY0 + slot));
                if (w == 0) PT(2, it, 3 * c + 2);
            }
        }
    } else {
        // ================================ epilogue warps ===================================
        const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        for (int it = 0; it < n_tiles; ++it) {
            const int n = r0 + it * TM + row;
            const bool live = n < r1;
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
            if (warp == 0) PT(1, it, 0);
            constexpr int CPT = H / 2;                             // this thread's columns of H': 32 or 16
            float v[CPT];
            const int cb = hf * CPT;
            if (!INPUT) {
                mbar_wait(bar(C::D2_FULL), it & 1);
                if (warp == 0) PT(1, it, 1);
                tc_fence_after();
                if constexpr (CPT == 32) {
                    tmem_ld32(lane_base + C::C_D2 + cb, v);
                } else {
                    tmem_ld16(lane_base + C::C_D2 + cb, v);
                }
#pragma unroll
                for (int i = 0; i < CPT; ++i) v[i] = tanh_node(v[i] + sB4[cb + i]);
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    float t = sWin[4 * H + cb + i];
                    t = fmaf(x.x, sWin[0 * H + cb + i], t);
                    t = fmaf(x.y, sWin[1 * H + cb + i], t);
                    t = fmaf(x.z, sWin[2 * H + cb + i], t);
                    t = fmaf(x.w, sWin[3 * H + cb + i], t);
                    v[i] = tanh_node(t);
                }
            }
            if (H_save && live) {
#pragma unroll
                for (int i = 0; i < CPT / 4; ++i)
                    st4(H_save + (size_t)n * H + cb + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
            // H' is complete in registers, its first 16 columns already split, BEFORE the wait (the compiler would otherwise
            // sink the arithmetic below it)
            float hi0[16], lo0[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                split3(v[i], hi0[i], lo0[i]);
                asm volatile("" : "+f"(hi0[i]), "+f"(lo0[i]));
                if constexpr (CPT == 32) asm volatile("" : "+f"(v[16 + i]));
            }
            if (warp == 0) PT(1, it, 4);
            mbar_wait(bar(C::A3_EMPTY), (it & 1) ^ 1);             // GEMM3 of the previous tile has read A3 and the X slab
            if (warp == 0) PT(1, it, 2);
            tc_fence_after();
            tmem_st16(lane_base + C::C_A3H + cb, hi0);
            tmem_st16(lane_base + C::C_A3L + cb, lo0);
            if constexpr (CPT == 32) {
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split3(v[16 + i], hi[i], lo[i]);
                tmem_st16(lane_base + C::C_A3H + cb + 16, hi);
                tmem_st16(lane_base + C::C_A3L + cb + 16, lo);
            }
            if (hf == 0) {                                         // [X | 1 | 0 0 0]: times the X rows and the bias row of the weight image
                float4 xh, xl;
                split3(x.x, xh.x, xl.x); split3(x.y, xh.y, xl.y); split3(x.z, xh.z, xl.z); split3(x.w, xh.w, xl.w);
                const int off = canon_off(row, 0, C::SBO_X);
                *reinterpret_cast<float4*>(smem + C::O_XS + off) = xh;
                *reinterpret_cast<float4*>(smem + C::O_XS + off + 128) = make_float4(1.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(smem + C::O_XS + C::XS_BYTES + off) = xl;
                *reinterpret_cast<float4*>(smem + C::O_XS + C::XS_BYTES + off + 128) = make_float4(0.f, 0.f, 0.f, 0.f);
                fence_async_smem();
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar(C::A3_FULL));
            if (warp == 0) PT(1, it, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    PT_CTA(1);
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

#ifdef GNNSEG_TRACE
extern "C" int gnnseg_debug_read_pipe_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_ptrace, sizeof(long long) * 5 * 16 * 16) == cudaSuccess ? 0 : -4;
}
extern "C" int gnnseg_debug_read_pipe_cta(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_pcta, sizeof(unsigned long long) * 160 * 2) == cudaSuccess ? 0 : -4;
}
#endif

// one contiguous node range per CTA, one CTA per SM
static int pipe_grid(const int n_nodes, const int sms, int& rows_per_cta) {
    const int n_tiles = (n_nodes + 127) / 128;
    const int grid = n_tiles < sms ? n_tiles : sms;
    rows_per_cta = (n_nodes + grid - 1) / grid;
    return grid;
}

template <int H>
static int launch_mlp(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, const ProjOut& out,
                      float* H_save, bool pdl, cudaStream_t st) {
    using C = Pipe<H>;
    if (n_nodes == 0) return GNNSEG_OK;
    if (!ensure_dynamic_smem<node_mlp_kernel_pipe<H, false>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    int rows = 0;
    const int grid = pipe_grid(n_nodes, sms, rows);
    if (launch_pdl(node_mlp_kernel_pipe<H, false>, grid, C::NT, C::SMEM_BYTES, st, pdl, blob, const_cast<float*>(X4), h1, ld_h1, n_nodes,
                   rows, out, H_save, (const float*)nullptr, 0) != cudaSuccess)
        return GNNSEG_ECUDA;
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

template <int H>
static int launch_input(const float* blob, const float* X, int n_nodes, int F, float* X4, const ProjOut& out, float* H_save,
                        cudaStream_t st) {
    using C = Pipe<H>;
    if (n_nodes == 0) return GNNSEG_OK;
    if (!ensure_dynamic_smem<node_mlp_kernel_pipe<H, true>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    int rows = 0;
    const int grid = pipe_grid(n_nodes, sms, rows);
    // the first kernel of a forward: launched fully serialised (it reads the blob the pack kernels wrote)
    node_mlp_kernel_pipe<H, true><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, X4, nullptr, 0, n_nodes, rows, out, H_save, X, F);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int launch_node_mlp_pipe(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, int h, const ProjOut& out,
                         float* H_save, bool pdl, cudaStream_t st) {
    return h == 32 ? launch_mlp<32>(blob, X4, h1, ld_h1, n_nodes, out, H_save, pdl, st)
                   : launch_mlp<64>(blob, X4, h1, ld_h1, n_nodes, out, H_save, pdl, st);
}
int launch_input_pipe(const float* blob, const float* X, int n_nodes, int F, int h, float* X4, const ProjOut& out, float* H_save,
                      cudaStream_t st) {
    return h == 32 ? launch_input<32>(blob, X, n_nodes, F, X4, out, H_save, st)
                   : launch_input<64>(blob, X, n_nodes, F, X4, out, H_save, st);
}

}  // namespace gnnseg
