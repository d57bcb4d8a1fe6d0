// Forward kernels of the segment classifier for sm_100a: weight packing, input step,
// edge step, node step.  The math follows gnn/model.py:140-156 of the reference with the
// dense incidence bmm replaced by int32 gathers and an ordered CSR segment sum:
//
//   edge step (gnn/model.py:69-81)   e_j = sigmoid(W2 . tanh(W1 . [HX[src_j]; HX[dst_j]] + b1) + b2)
//     computed "projection first": W1.[a;b] = W1a.a + W1b.b, and P[n] = [W1a.HX[n]+b1 | W1b.HX[n]]
//     is produced once per node by the kernel that wrote HX[n]; the edge kernel gathers two
//     h-wide rows of P with 16-byte loads, adds, tanh, dots with W2, sigmoid.
//   node step (gnn/model.py:113-125) mi[n] = sum_{dst_j = n} e_j HX[src_j]  (ascending j)
//                                    mo[n] = sum_{src_j = n} e_j HX[dst_j]  (ascending j)
//                                    H'[n] = tanh(W4 . tanh(W3 . [mi; mo; HX[n]] + b3) + b4)
//     persistent warp-specialised CTAs: producer warps walk the two CSR rows of a tile of
//     nodes (no atomics, fixed order) into shared memory while consumer warps run the two
//     layers + the next step's projection of the previous tile as register-tiled fp32 GEMMs
//     against weights resident in shared memory.  HX' = [H' | X] is written back (the
//     reference's cat([H, X]), gnn/model.py:146,154) together with P'.
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {

// ------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(GnnsegParams p, int F, int H, float* __restrict__ blob) {
    const int D = F + H, D4 = H + 4;
    const int o_bin = 4 * H, o_w1 = o_bin + H, o_b1 = o_w1 + D4 * 2 * H, o_w2 = o_b1 + H,
              o_b2 = o_w2 + H, o_w3 = o_b2 + 4, o_b3 = o_w3 + 3 * D4 * H, o_w4 = o_b3 + H,
              o_b4 = o_w4 + H * H, total = o_b4 + H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < o_bin) {                       // Win^T [4][H]
            const int f = i / H, j = i % H;
            if (f < F) v = p.w_in[j * F + f];
        } else if (i < o_w1) {
            v = p.b_in[i - o_bin];
        } else if (i < o_b1) {                 // W1^T [D4][2H]
            const int r = i - o_w1, k = r / (2 * H), c = r % (2 * H), part = c / H, j = c % H;
            if (k < D) {
                const int idx = j * (2 * D) + part * D + k;
                v = p.w_e1[idx];
                if (p.m_e1) v *= p.m_e1[idx];
            }
        } else if (i < o_w2) {
            v = p.b_e1[i - o_b1];
        } else if (i < o_b2) {
            const int j = i - o_w2;
            v = p.w_e2[j];
            if (p.m_e2) v *= p.m_e2[j];
        } else if (i < o_w3) {
            if (i == o_b2) v = p.b_e2[0];
        } else if (i < o_b3) {                 // W3^T [3*D4][H]
            const int r = i - o_w3, row = r / H, j = r % H, part = row / D4, k = row % D4;
            if (k < D) {
                const int idx = j * (3 * D) + part * D + k;
                v = p.w_n1[idx];
                if (p.m_n1) v *= p.m_n1[idx];
            }
        } else if (i < o_w4) {
            v = p.b_n1[i - o_b3];
        } else if (i < o_b4) {                 // W4^T [H][H]
            const int r = i - o_w4, k = r / H, j = r % H;
            const int idx = j * H + k;
            v = p.w_n2[idx];
            if (p.m_n2) v *= p.m_n2[idx];
        } else {
            v = p.b_n2[i - o_b4];
        }
        blob[i] = v;
    }
}

// ------------------------------------------------------------------------------------
// register-tiled shared-memory GEMM:  out[node][o] = sum_k A[node][k] * W[k][o]
// A: [TN][lda] fp32 in shared memory, W: [K][HOUT] fp32 in shared memory.
// A thread owns RN nodes (ng, ng+NG, ...) x RC consecutive outputs.  Lanes run over the
// output groups first, so W reads are contiguous and the (few) distinct A rows of a warp
// are consecutive rows, which tile_stride() keeps on distinct banks.  NT = threads taking
// part (they must be threads 0..NT-1 of the CTA).
// ------------------------------------------------------------------------------------
template <int K, int HOUT, int TN, int NT, int RN, int RC, typename Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ sA, const int lda,
                                          const float* __restrict__ sW, Epi epi) {
    constexpr int OG = HOUT / RC, NG = TN / RN, TILES = OG * NG;
    static_assert(K % 4 == 0 && RC % 4 == 0 && HOUT % RC == 0 && TN % RN == 0, "tile shape");
    for (int t = threadIdx.x; t < TILES; t += NT) {
        const int og = t % OG, ng = t / OG;
        float acc[RN][RC];
#pragma unroll
        for (int i = 0; i < RN; ++i)
#pragma unroll
            for (int c = 0; c < RC; ++c) acc[i][c] = 0.f;
        const float* a0 = sA + ng * lda;
        const float* w0 = sW + og * RC;
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            float4 a[RN];
#pragma unroll
            for (int i = 0; i < RN; ++i) a[i] = lds4(a0 + i * NG * lda + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int c4 = 0; c4 < RC / 4; ++c4) {
                    const float4 w = lds4(w0 + (k + kk) * HOUT + 4 * c4);
#pragma unroll
                    for (int i = 0; i < RN; ++i) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        acc[i][4 * c4 + 0] = fmaf(av, w.x, acc[i][4 * c4 + 0]);
                        acc[i][4 * c4 + 1] = fmaf(av, w.y, acc[i][4 * c4 + 1]);
                        acc[i][4 * c4 + 2] = fmaf(av, w.z, acc[i][4 * c4 + 2]);
                        acc[i][4 * c4 + 3] = fmaf(av, w.w, acc[i][4 * c4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RN; ++i)
#pragma unroll
            for (int c4 = 0; c4 < RC / 4; ++c4) epi(ng + i * NG, og * RC + 4 * c4, &acc[i][4 * c4]);
    }
}

template <int NT>
__device__ __forceinline__ void copy_to_smem(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
    for (int i = threadIdx.x * 4; i < n_floats; i += NT * 4) st4(dst + i, ldg4(src + i));
}

// Smallest stride >= k (floats) whose rows, read 8 bytes per lane by the mma fragments
// (4 rows x 4 lanes per half warp), fall on disjoint bank groups: stride mod 32 in {8, 24}.
__host__ __device__ constexpr int mma_stride(int k) {
    int s = (k + 7) / 8 * 8;
    while (s % 32 != 8 && s % 32 != 24) s += 8;
    return s;
}

// Kernel shape per hidden size.
template <int H>
struct NodeCfg {
    static constexpr int TN   = 64;                  // nodes per tile
    static constexpr int CW   = 8;                   // consumer warps (MLP)
    static constexpr int PW   = (H >= 64) ? 16 : 8;  // producer warps (CSR gather)
    static constexpr int CT   = CW * 32;
    static constexpr int PT   = PW * 32;
    static constexpr int NT   = CT + PT;             // threads per CTA
    static constexpr bool MMA = H >= 32;             // tensor-core (3xTF32) MLP; SIMT below that
    static constexpr int NBUF = MMA ? 1 : 2;         // [mi|mo|self] tile buffers
    static constexpr int MINB = (H >= 64) ? 1 : 2;   // CTAs per SM the register budget allows
    static constexpr int RN1  = 2;                   // SIMT path: nodes per thread, H outputs
    static constexpr int RNP  = 4;                   // SIMT path: nodes per thread, projection
    static constexpr int RCP  = 4;                   // SIMT path: outputs per thread, projection
    static constexpr int D4   = H + 4;
    static constexpr int K1   = 3 * D4;
    static constexpr int K1P  = (K1 + 7) / 8 * 8;    // MMA path: K padded to the k8 step
    static constexpr int D4P  = (D4 + 7) / 8 * 8;
    static constexpr int SM   = MMA ? mma_stride(K1P) : tile_stride(K1);   // [mi|mo|self] tile stride
    static constexpr int SH   = MMA ? mma_stride(H) : tile_stride(H);      // hidden-layer tile stride
    static constexpr int SD   = MMA ? mma_stride(D4P) : tile_stride(D4);   // HX tile stride
    static constexpr int CAP  = (H >= 64) ? 1536 : 768;   // staged CSR slots per direction per tile
    // weights in shared memory: SIMT keeps the blob's [k][out]; MMA wants [out][k] rows with the
    // same padded strides as the A tiles
    static constexpr int W_FLOATS = MMA ? (H * SM + H * SH + 2 * H * SD + 3 * H)
                                        : (K1 * H + H * H + D4 * 2 * H + 3 * H);
    static constexpr int STAGE_WORDS = 2 * 2 * CAP + 2 * (TN + 4);   // (nbr,w) pairs + row pointers
    static constexpr int SMEM_FLOATS = W_FLOATS + TN * (NBUF * SM + SH + SD) + STAGE_WORDS;
    static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;
};

template <int H>
struct InputCfg {
    static constexpr int TN = 64, NT = 256, RNP = 4, RCP = 4;
    static constexpr int D4 = H + 4;
    static constexpr int SD = tile_stride(D4);
    static constexpr int SMEM_FLOATS = 4 * H + H + D4 * 2 * H + H + TN * SD + TN * 4;
    static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;
};

// Write the tile's HX rows and its projection P = [W1a.HX+b1 | W1b.HX] to global memory.
// Runs on threads 0..NT-1.
template <int H, int TN, int NT, int RNP, int RCP>
__device__ __forceinline__ void store_hx_and_project(const float* __restrict__ sHX, const int sd,
                                                     const float* __restrict__ sW1,
                                                     const float* __restrict__ sB1,
                                                     const int node0, const int n_nodes,
                                                     float* __restrict__ HX_out,
                                                     float* __restrict__ P_out) {
    constexpr int D4 = H + 4, C4 = D4 / 4;
    for (int i = threadIdx.x; i < TN * C4; i += NT) {
        const int ln = i / C4, c = i % C4, n = node0 + ln;
        if (n < n_nodes) st4(HX_out + (size_t)n * D4 + 4 * c, lds4(sHX + ln * sd + 4 * c));
    }
    tile_gemm<D4, 2 * H, TN, NT, RNP, RCP>(sHX, sd, sW1, [&](int ln, int o, const float* acc) {
        const int n = node0 + ln;
        if (n < n_nodes) {
            float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
            if (o < H) {
                const float4 b = lds4(sB1 + o);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            st4(P_out + (size_t)n * 2 * H + o, v);
        }
    });
}

// ------------------------------------------------------------------------------------
// input step: H0 = tanh(Win.X + bin), HX = [H0 | X], P = projection   (gnn/model.py:144-146)
// ------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(InputCfg<H>::NT)
input_kernel(const float* __restrict__ blob, const float* __restrict__ X, const int F,
             const int n_nodes, const int n_tiles, float* __restrict__ HX_out,
             float* __restrict__ P_out) {
    using C = InputCfg<H>;
    using B = Blob<H>;
    constexpr int TN = C::TN, NT = C::NT, D4 = C::D4, SD = C::SD;
    extern __shared__ __align__(16) float smem[];
    float* sWin = smem;                 // [4][H]
    float* sBin = sWin + 4 * H;         // [H]
    float* sW1  = sBin + H;             // [D4][2H]
    float* sB1  = sW1 + D4 * 2 * H;     // [H]
    float* sHX  = sB1 + H;              // [TN][SD]
    float* sX   = sHX + TN * SD;        // [TN][4]
    copy_to_smem<NT>(sWin, blob + B::WIN, 4 * H + H);
    copy_to_smem<NT>(sW1, blob + B::W1, D4 * 2 * H + H);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int node0 = tile * TN;
        __syncthreads();   // weights visible (first pass) / previous tile's readers done
        for (int i = threadIdx.x; i < TN * 4; i += NT) {
            const int ln = i >> 2, f = i & 3, n = node0 + ln;
            sX[i] = (n < n_nodes && f < F) ? __ldg(X + (size_t)n * F + f) : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TN * (H / 4 + 1); i += NT) {
            const int ln = i / (H / 4 + 1), c = i % (H / 4 + 1);
            const float4 x = lds4(sX + ln * 4);
            float4 v;
            if (c < H / 4) {
                v = lds4(sBin + 4 * c);
                fma4(v, x.x, lds4(sWin + 0 * H + 4 * c));
                fma4(v, x.y, lds4(sWin + 1 * H + 4 * c));
                fma4(v, x.z, lds4(sWin + 2 * H + 4 * c));
                fma4(v, x.w, lds4(sWin + 3 * H + 4 * c));
                v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
            } else {
                v = x;
            }
            st4(sHX + ln * SD + 4 * c, v);
        }
        __syncthreads();
        store_hx_and_project<H, TN, NT, C::RNP, C::RCP>(sHX, SD, sW1, sB1, node0, n_nodes, HX_out, P_out);
    }
}

// ------------------------------------------------------------------------------------
// edge step
// ------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256)
edge_kernel(const float* __restrict__ blob, const float* __restrict__ P,
            const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
            const int n_slots, float* __restrict__ e_out) {
    using B = Blob<H>;
    constexpr int G = H / 4;        // lanes that share one edge (one float4 of P each)
    constexpr int EPP = 32 / G;     // edges a warp handles per pass
    const int lane = threadIdx.x & 31;
    const int c = lane % G, g = lane / G;
    const float4 w2 = ldg4(blob + B::W2 + 4 * c);
    const float4 b1 = ldg4(blob + B::B1 + 4 * c);
    const float b2 = __ldg(blob + B::B2);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;

    for (int base = warp * 32; base < n_slots; base += n_warps * 32) {
        const int j = base + lane;
        int s = -1, d = -1;
        if (j < n_slots) { s = __ldg(src + j); d = __ldg(dst + j); }
        float mine = 0.f;
#pragma unroll (G > 8 ? 8 : G)
        for (int p = 0; p < G; ++p) {
            const int k = p * EPP + g;                       // which of the warp's 32 edges
            const int ss = __shfl_sync(0xffffffffu, s, k);
            const int dd = __shfl_sync(0xffffffffu, d, k);
            float4 a = b1;                                   // absent start: W1a.0 + b1
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);      // absent end:   W1b.0
            if (ss >= 0) a = ldg4(P + (size_t)ss * (2 * H) + 4 * c);
            if (dd >= 0) b = ldg4(P + (size_t)dd * (2 * H) + H + 4 * c);
            float z = w2.x * tanhf(a.x + b.x);
            z = fmaf(w2.y, tanhf(a.y + b.y), z);
            z = fmaf(w2.z, tanhf(a.z + b.z), z);
            z = fmaf(w2.w, tanhf(a.w + b.w), z);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            const float v = __shfl_sync(0xffffffffu, z, (lane % EPP) * G);
            if (lane / EPP == p) mine = v;
        }
        if (j < n_slots) e_out[j] = 1.f / (1.f + expf(-(mine + b2)));
    }
}

// ------------------------------------------------------------------------------------
// node step: warp-specialised persistent kernel.
//   producer warps  walk the two CSR rows of every node of a tile and leave the tile's
//                   [mi | mo | self] rows in shared memory (latency bound: many loads in flight)
//   consumer warps  run layer 0, layer 2 and the projection as register-tiled fp32 GEMMs
//                   out of shared memory (FMA bound) and write HX' and P'
// The two halves meet through named barriers (full/empty per tile buffer), so the gather of
// tile t+1 overlaps the MLP of tile t on the same SM.
// ------------------------------------------------------------------------------------
// ---- 3xTF32 tensor-core GEMM on fp32 operands held in shared memory --------------------------
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi); a.b ~= lo_a.hi_b + hi_a.lo_b + hi_a.hi_b with
// fp32 accumulation, which keeps the result within a few 1e-7 of the fp32 FMA chain (the
// dropped lo.lo term is 2^-22 relative).  One warp computes MT x NT tiles of 16 nodes x 8 outputs.
// A: [rows][lda] fp32 (row = node, k contiguous); B: [cols][ldb] fp32 (row = output, k contiguous).
// Fragment slots: lane (g = lane/4, t = lane%4) supplies logical k = 8q+2t and 8q+2t+1 for the
// instruction's k-slots t and t+4 of both A and B, so each fragment is one 8-byte load.
__device__ __forceinline__ uint32_t to_tf32(const float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(const float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }

template <int KB, int MT, int NT>
__device__ __forceinline__ void warp_gemm_3xtf32(const float* __restrict__ sA, const int lda,
                                                 const float* __restrict__ sB, const int ldb,
                                                 float (&acc)[MT][NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* pa = sA + g * lda + 2 * t;
    const float* pb = sB + g * ldb + 2 * t;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = acc[mi][ni][2] = acc[mi][ni][3] = 0.f;
#pragma unroll 2
    for (int q = 0; q < KB; ++q) {
        uint32_t ahi[MT][4], alo[MT][4], bhi[NT][2], blo[NT][2];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const float2 r0 = lds2(pa + (mi * 16) * lda + 8 * q);       // row g
            const float2 r1 = lds2(pa + (mi * 16 + 8) * lda + 8 * q);   // row g + 8
            split_tf32(r0.x, ahi[mi][0], alo[mi][0]);
            split_tf32(r1.x, ahi[mi][1], alo[mi][1]);
            split_tf32(r0.y, ahi[mi][2], alo[mi][2]);
            split_tf32(r1.y, ahi[mi][3], alo[mi][3]);
        }
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            const float2 r = lds2(pb + (ni * 8) * ldb + 8 * q);
            split_tf32(r.x, bhi[ni][0], blo[ni][0]);
            split_tf32(r.y, bhi[ni][1], blo[ni][1]);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                mma_m16n8k8_tf32(acc[mi][ni], alo[mi], bhi[ni]);
                mma_m16n8k8_tf32(acc[mi][ni], ahi[mi], blo[ni]);
                mma_m16n8k8_tf32(acc[mi][ni], ahi[mi], bhi[ni]);
            }
    }
}

// Tile GEMM of the 8 consumer warps: out[TN=64][HOUT] = A[64][8*KB] . B[HOUT][8*KB]^T.  The warps
// form a WM x WN grid; epi(row, col, v0, v1) receives two adjacent outputs (col, col+1).
template <int KB, int HOUT, typename Epi>
__device__ __forceinline__ void tile_gemm_mma(const float* __restrict__ sA, const int lda,
                                              const float* __restrict__ sB, const int ldb, Epi epi) {
    constexpr int MTILES = 4, NTILES = HOUT / 8;          // 16-node x 8-output tiles
    constexpr int WN = NTILES >= 8 ? 4 : 2, WM = 8 / WN;  // warp grid
    constexpr int MT = MTILES / WM, NT = NTILES / WN;
    static_assert(MT >= 1 && NT >= 1 && MT * WM == MTILES && NT * WN == NTILES, "warp tiling");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int m0 = (warp / WN) * MT * 16, n0 = (warp % WN) * NT * 8;
    float acc[MT][NT][4];
    warp_gemm_3xtf32<KB, MT, NT>(sA + m0 * lda, lda, sB + n0 * ldb, ldb, acc);
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            const int r = m0 + mi * 16 + g, cidx = n0 + ni * 8 + 2 * t;
            epi(r, cidx, acc[mi][ni][0], acc[mi][ni][1]);
            epi(r + 8, cidx, acc[mi][ni][2], acc[mi][ni][3]);
        }
}

__device__ __forceinline__ void bar_sync(const int id, const int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(const int id, const int n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// One CSR row, G lanes per node: acc += w_s * HX[nbr_s] for the row's slots in ascending
// order.  Lane c owns float4 chunk c of the hidden part; the X chunk (4 floats) is spread over
// the lanes as scalars when the group has >= 4 lanes (lane c takes X[c & 3]), else lane 0 takes
// it as one float4.  STAGED: the (neighbour, weight) pairs of the whole tile were fetched into
// shared memory beforehand by coalesced loads, so the only global latency left in this loop is
// the row gather itself.  Slots go in batches of U: all U row loads are issued before the
// first FMA (no branch in between: an absent neighbour loads row 0 and is dropped by
// predication), the FMAs then run in ascending slot order.
template <int H>
struct XAcc {   // per-lane accumulator of the X chunk
    static constexpr bool SCALAR = (H / 4) >= 4;
    float4 v;
};

template <int H, bool STAGED>
__device__ __forceinline__ void csr_row_sum(const int2* __restrict__ pairs, const int32_t* __restrict__ eid,
                                            const int32_t* __restrict__ nbr, const float* __restrict__ e,
                                            const float* __restrict__ HX, const int beg, const int end,
                                            const int c, float4& acc_h, float4& acc_x) {
    constexpr bool SCALAR_X = XAcc<H>::SCALAR;
    constexpr int D4 = H + 4, U = SCALAR_X ? 4 : 2;
    for (int s0 = beg; s0 < end; s0 += U) {
        float w[U];
        bool ok[U];
        float4 vh[U];
        float4 vx4[SCALAR_X ? 1 : U];
        float vx1[SCALAR_X ? U : 1];
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
                w[u] = __ldg(e + __ldg(eid + s));
            }
            ok[u] = (s0 + u < end) && nb >= 0;   // nb < 0: half edge, gathers the zero row
            const float* row = HX + (size_t)max(nb, 0) * D4;
            vh[u] = ldg4(row + 4 * c);
            if (SCALAR_X) vx1[u] = __ldg(row + H + (c & 3));
            else if (c == 0) vx4[u] = ldg4(row + H);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[u]) {
                fma4(acc_h, w[u], vh[u]);
                if (SCALAR_X) acc_x.x = fmaf(w[u], vx1[u], acc_x.x);
                else if (c == 0) fma4(acc_x, w[u], vx4[u]);
            }
        }
    }
}

// Store the X-chunk accumulator of csr_row_sum into a tile row.
template <int H>
__device__ __forceinline__ void store_x_acc(float* __restrict__ dst, const int c, const float4& acc_x) {
    if (XAcc<H>::SCALAR) {
        if (c < 4) dst[c] = acc_x.x;
    } else if (c == 0) {
        st4(dst, acc_x);
    }
}

template <int H>
__global__ void __launch_bounds__(NodeCfg<H>::NT, NodeCfg<H>::MINB)
node_kernel(const float* __restrict__ blob, const GnnsegGraph g,
            const float* __restrict__ HX_in, const float* __restrict__ e, const int n_tiles,
            float* __restrict__ HX_out, float* __restrict__ P_out, const int dbg) {
    using C = NodeCfg<H>;
    using B = Blob<H>;
    constexpr int TN = C::TN, NT = C::NT, CT = C::CT, PT = C::PT, NBUF = C::NBUF;
    constexpr int D4 = C::D4, K1 = C::K1, SM = C::SM, SH = C::SH, SD = C::SD;
    constexpr bool MMA = C::MMA;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 1 + NBUF, BAR_CONS = 1 + 2 * NBUF, BAR_PROD = 2 + 2 * NBUF;
    extern __shared__ __align__(16) float smem[];
    // weights: SIMT path [k][out] as in the blob, MMA path [out][k] with padded row strides
    float* sW3 = smem;                                     // SIMT [K1][H]    | MMA [H][SM]
    float* sB3 = sW3 + (MMA ? H * SM : K1 * H);            // [H]
    float* sW4 = sB3 + H;                                  // SIMT [H][H]     | MMA [H][SH]
    float* sB4 = sW4 + (MMA ? H * SH : H * H);             // [H]
    float* sW1 = sB4 + H;                                  // SIMT [D4][2H]   | MMA [2H][SD]
    float* sB1 = sW1 + (MMA ? 2 * H * SD : D4 * 2 * H);    // [H]
    float* sM  = sB1 + H;               // [NBUF][TN][SM]   [mi | mo | self]
    float* sH1 = sM + NBUF * TN * SM;   // [TN][SH]
    float* sHX = sH1 + TN * SH;         // [TN][SD]
    int2* sPair = reinterpret_cast<int2*>(sHX + TN * SD);          // [2][CAP] (nbr, w) in / out
    int*  sPtr  = reinterpret_cast<int*>(sPair + 2 * C::CAP);      // [2][TN+4] row pointers in / out
    if (MMA) {
        // transpose the blob's [k][out] matrices into [out][k] rows; k beyond the real width is 0
        for (int i = threadIdx.x; i < H * SM; i += NT) {
            const int j = i / SM, k = i % SM;
            sW3[i] = k < K1 ? __ldg(blob + B::W3 + k * H + j) : 0.f;
        }
        for (int i = threadIdx.x; i < H * SH; i += NT) {
            const int j = i / SH, k = i % SH;
            sW4[i] = k < H ? __ldg(blob + B::W4 + k * H + j) : 0.f;
        }
        for (int i = threadIdx.x; i < 2 * H * SD; i += NT) {
            const int j = i / SD, k = i % SD;
            sW1[i] = k < D4 ? __ldg(blob + B::W1 + k * 2 * H + j) : 0.f;
        }
        for (int i = threadIdx.x; i < H; i += NT) {
            sB3[i] = __ldg(blob + B::B3 + i);
            sB4[i] = __ldg(blob + B::B4 + i);
            sB1[i] = __ldg(blob + B::B1 + i);
        }
        // zero the K padding of the A tiles once: nobody writes those columns afterwards
        for (int i = threadIdx.x; i < NBUF * TN * (SM - K1); i += NT) sM[(i / (SM - K1)) * SM + K1 + i % (SM - K1)] = 0.f;
        for (int i = threadIdx.x; i < TN * (SD - D4); i += NT) sHX[(i / (SD - D4)) * SD + D4 + i % (SD - D4)] = 0.f;
    } else {
        copy_to_smem<NT>(sW3, blob + B::W3, K1 * H + H + H * H + H);   // W3,b3,W4,b4 contiguous
        copy_to_smem<NT>(sW1, blob + B::W1, D4 * 2 * H + H);           // W1,b1 contiguous
    }
    __syncthreads();
    const int n_nodes = g.n_nodes;

    if (threadIdx.x >= CT) {
        // ================================ producers ====================================
        constexpr int G = H / 4;            // lanes per node
        constexpr int NGRP = PT / G;        // nodes gathered concurrently
        constexpr int CAP = C::CAP;
        const int pt = threadIdx.x - CT;
        const int grp = pt / G, c = pt % G;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it % NBUF;
            const int node0 = tile * TN;
            // ---- stage the tile's CSR slices: row pointers, then (neighbour, weight) pairs ----
            bar_sync(BAR_PROD, PT);                            // previous tile's readers are done
            if (pt <= TN) {
                const int n = min(node0 + pt, n_nodes);
                sPtr[pt] = __ldg(g.in_ptr + n);
                sPtr[TN + 4 + pt] = __ldg(g.out_ptr + n);
            }
            bar_sync(BAR_PROD, PT);
            const int ib = sPtr[0], ic = sPtr[TN] - ib;
            const int ob = sPtr[TN + 4], oc = sPtr[TN + 4 + TN] - ob;
            const bool staged = ic <= CAP && oc <= CAP;       // CTA-uniform
            if (staged) {
                for (int s = pt; s < ic; s += PT)
                    sPair[s] = make_int2(__ldg(g.in_nbr + ib + s), __float_as_int(__ldg(e + __ldg(g.in_eid + ib + s))));
                for (int s = pt; s < oc; s += PT)
                    sPair[CAP + s] = make_int2(__ldg(g.out_nbr + ob + s), __float_as_int(__ldg(e + __ldg(g.out_eid + ob + s))));
            }
            bar_sync(BAR_PROD, PT);
            if (it >= NBUF) bar_sync(BAR_EMPTY + buf, NT);     // consumers are done with this buffer
            float* sMb = sM + buf * TN * SM;
            for (int ln = grp; ln < TN; ln += NGRP) {
                const int n = node0 + ln;
                const bool live = n < n_nodes && !(dbg & 1);
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                float* m = sMb + ln * SM;
                const int i0 = sPtr[ln], i1 = live ? sPtr[ln + 1] : i0;
                const int o0 = sPtr[TN + 4 + ln], o1 = live ? sPtr[TN + 4 + ln + 1] : o0;
                {   // mi: sum over in-edges of e * HX[src]
                    float4 a_h = zero, a_x = zero;
                    if (staged) csr_row_sum<H, true>(sPair, nullptr, nullptr, nullptr, HX_in, i0 - ib, i1 - ib, c, a_h, a_x);
                    else        csr_row_sum<H, false>(nullptr, g.in_eid, g.in_nbr, e, HX_in, i0, i1, c, a_h, a_x);
                    st4(m + 4 * c, a_h);
                    store_x_acc<H>(m + H, c, a_x);
                }
                {   // mo: sum over out-edges of e * HX[dst]
                    float4 a_h = zero, a_x = zero;
                    if (staged) csr_row_sum<H, true>(sPair + CAP, nullptr, nullptr, nullptr, HX_in, o0 - ob, o1 - ob, c, a_h, a_x);
                    else        csr_row_sum<H, false>(nullptr, g.out_eid, g.out_nbr, e, HX_in, o0, o1, c, a_h, a_x);
                    st4(m + D4 + 4 * c, a_h);
                    store_x_acc<H>(m + D4 + H, c, a_x);
                }
                {   // the node's own row
                    float4 a_h = zero, a_x = zero;
                    if (live) {
                        const float* row = HX_in + (size_t)n * D4;
                        a_h = ldg4(row + 4 * c);
                        if (c == 0) a_x = ldg4(row + H);
                    }
                    st4(m + 2 * D4 + 4 * c, a_h);
                    if (c == 0) st4(m + 2 * D4 + H, a_x);
                }
            }
            bar_arrive(BAR_FULL + buf, NT);
        }
    } else {
        // ================================ consumers ====================================
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it % NBUF;
            const int node0 = tile * TN;
            const float* sMb = sM + buf * TN * SM;
            bar_sync(BAR_FULL + buf, NT);                      // the tile's rows are in sMb
            if (dbg & 2) {
                bar_sync(BAR_CONS, CT);
                if (tile + NBUF * (int)gridDim.x < n_tiles) bar_arrive(BAR_EMPTY + buf, NT);
                continue;
            }
            // ---- layer 0: h1 = tanh(W3 . [mi; mo; self] + b3) ---------------------------
            if constexpr (MMA) {
                tile_gemm_mma<C::K1P / 8, H>(sMb, SM, sW3, SM, [&](int ln, int o, float v0, float v1) {
                    *reinterpret_cast<float2*>(sH1 + ln * SH + o) =
                        make_float2(tanhf(v0 + sB3[o]), tanhf(v1 + sB3[o + 1]));
                });
            } else {
                tile_gemm<K1, H, TN, CT, C::RN1, 4>(sMb, SM, sW3, [&](int ln, int o, const float* acc) {
                    const float4 b = lds4(sB3 + o);
                    st4(sH1 + ln * SH + o, make_float4(tanhf(acc[0] + b.x), tanhf(acc[1] + b.y),
                                                       tanhf(acc[2] + b.z), tanhf(acc[3] + b.w)));
                });
            }
            // the X part of the new HX row is the old one (self block of the tile)
            for (int i = threadIdx.x; i < TN; i += CT) st4(sHX + i * SD + H, lds4(sMb + i * SM + 2 * D4 + H));
            bar_sync(BAR_CONS, CT);
            if (tile + NBUF * (int)gridDim.x < n_tiles) bar_arrive(BAR_EMPTY + buf, NT);   // release sMb
            // ---- layer 2: H' = tanh(W4 . h1 + b4) ---------------------------------------
            if constexpr (MMA) {
                tile_gemm_mma<H / 8, H>(sH1, SH, sW4, SH, [&](int ln, int o, float v0, float v1) {
                    *reinterpret_cast<float2*>(sHX + ln * SD + o) =
                        make_float2(tanhf(v0 + sB4[o]), tanhf(v1 + sB4[o + 1]));
                });
            } else {
                tile_gemm<H, H, TN, CT, C::RN1, 4>(sH1, SH, sW4, [&](int ln, int o, const float* acc) {
                    const float4 b = lds4(sB4 + o);
                    st4(sHX + ln * SD + o, make_float4(tanhf(acc[0] + b.x), tanhf(acc[1] + b.y),
                                                       tanhf(acc[2] + b.z), tanhf(acc[3] + b.w)));
                });
            }
            bar_sync(BAR_CONS, CT);
            // ---- write HX' and the projection for the next edge step ---------------------
            if constexpr (MMA) {
                constexpr int C4 = D4 / 4;
                for (int i = threadIdx.x; i < TN * C4; i += CT) {
                    const int ln = i / C4, c = i % C4, n = node0 + ln;
                    if (n < n_nodes) st4(HX_out + (size_t)n * D4 + 4 * c, lds4(sHX + ln * SD + 4 * c));
                }
                tile_gemm_mma<C::D4P / 8, 2 * H>(sHX, SD, sW1, SD, [&](int ln, int o, float v0, float v1) {
                    const int n = node0 + ln;
                    if (n < n_nodes) {
                        if (o < H) { v0 += sB1[o]; v1 += sB1[o + 1]; }
                        *reinterpret_cast<float2*>(P_out + (size_t)n * 2 * H + o) = make_float2(v0, v1);
                    }
                });
            } else {
                store_hx_and_project<H, TN, CT, C::RNP, C::RCP>(sHX, SD, sW1, sB1, node0, n_nodes, HX_out, P_out);
            }
            bar_sync(BAR_CONS, CT);   // sHX / sH1 are rewritten by the next tile
        }
    }
}

// ------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------
static inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

template <typename Kern>
static inline int persistent_grid(Kern kern, int threads, size_t smem, int n_tiles, int* grid) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return GNNSEG_ECUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1)
        return GNNSEG_ECUDA;
    const int sms = sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int cap = sms * occ;
    *grid = n_tiles < cap ? n_tiles : cap;
    return GNNSEG_OK;
}

static inline int check_launch() {
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

template <int H>
static int launch_input(const float* blob, const float* X, int n_nodes, int F, float* HX, float* P,
                        cudaStream_t st) {
    using C = InputCfg<H>;
    if (n_nodes == 0) return GNNSEG_OK;
    const int n_tiles = (n_nodes + C::TN - 1) / C::TN;
    int grid = 0;
    const int rc = persistent_grid(input_kernel<H>, C::NT, C::SMEM_BYTES, n_tiles, &grid);
    if (rc) return rc;
    input_kernel<H><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, X, F, n_nodes, n_tiles, HX, P);
    return check_launch();
}

template <int H>
static int launch_edge(const float* blob, const GnnsegGraph* g, const float* P, float* e, cudaStream_t st) {
    if (g->n_slots == 0) return GNNSEG_OK;
    const int sms = sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int warps_needed = (g->n_slots + 31) / 32;
    int grid = (warps_needed + 7) / 8;
    const int cap = sms * 8;   // 8 CTAs of 256 threads per SM
    if (grid > cap) grid = cap;
    edge_kernel<H><<<grid, 256, 0, st>>>(blob, P, g->src, g->dst, g->n_slots, e);
    return check_launch();
}

int launch_node_tc32(const float* blob, const GnnsegGraph* g, const float* HX_in, const float* e,
                     float* HX_out, float* P_out, cudaStream_t st);   // gnnseg_node_tc.cu (tcgen05)

template <int H>
static int launch_node(const float* blob, const GnnsegGraph* g, const float* HX_in, const float* e,
                       float* HX_out, float* P_out, cudaStream_t st) {
    using C = NodeCfg<H>;
    if (g->n_nodes == 0) return GNNSEG_OK;
    if (H == 32) {
        const char* impl = getenv("GNNSEG_NODE_IMPL");   // "mma": legacy mma.sync path (for A/B runs)
        if (!impl || impl[0] != 'm') return launch_node_tc32(blob, g, HX_in, e, HX_out, P_out, st);
    }
    const int n_tiles = (g->n_nodes + C::TN - 1) / C::TN;
    int grid = 0;
    const int rc = persistent_grid(node_kernel<H>, C::NT, C::SMEM_BYTES, n_tiles, &grid);
    if (rc) return rc;
    const char* dbg_env = getenv("GNNSEG_DBG");
    node_kernel<H><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, *g, HX_in, e, n_tiles, HX_out, P_out, dbg_env ? atoi(dbg_env) : 0);
    return check_launch();
}

#define GNNSEG_DISPATCH_H(h, CALL)                 \
    switch (h) {                                   \
        case 4:  { constexpr int HH = 4;  return CALL; } \
        case 8:  { constexpr int HH = 8;  return CALL; } \
        case 16: { constexpr int HH = 16; return CALL; } \
        case 32: { constexpr int HH = 32; return CALL; } \
        case 64: { constexpr int HH = 64; return CALL; } \
        default: return GNNSEG_EUNSUPPORTED;       \
    }

int input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* HX, float* P,
               cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_input<HH>(blob, X, n_nodes, F, HX, P, st));
}
int edge_step(const float* blob, const GnnsegGraph* g, const float* P, int h, float* e, cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_edge<HH>(blob, g, P, e, st));
}
int node_step(const float* blob, const GnnsegGraph* g, const float* HX_in, const float* e, int h,
              float* HX_out, float* P_out, cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_node<HH>(blob, g, HX_in, e, HX_out, P_out, st));
}
int pack_weights(const GnnsegParams* p, int F, int h, float* blob, cudaStream_t st) {
    const int total = blob_total(h);
    pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(*p, F, h, blob);
    return check_launch();
}

}  // namespace gnnseg
