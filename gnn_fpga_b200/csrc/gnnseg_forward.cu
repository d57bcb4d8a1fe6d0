// Forward kernels of the segment classifier for sm_100a: weight packing, input step, edge step,
// node step (generic path; hidden_dim = 32 runs the tcgen05 kernel in gnnseg_node_tc.cu).
// The math is gnn/model.py:140-156 of the reference with the dense incidence bmm replaced by int32
// gathers and an ordered CSR segment sum, in the projection-first form described in
// gnnseg_common.cuh:
//
//   input step (gnn/model.py:144-146)  H0 = tanh(Win.X + bin); projections of [H0 | X]
//   edge step  (gnn/model.py:69-81)    e_j = sigmoid(W2 . tanh(Ps[src_j] + Pd[dst_j]) + b2)
//     two aligned H-float rows gathered with 16-byte loads per lane, tanh, dot, sigmoid.
//   node step  (gnn/model.py:113-125)  h1 = tanh(Qs[n] + sum_in e Qi[src] + sum_out e Qo[dst])
//                                      H' = tanh(W4 . h1 + b4); projections of [H' | X]
//     persistent warp-specialised CTAs: producer warps walk the two CSR rows of every node of a
//     tile (no atomics, ascending slot order, own term first, then in-edges, then out-edges) and
//     leave h1 in shared memory; consumer warps run layer 2 and the five projections as
//     register-tiled GEMMs (SIMT fp32, or 3xTF32 mma.sync for hidden_dim >= 32) against weights
//     resident in shared memory and write P' and Q'.
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {

// ------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------
// head_w != nullptr packs the blob of the NodeClassifier's LAST producing step
// (gnn/MPNN_HitClassifier.ipynb c21: output_network = Linear(D, 1) + Sigmoid on [H | X]): the
// logit Wo.[H | X] + bo is one more linear map of [H | X], so it takes the place of column 0 of
// the start-node projection (the final step's edge projections have no reader in that model);
// the other edge-projection columns are zero.  The node-network part is packed as usual.
__global__ void pack_weights_kernel(GnnsegParams p, int F, int H, float* __restrict__ blob,
                                    const float* __restrict__ head_w, const float* __restrict__ head_b) {
    const int D = F + H, D4 = H + 4;
    const int o_bin = 4 * H, o_wp = o_bin + H, o_bp = o_wp + D4 * 5 * H, o_w2 = o_bp + 5 * H,
              o_b2 = o_w2 + H, o_w4 = o_b2 + 4, o_b4 = o_w4 + H * H, total = o_b4 + H;   // fp32 part
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < o_bin) {                       // Win^T [4][H]
            const int f = i / H, j = i % H;
            if (f < F) v = p.w_in[j * F + f];
        } else if (i < o_wp) {
            v = p.b_in[i - o_bin];
        } else if (i < o_bp) {                 // WP^T [D4][5H] = [W1a | W1b | W3a | W3b | W3c]
            const int r = i - o_wp, k = r / (5 * H), c = r % (5 * H), blk = c / H, j = c % H;
            if (k < D) {
                if (blk < 2 && head_w) {
                    v = (blk == 0 && j == 0) ? head_w[k] : 0.f;
                } else if (blk < 2) {
                    const int idx = j * (2 * D) + blk * D + k;
                    v = p.w_e1[idx];
                    if (p.m_e1) v *= p.m_e1[idx];
                } else {
                    const int idx = j * (3 * D) + (blk - 2) * D + k;
                    v = p.w_n1[idx];
                    if (p.m_n1) v *= p.m_n1[idx];
                }
            }
        } else if (i < o_w2) {                 // BP [5H] = [b1 | 0 | 0 | 0 | b3]
            const int c = i - o_bp;
            if (c < H) v = head_w ? (c == 0 ? head_b[0] : 0.f) : p.b_e1[c];
            else if (c >= 4 * H) v = p.b_n1[c - 4 * H];
        } else if (i < o_b2) {
            const int j = i - o_w2;
            v = p.w_e2[j];
            if (p.m_e2) v *= p.m_e2[j];
        } else if (i < o_w4) {
            if (i == o_b2) v = p.b_e2[0];
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

// tf32 hi / lo operand images of W4 ([H][H]) and WP ([5H][D4P]) in the canonical K-major
// core-matrix layout, from the fp32 [k][out] matrices written by pack_weights_kernel.
__global__ void pack_tc_images_kernel(int H, float* __restrict__ blob) {
    const int D4 = H + 4, D4P = (D4 + 7) / 8 * 8;
    const int o_wp = 5 * H, o_bp = o_wp + D4 * 5 * H, o_w2 = o_bp + 5 * H, o_w4 = o_w2 + H + 4, o_img = o_w4 + H * H + H;
    const int n_w4 = H * H, n_wp = 5 * H * D4P;
    const int o_w2n = o_img + 2 * n_w4 + 2 * n_wp, o_z0 = o_w2n + H, o_sb1 = o_z0 + 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_w4 + n_wp + 2 * H + 4; i += gridDim.x * blockDim.x) {
        if (i >= n_w4 + n_wp) {
            // the edge network's second layer for gnnseg_fused.cu: -2 w2, b2 + sum(w2), 2^(log2e b1)
            const int r = i - (n_w4 + n_wp);
            if (r < H) blob[o_w2n + r] = -2.f * blob[o_w2 + r];
            else if (r < H + 4) {
                float z = 0.f;
                if (r == H) { z = blob[o_w2 + H]; for (int k = 0; k < H; ++k) z += blob[o_w2 + k]; }
                blob[o_z0 + (r - H)] = z;
            } else {
                const float a = fminf(fmaxf(blob[o_bp + (r - H - 4)] * 1.4426950408889634f, -63.f), 63.f);
                blob[o_sb1 + (r - H - 4)] = exp2f(a);
            }
            continue;
        }
        const bool is_w4 = i < n_w4;
        const int f = is_w4 ? i : i - n_w4;                       // float index inside the image
        const int K = is_w4 ? H : D4P;                            // padded K of this operand
        const int sbo = (K / 4) * 128;                            // bytes between 8-row groups
        const int bytes = f * 4;
        const int grp = bytes / sbo, rem = bytes % sbo;
        const int row = grp * 8 + (rem % 128) / 16;
        const int k = (rem / 128) * 4 + (rem % 16) / 4;
        float v;
        // WP image, K index D4 (the first padding column): the bias of the projections.  The tensor-core
        // kernels put a constant 1 in that column of the A operand, so the GEMM adds the bias.
        if (is_w4) v = blob[o_w4 + k * H + row];
        else       v = k < D4 ? blob[o_wp + k * 5 * H + row] : (k == D4 ? blob[o_wp + D4 * 5 * H + row] : 0.f);
        uint32_t hi_bits;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_bits) : "f"(v));
        const float hi = __uint_as_float(hi_bits);
        uint32_t lo_bits;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_bits) : "f"(v - hi));
        const int o_hi = is_w4 ? o_img : o_img + 2 * n_w4;
        const int o_lo = is_w4 ? o_img + n_w4 : o_img + 2 * n_w4 + n_wp;
        blob[o_hi + f] = hi;
        blob[o_lo + f] = __uint_as_float(lo_bits);
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

// Store one float4 of the projection row of node n: columns [0,2H) go to P, [2H,5H) to Q.
template <int H>
__device__ __forceinline__ void store_proj4(float* __restrict__ P, float* __restrict__ Q, const int n,
                                            const int o, const float4 v, const bool write_q) {
    if (o < 2 * H) st4(P + (size_t)n * 2 * H + o, v);
    else if (write_q) st4(Q + (size_t)n * 3 * H + (o - 2 * H), v);
}
template <int H>
__device__ __forceinline__ void store_proj2(float* __restrict__ P, float* __restrict__ Q, const int n,
                                            const int o, const float v0, const float v1, const bool write_q) {
    if (o < 2 * H) *reinterpret_cast<float2*>(P + (size_t)n * 2 * H + o) = make_float2(v0, v1);
    else if (write_q) *reinterpret_cast<float2*>(Q + (size_t)n * 3 * H + (o - 2 * H)) = make_float2(v0, v1);
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
    static constexpr bool MMA = H >= 32;             // 3xTF32 mma.sync MLP; SIMT fp32 below that
    static constexpr int NBUF = 2;                   // h1 tile buffers
    static constexpr int MINB = (H >= 64) ? 1 : 2;   // CTAs per SM the register budget allows
    static constexpr int RN   = 2;                   // SIMT path: nodes per thread
    static constexpr int D4   = H + 4;
    static constexpr int D4P  = (D4 + 7) / 8 * 8;    // MMA path: K padded to the k8 step
    static constexpr int SH   = MMA ? mma_stride(H) : tile_stride(H);      // h1 tile stride
    static constexpr int SD   = MMA ? mma_stride(D4P) : tile_stride(D4);   // [H'|X] tile stride
    static constexpr int CAP  = (H >= 64) ? 1536 : 768;   // staged CSR slots per direction per tile
    // weights in shared memory: SIMT keeps the blob's [k][out]; MMA wants [out][k] rows with the
    // same padded strides as the A tiles
    static constexpr int W_FLOATS = (MMA ? (H * SH + 5 * H * SD) : (H * H + D4 * 5 * H)) + 6 * H;
    static constexpr int STAGE_WORDS = 2 * 2 * CAP + 2 * (TN + 4);   // (nbr,w) pairs + row pointers
    static constexpr int SMEM_FLOATS = W_FLOATS + TN * (NBUF * SH + SD) + STAGE_WORDS;
    static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;
};

template <int H>
struct InputCfg {
    static constexpr int TN = 64, NT = 256, RN = 2;
    static constexpr int D4 = H + 4;
    static constexpr int SD = tile_stride(D4);
    static constexpr int SMEM_FLOATS = 4 * H + H + D4 * 5 * H + 5 * H + TN * SD + TN * 4;
    static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;
};

// ------------------------------------------------------------------------------------
// input step: H0 = tanh(Win.X + bin), then X4, P, Q of [H0 | X]        (gnn/model.py:144-146)
// ------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(InputCfg<H>::NT)
input_kernel(const float* __restrict__ blob, const float* __restrict__ X, const int F,
             const int n_nodes, const int n_tiles, float* __restrict__ X4,
             float* __restrict__ P_out, float* __restrict__ Q_out, float* __restrict__ H_save) {
    using C = InputCfg<H>;
    using B = Blob<H>;
    constexpr int TN = C::TN, NT = C::NT, D4 = C::D4, SD = C::SD;
    extern __shared__ __align__(16) float smem[];
    float* sWin = smem;                 // [4][H]
    float* sBin = sWin + 4 * H;         // [H]
    float* sWP  = sBin + H;             // [D4][5H]
    float* sBP  = sWP + D4 * 5 * H;     // [5H]
    float* sHX  = sBP + 5 * H;          // [TN][SD]
    float* sX   = sHX + TN * SD;        // [TN][4]
    pdl_launch_dependents();
    copy_to_smem<NT>(sWin, blob + B::WIN, 4 * H + H);
    copy_to_smem<NT>(sWP, blob + B::WP, D4 * 5 * H + 5 * H);

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
                v.x = tanh_node(v.x); v.y = tanh_node(v.y); v.z = tanh_node(v.z); v.w = tanh_node(v.w);
                if (H_save && node0 + ln < n_nodes) st4(H_save + (size_t)(node0 + ln) * H + 4 * c, v);   // training: H0
            } else {
                v = x;
                if (node0 + ln < n_nodes) st4(X4 + (size_t)(node0 + ln) * 4, x);
            }
            st4(sHX + ln * SD + 4 * c, v);
        }
        __syncthreads();
        tile_gemm<D4, 5 * H, TN, NT, C::RN, 4>(sHX, SD, sWP, [&](int ln, int o, const float* acc) {
            const int n = node0 + ln;
            if (n < n_nodes) {
                const float4 b = lds4(sBP + o);
                store_proj4<H>(P_out, Q_out, n, o, make_float4(acc[0] + b.x, acc[1] + b.y, acc[2] + b.z, acc[3] + b.w), true);
            }
        });
    }
}

// ------------------------------------------------------------------------------------
// edge step
// ------------------------------------------------------------------------------------
// Measured on B200 (scripts/micro/gather_bench.cu): the rate at which an SM gathers random 128-byte
// rows with LDG.128 grows with the number of resident WARPS, not with the loads a thread keeps in
// flight (5 TB/s at 16 warps per SM, 9 at 32, 15 at 64, the same for 4, 8 or 16 loads per thread).
// In this kernel the product (resident warps x rows in flight per warp) is flat around its optimum:
// at hidden_dim 32, two slots per lane group at 48 warps (40 registers) 44.5 us, four slots at 32 warps
// 44.0, four at 40 warps 47.1, one slot at 64 warps (32 registers, nothing spilled) 50.9.  PAIRS slots
// are in flight per lane group, indices are fetched per group of PAIRS.
template <int H, int PAIRS = 2, int MINB = 6>
__global__ void __launch_bounds__(256, MINB)
edge_kernel(const float* __restrict__ blob, const float* __restrict__ P,
            const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
            const int32_t* __restrict__ in_pos, const int32_t* __restrict__ out_pos,
            const int n_slots, float* __restrict__ e_slot, float* __restrict__ e_in,
            float* __restrict__ e_out) {
    using B = Blob<H>;
    constexpr int G = H / 4;        // lanes that share one edge (one float4 of P each)
    constexpr int EPP = 32 / G;     // edges a warp handles per iteration; G iterations cover 32 slots
    constexpr int PAIR = G >= PAIRS ? PAIRS : (G >= 2 ? 2 : 1);
    const int lane = threadIdx.x & 31;
    const int c = lane % G, g = lane / G;
    const uint64_t keep = l2_policy_evict_last();   // every P row is gathered ~deg times
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const float* Pa = P + 4 * c;                    // this lane's chunk of the start-node half
    const float* Pb = P + H + 4 * c;                // ... and of the end-node half
    pdl_launch_dependents();
    pdl_wait();                                     // P comes from the kernel before

    for (int base = warp * 32; base < n_slots; base += n_warps * 32) {
        // iteration p of the group (g, *) works on slot base + p*EPP + g: the G lanes of a group read
        // the same endpoint words (one broadcast load), no index shuffles
        const int j = base + c * EPP + g;               // the slot whose score this lane will hold
        int pi = -1, po = -1;                           // its CSR positions: fetched now, needed after the MLP
        if (j < n_slots) {
            if (e_in) pi = __ldg(in_pos + j);
            if (e_out) po = __ldg(out_pos + j);
        }
        float z[G];
#pragma unroll
        for (int p0 = 0; p0 < G; p0 += PAIR) {
            int s[PAIR], d[PAIR];
#pragma unroll
            for (int q = 0; q < PAIR; ++q) {
                const int jj = base + (p0 + q) * EPP + g;
                s[q] = d[q] = -1;
                if (jj < n_slots) { s[q] = __ldg(src + jj); d[q] = __ldg(dst + jj); }
            }
            float4 a[PAIR], b[PAIR];
#pragma unroll
            for (int q = 0; q < PAIR; ++q) {
                // absent start: W1a.0 + b1; absent end: W1b.0.  Branch free: row 0 is loaded and dropped.
                // (Selecting the pointer instead of the values, or a warp-uniform branch around the selects,
                // costs registers: 72 - 200 bytes of spills at the same launch bounds.)
                a[q] = ldg4_hint(Pa + (size_t)max(s[q], 0) * (2 * H), keep);
                b[q] = ldg4_hint(Pb + (size_t)max(d[q], 0) * (2 * H), keep);
            }
            const float4 w2 = ldg4(blob + B::W2 + 4 * c);      // L1-resident: cheaper than four live registers
#pragma unroll
            for (int q = 0; q < PAIR; ++q) {
                if (s[q] < 0) a[q] = ldg4(blob + B::BP + 4 * c);
                if (d[q] < 0) b[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                float t = w2.x * tanh_edge(a[q].x + b[q].x);
                t = fmaf(w2.y, tanh_edge(a[q].y + b[q].y), t);
                t = fmaf(w2.z, tanh_edge(a[q].z + b[q].z), t);
                z[p0 + q] = fmaf(w2.w, tanh_edge(a[q].w + b[q].w), t);
            }
        }
        // transposed butterfly: G partial sums on each of G lanes -> lane c holds the total of
        // iteration p = c (G - 1 shuffles instead of G log2 G)
#pragma unroll
        for (int o = G / 2, cnt = G; o > 0; o >>= 1, cnt >>= 1) {
            const bool upper = (c & o) != 0;
#pragma unroll
            for (int i = 0; i < cnt / 2; ++i) {
                const float send = upper ? z[i] : z[i + cnt / 2];
                const float mine = upper ? z[i + cnt / 2] : z[i];
                z[i] = mine + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        if (j < n_slots) {
            const float score = 1.f / (1.f + expf(-(z[0] + __ldg(blob + B::B2))));
            if (e_slot) e_slot[j] = score;
            if (pi >= 0) e_in[pi] = score;              // the node step reads the scores in CSR order
            if (po >= 0) e_out[po] = score;
        }
    }
}

// ------------------------------------------------------------------------------------
// node step, first half: the ordered CSR sum.  h1[n] = tanh(Qs[n] + sum_in e Qi[src] + sum_out e Qo[dst]),
// own term first, then in-slots, then out-slots, ascending: one thread per (node, float4 chunk), no
// atomics.  Written for 64 resident warps per SM (32 registers) with three rows in flight per thread
// (see edge_kernel for the measured trade).  h1 rows are written with stride ld_out; the tensor-core
// MLP kernel (node_mlp_kernel_tc) reads them back.
// ------------------------------------------------------------------------------------
template <int H, int U = 3, int MINB = 8, int NTHR = 256, bool RANGES = false>
__global__ void __launch_bounds__(NTHR, MINB)
node_gather_kernel(const GnnsegGraph g, const float* __restrict__ Q_in, const float* __restrict__ e_in,
                   const float* __restrict__ e_out, float* __restrict__ h1_out, const int ld_out,
                   float* __restrict__ h1_save) {
    constexpr int G = H / 4;
    const uint64_t keep = l2_policy_evict_last();
    const long long total = (long long)g.n_nodes * G;
    pdl_launch_dependents();
    pdl_wait();                                     // Q_in, e_in, e_out come from the kernels before
    // U slots of a CSR row at a time: all U rows are requested before the first FMA, added in slot order.
    // Measured (acts64 / mu200, per launch): U = 3 at 64 warps 45.8 / 45.5 us, U = 4 at 48 warps (40 registers)
    // 46.5 / 47.4 us, U = 2 at 64 warps 47.6 / 46.3 us.
    auto row_sum = [&](const int32_t* __restrict__ nbr, const float* __restrict__ ew, const float* __restrict__ Qcol,
                       const int beg, const int end, float4& acc) {
#pragma unroll 1
        for (int s = beg; s < end; s += U) {
            int nb[U];
            float w[U];
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in = s + u < end;
                nb[u] = in ? __ldg(nbr + s + u) : -1;
                w[u] = in ? __ldg(ew + s + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nb[u] >= 0) v[u] = ldg4_hint(Qcol + (size_t)nb[u] * 3 * H, keep);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (nb[u] >= 0) fma4(acc, w[u], v[u]);                        // nb < 0: half edge, the zero row
        }
    };
    // RANGES: every CTA sweeps one contiguous range of nodes (neighbouring nodes share neighbours when the
    // batch was renumbered for locality: the rows stay in this SM's L1); otherwise grid-stride
    long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, idx1 = total, step = (long long)gridDim.x * blockDim.x;
    if (RANGES) {
        const long long chunk = ((total + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
        idx0 = (long long)blockIdx.x * chunk + threadIdx.x;
        idx1 = min(total, (long long)(blockIdx.x + 1) * chunk);
        step = blockDim.x;
    }
    for (long long idx = idx0; idx < idx1; idx += step) {
        const int n = (int)(idx / G), c = (int)(idx % G);
        const int i0 = __ldg(g.in_ptr + n), i1 = __ldg(g.in_ptr + n + 1);
        const int o0 = __ldg(g.out_ptr + n), o1 = __ldg(g.out_ptr + n + 1);
        float4 acc = ldg4(Q_in + (size_t)n * 3 * H + 2 * H + 4 * c);             // Qs[n] (holds b3)
        row_sum(g.in_nbr, e_in, Q_in + 4 * c, i0, i1, acc);
        row_sum(g.out_nbr, e_out, Q_in + H + 4 * c, o0, o1, acc);
        acc.x = tanh_node(acc.x); acc.y = tanh_node(acc.y); acc.z = tanh_node(acc.z); acc.w = tanh_node(acc.w);
        st4(h1_out + (size_t)n * ld_out + 4 * c, acc);
        if (h1_save) st4(h1_save + (size_t)n * H + 4 * c, acc);
    }
}

// ------------------------------------------------------------------------------------
// node step (generic path): warp-specialised persistent kernel.
//   producer warps  stage the tile's CSR slices, gather and sum the aligned Q rows and leave
//                   h1 = tanh(...) of the tile in shared memory (latency bound: many loads in flight)
//   consumer warps  run layer 2 and the projections as register-tiled GEMMs out of shared memory
//                   and write P' and Q'
// The two halves meet through named barriers (full/empty per tile buffer), so the gather of
// tile t+1 overlaps the MLP of tile t on the same SM.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync(const int id, const int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(const int id, const int n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// One CSR row, one float4 chunk per lane: acc += w_s * Qrow[nbr_s] for the row's slots in
// ascending order.  Slots go in batches of U: all U row loads are issued before the first FMA
// (no branch in between: an absent neighbour loads row 0 and is dropped by predication).
// STAGED: the (neighbour, weight) pairs of the whole tile were fetched into shared memory
// beforehand by coalesced loads.
template <bool STAGED>
__device__ __forceinline__ void csr_row_sum(const int2* __restrict__ pairs, const int32_t* __restrict__ nbr,
                                            const float* __restrict__ ew, const float* __restrict__ Qcol,
                                            const int row_floats, const int beg, const int end, float4& acc) {
    constexpr int U = 4;
    const uint64_t keep = l2_policy_evict_last();
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
            ok[u] = (s0 + u < end) && nb >= 0;   // nb < 0: half edge, gathers the zero row
            v[u] = ldg4_hint(Qcol + (size_t)max(nb, 0) * row_floats, keep);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) fma4(acc, w[u], v[u]);
    }
}

template <int H>
__global__ void __launch_bounds__(NodeCfg<H>::NT, NodeCfg<H>::MINB)
node_kernel(const float* __restrict__ blob, const GnnsegGraph g, const float* __restrict__ X4,
            const float* __restrict__ Q_in, const float* __restrict__ e_in, const float* __restrict__ e_out, const int n_tiles,
            float* __restrict__ P_out, float* __restrict__ Q_out, const int write_q,
            float* __restrict__ h1_save, float* __restrict__ H_save) {
    using C = NodeCfg<H>;
    using B = Blob<H>;
    constexpr int TN = C::TN, NT = C::NT, CT = C::CT, PT = C::PT, NBUF = C::NBUF;
    constexpr int D4 = C::D4, SH = C::SH, SD = C::SD;
    constexpr bool MMA = C::MMA;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 1 + NBUF, BAR_CONS = 1 + 2 * NBUF, BAR_PROD = 2 + 2 * NBUF;
    extern __shared__ __align__(16) float smem[];
    // weights: SIMT path [k][out] as in the blob, MMA path [out][k] with padded row strides
    float* sW4 = smem;                                     // SIMT [H][H]     | MMA [H][SH]
    float* sWP = sW4 + (MMA ? H * SH : H * H);             // SIMT [D4][5H]   | MMA [5H][SD]
    float* sBP = sWP + (MMA ? 5 * H * SD : D4 * 5 * H);    // [5H]
    float* sB4 = sBP + 5 * H;                              // [H]
    float* sH1 = sB4 + H;               // [NBUF][TN][SH]   h1 tiles
    float* sHX = sH1 + NBUF * TN * SH;  // [TN][SD]         [H' | X | 0]
    int2* sPair = reinterpret_cast<int2*>(sHX + TN * SD);          // [2][CAP] (nbr, w) in / out
    int*  sPtr  = reinterpret_cast<int*>(sPair + 2 * C::CAP);      // [2][TN+4] row pointers in / out
    if constexpr (MMA) {
        // transpose the blob's [k][out] matrices into [out][k] rows; k beyond the real width is 0
        for (int i = threadIdx.x; i < H * SH; i += NT) {
            const int j = i / SH, k = i % SH;
            sW4[i] = k < H ? __ldg(blob + B::W4 + k * H + j) : 0.f;
        }
        for (int i = threadIdx.x; i < 5 * H * SD; i += NT) {
            const int j = i / SD, k = i % SD;
            sWP[i] = k < D4 ? __ldg(blob + B::WP + k * 5 * H + j) : 0.f;
        }
        // zero the K padding of the [H'|X] tile once: nobody writes those columns afterwards
        for (int i = threadIdx.x; i < TN * (SD - D4); i += NT) sHX[(i / (SD - D4)) * SD + D4 + i % (SD - D4)] = 0.f;
    } else {
        copy_to_smem<NT>(sW4, blob + B::W4, H * H);
        copy_to_smem<NT>(sWP, blob + B::WP, D4 * 5 * H);
    }
    for (int i = threadIdx.x; i < 5 * H; i += NT) sBP[i] = __ldg(blob + B::BP + i);
    for (int i = threadIdx.x; i < H; i += NT) sB4[i] = __ldg(blob + B::B4 + i);
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();                         // the weights are staged; Q_in / e_in / e_out come from the kernels before
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
                    sPair[s] = make_int2(__ldg(g.in_nbr + ib + s), __float_as_int(__ldg(e_in + ib + s)));
                for (int s = pt; s < oc; s += PT)
                    sPair[CAP + s] = make_int2(__ldg(g.out_nbr + ob + s), __float_as_int(__ldg(e_out + ob + s)));
            }
            bar_sync(BAR_PROD, PT);
            if (it >= NBUF) bar_sync(BAR_EMPTY + buf, NT);     // consumers are done with this buffer
            float* sHb = sH1 + buf * TN * SH;
            for (int ln = grp; ln < TN; ln += NGRP) {
                const int n = node0 + ln;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < n_nodes) {
                    acc = ldg4(Q_in + (size_t)n * 3 * H + 2 * H + 4 * c);             // Qs[n] (holds b3)
                    const int i0 = sPtr[ln], i1 = sPtr[ln + 1], o0 = sPtr[TN + 4 + ln], o1 = sPtr[TN + 4 + ln + 1];
                    if (staged) {
                        csr_row_sum<true>(sPair, nullptr, nullptr, Q_in + 4 * c, 3 * H, i0 - ib, i1 - ib, acc);
                        csr_row_sum<true>(sPair + CAP, nullptr, nullptr, Q_in + H + 4 * c, 3 * H, o0 - ob, o1 - ob, acc);
                    } else {
                        csr_row_sum<false>(nullptr, g.in_nbr, e_in, Q_in + 4 * c, 3 * H, i0, i1, acc);
                        csr_row_sum<false>(nullptr, g.out_nbr, e_out, Q_in + H + 4 * c, 3 * H, o0, o1, acc);
                    }
                    acc.x = tanh_node(acc.x); acc.y = tanh_node(acc.y); acc.z = tanh_node(acc.z); acc.w = tanh_node(acc.w);
                    if (h1_save) st4(h1_save + (size_t)n * H + 4 * c, acc);       // training: kept for the backward pass
                }
                st4(sHb + ln * SH + 4 * c, acc);
            }
            bar_arrive(BAR_FULL + buf, NT);
        }
    } else {
        // ================================ consumers ====================================
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it % NBUF;
            const int node0 = tile * TN;
            const float* sHb = sH1 + buf * TN * SH;
            // the X part of the [H'|X] rows (independent of the producers)
            for (int i = threadIdx.x; i < TN; i += CT) {
                const int n = node0 + i;
                st4(sHX + i * SD + H, n < n_nodes ? ldg4(X4 + (size_t)n * 4) : make_float4(0.f, 0.f, 0.f, 0.f));
            }
            bar_sync(BAR_FULL + buf, NT);                      // the tile's h1 rows are in sHb
            // ---- layer 2: H' = tanh(W4 . h1 + b4) ---------------------------------------
            if constexpr (MMA) {
                tile_gemm_mma<H / 8, H>(sHb, SH, sW4, SH, [&](int ln, int o, float v0, float v1) {
                    const float2 hv = make_float2(tanh_node(v0 + sB4[o]), tanh_node(v1 + sB4[o + 1]));
                    *reinterpret_cast<float2*>(sHX + ln * SD + o) = hv;
                    if (H_save && node0 + ln < n_nodes) *reinterpret_cast<float2*>(H_save + (size_t)(node0 + ln) * H + o) = hv;
                });
            } else {
                tile_gemm<H, H, TN, CT, C::RN, 4>(sHb, SH, sW4, [&](int ln, int o, const float* acc) {
                    const float4 b = lds4(sB4 + o);
                    const float4 hv = make_float4(tanh_node(acc[0] + b.x), tanh_node(acc[1] + b.y),
                                                  tanh_node(acc[2] + b.z), tanh_node(acc[3] + b.w));
                    st4(sHX + ln * SD + o, hv);
                    if (H_save && node0 + ln < n_nodes) st4(H_save + (size_t)(node0 + ln) * H + o, hv);
                });
            }
            bar_sync(BAR_CONS, CT);
            if (tile + NBUF * (int)gridDim.x < n_tiles) bar_arrive(BAR_EMPTY + buf, NT);   // release sHb
            // ---- projections of [H'|X] for the next edge and node steps ---------------------
            if constexpr (MMA) {
#pragma unroll 1
                for (int blk = 0; blk < 5; ++blk) {
                    if (blk >= 2 && !write_q) break;
                    tile_gemm_mma<C::D4P / 8, H>(sHX, SD, sWP + blk * H * SD, SD, [&](int ln, int o, float v0, float v1) {
                        const int n = node0 + ln, oc = blk * H + o;
                        if (n < n_nodes) store_proj2<H>(P_out, Q_out, n, oc, v0 + sBP[oc], v1 + sBP[oc + 1], true);
                    });
                }
            } else {
                tile_gemm<D4, 5 * H, TN, CT, C::RN, 4>(sHX, SD, sWP, [&](int ln, int o, const float* acc) {
                    const int n = node0 + ln;
                    if (n < n_nodes) {
                        const float4 b = lds4(sBP + o);
                        store_proj4<H>(P_out, Q_out, n, o, make_float4(acc[0] + b.x, acc[1] + b.y, acc[2] + b.z, acc[3] + b.w), write_q != 0);
                    }
                });
            }
            bar_sync(BAR_CONS, CT);   // sHX is rewritten by the next tile
        }
    }
}

// ------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------
static inline int sm_count() { return cached_sm_count(); }

template <auto Kern>
static inline int persistent_grid(int threads, size_t smem, int n_tiles, int* grid) {
    if (!ensure_dynamic_smem<Kern>((int)smem)) return GNNSEG_ECUDA;
    const int occ = cached_occupancy<Kern>(threads, smem);
    if (occ < 1) return GNNSEG_ECUDA;
    const int sms = sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int cap = sms * occ;
    *grid = n_tiles < cap ? n_tiles : cap;
    return GNNSEG_OK;
}

// Programmatic dependent launch pays where the forward is launch bound (Toy2D batch: 27.9 -> 26.1 us);
// on the large batches it measured neutral to negative for the forward as a whole (round 2: acts64 0.562 ->
// 0.569 ms, mu200 1.10 -> 1.31 ms), so it is used for small graphs only.  GNNSEG_PDL=0 / 1 forces it off / on (A/B).
static int pdl_forced() {
    static const int forced = [] { const char* v = getenv("GNNSEG_PDL"); return v ? (v[0] == '0' ? 0 : 1) : -1; }();
    return forced;
}
bool use_pdl(const int n_slots) { return pdl_forced() >= 0 ? pdl_forced() == 1 : n_slots < (1 << 17); }
// The exception, at every size: the MLP launch of the fused path behind the fused gather.  The one-CTA-per-SM MLP kernel
// gets onto an SM as soon as that SM's gather CTAs have left and runs its prologue (tensor-memory allocation, barriers,
// weights) under the gather's tail: acts64 0.562 -> 0.556 ms, mu200 1.095 -> 1.074 ms.  The other direction (a gather
// or the final edge step as the dependent of an MLP kernel) is what loses: mu200 1.095 -> 1.32 ms.
bool use_pdl_mlp() { return pdl_forced() != 0; }

static inline int check_launch() {
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int launch_input_tc32(const float* blob, const float* X, int n_nodes, int F, float* X4, float* P, float* Q,
                      float* H_save, cudaStream_t st);   // gnnseg_node_tc.cu (tcgen05)
int launch_input_tc64(const float* blob, const float* X, int n_nodes, int F, float* X4, float* P, float* Q,
                      float* H_save, cudaStream_t st);   // gnnseg_node_tc.cu (tcgen05, weights streamed)

template <int H>
static int launch_input(const float* blob, const float* X, int n_nodes, int F, float* X4, float* P, float* Q,
                        float* H_save, cudaStream_t st) {
    using C = InputCfg<H>;
    if (n_nodes == 0) return GNNSEG_OK;
    if (H == 32 || H == 64) {
        const char* impl = getenv("GNNSEG_NODE_IMPL");   // "mma": generic path (for A/B runs)
        if (!impl || impl[0] != 'm')
            return H == 32 ? launch_input_tc32(blob, X, n_nodes, F, X4, P, Q, H_save, st)
                           : launch_input_tc64(blob, X, n_nodes, F, X4, P, Q, H_save, st);
    }
    const int n_tiles = (n_nodes + C::TN - 1) / C::TN;
    int grid = 0;
    const int rc = persistent_grid<input_kernel<H>>(C::NT, C::SMEM_BYTES, n_tiles, &grid);
    if (rc) return rc;
    input_kernel<H><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, X, F, n_nodes, n_tiles, X4, P, Q, H_save);
    return check_launch();
}

template <int H>
static int launch_edge(const float* blob, const GnnsegGraph* g, const float* P, float* e, float* e_in,
                       float* e_out, cudaStream_t st) {
    if (g->n_slots == 0) return GNNSEG_OK;
    const int sms = sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int warps_needed = (g->n_slots + 31) / 32;
    int grid = (warps_needed + 7) / 8;
    // hidden_dim 64: four slots (eight rows) in flight per lane group at 40 resident warps measured 55.6 us
    // against 59.8 us for the two-slot form at mu200 size; at hidden_dim 32 the two forms are level
    // (44.0 - 47.1 us against 44.5 us), so the narrower rows keep the two-slot form and 48 warps.
    constexpr int PAIRS = H >= 64 ? 4 : 2, MINB = H >= 64 ? 5 : 6;
    const int occ = cached_occupancy<edge_kernel<H, PAIRS, MINB>>(256, 0);   // resident CTAs per SM: one full wave, the warps stride over the slots
    if (occ < 1) return GNNSEG_ECUDA;
    const int cap = sms * occ;
    if (grid > cap) grid = cap;
    if (launch_pdl(edge_kernel<H, PAIRS, MINB>, grid, 256, 0, st, use_pdl(g->n_slots), blob, P, g->src, g->dst, g->in_pos, g->out_pos,
                   g->n_slots, e, e_in, e_out) != cudaSuccess)
        return GNNSEG_ECUDA;
    return check_launch();
}

int launch_node_tc32(const float* blob, const GnnsegGraph* g, const float* X4, const float* Q_in, const float* e_in,
                     const float* e_out, float* P_out, float* Q_out, int write_q, float* h1_save, float* H_save,
                     cudaStream_t st);   // gnnseg_node_tc.cu (tcgen05, gather fused)
int launch_node_mlp_tc32(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, float* P_out,
                         float* Q_out, int write_q, float* H_save, bool pdl, cudaStream_t st);   // gnnseg_node_tc.cu (tcgen05, MLP only)
int launch_node_mlp_tc64(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, float* P_out,
                         float* Q_out, int write_q, float* H_save, bool pdl, cudaStream_t st);   // hidden_dim = 64, weights streamed

template <int H>
static int launch_gather(const GnnsegGraph* g, const float* Q_in, const float* e_in, const float* e_out, float* h1_out,
                         int ld_out, float* h1_save, cudaStream_t st) {
    if (g->n_nodes == 0) return GNNSEG_OK;
    const int sms = sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    long long blocks = ((long long)g->n_nodes * (H / 4) + 255) / 256;
    static const int ranges = [] { const char* v = getenv("GNNSEG_GATHER_RANGES"); return v ? atoi(v) : 0; }();   // experiment switch
    if (ranges == 1) {            // contiguous ranges, 256-thread CTAs (8 per SM)
        if (blocks > sms * 8) blocks = sms * 8;
        node_gather_kernel<H, 3, 8, 256, true><<<(int)blocks, 256, 0, st>>>(*g, Q_in, e_in, e_out, h1_out, ld_out, h1_save);
        return check_launch();
    }
    if (ranges == 2) {            // contiguous ranges, 1024-thread CTAs (2 per SM)
        blocks = ((long long)g->n_nodes * (H / 4) + 1023) / 1024;
        if (blocks > sms * 2) blocks = sms * 2;
        node_gather_kernel<H, 3, 2, 1024, true><<<(int)blocks, 1024, 0, st>>>(*g, Q_in, e_in, e_out, h1_out, ld_out, h1_save);
        return check_launch();
    }
    if (blocks > sms * 8) blocks = sms * 8;                      // one resident wave of 64 warps per SM
    if (launch_pdl(node_gather_kernel<H>, (int)blocks, 256, 0, st, use_pdl(g->n_slots), *g, Q_in, e_in, e_out, h1_out, ld_out,
                   h1_save) != cudaSuccess)
        return GNNSEG_ECUDA;
    return check_launch();
}

template <int H>
static int launch_node(const float* blob, const GnnsegGraph* g, const float* X4, const float* Q_in, const float* e_in,
                       const float* e_out, float* P_out, float* Q_out, int write_q, float* h1_save, float* H_save,
                       cudaStream_t st) {
    using C = NodeCfg<H>;
    if (g->n_nodes == 0) return GNNSEG_OK;
    if (H == 32) {
        // "mma": generic mma.sync path, "fused": one tcgen05 kernel with the gather inside (for A/B runs);
        // default: gather kernel at full occupancy + tcgen05 MLP kernel, two CTAs per SM.  The h1 rows
        // travel through P_out's own rows (row n of P' is written after row n of h1 has been read).
        const char* impl = getenv("GNNSEG_NODE_IMPL");
        if (impl && impl[0] == 'f') return launch_node_tc32(blob, g, X4, Q_in, e_in, e_out, P_out, Q_out, write_q, h1_save, H_save, st);
        if (!impl || impl[0] != 'm') {
            const int rc = launch_gather<H>(g, Q_in, e_in, e_out, P_out, 2 * H, h1_save, st);
            if (rc) return rc;
            return launch_node_mlp_tc32(blob, X4, P_out, 2 * H, g->n_nodes, P_out, Q_out, write_q, H_save, use_pdl(g->n_slots), st);
        }
    }
    if (H == 64) {
        // gather kernel + tcgen05 MLP kernel with streamed weight images; "mma": the fused mma.sync kernel (A/B runs)
        const char* impl = getenv("GNNSEG_NODE_IMPL");
        if (!impl || impl[0] != 'm') {
            const int rc = launch_gather<H>(g, Q_in, e_in, e_out, P_out, 2 * H, h1_save, st);
            if (rc) return rc;
            return launch_node_mlp_tc64(blob, X4, P_out, 2 * H, g->n_nodes, P_out, Q_out, write_q, H_save, use_pdl(g->n_slots), st);
        }
    }
    const int n_tiles = (g->n_nodes + C::TN - 1) / C::TN;
    int grid = 0;
    const int rc = persistent_grid<node_kernel<H>>(C::NT, C::SMEM_BYTES, n_tiles, &grid);
    if (rc) return rc;
    if (launch_pdl(node_kernel<H>, grid, C::NT, C::SMEM_BYTES, st, use_pdl(g->n_slots), blob, *g, X4, Q_in, e_in, e_out, n_tiles, P_out,
                   Q_out, write_q, h1_save, H_save) != cudaSuccess)
        return GNNSEG_ECUDA;
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

int input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* X4, float* P, float* Q,
               float* H_save, cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_input<HH>(blob, X, n_nodes, F, X4, P, Q, H_save, st));
}
int edge_step(const float* blob, const GnnsegGraph* g, const float* P, int h, float* e, float* e_in, float* e_out,
              cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_edge<HH>(blob, g, P, e, e_in, e_out, st));
}
int node_step(const float* blob, const GnnsegGraph* g, const float* X4, const float* Q_in, const float* e_in,
              const float* e_out, int h, float* P_out, float* Q_out, int write_q, float* h1_save, float* H_save,
              cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_node<HH>(blob, g, X4, Q_in, e_in, e_out, P_out, Q_out, write_q, h1_save, H_save, st));
}
int node_gather_step(const GnnsegGraph* g, const float* Q_in, const float* e_in, const float* e_out, int h, float* h1_out,
                     int ld_out, cudaStream_t st) {
    GNNSEG_DISPATCH_H(h, launch_gather<HH>(g, Q_in, e_in, e_out, h1_out, ld_out, nullptr, st));
}
int node_mlp_step(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, int h, float* P_out,
                  float* Q_out, cudaStream_t st) {
    if (h == 64) return launch_node_mlp_tc64(blob, X4, h1, ld_h1, n_nodes, P_out, Q_out, Q_out != nullptr, nullptr, false, st);
    if (h != 32) return GNNSEG_EUNSUPPORTED;     // the other widths run the fused generic kernel (gnnseg_node_step)
    return launch_node_mlp_tc32(blob, X4, h1, ld_h1, n_nodes, P_out, Q_out, Q_out != nullptr, nullptr, false, st);
}
int pack_weights(const GnnsegParams* p, int F, int h, float* blob, cudaStream_t st) {
    const int total = blob_total(h);
    pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(*p, F, h, blob, nullptr, nullptr);
    pack_tc_images_kernel<<<(total + 255) / 256, 256, 0, st>>>(h, blob);
    return check_launch();
}
int pack_head_weights(const GnnsegParams* p, const float* w_out, const float* b_out, int F, int h, float* head_blob,
                      cudaStream_t st) {
    const int total = blob_total(h);
    pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(*p, F, h, head_blob, w_out, b_out);
    pack_tc_images_kernel<<<(total + 255) / 256, 256, 0, st>>>(h, head_blob);
    return check_launch();
}

// NodeClassifier output: score[n] = sigmoid(logit[n]); the head blob left the logit in column 0 of P.
__global__ void node_head_kernel(const float* __restrict__ P, const int ld, const int n_nodes, float* __restrict__ out) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x)
        out[n] = 1.f / (1.f + expf(-__ldg(P + (size_t)n * ld)));
}
int node_head(const float* P, int ld, int n_nodes, float* out, cudaStream_t st) {
    if (n_nodes == 0) return GNNSEG_OK;
    int grid = (n_nodes + 255) / 256;
    const int cap = cached_sm_count() * 8;
    if (grid > cap) grid = cap;
    node_head_kernel<<<grid, 256, 0, st>>>(P, ld, n_nodes, out);
    return check_launch();
}

}  // namespace gnnseg
