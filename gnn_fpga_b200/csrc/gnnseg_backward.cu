// Backward pass of the segment classifier for sm_100a: what loss.backward() computes in
// Estimator.training_step (gnn/estimator.py:49-60) through SegmentClassifier.forward
// (gnn/model.py:140-156), written against the projection-first dataflow of the forward kernels
// (gnnseg_common.cuh).  No atomics anywhere: every scatter of the chain rule is turned into a CSR
// gather, and every weight gradient is a per-CTA partial sum that a last kernel adds up in a
// fixed order, so two runs give bit-identical gradients.
//
// Forward, iteration t (HX_t = [H_t | X]):
//   P_t = [W1a.HX_t + b1 | W1b.HX_t]                  Q_t = [W3a.HX_t | W3b.HX_t | W3c.HX_t + b3]
//   u_j = tanh(Ps_t[src_j] + Pd_t[dst_j])             e_j = sigmoid(w2.u_j + b2)
//   g_n = Qs_t[n] + sum_{dst_j=n} e_j Qi_t[src_j] + sum_{src_j=n} e_j Qo_t[dst_j]
//   h1_n = tanh(g_n)                                  H_{t+1}[n] = tanh(W4.h1_n + b4)
// Backward, iteration t, from dg_t = dL/dg (n x h):
//   edge_bwd     (per slot)  de_j = dg[dst_j].Qi[src_j] + dg[src_j].Qo[dst_j];  ds_j = de_j e_j (1-e_j)
//                            dw2 += ds_j u_j;  db2 += ds_j;  ds_j written in both CSR orders
//   gather_bwd   (per node)  dPs[n] = sum_{src_j=n} ds_j w2*(1-u_j^2)     dQi[n] = sum_{src_j=n} e_j dg[dst_j]
//                            dPd[n] = sum_{dst_j=n} ds_j w2*(1-u_j^2)     dQo[n] = sum_{dst_j=n} e_j dg[src_j]
//                            dQs[n] = dg[n]                                 (u_j recomputed from P_t)
//   dense_bwd    (per node tile)  dWP += HX_t^T.[dP|dQ], dbias += colsum;  dH_t = [dP|dQ].WP^T;
//                            dz = dH_t*(1-H_t^2);  dW4 += dz^T.h1_{t-1};  db4 += colsum(dz);
//                            dg_{t-1} = (dz.W4)*(1-h1_{t-1}^2)            (t = 0: dWin, dbin instead)
// The last edge step (scores) enters the same kernels with ds_j = dL/dscore_j * p_j (1-p_j).
#include <cmath>
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {

// per-CTA partial sums of the edge backward kernel
template <int H>
struct EdgePart {
    static constexpr int DW2 = 0, DB1X = H, DB2 = 2 * H, SIZE = 2 * H + 4;
};
// per-CTA partial sums of the dense backward kernel
template <int H>
struct NodePart {
    static constexpr int D4 = H + 4;
    static constexpr int WP = 0;                  // [D4][5H]  d[W1a|W1b|W3a|W3b|W3c]^T
    static constexpr int BP = WP + D4 * 5 * H;    // [5H]      column sums of [dP|dQ]
    static constexpr int W4 = BP + 5 * H;         // [H][H]    dW4[o][k]
    static constexpr int B4 = W4 + H * H;         // [H]
    static constexpr int WIN = B4 + H;            // [4][H]    dWin^T
    static constexpr int BIN = WIN + 4 * H;       // [H]
    static constexpr int SIZE = BIN + H;
};
constexpr int EDGE_BWD_MAX_BLOCKS = 2048;
constexpr int DENSE_BWD_MAX_BLOCKS = 320;

size_t edge_part_floats(int h) { return (size_t)EDGE_BWD_MAX_BLOCKS * (2 * h + 4); }
size_t node_part_floats(int h) {
    return (size_t)DENSE_BWD_MAX_BLOCKS * ((h + 4) * 5 * h + 5 * h + h * h + h + 4 * h + h);
}

__device__ __forceinline__ float4 tanh4(const float4 a, const float4 b) {
    return make_float4(tanh_fast(a.x + b.x), tanh_fast(a.y + b.y), tanh_fast(a.z + b.z), tanh_fast(a.w + b.w));
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
// acc += s * w * (1 - u^2)
__device__ __forceinline__ void dtanh_acc(float4& acc, const float s, const float4 w, const float4 u) {
    acc.x = fmaf(s * w.x, fmaf(-u.x, u.x, 1.f), acc.x);
    acc.y = fmaf(s * w.y, fmaf(-u.y, u.y, 1.f), acc.y);
    acc.z = fmaf(s * w.z, fmaf(-u.z, u.z, 1.f), acc.z);
    acc.w = fmaf(s * w.w, fmaf(-u.w, u.w, 1.f), acc.w);
}

// ------------------------------------------------------------------------------------------
// edge backward: one slot per group of H/4 lanes.  FINAL: ds comes from dL/dscore; otherwise from
// the node step's dg.  Slots with an absent start carry b1 in place of Ps (forward), so their
// share of db1 is summed here (DB1X); the rest of db1 arrives through the column sum of dPs.
// ------------------------------------------------------------------------------------------
template <int H, bool FINAL>
__global__ void __launch_bounds__(256)
edge_bwd_kernel(const float* __restrict__ blob, const float* __restrict__ P, const float* __restrict__ Q,
                const float* __restrict__ dg, const float* __restrict__ dscore,
                const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                const int32_t* __restrict__ in_pos, const int32_t* __restrict__ out_pos, const int n_slots,
                float* __restrict__ ds_in, float* __restrict__ ds_out, float* __restrict__ part,
                const int accumulate) {
    using B = Blob<H>;
    using EP = EdgePart<H>;
    constexpr int G = H / 4, GPW = 32 / G;
    const int lane = threadIdx.x & 31, c = lane % G, warp_in_block = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const float4 w2 = ldg4(blob + B::W2 + 4 * c);
    const float4 b1 = ldg4(blob + B::BP + 4 * c);
    const float b2 = __ldg(blob + B::B2);
    float4 a_w2 = make_float4(0.f, 0.f, 0.f, 0.f), a_b1 = a_w2;
    float a_b2 = 0.f;

    for (int base = warp * GPW; base < n_slots; base += n_warps * GPW) {     // warp-uniform trip count
        const int j = base + lane / G;
        const bool valid = j < n_slots;
        int ss = -1, dd = -1;
        if (valid) { ss = __ldg(src + j); dd = __ldg(dst + j); }
        float4 a = b1, b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ss >= 0) a = ldg4(P + (size_t)ss * (2 * H) + 4 * c);
        if (dd >= 0) b = ldg4(P + (size_t)dd * (2 * H) + H + 4 * c);
        float de = 0.f;
        if (!FINAL && ss >= 0 && dd >= 0) {
            const float4 gd = ldg4(dg + (size_t)dd * H + 4 * c), qi = ldg4(Q + (size_t)ss * (3 * H) + 4 * c);
            const float4 gs = ldg4(dg + (size_t)ss * H + 4 * c), qo = ldg4(Q + (size_t)dd * (3 * H) + H + 4 * c);
            de = dot4(gd, qi) + dot4(gs, qo);
        }
        const float4 u = tanh4(a, b);
        float z = dot4(w2, u);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            z += __shfl_xor_sync(0xffffffffu, z, o);
            if (!FINAL) de += __shfl_xor_sync(0xffffffffu, de, o);
        }
        const float e = 1.f / (1.f + expf(-(z + b2)));
        float ds = 0.f;
        if (valid) ds = (FINAL ? __ldg(dscore + j) : de) * e * (1.f - e);
        fma4(a_w2, ds, u);
        if (c == 0) a_b2 += ds;
        if (valid && ss < 0) dtanh_acc(a_b1, ds, w2, u);
        if (valid && c == 0) {
            const int pi = __ldg(in_pos + j), po = __ldg(out_pos + j);
            if (pi >= 0) ds_in[pi] = ds;
            if (po >= 0) ds_out[po] = ds;
        }
    }
    // lanes holding the same chunk -> lanes 0..G-1; warps -> shared memory -> fixed-order sum
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) {
        a_w2.x += __shfl_xor_sync(0xffffffffu, a_w2.x, o); a_w2.y += __shfl_xor_sync(0xffffffffu, a_w2.y, o);
        a_w2.z += __shfl_xor_sync(0xffffffffu, a_w2.z, o); a_w2.w += __shfl_xor_sync(0xffffffffu, a_w2.w, o);
        a_b1.x += __shfl_xor_sync(0xffffffffu, a_b1.x, o); a_b1.y += __shfl_xor_sync(0xffffffffu, a_b1.y, o);
        a_b1.z += __shfl_xor_sync(0xffffffffu, a_b1.z, o); a_b1.w += __shfl_xor_sync(0xffffffffu, a_b1.w, o);
        a_b2 += __shfl_xor_sync(0xffffffffu, a_b2, o);
    }
    __shared__ float sRed[8][EP::SIZE];
    if (lane < G) {
        st4(&sRed[warp_in_block][EP::DW2 + 4 * lane], a_w2);
        st4(&sRed[warp_in_block][EP::DB1X + 4 * lane], a_b1);
        if (lane == 0) sRed[warp_in_block][EP::DB2] = a_b2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= EP::DB2; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += sRed[w][i];
        float* dstp = part + (size_t)blockIdx.x * EP::SIZE + i;
        *dstp = accumulate ? *dstp + s : s;
    }
}

// ------------------------------------------------------------------------------------------
// gather backward: the chain rule's scatters written as CSR gathers, one float4 chunk per thread.
// dproj row = [dPs | dPd] (FINAL) or [dPs | dPd | dQi | dQo | dQs].
// ------------------------------------------------------------------------------------------
template <int H, bool FINAL>
__global__ void __launch_bounds__(256)
gather_bwd_kernel(const float* __restrict__ blob, const GnnsegGraph g, const float* __restrict__ P,
                  const float* __restrict__ dg, const float* __restrict__ e_in, const float* __restrict__ e_out,
                  const float* __restrict__ ds_in, const float* __restrict__ ds_out, float* __restrict__ dproj) {
    using B = Blob<H>;
    constexpr int G = H / 4, NB = FINAL ? 2 : 5;
    const long long total = (long long)g.n_nodes * G;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / G), c = (int)(idx % G);
        const float4 w2 = ldg4(blob + B::W2 + 4 * c);
        const float4 b1 = ldg4(blob + B::BP + 4 * c);
        const float4 ps = ldg4(P + (size_t)n * (2 * H) + 4 * c), pd = ldg4(P + (size_t)n * (2 * H) + H + 4 * c);
        float4 dPs = make_float4(0.f, 0.f, 0.f, 0.f), dPd = dPs, dQi = dPs, dQo = dPs;
        const int o0 = __ldg(g.out_ptr + n), o1 = __ldg(g.out_ptr + n + 1);
        const int i0 = __ldg(g.in_ptr + n), i1 = __ldg(g.in_ptr + n + 1);
#pragma unroll 2
        for (int s = o0; s < o1; ++s) {                       // slots that start at n: neighbour = end node
            const int nb = __ldg(g.out_nbr + s);
            const float ds = __ldg(ds_out + s);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nb >= 0) b = ldg4(P + (size_t)nb * (2 * H) + H + 4 * c);
            dtanh_acc(dPs, ds, w2, tanh4(ps, b));
            if (!FINAL && nb >= 0) fma4(dQi, __ldg(e_out + s), ldg4(dg + (size_t)nb * H + 4 * c));
        }
#pragma unroll 2
        for (int s = i0; s < i1; ++s) {                       // slots that end at n: neighbour = start node
            const int nb = __ldg(g.in_nbr + s);
            const float ds = __ldg(ds_in + s);
            float4 a = b1;
            if (nb >= 0) a = ldg4(P + (size_t)nb * (2 * H) + 4 * c);
            dtanh_acc(dPd, ds, w2, tanh4(a, pd));
            if (!FINAL && nb >= 0) fma4(dQo, __ldg(e_in + s), ldg4(dg + (size_t)nb * H + 4 * c));
        }
        float* row = dproj + (size_t)n * (NB * H) + 4 * c;
        st4(row, dPs);
        st4(row + H, dPd);
        if (!FINAL) {
            st4(row + 2 * H, dQi);
            st4(row + 3 * H, dQo);
            st4(row + 4 * H, ldg4(dg + (size_t)n * H + 4 * c));
        }
    }
}

// ------------------------------------------------------------------------------------------
// dense backward
// ------------------------------------------------------------------------------------------
template <int H>
struct DenseCfg {
    static constexpr int TN  = (H >= 64) ? 32 : 64;        // nodes per tile
    static constexpr int NT  = (H >= 64) ? 512 : 256;
    static constexpr int MINB = (H >= 64) ? 1 : 2;         // CTAs per SM (shared memory: 174 KB at H = 64, <= 99 KB below)
    static constexpr int D4  = H + 4;
    static constexpr int LW  = 5 * H + 4;                  // row strides (floats): +4 keeps 8 consecutive rows on distinct banks
    static constexpr int LH  = H + 4;
    static constexpr int LX  = D4 + 4;
    static constexpr int KG  = H / 4;
    static constexpr int RN  = (TN * KG / NT) < 1 ? 1 : (TN * KG / NT);      // nodes per thread in the dot-product GEMMs
    static_assert(TN % RN == 0, "tile");
    static constexpr int SMEM_FLOATS = D4 * LW + H * LH + TN * (LW + LX + 2 * LH);
    static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * 4;
};

// acc[t][i][c] += sum_n A[n][4 kg + i] * Bm[n][4 og + c] over the tile; thread tile ids t = tid, tid+NT, ...
template <int KGN, int OGN, int TN, int NT, int MAXT>
__device__ __forceinline__ void outer_acc(const float* __restrict__ sA, const int lda, const float* __restrict__ sB,
                                          const int ldb, float (&acc)[MAXT][4][4]) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int id = threadIdx.x + t * NT;
        if (id < KGN * OGN) {
            const int og = id % OGN, kg = id / OGN;
            const float* a = sA + 4 * kg;
            const float* b = sB + 4 * og;
#pragma unroll 4
            for (int n = 0; n < TN; ++n) {
                const float4 av = lds4(a + n * lda), bv = lds4(b + n * ldb);
                const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[t][i][0] = fmaf(ar[i], bv.x, acc[t][i][0]);
                    acc[t][i][1] = fmaf(ar[i], bv.y, acc[t][i][1]);
                    acc[t][i][2] = fmaf(ar[i], bv.z, acc[t][i][2]);
                    acc[t][i][3] = fmaf(ar[i], bv.w, acc[t][i][3]);
                }
            }
        }
    }
}
// partial[(4 kg + i) * ld + 4 og + c] (+)= acc
template <int KGN, int OGN, int NT, int MAXT>
__device__ __forceinline__ void outer_store(float* __restrict__ dstp, const int ld, const float (&acc)[MAXT][4][4],
                                            const bool accumulate) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int id = threadIdx.x + t * NT;
        if (id < KGN * OGN) {
            const int og = id % OGN, kg = id / OGN;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float* p = dstp + (size_t)(4 * kg + i) * ld + 4 * og;
                float4 v = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
                if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(p); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                st4(p, v);
            }
        }
    }
}

// out(n, k) = sum_{o < OLEN} A[n][o] * W[k][o] for the tile; a thread owns RN nodes x 4 outputs
// (k = kg + i*KG: consecutive lanes read consecutive W rows, which the +4 row padding keeps on
// distinct banks; the A row of a quarter warp is a broadcast).  epi(n_local, k, value).
template <int OLEN, int KG, int TN, int NT, int RN, typename Epi>
__device__ __forceinline__ void dot_gemm(const float* __restrict__ sA, const int lda, const float* __restrict__ sW,
                                         const int ldw, Epi epi) {
    constexpr int NG = TN / RN, TILES = KG * NG;
    for (int t = threadIdx.x; t < TILES; t += NT) {
        const int kg = t % KG, ng = t / KG;
        float acc[RN][4];
#pragma unroll
        for (int r = 0; r < RN; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
#pragma unroll 2
        for (int o = 0; o < OLEN; o += 4) {
            float4 w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = lds4(sW + (kg + i * KG) * ldw + o);
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                const float4 a = lds4(sA + (ng + r * NG) * lda + o);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[r][i] += dot4(a, w[i]);
            }
        }
#pragma unroll
        for (int r = 0; r < RN; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) epi(ng + r * NG, kg + i * KG, acc[r][i]);
    }
}

// NB = 2: dproj rows are [dPs|dPd] (after the final edge step), 5: all five projections.
// FIRST: H_in is H_0 (input network: dWin, dbin); otherwise H_in = H_t, t > 0, produced by node
// step t-1 from h1_prev: dW4, db4 and dg_out = dL/dg_{t-1}.
// WGRAD = false (hidden_dim 32 / 64): the weight gradients are left to wgrad_tc_kernel (gnnseg_wgrad_tc.cu, tcgen05); this
// kernel then only propagates: dz -> dz_out (which that kernel reads), dg_out.
template <int H, int NB, bool FIRST, bool WGRAD = true>
__global__ void __launch_bounds__(DenseCfg<H>::NT, DenseCfg<H>::MINB)
dense_bwd_kernel(const float* __restrict__ blob, const float* __restrict__ dproj, const float* __restrict__ H_in,
                 const float* __restrict__ X4, const float* __restrict__ h1_prev, const int n_nodes,
                 const int n_tiles, float* __restrict__ dg_out, float* __restrict__ part, const int accumulate,
                 float* __restrict__ dz_out = nullptr) {
    using C = DenseCfg<H>;
    using B = Blob<H>;
    using NP = NodePart<H>;
    constexpr int TN = C::TN, NT = C::NT, D4 = C::D4, LW = C::LW, LH = C::LH, LX = C::LX, KG = C::KG, RN = C::RN;
    constexpr int NO = NB * H;                                        // live projection columns
    constexpr int T_WP = WGRAD ? ((D4 / 4) * (NO / 4) + NT - 1) / NT : 1;         // register tiles per thread
    constexpr int T_W4 = WGRAD ? (KG * KG + NT - 1) / NT : 1;
    constexpr int T_WIN = WGRAD ? (KG + NT - 1) / NT : 1;
    extern __shared__ __align__(16) float smem[];
    float* sWP = smem;                 // [D4][LW]   WP[k][o]
    float* sW4 = sWP + D4 * LW;        // [H][LH]    W4^T: [k][o]
    float* sDP = sW4 + H * LH;         // [TN][LW]   dproj tile
    float* sHX = sDP + TN * LW;        // [TN][LX]   [H_in | X]
    float* sDZ = sHX + TN * LX;        // [TN][LH]   dz
    float* sH1 = sDZ + TN * LH;        // [TN][LH]   h1_prev
    for (int i = threadIdx.x; i < D4 * (5 * H / 4); i += NT) {
        const int k = i / (5 * H / 4), o4 = i % (5 * H / 4);
        st4(sWP + k * LW + 4 * o4, ldg4(blob + B::WP + k * 5 * H + 4 * o4));
    }
    for (int i = threadIdx.x; i < H * KG; i += NT) {
        const int k = i / KG, o4 = i % KG;
        st4(sW4 + k * LH + 4 * o4, ldg4(blob + B::W4 + k * H + 4 * o4));
    }
    float aWP[T_WP][4][4], aW4[T_W4][4][4], aWin[T_WIN][4][4];
    float aBP = 0.f, aB = 0.f;
#pragma unroll
    for (int t = 0; t < T_WP; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) aWP[t][i][0] = aWP[t][i][1] = aWP[t][i][2] = aWP[t][i][3] = 0.f;
#pragma unroll
    for (int t = 0; t < T_W4; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) aW4[t][i][0] = aW4[t][i][1] = aW4[t][i][2] = aW4[t][i][3] = 0.f;
#pragma unroll
    for (int t = 0; t < T_WIN; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) aWin[t][i][0] = aWin[t][i][1] = aWin[t][i][2] = aWin[t][i][3] = 0.f;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int node0 = tile * TN;
        __syncthreads();                                   // weights visible / previous tile's readers done
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < TN * (NO / 4); i += NT) {
            const int ln = i / (NO / 4), o4 = i % (NO / 4), n = node0 + ln;
            st4(sDP + ln * LW + 4 * o4, n < n_nodes ? ldg4(dproj + (size_t)n * NO + 4 * o4) : zero);
        }
        for (int i = threadIdx.x; i < TN * (KG + 1); i += NT) {
            const int ln = i / (KG + 1), c = i % (KG + 1), n = node0 + ln;
            float4 v = zero;
            if (n < n_nodes) v = c < KG ? ldg4(H_in + (size_t)n * H + 4 * c) : ldg4(X4 + (size_t)n * 4);
            st4(sHX + ln * LX + 4 * c, v);
        }
        if (!FIRST) {
            for (int i = threadIdx.x; i < TN * KG; i += NT) {
                const int ln = i / KG, c = i % KG, n = node0 + ln;
                st4(sH1 + ln * LH + 4 * c, n < n_nodes ? ldg4(h1_prev + (size_t)n * H + 4 * c) : zero);
            }
        }
        __syncthreads();
        // weight gradients of the projections and their bias
        if (WGRAD) outer_acc<D4 / 4, NO / 4, TN, NT, T_WP>(sHX, LX, sDP, LW, aWP);
        if (WGRAD && threadIdx.x < NO) {
#pragma unroll 8
            for (int n = 0; n < TN; ++n) aBP += sDP[n * LW + threadIdx.x];
        }
        // dH = dproj . WP^T (hidden columns only: X has no gradient), dz = dH * (1 - H^2)
        dot_gemm<NO, KG, TN, NT, RN>(sDP, LW, sWP, LW, [&](const int ln, const int k, const float v) {
            const float hv = sHX[ln * LX + k];
            sDZ[ln * LH + k] = v * fmaf(-hv, hv, 1.f);
        });
        __syncthreads();
        if (!WGRAD) {
            for (int i = threadIdx.x; i < TN * KG; i += NT) {
                const int ln = i / KG, c = i % KG, n = node0 + ln;
                if (n < n_nodes) st4(dz_out + (size_t)n * H + 4 * c, lds4(sDZ + ln * LH + 4 * c));
            }
        }
        if (WGRAD && threadIdx.x < H) {
#pragma unroll 8
            for (int n = 0; n < TN; ++n) aB += sDZ[n * LH + threadIdx.x];
        }
        if (FIRST) {
            // H_0 = tanh(Win.X + bin): dWin^T[f][o] += X[n][f] dz[n][o]
            if (WGRAD) outer_acc<1, KG, TN, NT, T_WIN>(sHX + H, LX, sDZ, LH, aWin);
        } else {
            // H_t = tanh(W4.h1 + b4): dW4[o][k] += dz[n][o] h1[n][k];  dg = (dz . W4) * (1 - h1^2)
            if (WGRAD) outer_acc<KG, KG, TN, NT, T_W4>(sDZ, LH, sH1, LH, aW4);
            dot_gemm<H, KG, TN, NT, RN>(sDZ, LH, sW4, LH, [&](const int ln, const int k, const float v) {
                const int n = node0 + ln;
                const float h1v = sH1[ln * LH + k];
                if (n < n_nodes) dg_out[(size_t)n * H + k] = v * fmaf(-h1v, h1v, 1.f);
            });
        }
    }
    if (!WGRAD) return;
    float* mine = part + (size_t)blockIdx.x * NP::SIZE;
    outer_store<D4 / 4, NO / 4, NT, T_WP>(mine + NP::WP, 5 * H, aWP, accumulate != 0);
    if (threadIdx.x < NO) mine[NP::BP + threadIdx.x] = accumulate ? mine[NP::BP + threadIdx.x] + aBP : aBP;
    if (FIRST) {
        outer_store<1, KG, NT, T_WIN>(mine + NP::WIN, H, aWin, false);
        if (threadIdx.x < H) mine[NP::BIN + threadIdx.x] = aB;
    } else {
        outer_store<KG, KG, NT, T_W4>(mine + NP::W4, H, aW4, accumulate > 1);
        if (threadIdx.x < H) mine[NP::B4 + threadIdx.x] = accumulate > 1 ? mine[NP::B4 + threadIdx.x] + aB : aB;
    }
}

__global__ void zero_kernel(float* __restrict__ p, const size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// ------------------------------------------------------------------------------------------
// last step: fixed-order sum of the per-CTA partials, reference layouts, masks
// (MaskedLinear: y = F.linear(x, W*mask) => dW = dW_eff * mask, gnn/model.py:28-31)
// ------------------------------------------------------------------------------------------
struct GradOut {
    float* w_in; float* b_in; float* w_e1; float* b_e1; float* w_e2; float* b_e2;
    float* w_n1; float* b_n1; float* w_n2; float* b_n2;
    const float* m_e1; const float* m_e2; const float* m_n1; const float* m_n2;
};

template <int H>
__global__ void finalize_grads_kernel(const float* __restrict__ partE, const int nE, const float* __restrict__ partN,
                                      const int nN, const int F, const int has_node, const GradOut go) {
    using EP = EdgePart<H>;
    using NP = NodePart<H>;
    const int D = F + H;
    const int n_win = H * F, n_e1 = H * 2 * D, n_n1 = H * 3 * D, n_n2 = H * H;
    const int o_bin = n_win, o_e1 = o_bin + H, o_be1 = o_e1 + n_e1, o_e2 = o_be1 + H, o_be2 = o_e2 + H,
              o_n1 = o_be2 + 1, o_bn1 = o_n1 + n_n1, o_n2 = o_bn1 + H, o_bn2 = o_n2 + n_n2, total = o_bn2 + H;
    // one warp per gradient element: lane l adds CTAs l, l+32, ... in order, then a butterfly
    const int lane = threadIdx.x & 31;
    auto wsum = [&](float s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        return s;
    };
    auto sumN = [&](const int off) { float s = 0.f; for (int b = lane; b < nN; b += 32) s += partN[(size_t)b * NP::SIZE + off]; return wsum(s); };
    auto sumE = [&](const int off) { float s = 0.f; for (int b = lane; b < nE; b += 32) s += partE[(size_t)b * EP::SIZE + off]; return wsum(s); };
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += (gridDim.x * blockDim.x) >> 5) {
        if (i < o_bin) {                                   // input_network.0.weight (h, F)
            const int j = i / F, f = i % F;
            go.w_in[i] = sumN(NP::WIN + f * H + j);
        } else if (i < o_e1) {
            go.b_in[i - o_bin] = sumN(NP::BIN + (i - o_bin));
        } else if (i < o_be1) {                            // edge_network.network.0.weight (h, 2D)
            const int r = i - o_e1, j = r / (2 * D), cc = r % (2 * D), blk = cc / D, k = cc % D;
            float v = sumN(NP::WP + k * 5 * H + blk * H + j);
            if (go.m_e1) v *= go.m_e1[r];
            go.w_e1[r] = v;
        } else if (i < o_e2) {
            const int j = i - o_be1;
            go.b_e1[j] = sumN(NP::BP + j) + sumE(EP::DB1X + j);
        } else if (i < o_be2) {                            // edge_network.network.2.weight (1, h)
            const int j = i - o_e2;
            float v = sumE(EP::DW2 + j);
            if (go.m_e2) v *= go.m_e2[j];
            go.w_e2[j] = v;
        } else if (i < o_n1) {
            go.b_e2[0] = sumE(EP::DB2);
        } else if (i < o_bn1) {                            // node_network.network.0.weight (h, 3D)
            const int r = i - o_n1, j = r / (3 * D), cc = r % (3 * D), blk = cc / D, k = cc % D;
            float v = has_node ? sumN(NP::WP + k * 5 * H + (2 + blk) * H + j) : 0.f;
            if (go.m_n1) v *= go.m_n1[r];
            go.w_n1[r] = v;
        } else if (i < o_n2) {
            const int j = i - o_bn1;
            go.b_n1[j] = has_node ? sumN(NP::BP + 4 * H + j) : 0.f;
        } else if (i < o_bn2) {                            // node_network.network.2.weight (h, h)
            const int r = i - o_n2;
            float v = has_node ? sumN(NP::W4 + r) : 0.f;
            if (go.m_n2) v *= go.m_n2[r];
            go.w_n2[r] = v;
        } else {
            go.b_n2[i - o_bn2] = has_node ? sumN(NP::B4 + (i - o_bn2)) : 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------
// NodeClassifier head (gnn/MPNN_HitClassifier.ipynb c21: sigmoid(Wo.[H_T | X] + bo) per node).  The
// forward leaves the logit in column 0 of P_T (pack_weights_kernel, head blob), so its backward is
// the dense kernel with NB = 2 on dproj = [dlogit | 0 ...]: dWo lands in column 0 of the WP partial,
// dbo in BP[0].  head_grad_kernel sums them over the CTAs in a fixed order and clears the slots,
// which the launches of the iterations below use for the edge network's first layer.
// ------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ P_T, const float* __restrict__ dnode, const int n_nodes,
                float* __restrict__ dproj) {
    constexpr int C2 = 2 * H / 4;
    const long long total = (long long)n_nodes * C2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / C2), c = (int)(i % C2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c == 0) {
            const float sg = 1.f / (1.f + expf(-__ldg(P_T + (size_t)n * 2 * H)));
            v.x = __ldg(dnode + n) * sg * (1.f - sg);
        }
        st4(dproj + (size_t)n * 2 * H + 4 * c, v);
    }
}

template <int H>
__global__ void __launch_bounds__(128)
head_grad_kernel(float* __restrict__ partN, const int nN, const int D, float* __restrict__ g_w, float* __restrict__ g_b) {
    using NP = NodePart<H>;
    const int k = threadIdx.x;                        // k < D: dWo[k]; k == D: dbo
    if (k > D) return;
    const int slot = k < D ? NP::WP + k * 5 * H : NP::BP;
    float s = 0.f;
    for (int c = 0; c < nN; ++c) {
        float* q = partN + (size_t)c * NP::SIZE + slot;
        s += *q;
        *q = 0.f;
    }
    if (k < D) g_w[k] = s; else g_b[0] = s;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int bwd_sm_count() { return cached_sm_count(); }
static inline int bwd_check() { return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA; }

struct TrainState {          // device arrays saved by the training forward + backward scratch
    const float* x4;
    const float* const* Hs;      // [T+1] n x h
    const float* const* h1s;     // [T]   n x h
    const float* const* Ps;      // [T+1] n x 2h
    const float* const* Qs;      // [T]   n x 3h
    const float* const* e_in;    // [T]
    const float* const* e_out;   // [T]
    float* dg;                   // n x h
    float* dproj;                // n x 5h
    float* ds_in;
    float* ds_out;
    float* partE;
    float* partN;
    float* dz;                   // n x h (tensor-core weight-gradient path)
};

// gnnseg_wgrad_tc.cu
bool wgrad_tc_width(int h);
int wgrad_tc_grid(int n_nodes, int h, int sms);
int wgrad_tc(int h, int nb, bool first, const float* dproj, const float* H_in, const float* X4, const float* dz, const float* h1_prev,
             int n_nodes, float* part, int accumulate, int grid, cudaStream_t st);
// gnnseg_dprop_tc.cu
bool dprop_tc_width(int h);
int dprop_tc(int h, int nb, bool first, const float* blob, const float* dproj, const float* H_in, const float* h1_prev, int n_nodes,
             float* dz_out, float* dg_out, int sms, cudaStream_t st);
// GNNSEG_DENSE_BWD (A/B runs): simt = the all-SIMT dense backward kernel at every width; w = tcgen05 weight gradients, the
// propagating half on CUDA cores; default = both halves on tcgen05 where compiled (hidden_dim 32; 64: weight gradients)
static int dense_bwd_mode() {
    static const int mode = [] { const char* v = getenv("GNNSEG_DENSE_BWD"); return !v ? 2 : v[0] == 's' ? 0 : v[0] == 'w' ? 1 : 2; }();
    return mode;
}
static bool wgrad_tc_enabled() { return dense_bwd_mode() >= 1; }

template <int H>
static int backward_impl(const float* blob, const GnnsegGraph* g, const int F, const int T, const float* dscores,
                         const TrainState& s, const GradOut& go, const float* head_blob, float* g_wout, float* g_bout,
                         cudaStream_t st) {
    using C = DenseCfg<H>;
    const int sms = bwd_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const int n = g->n_nodes, m = g->n_slots;
    constexpr int G = H / 4;
    // grids are functions of (n, m, SM count) only: every launch of a kind uses the same grid, so a
    // CTA accumulates into its own partial across launches in a fixed order
    int gridE = (int)(((long long)m * G + 255) / 256);
    if (gridE > sms * 8) gridE = sms * 8;
    if (gridE > EDGE_BWD_MAX_BLOCKS) gridE = EDGE_BWD_MAX_BLOCKS;
    if (gridE < 1) gridE = 1;
    int gridG = (int)(((long long)n * G + 255) / 256);
    if (gridG > sms * 16) gridG = sms * 16;
    if (gridG < 1) gridG = 1;
    const int n_tiles = (n + C::TN - 1) / C::TN;
    int gridD = n_tiles < sms * C::MINB ? n_tiles : sms * C::MINB;
    if (gridD > DENSE_BWD_MAX_BLOCKS) gridD = DENSE_BWD_MAX_BLOCKS;
    if (gridD < 1) gridD = 1;
    // hidden_dim 32 / 64: the propagating half of the dense step on CUDA cores (dz, dg), the weight gradients on tcgen05
    constexpr bool TCW = (H == 32 || H == 64);
    const bool tcw = TCW && wgrad_tc_enabled();
    const int gridW = tcw ? wgrad_tc_grid(n, H, sms) : 0;
    const int n_part = tcw ? gridW : gridD;                  // CTAs that own a slot of partN
    if (tcw && gridW > DENSE_BWD_MAX_BLOCKS) return GNNSEG_EINVAL;
    auto kF2 = dense_bwd_kernel<H, 2, true, !TCW>, kF2s = dense_bwd_kernel<H, 2, true, true>;
    auto kI2 = dense_bwd_kernel<H, 2, false, !TCW>, kI2s = dense_bwd_kernel<H, 2, false, true>;
    auto kF5 = dense_bwd_kernel<H, 5, true, !TCW>, kF5s = dense_bwd_kernel<H, 5, true, true>;
    auto kI5 = dense_bwd_kernel<H, 5, false, !TCW>, kI5s = dense_bwd_kernel<H, 5, false, true>;
    if (!ensure_dynamic_smem<dense_bwd_kernel<H, 2, true, true>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 2, false, true>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 5, true, true>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 5, false, true>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 2, true, !TCW>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 2, false, !TCW>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 5, true, !TCW>>((int)C::SMEM_BYTES) ||
        !ensure_dynamic_smem<dense_bwd_kernel<H, 5, false, !TCW>>((int)C::SMEM_BYTES))
        return GNNSEG_ECUDA;
    int rc_w = GNNSEG_OK;
    // one dense step: nb = 2 / 5 projections live in dproj, first = the input step
    auto dense = [&](const float* wblob, const int nb, const bool first, const float* Hin, const float* h1p, float* dg_out, const int acc) {
        if (tcw && dprop_tc_width(H) && dense_bwd_mode() >= 2) {
            int rc = dprop_tc(H, nb, first, wblob, s.dproj, Hin, h1p, n, s.dz, dg_out, sms, st);
            if (rc == GNNSEG_OK) rc = wgrad_tc(H, nb, first, s.dproj, Hin, s.x4, s.dz, h1p, n, s.partN, acc, gridW, st);
            if (rc != GNNSEG_OK) rc_w = rc;
        } else if (tcw) {
            if (nb == 2) {
                if (first) kF2<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, nullptr, n, n_tiles, nullptr, s.partN, acc, s.dz);
                else kI2<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, h1p, n, n_tiles, dg_out, s.partN, acc, s.dz);
            } else {
                if (first) kF5<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, nullptr, n, n_tiles, nullptr, s.partN, acc, s.dz);
                else kI5<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, h1p, n, n_tiles, dg_out, s.partN, acc, s.dz);
            }
            const int rc = wgrad_tc(H, nb, first, s.dproj, Hin, s.x4, s.dz, h1p, n, s.partN, acc, gridW, st);
            if (rc != GNNSEG_OK) rc_w = rc;
        } else {
            if (nb == 2) {
                if (first) kF2s<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, nullptr, n, n_tiles, nullptr, s.partN, acc, nullptr);
                else kI2s<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, h1p, n, n_tiles, dg_out, s.partN, acc, nullptr);
            } else {
                if (first) kF5s<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, nullptr, n, n_tiles, nullptr, s.partN, acc, nullptr);
                else kI5s<<<gridD, C::NT, C::SMEM_BYTES, st>>>(wblob, s.dproj, Hin, s.x4, h1p, n, n_tiles, dg_out, s.partN, acc, nullptr);
            }
        }
    };
    // the dense kernel's partial holds slots (W4/B4 or WIN/BIN) that some launches never write
    zero_kernel<<<64, 256, 0, st>>>(s.partN, (size_t)n_part * NodePart<H>::SIZE);

    const bool nodes = head_blob != nullptr;     // NodeClassifier: dscores is per node, no final edge step
    if (nodes) {
        zero_kernel<<<64, 256, 0, st>>>(s.partE, (size_t)gridE * EdgePart<H>::SIZE);
        if (n > 0) {
            head_bwd_kernel<H><<<gridG, 256, 0, st>>>(s.Ps[T], dscores, n, s.dproj);
            if (T == 0) dense(head_blob, 2, true, s.Hs[0], nullptr, nullptr, 0);
            else dense(head_blob, 2, false, s.Hs[T], s.h1s[T - 1], s.dg, 0);
        }
        head_grad_kernel<H><<<1, 128, 0, st>>>(s.partN, n > 0 ? n_part : 0, F + H, g_wout, g_bout);
    } else {
        // ---- final edge step: scores = edge(P_T) ------------------------------------------------
        if (m > 0)
            edge_bwd_kernel<H, true><<<gridE, 256, 0, st>>>(blob, s.Ps[T], nullptr, nullptr, dscores, g->src, g->dst,
                                                             g->in_pos, g->out_pos, m, s.ds_in, s.ds_out, s.partE, 0);
        else
            zero_kernel<<<1, 256, 0, st>>>(s.partE, (size_t)gridE * EdgePart<H>::SIZE);
        if (n > 0) {
            gather_bwd_kernel<H, true><<<gridG, 256, 0, st>>>(blob, *g, s.Ps[T], nullptr, nullptr, nullptr, s.ds_in,
                                                               s.ds_out, s.dproj);
            if (T == 0) dense(blob, 2, true, s.Hs[0], nullptr, nullptr, 0);
            else dense(blob, 2, false, s.Hs[T], s.h1s[T - 1], s.dg, 0);
        }
    }
    // ---- iterations T-1 .. 0 ------------------------------------------------------------------
    for (int t = T - 1; t >= 0 && n > 0; --t) {
        if (m > 0)
            edge_bwd_kernel<H, false><<<gridE, 256, 0, st>>>(blob, s.Ps[t], s.Qs[t], s.dg, nullptr, g->src, g->dst,
                                                              g->in_pos, g->out_pos, m, s.ds_in, s.ds_out, s.partE, 1);   // accumulates (final step or the zeroing above came first)
        gather_bwd_kernel<H, false><<<gridG, 256, 0, st>>>(blob, *g, s.Ps[t], s.dg, s.e_in[t], s.e_out[t], s.ds_in,
                                                            s.ds_out, s.dproj);
        // accumulate: 1 = projections only (the W4 slots are first written by launch T-1 ... ), 2 = W4 too
        if (t == 0) dense(blob, 5, true, s.Hs[0], nullptr, nullptr, 1);
        else dense(blob, 5, false, s.Hs[t], s.h1s[t - 1], s.dg, 2);
    }
    finalize_grads_kernel<H><<<sms * 4, 256, 0, st>>>(s.partE, gridE, s.partN, n > 0 ? n_part : 0, F, T > 0, go);
    if (rc_w != GNNSEG_OK) return rc_w;
    return bwd_check();
}

int backward(const float* blob, const GnnsegGraph* g, int F, int h, int T, const float* dscores,
             const TrainState& s, const GradOut& go, const float* head_blob, float* g_wout, float* g_bout,
             cudaStream_t st) {
    switch (h) {
        case 4:  return backward_impl<4>(blob, g, F, T, dscores, s, go, head_blob, g_wout, g_bout, st);
        case 8:  return backward_impl<8>(blob, g, F, T, dscores, s, go, head_blob, g_wout, g_bout, st);
        case 16: return backward_impl<16>(blob, g, F, T, dscores, s, go, head_blob, g_wout, g_bout, st);
        case 32: return backward_impl<32>(blob, g, F, T, dscores, s, go, head_blob, g_wout, g_bout, st);
        case 64: return backward_impl<64>(blob, g, F, T, dscores, s, go, head_blob, g_wout, g_bout, st);
        default: return GNNSEG_EUNSUPPORTED;
    }
}

// ------------------------------------------------------------------------------------------
// loss and optimiser of Estimator.training_step (gnn/estimator.py:49-60): nn.BCELoss() (mean over
// every slot of the padded (B, E_max) batch), the L1 penalty over the edge / node network weights,
// torch.optim.Adam.  Two-stage fixed-order reductions: bit-identical from run to run.
// ------------------------------------------------------------------------------------------
constexpr int LOSS_BLOCKS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
    return s;       // valid in thread 0
}

// torch.nn.functional.binary_cross_entropy: log terms clamped at -100; backward
// (p - y) / max(p (1 - p), 1e-12), scaled by 1/n for the mean.
__global__ void __launch_bounds__(256)
bce_partial_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ w,
                   const int n, float* __restrict__ dp, float* __restrict__ part) {
    __shared__ float sred[8];
    const float inv_n = 1.f / (float)n;
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float pi = __ldg(p + i), yi = __ldg(y + i), wi = w ? __ldg(w + i) : 1.f;
        const float l = -(yi * fmaxf(logf(pi), -100.f) + (1.f - yi) * fmaxf(log1pf(-pi), -100.f));
        acc += wi * l;
        if (dp) dp[i] = wi * (pi - yi) / fmaxf(pi * (1.f - pi), 1e-12f) * inv_n;
    }
    const float s = block_sum_256(acc, sred);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256)
bce_final_kernel(const float* __restrict__ part, const int n_part, const int n, float* __restrict__ loss) {
    __shared__ float sred[8];
    const float v = (int)threadIdx.x < n_part ? part[threadIdx.x] : 0.f;
    const float s = block_sum_256(v, sred);
    if (threadIdx.x == 0) *loss = n > 0 ? s / (float)n : 0.f;
}

int bce_loss(const float* scores, const float* targets, const float* weights, int n, float* loss, float* dscores,
             float* ws, cudaStream_t st) {
    int grid = (n + 255) / 256;
    if (grid > LOSS_BLOCKS) grid = LOSS_BLOCKS;
    if (grid < 1) grid = 1;
    bce_partial_kernel<<<grid, 256, 0, st>>>(scores, targets, weights, n, dscores, ws);
    bce_final_kernel<<<1, 256, 0, st>>>(ws, grid, n, loss);
    return bwd_check();
}

// loss += l1 * sum |W| and grad += l1 * sign(W) over the four weight matrices that
// gnn/estimator.py:54-56 walks (node_network.network[0,2], edge_network.network[0,2]; raw
// weights, not weight*mask).  One CTA: 26 k elements at most.
__global__ void __launch_bounds__(256)
l1_kernel(const float* w0, const int n0, const float* w1, const int n1, const float* w2, const int n2,
          const float* w3, const int n3, float* g0, float* g1, float* g2, float* g3, const float l1,
          float* __restrict__ loss) {
    __shared__ float sred[8];
    const float* ws[4] = {w0, w1, w2, w3};
    float* gs[4] = {g0, g1, g2, g3};
    const int ns[4] = {n0, n1, n2, n3};
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
        for (int i = threadIdx.x; i < ns[t]; i += blockDim.x) {
            const float v = ws[t][i];
            acc += fabsf(v);
            if (gs[t]) gs[t][i] += l1 * (v > 0.f ? 1.f : v < 0.f ? -1.f : 0.f);
        }
    const float s = block_sum_256(acc, sred);
    if (threadIdx.x == 0 && loss) *loss += l1 * s;
}

int l1_penalty(const GnnsegParams* p, int F, int h, float l1, float* loss, const GnnsegGrads* g, cudaStream_t st) {
    const int D = F + h;
    l1_kernel<<<1, 256, 0, st>>>(p->w_n1, h * 3 * D, p->w_n2, h * h, p->w_e1, h * 2 * D, p->w_e2, h,
                                 g ? g->w_n1 : nullptr, g ? g->w_n2 : nullptr, g ? g->w_e1 : nullptr,
                                 g ? g->w_e2 : nullptr, l1, loss);
    return bwd_check();
}

// torch.optim.Adam (single tensor, amsgrad off): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2 v + (1-b2) g g;
// denom = sqrt(v) / sqrt(1 - b2^t) + eps; p -= lr / (1 - b1^t) * m / denom.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const int n, const float step_size, const float bc2_sqrt,
                            const float beta1, const float beta2, const float eps, const float wd) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float gi = g[i];
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (1.f - beta1) * (gi - m[i]);
        const float vi = fmaf(1.f - beta2, gi * gi, beta2 * v[i]);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
}

int adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n, int step, float lr,
              float beta1, float beta2, float eps, float weight_decay, cudaStream_t st) {
    if (n == 0) return GNNSEG_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    int grid = (n + 255) / 256;
    if (grid > 1024) grid = 1024;
    adam_kernel<<<grid, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, (float)((double)lr / bc1), (float)sqrt(bc2),
                                      beta1, beta2, eps, weight_decay);
    return bwd_check();
}

}  // namespace gnnseg
