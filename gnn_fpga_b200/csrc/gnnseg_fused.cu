// Fused inference path (hidden_dim 32 / 64): the edge step evaluated inside the node step's CSR walk.
//
// Reference: EdgeNetwork.forward + NodeNetwork.forward, gnn/model.py:69-81,113-125, in the projection-first
// form of gnnseg_common.cuh.  Per node n the state row is
//     S[n] = [ SPs (H) | Qi (H) | SPd (H) | Qo (H) | Qs (H) ]        5H floats, 128-byte aligned
// with SPs = 2^(log2e * Ps), SPd = 2^(log2e * Pd): the exponentials of the edge network's first-layer
// projections, taken ONCE per node by the kernel that produced them.  For an edge s -> d
//     tanh(Ps[s] + Pd[d]) = 1 - 2 / ((SPs[s] * SPd[d])^2 + 1)
// costs one reciprocal per hidden unit instead of an exponential and a reciprocal (the MUFU pipe is what
// bounds this kernel), and the row a visit fetches serves both halves of the work:
//     in-edge  s -> n :  row S[s][0 : 2H]   = [SPs[s] | Qi[s]]   e = sigmoid(w2 . tanh(Ps[s] + Pd[n]) + b2),  acc += e * Qi[s]
//     out-edge n -> d :  row S[d][2H : 4H]  = [SPd[d] | Qo[d]]   e = sigmoid(w2 . tanh(Ps[n] + Pd[d]) + b2),  acc += e * Qo[d]
//     h1[n] = tanh(Qs[n] + acc)          own term first, then in-edges, then out-edges, ascending slot order
// Every edge is evaluated twice (once from each end) and nothing is written per edge: no e_in / e_out arrays,
// no inverse maps, no separate edge launch, two 256-byte (H = 32) row requests per edge and iteration
// instead of four 128-byte ones.  No atomics; the order of every sum is fixed => bit-reproducible.
//
// Range of the exponentials: log2e * P is clamped to [-63, 63] (|P| <= 43.6) so that the product of two
// stays a normal fp32 number; beyond, the squared product saturates to inf / 0 and tanh to +-1 exactly as
// it should, EXCEPT when two clamped values of opposite sign would have cancelled.  The producing kernel
// raises a flag word when it clamps; the host side turns that into an error (gnnseg_forward's `flags`).
//
// Work distribution: the warps of the whole grid take consecutive groups of nodes (grid stride), so that at
// any moment all SMs work on the same window of the batch and the rows they gather are L2 hits.  (Measured
// on the plain gather kernel: one contiguous node range per CTA, with or without a locality renumbering of
// the nodes, is 20 - 35 % SLOWER, profiles/r2/locality_experiment.txt; RANGES keeps that variant for A/B runs.)
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float rcp_approx(const float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(const float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// sigmoid(z) = 1 / (1 + 2^(-z log2e)), Newton-refined reciprocal (one evaluation per edge, not per hidden unit)
__device__ __forceinline__ float sigmoid_fast(const float z) {
    const float d = 1.f + ex2_approx(fminf(-LOG2E * z, 126.f));   // d stays finite: the Newton step would turn inf * 0 into NaN
    float r = rcp_approx(d);
    r = fmaf(r, fmaf(-d, r, 1.f), r);
    return z != z ? z : r;                                        // fminf drops a NaN: put it back
}

// partial sum over this lane's four hidden units of  -2 w2_k / ((a_k * own_k)^2 + 1)
__device__ __forceinline__ float edge_partial(const float4 a, const float4 own, const float4 w2n) {
    const float px = a.x * own.x, py = a.y * own.y, pz = a.z * own.z, pw = a.w * own.w;
    float z = w2n.x * rcp_approx(fmaf(px, px, 1.f));
    z = fmaf(w2n.y, rcp_approx(fmaf(py, py, 1.f)), z);
    z = fmaf(w2n.z, rcp_approx(fmaf(pz, pz, 1.f)), z);
    z = fmaf(w2n.w, rcp_approx(fmaf(pw, pw, 1.f)), z);
    return z;
}

// Four partial sums z[0..3] on each of the G lanes of a node -> the lane's own total: afterwards lane c holds
// the full sum of edge u(c) = 2 * bit(c, log2 G - 1) + bit(c, log2 G - 2)   (2 + 1 + log2(G / 4) shuffles
// instead of 4 log2 G).  Lanes that agree in those two bits hold the same value.
template <int G>
__device__ __forceinline__ float reduce4_transposed(const float (&z)[4], const int c) {
    constexpr unsigned FULL = 0xffffffffu;
    const bool up1 = (c & (G / 2)) != 0;
    const float y0 = (up1 ? z[2] : z[0]) + __shfl_xor_sync(FULL, up1 ? z[0] : z[2], G / 2);
    const float y1 = (up1 ? z[3] : z[1]) + __shfl_xor_sync(FULL, up1 ? z[1] : z[3], G / 2);
    const bool up2 = (c & (G / 4)) != 0;
    float w = (up2 ? y1 : y0) + __shfl_xor_sync(FULL, up2 ? y0 : y1, G / 4);
#pragma unroll
    for (int o = G / 8; o > 0; o >>= 1) w += __shfl_xor_sync(FULL, w, o);
    return w;
}
template <int G>
__device__ __forceinline__ int edge_of_lane(const int c) { return ((c & (G / 2)) ? 2 : 0) + ((c & (G / 4)) ? 1 : 0); }
template <int G>
__device__ __forceinline__ int lane_of_edge(const int u) { return ((u & 2) ? G / 2 : 0) + ((u & 1) ? G / 4 : 0); }

// This CTA's contiguous node range [lo, hi): the nodes are cut where the cumulative weight
// w(n) = ptr[n] + 2 n (entries to visit plus a constant per node) crosses multiples of total / gridDim.x.
__device__ __forceinline__ void cta_node_range(const int32_t* __restrict__ ptr, const int n_nodes, int* s_range) {
    if (threadIdx.x < 2) {
        const long long total = (long long)__ldg(ptr + n_nodes) + 2LL * n_nodes;
        const long long target = total * (blockIdx.x + threadIdx.x) / gridDim.x;
        int lo = 0, hi = n_nodes;                  // first n with w(n) >= target
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((long long)__ldg(ptr + mid) + 2LL * mid >= target) hi = mid; else lo = mid + 1;
        }
        s_range[threadIdx.x] = (blockIdx.x + threadIdx.x == gridDim.x) ? n_nodes : lo;
    }
    __syncthreads();
}

constexpr int ADJ_OUT = (int)0x80000000;      // adjacency entry: bit 31 = out-edge, bits 0..30 = neighbour node
constexpr int ADJ_NONE = 0x7fffffff;          // half edge (absent neighbour): contributes nothing

template <int H, int NTHR, int MINB, bool RANGES>
__global__ void __launch_bounds__(NTHR, MINB)
fused_gather_kernel(const float* __restrict__ blob, const float* __restrict__ S, const int32_t* __restrict__ adj_ptr,
                    const int32_t* __restrict__ adj, const int n_nodes, float* __restrict__ h1_out, const int ld_out) {
    using B = Blob<H>;
    constexpr int G = H / 4, NPW = 32 / G, LD = 5 * H;
    __shared__ int s_range[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, g = lane / G;
    pdl_launch_dependents();
    int lo = 0, hi = n_nodes, first = (blockIdx.x * (NTHR / 32) + warp) * NPW, step = gridDim.x * (NTHR / 32) * NPW;
    if (RANGES) {
        cta_node_range(adj_ptr, n_nodes, s_range);
        lo = s_range[0]; hi = s_range[1];
        first = lo + warp * NPW; step = (NTHR / 32) * NPW;
    }
    const float4 w2n = ldg4(blob + B::W2N + 4 * c);
    const float z0 = __ldg(blob + B::Z0);
    const int grp_lane0 = lane - c;
    pdl_wait();                                    // S comes from the kernel before

    for (int n0 = first; n0 < hi; n0 += step) {
        const int n = n0 + g;
        const bool live = n < hi;
        int k0 = 0, k1 = 0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), sps_own = acc, spd_own = acc;
        if (live) {
            k0 = __ldg(adj_ptr + n);
            k1 = __ldg(adj_ptr + n + 1);
            const float* row = S + (size_t)n * LD + 4 * c;
            sps_own = ldg4(row);
            spd_own = ldg4(row + 2 * H);
            acc = ldg4(row + 4 * H);               // Qs[n] (holds b3)
        }
        const int trips = __reduce_max_sync(0xffffffffu, k1 - k0);      // warp-uniform trip count
        for (int t = 0; t < trips; t += 4) {
            int ent[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ent[u] = (k0 + t + u < k1) ? __ldg(adj + k0 + t + u) : ADJ_NONE;
            float4 a[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int nb = ent[u] & 0x7fffffff;
                a[u] = q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nb != ADJ_NONE) {
                    const float* rp = S + (size_t)nb * LD + (ent[u] < 0 ? 2 * H : 0) + 4 * c;
                    a[u] = ldg4(rp);
                    q[u] = ldg4(rp + H);
                }
            }
            float z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = edge_partial(a[u], ent[u] < 0 ? sps_own : spd_own, w2n);
            const float e_mine = sigmoid_fast(z0 + reduce4_transposed<G>(z, c));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float e = __shfl_sync(0xffffffffu, e_mine, grp_lane0 + lane_of_edge<G>(u));
                fma4(acc, e, q[u]);                // absent entry: q = 0 (and e finite: a = 0 gives 1 / (0 + 1))
            }
        }
        if (live) {
            acc.x = tanh_node(acc.x); acc.y = tanh_node(acc.y); acc.z = tanh_node(acc.z); acc.w = tanh_node(acc.w);
            st4(h1_out + (size_t)n * ld_out + 4 * c, acc);
        }
    }
}

// Final edge step (gnn/model.py:156 -> 69-81): scores per slot.  Walks the destination-CSR (the end node's
// SPd is the CTA-local operand, the start node's SPs the gathered row), writes scores[in_eid[k]], then a
// sweep over the slots WITHOUT an end node (padding of merge_graphs, half edges), which the CSR does not list.
// State: rows of `ld` floats with SPs at column off_s and SPd at column off_d.
template <int H, int NTHR, int MINB, bool RANGES>
__global__ void __launch_bounds__(NTHR, MINB)
edge_final_kernel(const float* __restrict__ blob, const float* __restrict__ P, const int ld, const int off_s, const int off_d,
                  const GnnsegGraph gr, float* __restrict__ scores) {
    using B = Blob<H>;
    constexpr int G = H / 4, NPW = 32 / G;
    __shared__ int s_range[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, g = lane / G;
    pdl_launch_dependents();
    int lo = 0, hi = gr.n_nodes, first = (blockIdx.x * (NTHR / 32) + warp) * NPW, step = gridDim.x * (NTHR / 32) * NPW;
    if (RANGES) {
        cta_node_range(gr.in_ptr, gr.n_nodes, s_range);
        lo = s_range[0]; hi = s_range[1];
        first = lo + warp * NPW; step = (NTHR / 32) * NPW;
    }
    const float4 w2n = ldg4(blob + B::W2N + 4 * c);
    const float4 sb1 = ldg4(blob + B::SB1 + 4 * c);       // SPs of an absent start node: Ps = b1
    const float z0 = __ldg(blob + B::Z0);
    const int my_edge = edge_of_lane<G>(c);
    const bool writer = (c & (G / 4 - 1)) == 0;           // one lane per edge total writes
    pdl_wait();

    for (int n0 = first; n0 < hi; n0 += step) {
        const int n = n0 + g;
        const bool live = n < hi;
        int k0 = 0, k1 = 0;
        float4 spd_own = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            k0 = __ldg(gr.in_ptr + n);
            k1 = __ldg(gr.in_ptr + n + 1);
            spd_own = ldg4(P + (size_t)n * ld + off_d + 4 * c);
        }
        const int trips = __reduce_max_sync(0xffffffffu, k1 - k0);
        for (int t = 0; t < trips; t += 4) {
            float z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + t + u < k1) {
                    const int nb = __ldg(gr.in_nbr + k0 + t + u);
                    a = nb >= 0 ? ldg4(P + (size_t)nb * ld + off_s + 4 * c) : sb1;
                }
                z[u] = edge_partial(a, spd_own, w2n);
            }
            const float e_mine = sigmoid_fast(z0 + reduce4_transposed<G>(z, c));
            const int k = k0 + t + my_edge;
            if (writer && k < k1) scores[__ldg(gr.in_eid + k)] = e_mine;
        }
    }
    // slots without an end node: Pd contributes nothing (SPd = 1); without a start node either: the padding constant
    const float* w2s = blob + B::W2N;
    const float* sb1s = blob + B::SB1;
    for (int j = blockIdx.x * NTHR + threadIdx.x; j < gr.n_slots; j += gridDim.x * NTHR) {
        if (__ldg(gr.dst + j) >= 0) continue;
        const int s = __ldg(gr.src + j);
        const float* row = s >= 0 ? P + (size_t)s * ld + off_s : sb1s;
        float z = z0;
        for (int k = 0; k < H; ++k) {
            const float p = __ldg(row + k);
            z = fmaf(__ldg(w2s + k), rcp_approx(fmaf(p, p, 1.f)), z);
        }
        scores[j] = sigmoid_fast(z);
    }
}

// adj_ptr[n] = in_ptr[n] + out_ptr[n]; the node's entries: in-edges (neighbour = start node), then out-edges
// (neighbour = end node, bit 31 set), each in ascending slot order; absent neighbours become ADJ_NONE.
__global__ void __launch_bounds__(256)
build_adjacency_kernel(const GnnsegGraph g, int32_t* __restrict__ adj_ptr, int32_t* __restrict__ adj) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n <= g.n_nodes; n += gridDim.x * blockDim.x) {
        const int i0 = __ldg(g.in_ptr + n), o0 = __ldg(g.out_ptr + n);
        adj_ptr[n] = i0 + o0;
        if (n == g.n_nodes) break;
        const int i1 = __ldg(g.in_ptr + n + 1), o1 = __ldg(g.out_ptr + n + 1);
        int w = i0 + o0;
        for (int k = i0; k < i1; ++k) { const int nb = __ldg(g.in_nbr + k); adj[w++] = nb >= 0 ? nb : ADJ_NONE; }
        for (int k = o0; k < o1; ++k) { const int nb = __ldg(g.out_nbr + k); adj[w++] = nb >= 0 ? (nb | ADJ_OUT) : ADJ_NONE; }
    }
}

// ---- launchers ---------------------------------------------------------------------------------------
bool use_pdl(int n_slots);   // gnnseg_forward.cu

template <int H>
static int launch_fused_gather(const float* blob, const GnnsegGraph* g, const float* S, float* h1, int ld_h1, cudaStream_t st) {
    if (g->n_nodes == 0) return GNNSEG_OK;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    constexpr int NPW = 32 / (H / 4);
    // GNNSEG_FUSED_CFG (A/B runs): 0 = 256 threads x 4 CTAs per SM (64 registers), 1 = 256 x 3 (80 registers),
    // 2 = 256 x 5 (48 registers), 3 = one 768-thread CTA per SM sweeping a contiguous node range
    static const int cfg = [] { const char* v = getenv("GNNSEG_FUSED_CFG"); return v ? atoi(v) : 0; }();
    auto go = [&](auto kern, int nthr, int per_sm) {
        int grid = (g->n_nodes + (nthr / 32) * NPW - 1) / ((nthr / 32) * NPW);
        if (grid > sms * per_sm) grid = sms * per_sm;
        if (launch_pdl(kern, grid, nthr, 0, st, use_pdl(g->n_slots), blob, S, g->adj_ptr, g->adj, g->n_nodes, h1, ld_h1) != cudaSuccess)
            return (int)GNNSEG_ECUDA;
        return cudaGetLastError() == cudaSuccess ? (int)GNNSEG_OK : (int)GNNSEG_ECUDA;
    };
    if (cfg == 1) return go(fused_gather_kernel<H, 256, 3, false>, 256, 3);
    if (cfg == 2) return go(fused_gather_kernel<H, 256, 5, false>, 256, 5);
    if (cfg == 3) return go(fused_gather_kernel<H, 768, 1, true>, 768, 1);
    return go(fused_gather_kernel<H, 256, 4, false>, 256, 4);
}

template <int H>
static int launch_edge_final(const float* blob, const GnnsegGraph* g, const float* P, int ld, int off_s, int off_d, float* scores,
                             cudaStream_t st) {
    if (g->n_slots == 0) return GNNSEG_OK;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    constexpr int NTHR = 256, PER_SM = 6;
    int grid = (g->n_slots + NTHR - 1) / NTHR;
    if (grid > sms * PER_SM) grid = sms * PER_SM;
    if (launch_pdl(edge_final_kernel<H, NTHR, PER_SM, false>, grid, NTHR, 0, st, use_pdl(g->n_slots), blob, P, ld, off_s, off_d, *g,
                   scores) != cudaSuccess)
        return GNNSEG_ECUDA;
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int fused_gather_step(const float* blob, const GnnsegGraph* g, const float* S, int h, float* h1, int ld_h1, cudaStream_t st) {
    if (h == 32) return launch_fused_gather<32>(blob, g, S, h1, ld_h1, st);
    if (h == 64) return launch_fused_gather<64>(blob, g, S, h1, ld_h1, st);
    return GNNSEG_EUNSUPPORTED;
}
int edge_final_step(const float* blob, const GnnsegGraph* g, const float* P, int ld, int off_s, int off_d, int h, float* scores,
                    cudaStream_t st) {
    if (h == 32) return launch_edge_final<32>(blob, g, P, ld, off_s, off_d, scores, st);
    if (h == 64) return launch_edge_final<64>(blob, g, P, ld, off_s, off_d, scores, st);
    return GNNSEG_EUNSUPPORTED;
}
int build_adjacency(const GnnsegGraph* g, int32_t* adj_ptr, int32_t* adj, cudaStream_t st) {
    int grid = (g->n_nodes + 1 + 255) / 256;
    const int cap = 148 * 8;
    if (grid > cap) grid = cap;
    build_adjacency_kernel<<<grid, 256, 0, st>>>(*g, adj_ptr, adj);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
