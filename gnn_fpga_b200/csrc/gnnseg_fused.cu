// Fused inference path (hidden_dim 32 / 64): the edge step evaluated inside the node step's CSR walk.
//
// Reference: EdgeNetwork.forward + NodeNetwork.forward, gnn/model.py:69-81,113-125, in the projection-first
// form of gnnseg_common.cuh.  Per node n the state row is
//     S[n] = [ SPs (H) | SPd (H) | Qi (H) | Qo (H) | Qs (H) ]        5H floats, 128-byte aligned
// with SPs = 2^(log2e * Ps), SPd = 2^(log2e * Pd): the exponentials of the edge network's first-layer
// projections, taken ONCE per node by the kernel that produced them.  For an edge s -> d
//     tanh(Ps[s] + Pd[d]) = 1 - 2 / ((SPs[s] * SPd[d])^2 + 1)
// costs one reciprocal per hidden unit instead of an exponential and a reciprocal (the MUFU pipe is what
// bounds this kernel), and the row a visit fetches serves both halves of the work:
//     in-edge  s -> n :  SPs[s], Qi[s] (columns 0, 2H of row s)    e = sigmoid(w2 . tanh(Ps[s] + Pd[n]) + b2),  acc += e * Qi[s]
//     out-edge n -> d :  SPd[d], Qo[d] (columns H, 3H of row d)    e = sigmoid(w2 . tanh(Ps[n] + Pd[d]) + b2),  acc += e * Qo[d]
//     h1[n] = tanh(Qs[n] + acc)          own term first, then in-edges, then out-edges, ascending slot order
// Every edge is evaluated twice (once from each end) and nothing is written per edge: no e_in / e_out arrays,
// no inverse maps, no separate edge launch.  No atomics; the order of every sum is fixed => bit-reproducible.
//
// Range of the exponentials: log2e * P is clamped to [-63, 63] (|P| <= 43.6) so that the product of two
// stays a normal fp32 number; beyond, the squared product saturates to inf / 0 and tanh to +-1 exactly as
// it should, EXCEPT when two clamped values of opposite sign would have cancelled.  The producing kernel
// raises a flag word when it clamps; the host side turns that into an error (gnnseg_forward's `flags`).
//
// Work distribution: the warps of the whole grid take consecutive groups of nodes (grid stride), so that at
// any moment all SMs work on the same window of the batch and the rows they gather are L2 hits.  (Measured:
// one contiguous node range per CTA, with or without a locality renumbering of the nodes, is 20 - 35 % SLOWER,
// profiles/r2/locality_experiment.txt.)  The kernel is bound by instruction issue (ncu: 67 % issue slots, 16 %
// MUFU, 27 % DRAM at 203 instructions per batch of four entries in its first form), hence: padded adjacency
// lists read four entries per load with no bounds checks, an all-zero extra row instead of predicates for
// absent neighbours, packed fp32x2 arithmetic, and a degree-balanced node order inside every window.
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

// ---- adjacency ---------------------------------------------------------------------------------------
// Per node n its in-edges (neighbour = start node) then its out-edges (neighbour = end node, bit 31 set), each in
// ascending slot order, PADDED to a multiple of four entries; an absent neighbour (half edge) and the padding
// point at row n_nodes, an extra row every state buffer carries (zeros for the state rows: such an entry then
// contributes e * 0).  The list of node n starts at adj_offset(adj_ptr[n], n), a multiple of four computed from
// the unpadded cumulative count adj_ptr[n] = in_ptr[n] + out_ptr[n]: no scan is needed to place the padded
// lists, a batch of four entries is one aligned 16-byte load, and the kernels need no per-entry bounds checks.
constexpr int ADJ_OUT = (int)0x80000000;
__host__ __device__ __forceinline__ int adj_offset(const int cum, const int n) { return (((cum + 3) >> 2) << 2) + 4 * n; }

// `order` lists the nodes of every window of ORDER_WINDOW nodes by their number of four-entry batches: the
// NPW nodes a warp works on at a time then have the same trip count (measured on the ACTS-like events: 3.76
// -> 2.86 trips per node group), while all SMs still sweep the same window of the batch at any moment.
constexpr int ORDER_WINDOW = 4096;

// ENDPOINTS: the neighbour of an entry is looked up through its slot (src[in_eid[k]], dst[out_eid[k]]) instead of read from
// in_nbr / out_nbr: the lean batch assembly of the inference stream does not write those
template <bool ENDPOINTS>
__global__ void __launch_bounds__(256)
build_adjacency_kernel(const GnnsegGraph g, int32_t* __restrict__ adj_ptr, int32_t* __restrict__ adj) {
    const int pad = g.n_nodes;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n <= g.n_nodes; n += gridDim.x * blockDim.x) {
        const int i0 = __ldg(g.in_ptr + n), o0 = __ldg(g.out_ptr + n);
        adj_ptr[n] = i0 + o0;
        if (n == g.n_nodes) break;
        const int i1 = __ldg(g.in_ptr + n + 1), o1 = __ldg(g.out_ptr + n + 1);
        int w = adj_offset(i0 + o0, n);
        const int w_end = w + (((i1 - i0) + (o1 - o0) + 3) & ~3);
        for (int k = i0; k < i1; ++k) {
            const int nb = ENDPOINTS ? __ldg(g.src + __ldg(g.in_eid + k)) : __ldg(g.in_nbr + k);
            adj[w++] = nb >= 0 ? nb : pad;
        }
        for (int k = o0; k < o1; ++k) {
            const int nb = ENDPOINTS ? __ldg(g.dst + __ldg(g.out_eid + k)) : __ldg(g.out_nbr + k);
            adj[w++] = nb >= 0 ? (nb | ADJ_OUT) : pad;
        }
        while (w < w_end) adj[w++] = pad;
    }
}

// one CTA per window: counting sort of the window's nodes by min(trips, 31) (shared-memory counters; the order
// inside a class is whatever the atomics give: it only schedules work, no result depends on it)
__global__ void __launch_bounds__(256)
build_order_kernel(const GnnsegGraph g, int32_t* __restrict__ order) {
    __shared__ int cnt[32], base[32];
    const int w0 = blockIdx.x * ORDER_WINDOW, w1 = min(w0 + ORDER_WINDOW, g.n_nodes);
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    auto key = [&](const int n) {
        const int deg = (__ldg(g.in_ptr + n + 1) - __ldg(g.in_ptr + n)) + (__ldg(g.out_ptr + n + 1) - __ldg(g.out_ptr + n));
        return min((deg + 3) >> 2, 31);
    };
    for (int n = w0 + threadIdx.x; n < w1; n += blockDim.x) atomicAdd(&cnt[key(n)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 31; k >= 0; --k) { base[k] = run; run += cnt[k]; }      // long lists first
    }
    __syncthreads();
    for (int n = w0 + threadIdx.x; n < w1; n += blockDim.x) order[w0 + atomicAdd(&base[key(n)], 1)] = n;
}

// ---- the fused edge + node-gather step ---------------------------------------------------------------
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }

// partial sum over this lane's four hidden units of  -2 w2_k / ((a_k * own_k)^2 + 1), packed fp32x2 arithmetic
// (FMUL2 / FFMA2: this kernel is bound by instruction issue, not by the FMA pipe)
__device__ __forceinline__ float edge_partial(const float4 a, const float4 own, const float4 w2n) {
    const float2 one = make_float2(1.f, 1.f);
    const float2 p0 = __fmul2_rn(lo2(a), lo2(own)), p1 = __fmul2_rn(hi2(a), hi2(own));
    const float2 d0 = __ffma2_rn(p0, p0, one), d1 = __ffma2_rn(p1, p1, one);
    const float2 r0 = make_float2(rcp_approx(d0.x), rcp_approx(d0.y)), r1 = make_float2(rcp_approx(d1.x), rcp_approx(d1.y));
    const float2 z = __ffma2_rn(hi2(w2n), r1, __fmul2_rn(lo2(w2n), r0));
    return z.x + z.y;
}

template <int H, int MINB>
__global__ void __launch_bounds__(256, MINB)
fused_gather_kernel(const float* __restrict__ blob, const float* __restrict__ S, const int32_t* __restrict__ adj_ptr,
                    const int32_t* __restrict__ adj, const int32_t* __restrict__ order, const int n_nodes,
                    float* __restrict__ h1_out, const int ld_out) {
    using B = Blob<H>;
    constexpr int G = H / 4, NPW = 32 / G, LD4 = 5 * G;          // a state row is 5G float4
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, g = lane / G;
    const float4* __restrict__ S4 = reinterpret_cast<const float4*>(S);
    pdl_launch_dependents();
    const float4 w2n = ldg4(blob + B::W2N + 4 * c);
    const float z0 = __ldg(blob + B::Z0);
    const int grp_lane0 = lane - c;
    pdl_wait();                                    // S comes from the kernel before

    // The walk is a chain of dependent loads (node id -> list bounds -> entries -> rows); the kernel keeps one step of
    // it in flight ahead of the arithmetic: the header of the NEXT node group is fetched while this one is worked on,
    // and the entries of the next batch of four while the rows of this one are in flight.
    const int step = gridDim.x * 8 * NPW;
    auto node_of = [&](const int i) { return i < n_nodes ? (order ? __ldg(order + i) : i) : n_nodes; };   // spare lanes: the zero row
    int i0 = (blockIdx.x * 8 + warp) * NPW;
    int n = node_of(i0 + g), cum = 0, deg = 0;
    if (n < n_nodes) {
        cum = __ldg(adj_ptr + n);
        deg = __ldg(adj_ptr + n + 1) - cum;
    }
    while (i0 < n_nodes) {
        const int n_nx = node_of(i0 + step + g);                 // next group, step 1: the node id
        const int4* __restrict__ ap = reinterpret_cast<const int4*>(adj + adj_offset(cum, n));
        const int4 pad = make_int4(n_nodes, n_nodes, n_nodes, n_nodes);
        int4 ent_nx = pad;
        if (deg > 0) ent_nx = __ldg(ap);
        const float4* row = S4 + (size_t)n * LD4 + c;
        const float4 sps_own = __ldg(row), spd_own = __ldg(row + G);
        float4 acc = __ldg(row + 4 * G);           // Qs[n] (holds b3)
        int cum_nx = 0, deg_nx = 0;                // next group, step 2: its list bounds
        if (n_nx < n_nodes) {
            cum_nx = __ldg(adj_ptr + n_nx);
            deg_nx = __ldg(adj_ptr + n_nx + 1) - cum_nx;
        }
        const int trips = __reduce_max_sync(0xffffffffu, (deg + 3) >> 2);      // warp-uniform trip count
        for (int t = 0; t < trips; ++t) {
            const int4 ent = ent_nx;
            ent_nx = pad;
            if (4 * (t + 1) < deg) ent_nx = __ldg(ap + t + 1);
            const int e4[4] = {ent.x, ent.y, ent.z, ent.w};
            float4 a[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4* rp = S4 + (size_t)(e4[u] & 0x7fffffff) * LD4 + (e4[u] < 0 ? G : 0) + c;
                a[u] = __ldg(rp);
                q[u] = __ldg(rp + 2 * G);
            }
            float z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = edge_partial(a[u], e4[u] < 0 ? sps_own : spd_own, w2n);
            const float e_mine = sigmoid_fast(z0 + reduce4_transposed<G>(z, c));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float e = __shfl_sync(0xffffffffu, e_mine, grp_lane0 + lane_of_edge<G>(u));
                const float2 e2 = make_float2(e, e);
                const float2 lo = __ffma2_rn(e2, lo2(q[u]), lo2(acc)), hi = __ffma2_rn(e2, hi2(q[u]), hi2(acc));
                acc = make_float4(lo.x, lo.y, hi.x, hi.y);      // the zero row: q = 0, e finite
            }
        }
        if (n < n_nodes) {
            acc.x = tanh_node(acc.x); acc.y = tanh_node(acc.y); acc.z = tanh_node(acc.z); acc.w = tanh_node(acc.w);
            st4(h1_out + (size_t)n * ld_out + 4 * c, acc);
        }
        i0 += step; n = n_nx; cum = cum_nx; deg = deg_nx;
    }
}

// Final edge step (gnn/model.py:156 -> 69-81): scores per slot.  Walks the in-edges of every node (the end node's
// SPd is the node's own operand, the start node's SPs the gathered row; an absent start node reads the extra row,
// which holds 2^(log2e b1) there), writes scores[in_eid[k]], then a sweep over the slots WITHOUT an end node (padding
// of merge_graphs, half edges), which no node lists.  State: n_nodes + 1 rows of `ld` floats with SPs at column
// off_s and SPd at column off_d.
template <int H, int MINB>
__global__ void __launch_bounds__(256, MINB)
edge_final_kernel(const float* __restrict__ blob, const float* __restrict__ P, const int ld, const int off_s, const int off_d,
                  const GnnsegGraph gr, float* __restrict__ scores) {
    using B = Blob<H>;
    constexpr int G = H / 4, NPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, g = lane / G;
    const int n_nodes = gr.n_nodes, ld4 = ld >> 2;
    const float4* __restrict__ Ps4 = reinterpret_cast<const float4*>(P + off_s) + c;
    const float4* __restrict__ Pd4 = reinterpret_cast<const float4*>(P + off_d) + c;
    pdl_launch_dependents();
    const float4 w2n = ldg4(blob + B::W2N + 4 * c);
    const float z0 = __ldg(blob + B::Z0);
    const int my_edge = edge_of_lane<G>(c);
    const bool writer = (c & (G / 4 - 1)) == 0;           // one lane per edge total writes
    pdl_wait();

    const int step = gridDim.x * 8 * NPW;
    auto node_of = [&](const int i) { return i < n_nodes ? (gr.node_order ? __ldg(gr.node_order + i) : i) : n_nodes; };
    int i0 = (blockIdx.x * 8 + warp) * NPW;
    int n = node_of(i0 + g), cum = 0, k0 = 0, deg = 0;
    if (n < n_nodes) {
        cum = __ldg(gr.adj_ptr + n);
        k0 = __ldg(gr.in_ptr + n);
        deg = __ldg(gr.in_ptr + n + 1) - k0;               // the in-edges lead the node's adjacency list
    }
    while (i0 < n_nodes) {
        const int n_nx = node_of(i0 + step + g);
        const int4* __restrict__ ap = reinterpret_cast<const int4*>(gr.adj + adj_offset(cum, n));
        const int4 pad = make_int4(n_nodes, n_nodes, n_nodes, n_nodes);
        int4 ent_nx = pad;
        if (deg > 0) ent_nx = __ldg(ap);
        const float4 spd_own = __ldg(Pd4 + (size_t)n * ld4);
        int cum_nx = 0, k0_nx = 0, deg_nx = 0;
        if (n_nx < n_nodes) {
            cum_nx = __ldg(gr.adj_ptr + n_nx);
            k0_nx = __ldg(gr.in_ptr + n_nx);
            deg_nx = __ldg(gr.in_ptr + n_nx + 1) - k0_nx;
        }
        const int trips = __reduce_max_sync(0xffffffffu, (deg + 3) >> 2);
        for (int t = 0; t < trips; ++t) {
            const int4 ent = ent_nx;
            ent_nx = pad;
            if (4 * (t + 1) < deg) ent_nx = __ldg(ap + t + 1);
            const int e4[4] = {ent.x, ent.y, ent.z, ent.w};
            const int k = 4 * t + my_edge;
            int slot = -1;
            if (writer && k < deg) slot = __ldg(gr.in_eid + k0 + k);             // (entries past deg: out-edges or padding)
            float z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = edge_partial(__ldg(Ps4 + (size_t)(e4[u] & 0x7fffffff) * ld4), spd_own, w2n);
            const float e_mine = sigmoid_fast(z0 + reduce4_transposed<G>(z, c));
            if (slot >= 0) scores[slot] = e_mine;
        }
        i0 += step; n = n_nx; cum = cum_nx; k0 = k0_nx; deg = deg_nx;
    }
    // slots without an end node: Pd contributes nothing (SPd = 1); without a start node either: the padding constant
    const float* w2s = blob + B::W2N;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < gr.n_slots; j += gridDim.x * 256) {
        if (__ldg(gr.dst + j) >= 0) continue;
        const int s = __ldg(gr.src + j);
        const float* row = P + (size_t)(s >= 0 ? s : n_nodes) * ld + off_s;
        float z = z0;
        for (int k = 0; k < H; ++k) {
            const float p = __ldg(row + k);
            z = fmaf(__ldg(w2s + k), rcp_approx(fmaf(p, p, 1.f)), z);
        }
        scores[j] = sigmoid_fast(z);
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
    // GNNSEG_FUSED_CFG (A/B runs): resident 256-thread CTAs per SM: 0 -> 4 (64 registers), 1 -> 3 (80), 2 -> 5 (48), 3 -> 6 (40);
    // +10: without the degree-balanced node order
    static const int cfg = [] { const char* v = getenv("GNNSEG_FUSED_CFG"); return v ? atoi(v) : 0; }();
    const int32_t* order = cfg >= 10 ? nullptr : g->node_order;
    auto go = [&](auto kern, int per_sm) {
        int grid = (g->n_nodes + 8 * NPW - 1) / (8 * NPW);
        if (grid > sms * per_sm) grid = sms * per_sm;
        if (launch_pdl(kern, grid, 256, 0, st, use_pdl(g->n_slots), blob, S, g->adj_ptr, g->adj, order, g->n_nodes, h1, ld_h1) !=
            cudaSuccess)
            return (int)GNNSEG_ECUDA;
        return cudaGetLastError() == cudaSuccess ? (int)GNNSEG_OK : (int)GNNSEG_ECUDA;
    };
    switch (cfg % 10) {
        case 1: return go(fused_gather_kernel<H, 3>, 3);
        case 2: return go(fused_gather_kernel<H, 5>, 5);
        case 3: return go(fused_gather_kernel<H, 6>, 6);
        default: return go(fused_gather_kernel<H, 4>, 4);
    }
}

template <int H>
static int launch_edge_final(const float* blob, const GnnsegGraph* g, const float* P, int ld, int off_s, int off_d, float* scores,
                             cudaStream_t st) {
    if (g->n_slots == 0) return GNNSEG_OK;
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    constexpr int PER_SM = 6, NPW = 32 / (H / 4);
    int grid = (g->n_nodes + 8 * NPW - 1) / (8 * NPW);
    const int by_slots = (g->n_slots + 255) / 256;
    if (grid < by_slots) grid = by_slots;
    if (grid > sms * PER_SM) grid = sms * PER_SM;
    if (launch_pdl(edge_final_kernel<H, PER_SM>, grid, 256, 0, st, use_pdl(g->n_slots), blob, P, ld, off_s, off_d, *g, scores) !=
        cudaSuccess)
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
size_t adjacency_entries(int n_nodes, int n_slots) { return 2 * (size_t)n_slots + 4 * (size_t)n_nodes + 8; }

int build_adjacency(const GnnsegGraph* g, int32_t* adj_ptr, int32_t* adj, int32_t* order, cudaStream_t st, bool from_endpoints) {
    int grid = (g->n_nodes + 1 + 255) / 256;
    const int cap = 148 * 8;
    if (grid > cap) grid = cap;
    if (from_endpoints) build_adjacency_kernel<true><<<grid, 256, 0, st>>>(*g, adj_ptr, adj);
    else build_adjacency_kernel<false><<<grid, 256, 0, st>>>(*g, adj_ptr, adj);
    if (order && g->n_nodes > 0) build_order_kernel<<<(g->n_nodes + ORDER_WINDOW - 1) / ORDER_WINDOW, 256, 0, st>>>(*g, order);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
