// Segment (edge candidate) construction on the GPU: what construct_graph / construct_segments /
// select_segments do with a pandas cross join per layer pair (gnn/graph.py:44-142), i.e. the step
// in front of the classifier.  For every layer pair (l1, l2), every hit i on l1 (in row order) and
// every hit j on l2 (in row order):
//     dphi = phi_j - phi_i, wrapped once into [-pi, pi]      (calc_dphi, gnn/graph.py:37-42)
//     dz = z_j - z_i;  dr = r_j - r_i;  phi_slope = dphi / dr;  z0 = z_i - r_i * dz / dr
//     keep  <=>  |phi_slope| < (l1 < 5 ? phi_slope_max : phi_slope_outer_max)  and  |z0| < z0_max
// Kept pairs come out in the reference's order (layer pair, i, j), as positional hit indices, with
// the label y = (particle_id_i == particle_id_j).  Every operation is carried out in the dtype of the
// hit columns (float32 or float64) with round-to-nearest IEEE operations and no contraction, in the
// reference's order of evaluation, so the selection is bit-identical to pandas / numpy.
//
// Two passes, no atomics: count per (pair, i) row -> exclusive scan -> fill (warp ballots keep j order).
#include "gnnseg_common.cuh"

namespace gnnseg {

constexpr int SEG_MAX_PAIRS = 32;
constexpr int SEG_MAX_LAYERS = 32;

struct SegPairs {
    int n;
    int l1[SEG_MAX_PAIRS];
    int l2[SEG_MAX_PAIRS];
};

// ws layout (int32): layer_ptr [SEG_MAX_LAYERS + 1] | row_off [SEG_MAX_PAIRS + 1] | layer_hits [n_hits] | off [rows_cap + 1]
struct SegWs {
    int* layer_ptr;
    int* row_off;
    int* layer_hits;
    int* off;
};

// Which events a launch covers.  Single event (hit_off == nullptr): the columns hold one event of
// n_hits rows, node ids start at node_offset, slots at 0.  Batch: event b (blockIdx.y, or blockIdx.x
// of the one-block kernels) owns rows [hit_off[b], hit_off[b+1]) of the concatenated columns, its
// node ids are those row numbers, its slots start at b * capacity.  Both carve the same workspace:
//   [n_events][layer_ptr 33 | row_off 33 | 2] | layer_hits [hits_pad] | off [n_pairs * total_hits + n_events]
struct SegCtx {
    const int32_t* hit_off;
    int n_events, n_hits, node_offset, hits_pad, n_pairs;
    int* ws;
};
constexpr int SEG_FIXED = (SEG_MAX_LAYERS + 1) + (SEG_MAX_PAIRS + 1) + 2;

__device__ __forceinline__ SegWs seg_event(const SegCtx& c, const int b, int& h0, int& n, int& node0) {
    h0 = c.hit_off ? __ldg(c.hit_off + b) : 0;
    n = c.hit_off ? __ldg(c.hit_off + b + 1) - h0 : c.n_hits;
    node0 = c.hit_off ? h0 : c.node_offset;
    SegWs w;
    w.layer_ptr = c.ws + b * SEG_FIXED;
    w.row_off = w.layer_ptr + (SEG_MAX_LAYERS + 1);
    w.layer_hits = c.ws + c.n_events * SEG_FIXED + h0;
    w.off = c.ws + c.n_events * SEG_FIXED + c.hits_pad + c.n_pairs * h0 + b;
    return w;
}

// One block per event, one warp per layer: count the layer's hits, scan, then list them in row order
// (groupby('layer').get_group(l) keeps the frame's row order, gnn/graph.py:79-85).
__global__ void __launch_bounds__(32 * SEG_MAX_LAYERS)
seg_layer_lists_kernel(const int32_t* __restrict__ layer_all, const SegCtx ctx, const int n_layers, const SegPairs pairs) {
    __shared__ int s_cnt[SEG_MAX_LAYERS], s_ptr[SEG_MAX_LAYERS + 1];
    int h0, n_hits, node0;
    const SegWs ws = seg_event(ctx, blockIdx.x, h0, n_hits, node0);
    const int32_t* __restrict__ layer = layer_all + h0;
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int total = 0;
    if (l < n_layers)
        for (int base = 0; base < n_hits; base += 32) {
            const int i = base + lane;
            total += __popc(__ballot_sync(0xffffffffu, i < n_hits && __ldg(layer + i) == l));
        }
    if (lane == 0 && l < SEG_MAX_LAYERS) s_cnt[l] = l < n_layers ? total : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < SEG_MAX_LAYERS; ++k) { s_ptr[k] = acc; acc += s_cnt[k]; }
        s_ptr[SEG_MAX_LAYERS] = acc;
        for (int k = 0; k <= SEG_MAX_LAYERS; ++k) ws.layer_ptr[k] = s_ptr[k];
        int rows = 0;                                  // one row per hit on the first layer of every pair
        for (int p = 0; p < pairs.n; ++p) { ws.row_off[p] = rows; rows += s_cnt[pairs.l1[p]]; }
        for (int p = pairs.n; p <= SEG_MAX_PAIRS; ++p) ws.row_off[p] = rows;
    }
    __syncthreads();
    if (l < n_layers) {
        int pos = s_ptr[l];
        for (int base = 0; base < n_hits; base += 32) {
            const int i = base + lane;
            const bool mine = i < n_hits && __ldg(layer + i) == l;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (mine) ws.layer_hits[pos + __popc(m & ((1u << lane) - 1u))] = i;
            pos += __popc(m);
        }
    }
}

template <typename T> struct SegMath;
template <> struct SegMath<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};
template <> struct SegMath<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};

template <typename T>
struct SegCuts {
    T pi, two_pi, slope_in, slope_out, z0_max;
    int outer_from;
};

// select_segments' filter for one (i, j), gnn/graph.py:58-66
template <typename T>
__device__ __forceinline__ bool seg_keep(const T r1, const T phi1, const T z1, const T r2, const T phi2, const T z2,
                                         const T slope_cut, const SegCuts<T>& c) {
    using M = SegMath<T>;
    T dphi = M::sub(phi2, phi1);
    if (dphi > c.pi) dphi = M::sub(dphi, c.two_pi);
    if (dphi < -c.pi) dphi = M::add(dphi, c.two_pi);
    const T dz = M::sub(z2, z1), dr = M::sub(r2, r1);
    const T slope = M::div(dphi, dr);
    const T z0 = M::sub(z1, M::div(M::mul(r1, dz), dr));
    return M::abs(slope) < slope_cut && M::abs(z0) < c.z0_max;        // NaN (dr = 0 and dphi = 0) compares false, as in numpy
}

// One warp per row = (layer pair, hit i on the first layer).  FILL = false: write the row's count;
// FILL = true: write the kept pairs at off[row], j ascending.
template <typename T, bool FILL>
__global__ void __launch_bounds__(256)
seg_rows_kernel(const T* __restrict__ r_all, const T* __restrict__ phi_all, const T* __restrict__ z_all,
                const int64_t* __restrict__ pid_all, const SegPairs pairs, const SegCuts<T> cuts, const SegCtx ctx,
                const int capacity, int32_t* __restrict__ src_all, int32_t* __restrict__ dst_all,
                float* __restrict__ y_all) {
    int h0, n_hits, node_offset;
    const SegWs ws = seg_event(ctx, blockIdx.y, h0, n_hits, node_offset);
    const T* __restrict__ r = r_all + h0;
    const T* __restrict__ phi = phi_all + h0;
    const T* __restrict__ z = z_all + h0;
    const int64_t* __restrict__ pid = pid_all ? pid_all + h0 : nullptr;
    const size_t slot0 = ctx.hit_off ? (size_t)blockIdx.y * capacity : 0;
    int32_t* __restrict__ src = FILL ? src_all + slot0 : nullptr;
    int32_t* __restrict__ dst = FILL ? dst_all + slot0 : nullptr;
    float* __restrict__ y = (FILL && y_all) ? y_all + slot0 : nullptr;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int rows = __ldg(ws.row_off + SEG_MAX_PAIRS);
    for (int row = warp; row < rows; row += n_warps) {
        // which pair: row_off[p] <= row < row_off[p + 1] (empty pairs have equal offsets)
        const bool here = lane < pairs.n && __ldg(ws.row_off + lane) <= row && row < __ldg(ws.row_off + lane + 1);
        const int p = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
        const int l1 = pairs.l1[p], l2 = pairs.l2[p];
        const int i = __ldg(ws.layer_hits + __ldg(ws.layer_ptr + l1) + (row - __ldg(ws.row_off + p)));
        const int j0 = __ldg(ws.layer_ptr + l2), j1 = __ldg(ws.layer_ptr + l2 + 1);
        const T r1 = __ldg(r + i), phi1 = __ldg(phi + i), z1 = __ldg(z + i);
        const T slope_cut = l1 < cuts.outer_from ? cuts.slope_in : cuts.slope_out;
        const int64_t pid1 = (FILL && pid) ? __ldg(pid + i) : 0;
        int kept = FILL ? __ldg(ws.off + row) : 0;
        for (int jb = j0; jb < j1; jb += 32) {
            const int jj = jb + lane;
            bool keep = false;
            int j = 0;
            if (jj < j1) {
                j = __ldg(ws.layer_hits + jj);
                keep = seg_keep<T>(r1, phi1, z1, __ldg(r + j), __ldg(phi + j), __ldg(z + j), slope_cut, cuts);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) {
                const int k = kept + __popc(m & ((1u << lane) - 1u));
                if (k < capacity) {
                    src[k] = node_offset + i;
                    dst[k] = node_offset + j;
                    if (y) y[k] = (pid && __ldg(pid + j) == pid1) ? 1.f : 0.f;
                }
            }
            kept += __popc(m);
        }
        if (!FILL && lane == 0) ws.off[row] = kept;
    }
}

// exclusive scan of off[0..rows) in place, off[rows] = total; one block (rows <= n_pairs * n_hits)
__global__ void __launch_bounds__(1024)
seg_scan_kernel(const SegCtx ctx, const int capacity, int32_t* __restrict__ n_edges_all) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    int h0, n_hits, node0;
    const SegWs ws = seg_event(ctx, blockIdx.x, h0, n_hits, node0);
    int32_t* __restrict__ n_edges = n_edges_all + 2 * blockIdx.x;
    const int rows = ws.row_off[SEG_MAX_PAIRS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < rows ? ws.off[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;                           // inclusive over warps
        }
        __syncthreads();
        const int before = s_carry + (warp > 0 ? s_warp[warp - 1] : 0) + x - v;
        if (i < rows) ws.off[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws.off[rows] = s_carry;
        n_edges[0] = s_carry;
        n_edges[1] = s_carry > capacity ? 1 : 0;        // the output arrays were too small
    }
}

// X = (hits[feature_names].values / feature_scale).astype(np.float32), gnn/graph.py:118: float64 division
template <typename T>
__global__ void seg_features_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                                    const int n, const double sa, const double sb, const double sc,
                                    float* __restrict__ X) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        X[3 * i + 0] = (float)__ddiv_rn((double)a[i], sa);
        X[3 * i + 1] = (float)__ddiv_rn((double)b[i], sb);
        X[3 * i + 2] = (float)__ddiv_rn((double)c[i], sc);
    }
}

// ints of workspace for n_events events holding total_hits rows in all (see SegCtx)
static size_t seg_ws_ints(int n_events, int total_hits, int n_pairs) {
    return (size_t)n_events * SEG_FIXED + ((size_t)(total_hits + 3) & ~size_t(3)) + (size_t)n_pairs * total_hits + n_events + 4;
}
size_t segments_workspace_bytes(int n_hits, int n_pairs) { return seg_ws_ints(1, n_hits, n_pairs) * 4; }
size_t segments_batch_workspace_bytes(int n_events, int total_hits, int n_pairs) {
    return seg_ws_ints(n_events, total_hits, n_pairs) * 4;
}

template <typename T>
static SegCuts<T> seg_cuts(double slope_in, double slope_out, double z0_max, int outer_from) {
    SegCuts<T> cuts;
    cuts.pi = (T)3.141592653589793;            // np.pi, converted to the column dtype as numpy converts a Python float
    cuts.two_pi = (T)6.283185307179586;        // 2 * np.pi
    cuts.slope_in = (T)slope_in;
    cuts.slope_out = (T)slope_out;
    cuts.z0_max = (T)z0_max;
    cuts.outer_from = outer_from;
    return cuts;
}

static SegPairs seg_pairs(const int32_t* layer_pairs_host, int n_pairs) {
    SegPairs pairs;
    pairs.n = n_pairs;
    for (int p = 0; p < SEG_MAX_PAIRS; ++p) {
        pairs.l1[p] = p < n_pairs ? layer_pairs_host[2 * p] : 0;
        pairs.l2[p] = p < n_pairs ? layer_pairs_host[2 * p + 1] : 0;
    }
    return pairs;
}

// count = lists + per-row counts + scan (leaves the row offsets in the workspace); fill = the second pass.
template <typename T>
static int seg_run(const int32_t* layer, const T* r, const T* phi, const T* z, const int64_t* pid, const SegCtx& ctx,
                   int max_hits, const SegPairs& pairs, int n_layers, const SegCuts<T>& cuts, bool count, bool fill,
                   int capacity, int32_t* src, int32_t* dst, float* y, int32_t* n_edges, cudaStream_t st) {
    const int sms = cached_sm_count();
    if (sms < 1) return GNNSEG_ENODEVICE;
    const long long rows_cap = (long long)pairs.n * max_hits;
    long long gx = (rows_cap * 32 + 255) / 256;
    const long long cap = ctx.n_events > 1 ? (sms * 8 + ctx.n_events - 1) / ctx.n_events + 8 : sms * 8;   // a few waves in all
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)ctx.n_events);
    if (count) {
        seg_layer_lists_kernel<<<ctx.n_events, 32 * SEG_MAX_LAYERS, 0, st>>>(layer, ctx, n_layers, pairs);
        seg_rows_kernel<T, false><<<grid, 256, 0, st>>>(r, phi, z, pid, pairs, cuts, ctx, capacity, nullptr, nullptr, nullptr);
        seg_scan_kernel<<<ctx.n_events, 1024, 0, st>>>(ctx, capacity, n_edges);
    }
    if (fill)
        seg_rows_kernel<T, true><<<grid, 256, 0, st>>>(r, phi, z, pid, pairs, cuts, ctx, capacity, src, dst, y);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int build_segments(const int32_t* layer, const void* r, const void* phi, const void* z, int dtype_bytes,
                   const int64_t* pid, int n_hits, const int32_t* layer_pairs_host, int n_pairs, int n_layers,
                   double slope_in, double slope_out, double z0_max, int outer_from, int node_offset, int capacity,
                   int32_t* src, int32_t* dst, float* y, int32_t* n_edges, void* ws, cudaStream_t st) {
    const SegPairs pairs = seg_pairs(layer_pairs_host, n_pairs);
    SegCtx ctx;
    ctx.hit_off = nullptr; ctx.n_events = 1; ctx.n_hits = n_hits; ctx.node_offset = node_offset;
    ctx.hits_pad = (n_hits + 3) & ~3; ctx.n_pairs = n_pairs; ctx.ws = static_cast<int*>(ws);
    const bool fill = src && dst && capacity > 0;
    if (dtype_bytes == 4)
        return seg_run<float>(layer, static_cast<const float*>(r), static_cast<const float*>(phi), static_cast<const float*>(z),
                              pid, ctx, n_hits, pairs, n_layers, seg_cuts<float>(slope_in, slope_out, z0_max, outer_from),
                              true, fill, capacity, src, dst, y, n_edges, st);
    return seg_run<double>(layer, static_cast<const double*>(r), static_cast<const double*>(phi), static_cast<const double*>(z),
                           pid, ctx, n_hits, pairs, n_layers, seg_cuts<double>(slope_in, slope_out, z0_max, outer_from),
                           true, fill, capacity, src, dst, y, n_edges, st);
}

// All events of a batch in one set of launches.  e_max == 0: the counting pass (n_edges[2b] = edges of
// event b); e_max > 0: the filling pass alone, on the workspace the counting pass left behind.
int build_segments_batch(const int32_t* layer, const void* r, const void* phi, const void* z, int dtype_bytes,
                         const int64_t* pid, int n_events, const int32_t* hit_off, int total_hits, int max_hits,
                         const int32_t* layer_pairs_host, int n_pairs, int n_layers, double slope_in, double slope_out,
                         double z0_max, int outer_from, int e_max, int32_t* src, int32_t* dst, float* y, int32_t* n_edges,
                         void* ws, cudaStream_t st) {
    if (n_events == 0) return GNNSEG_OK;
    const SegPairs pairs = seg_pairs(layer_pairs_host, n_pairs);
    SegCtx ctx;
    ctx.hit_off = hit_off; ctx.n_events = n_events; ctx.n_hits = 0; ctx.node_offset = 0;
    ctx.hits_pad = (total_hits + 3) & ~3; ctx.n_pairs = n_pairs; ctx.ws = static_cast<int*>(ws);
    const bool fill = e_max > 0;
    if (dtype_bytes == 4)
        return seg_run<float>(layer, static_cast<const float*>(r), static_cast<const float*>(phi), static_cast<const float*>(z),
                              pid, ctx, max_hits, pairs, n_layers, seg_cuts<float>(slope_in, slope_out, z0_max, outer_from),
                              !fill, fill, e_max, src, dst, y, n_edges, st);
    return seg_run<double>(layer, static_cast<const double*>(r), static_cast<const double*>(phi), static_cast<const double*>(z),
                           pid, ctx, max_hits, pairs, n_layers, seg_cuts<double>(slope_in, slope_out, z0_max, outer_from),
                           !fill, fill, e_max, src, dst, y, n_edges, st);
}

int scale_features(const void* a, const void* b, const void* c, int dtype_bytes, int n, double sa, double sb, double sc,
                   float* X, cudaStream_t st) {
    if (n == 0) return GNNSEG_OK;
    const int grid = (n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024;
    if (dtype_bytes == 4)
        seg_features_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                         static_cast<const float*>(c), n, sa, sb, sc, X);
    else
        seg_features_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b),
                                                          static_cast<const double*>(c), n, sa, sb, sc, X);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
