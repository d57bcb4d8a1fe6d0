// Graph boundary kernels for sm_100a: dense incidence matrices -> int32 endpoints, and
// endpoints -> CSR in np.nonzero order.  Replaces make_sparse_graph / graph_from_sparse
// (gnn/graph.py:23-35) on the device.  Everything here is integer work and bit-exact.
#include "gnnseg_common.cuh"

namespace gnnseg {

__global__ void fill_i32_kernel(int32_t* __restrict__ p, int32_t v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// One warp per (b, n) row of Ri and of Ro: a coalesced scan of E floats.  A non-zero at
// column c makes node b*N+n the end (Ri) or start (Ro) of slot b*E+c (gnn/graph.py:132-135).
__global__ void __launch_bounds__(256)
dense_to_edges_kernel(const float* __restrict__ Ri, const float* __restrict__ Ro, const int B,
                      const int N, const int E, int32_t* __restrict__ src,
                      int32_t* __restrict__ dst, int32_t* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const long long n_rows = 2LL * B * N;   // rows of Ri then rows of Ro
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    int err = 0;
    for (long long r = warp0; r < n_rows; r += n_warps) {
        const bool is_ro = r >= (long long)B * N;
        const long long row = is_ro ? r - (long long)B * N : r;   // b*N + n
        const int b = (int)(row / N);
        const float* __restrict__ m = (is_ro ? Ro : Ri) + row * (long long)E;
        int32_t* __restrict__ out = (is_ro ? src : dst) + (long long)b * E;
        for (int c = lane; c < E; c += 32) {
            const float v = __ldg(m + c);
            if (v != 0.f) {
                if (v != 1.f) err |= GNNSEG_BAD_VALUE;
                if (atomicCAS(out + c, -1, (int32_t)row) != -1) err |= GNNSEG_BAD_HYPEREDGE;
            }
        }
    }
    if (err) atomicOr(err_flag, err);
}

// ---- CSR build -----------------------------------------------------------------------
// Every kernel handles one or two directions at once (blockIdx.y picks the direction), so that
// the destination-CSR and the source-CSR of a batch are built by the same eight launches.
struct CsrDir {
    const int32_t* key;      // [n_slots] row of each slot (dst for the in-CSR, src for the out-CSR), < 0 = absent
    const int32_t* other;    // [n_slots] the other endpoint
    int32_t* ptr;            // [n_nodes+1]
    int32_t* eid;            // [n_slots]
    int32_t* nbr;            // [n_slots]
    int32_t* pos;            // [n_slots] or null
    int32_t* count;          // workspace [n_nodes+1]
    int32_t* cursor;         // workspace [n_nodes+1]
    int32_t* block_sums;     // workspace [n_blocks+1]
    int32_t* grand_total;    // workspace [1]
};
struct CsrPair { CsrDir d[2]; };

__global__ void csr_init_kernel(const CsrPair p, const int n_slots, const int n_nodes) {
    const CsrDir& d = p.d[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(n_slots, n_nodes + 1); i += gridDim.x * blockDim.x) {
        if (i <= n_nodes) d.count[i] = 0;
        if (d.pos && i < n_slots) d.pos[i] = -1;
    }
}

__global__ void histogram_kernel(const CsrPair p, const int n_slots, const int n_nodes) {
    const CsrDir& d = p.d[blockIdx.y];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slots; j += gridDim.x * blockDim.x) {
        const int k = __ldg(d.key + j);
        if (k >= 0 && k < n_nodes) atomicAdd(d.count + k, 1);
    }
}

// Exclusive scan of count[0..n) into ptr[0..n], three passes.  SCAN_ITEMS per block.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_ITEMS = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;   // inclusive over warps
    }
    __syncthreads();
    const int warp_off = w == 0 ? 0 : warp_sums[w - 1];
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return warp_off + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_local_kernel(const CsrPair p, const int n) {
    const CsrDir& d = p.d[blockIdx.y];
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        v[i] = (base + i < n) ? d.count[base + i] : 0;
        sum += v[i];
    }
    int total;
    int off = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        if (base + i < n) d.ptr[base + i] = off;
        off += v[i];
    }
    if (threadIdx.x == 0) d.block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(const CsrPair p, const int n_blocks) {
    const CsrDir& d = p.d[blockIdx.y];
    int carry = 0;
    for (int base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? d.block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) d.block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *d.grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(const CsrPair p, const int n) {
    const CsrDir& d = p.d[blockIdx.y];
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
    const int off = d.block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i)
        if (base + i < n) {
            const int v = d.ptr[base + i] + off;
            d.ptr[base + i] = v;
            d.cursor[base + i] = v;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) d.ptr[n] = *d.grand_total;
}

__global__ void csr_fill_kernel(const CsrPair p, const int n_slots, const int n_nodes) {
    const CsrDir& d = p.d[blockIdx.y];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slots; j += gridDim.x * blockDim.x) {
        const int k = __ldg(d.key + j);
        if (k >= 0 && k < n_nodes) d.eid[atomicAdd(d.cursor + k, 1)] = j;
    }
}

// Put every row in ascending slot order (the atomics above fill rows in arbitrary order;
// ascending order is unique, so the result is deterministic and equals np.nonzero's).
// One thread per row: insertion sort for short rows, heapsort otherwise.
__device__ inline void sift_down(int32_t* a, int start, int end) {
    int root = start;
    while (2 * root + 1 <= end) {
        int child = 2 * root + 1, sw = root;
        if (a[sw] < a[child]) sw = child;
        if (child + 1 <= end && a[sw] < a[child + 1]) sw = child + 1;
        if (sw == root) return;
        const int32_t t = a[root]; a[root] = a[sw]; a[sw] = t;
        root = sw;
    }
}

__global__ void csr_sort_rows_kernel(const CsrPair p, const int n_nodes) {
    const CsrDir& d = p.d[blockIdx.y];
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        const int beg = d.ptr[n], len = d.ptr[n + 1] - beg;
        int32_t* a = d.eid + beg;
        if (len <= 24) {
            for (int i = 1; i < len; ++i) {
                const int32_t v = a[i];
                int k = i - 1;
                while (k >= 0 && a[k] > v) { a[k + 1] = a[k]; --k; }
                a[k + 1] = v;
            }
        } else {
            for (int s = (len - 2) / 2; s >= 0; --s) sift_down(a, s, len - 1);
            for (int end = len - 1; end > 0; --end) {
                const int32_t t = a[end]; a[end] = a[0]; a[0] = t;
                sift_down(a, 0, end - 1);
            }
        }
    }
}

__global__ void csr_nbr_kernel(const CsrPair p, const int n_nodes) {
    const CsrDir& d = p.d[blockIdx.y];
    const int n_used = d.ptr[n_nodes];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_used; s += gridDim.x * blockDim.x) {
        const int j = d.eid[s];
        d.nbr[s] = __ldg(d.other + j);
        if (d.pos) d.pos[j] = s;       // inverse map (pos was filled with -1)
    }
}

// ---- launchers -----------------------------------------------------------------------
static inline int grid_for(long long n, int threads, int cap) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

int dense_to_edges(const float* Ri, const float* Ro, int B, int N, int E, int32_t* src,
                   int32_t* dst, int32_t* err_flag, cudaStream_t st) {
    const size_t n_slots = (size_t)B * E;
    if (n_slots == 0) return GNNSEG_OK;
    fill_i32_kernel<<<grid_for((long long)n_slots, 256, 4096), 256, 0, st>>>(src, -1, n_slots);
    fill_i32_kernel<<<grid_for((long long)n_slots, 256, 4096), 256, 0, st>>>(dst, -1, n_slots);
    if ((long long)B * N > 0) {
        const long long rows = 2LL * B * N;
        dense_to_edges_kernel<<<grid_for(rows * 32, 256, 148 * 16), 256, 0, st>>>(Ri, Ro, B, N, E, src, dst, err_flag);
    }
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

size_t csr_workspace_bytes(int n_nodes, int n_slots) {
    (void)n_slots;
    const size_t n_blocks = ((size_t)n_nodes + SCAN_ITEMS - 1) / SCAN_ITEMS + 1;
    // count[n_nodes+1] | cursor[n_nodes+1] | block_sums[n_blocks] | grand_total[1]
    return (2 * ((size_t)n_nodes + 1) + n_blocks + 1 + 4) * sizeof(int32_t);
}

static void carve_dir(CsrDir& d, void* ws, int n_nodes) {
    const int n_blocks = (n_nodes + SCAN_ITEMS - 1) / SCAN_ITEMS;
    d.count = static_cast<int32_t*>(ws);
    d.cursor = d.count + (n_nodes + 1);
    d.block_sums = d.cursor + (n_nodes + 1);
    d.grand_total = d.block_sums + n_blocks + 1;
}

// n_dirs = 1 or 2 directions in the same launches
static int build_csr_dirs(const CsrPair& p, int n_dirs, int n_slots, int n_nodes, cudaStream_t st) {
    const int n_blocks = (n_nodes + SCAN_ITEMS - 1) / SCAN_ITEMS;
    const dim3 by(1, n_dirs);
    auto g2 = [&](int gx) { return dim3(gx, n_dirs); };
    csr_init_kernel<<<g2(grid_for(n_slots > n_nodes + 1 ? n_slots : n_nodes + 1, 256, 2048)), 256, 0, st>>>(p, n_slots, n_nodes);
    if (n_slots > 0) histogram_kernel<<<g2(grid_for(n_slots, 256, 2048)), 256, 0, st>>>(p, n_slots, n_nodes);
    if (n_blocks > 0) {
        scan_local_kernel<<<g2(n_blocks), SCAN_THREADS, 0, st>>>(p, n_nodes);
        scan_block_sums_kernel<<<by, SCAN_THREADS, 0, st>>>(p, n_blocks);
        scan_add_kernel<<<g2(n_blocks), SCAN_THREADS, 0, st>>>(p, n_nodes);
    } else {
        for (int i = 0; i < n_dirs; ++i) fill_i32_kernel<<<1, 32, 0, st>>>(p.d[i].ptr, 0, 1);
    }
    if (n_slots > 0 && n_nodes > 0) {
        csr_fill_kernel<<<g2(grid_for(n_slots, 256, 2048)), 256, 0, st>>>(p, n_slots, n_nodes);
        csr_sort_rows_kernel<<<g2(grid_for(n_nodes, 128, 4096)), 128, 0, st>>>(p, n_nodes);
        csr_nbr_kernel<<<g2(grid_for(n_slots, 256, 2048)), 256, 0, st>>>(p, n_nodes);
    }
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int build_csr(const int32_t* key, const int32_t* other, int n_slots, int n_nodes, int32_t* ptr,
              int32_t* eid, int32_t* nbr, int32_t* pos, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < csr_workspace_bytes(n_nodes, n_slots)) return GNNSEG_EWORKSPACE;
    CsrPair p{};
    p.d[0] = CsrDir{key, other, ptr, eid, nbr, pos, nullptr, nullptr, nullptr, nullptr};
    carve_dir(p.d[0], ws, n_nodes);
    return build_csr_dirs(p, 1, n_slots, n_nodes, st);
}

// both CSRs of a graph: destination-CSR (key = dst) into in_*, source-CSR (key = src) into out_*
int build_graph(const int32_t* src, const int32_t* dst, int n_slots, int n_nodes, int32_t* in_ptr,
                int32_t* in_eid, int32_t* in_nbr, int32_t* in_pos, int32_t* out_ptr, int32_t* out_eid,
                int32_t* out_nbr, int32_t* out_pos, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t one = csr_workspace_bytes(n_nodes, n_slots);
    if (ws_bytes < 2 * one) return GNNSEG_EWORKSPACE;
    CsrPair p{};
    p.d[0] = CsrDir{dst, src, in_ptr, in_eid, in_nbr, in_pos, nullptr, nullptr, nullptr, nullptr};
    p.d[1] = CsrDir{src, dst, out_ptr, out_eid, out_nbr, out_pos, nullptr, nullptr, nullptr, nullptr};
    carve_dir(p.d[0], ws, n_nodes);
    carve_dir(p.d[1], static_cast<char*>(ws) + one, n_nodes);
    return build_csr_dirs(p, 2, n_slots, n_nodes, st);
}

// ---- batch assembly from the host event store (gnnseg_store.cpp) -----------------------------------
// A batch of B consecutive events arrives as contiguous slices of the store's arrays: per-event local
// CSR row pointers and the edge columns of the incidence entries in CSR order.  The flattened batch
// graph (endpoints per slot, both CSRs, inverse maps) follows without sorting and without atomics:
// the store was validated at load time (no column listed twice), so every slot has one writer.
// meta = [node_off (B+1) | in_base (B+1) | out_base (B+1)]: first node / first in-entry / first out-entry
// of every event relative to the batch.
// LEAN: only what the fused inference path reads (endpoints per slot, row pointers, slot ids in CSR order): no inverse maps
template <typename ColT, bool LEAN>
__global__ void __launch_bounds__(256)
assemble_rows_kernel(const int32_t* __restrict__ meta, const int B, const int n_nodes, const int e_max,
                     const int32_t* __restrict__ in_ptr_l, const int32_t* __restrict__ out_ptr_l,
                     const ColT* __restrict__ in_col, const ColT* __restrict__ out_col, GnnsegGraphMut g) {
    const int32_t* node_off = meta;
    const int32_t* in_base = meta + (B + 1);
    const int32_t* out_base = meta + 2 * (B + 1);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        int lo = 0, hi = B;                                  // event of node n: last b with node_off[b] <= n
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(node_off + mid) <= n) lo = mid; else hi = mid;
        }
        const int b = lo;
        const int lp = n + b;                                // an event owns n_b + 1 local pointer entries
        const int slot0 = b * e_max;
        {
            const int beg = in_base[b] + __ldg(in_ptr_l + lp), end = in_base[b] + __ldg(in_ptr_l + lp + 1);
            g.in_ptr[n] = beg;
            if (n == n_nodes - 1) g.in_ptr[n_nodes] = end;
            for (int k = beg; k < end; ++k) {
                const int slot = slot0 + (int)in_col[k];
                g.dst[slot] = n;
                g.in_eid[k] = slot;
                if (!LEAN) g.in_pos[slot] = k;
            }
        }
        {
            const int beg = out_base[b] + __ldg(out_ptr_l + lp), end = out_base[b] + __ldg(out_ptr_l + lp + 1);
            g.out_ptr[n] = beg;
            if (n == n_nodes - 1) g.out_ptr[n_nodes] = end;
            for (int k = beg; k < end; ++k) {
                const int slot = slot0 + (int)out_col[k];
                g.src[slot] = n;
                g.out_eid[k] = slot;
                if (!LEAN) g.out_pos[slot] = k;
            }
        }
    }
}

// The lean assembly, one warp per 32 consecutive nodes.  The entries of those nodes are one contiguous range of the
// batch's entry arrays (events are concatenated), so the lanes sweep it 32 entries at a time -- coalesced reads of the
// columns, coalesced writes of the slot ids -- and find the owning node of an entry by a binary search over the 32 row
// starts held one per lane (five shuffles).  One thread per node walking its own five entries was 41 us at acts64.
template <typename ColT>
__global__ void __launch_bounds__(256)
assemble_rows_warp_kernel(const int32_t* __restrict__ meta, const int B, const int n_nodes, const int e_max,
                          const int32_t* __restrict__ in_ptr_l, const int32_t* __restrict__ out_ptr_l,
                          const ColT* __restrict__ in_col, const ColT* __restrict__ out_col, GnnsegGraphMut g) {
    constexpr unsigned FULL = 0xffffffffu;
    const int32_t* node_off = meta;
    const int32_t* in_base = meta + (B + 1);
    const int32_t* out_base = meta + 2 * (B + 1);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int n0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; n0 < n_nodes; n0 += warps * 32) {
        const int n = min(n0 + lane, n_nodes - 1);              // spare lanes repeat the last node (empty ranges below)
        int lo = 0, hi = B;                                    // event of node n: last b with node_off[b] <= n
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(node_off + mid) <= n) lo = mid; else hi = mid;
        }
        const int b = lo, lp = n + b;                          // an event owns n_b + 1 local pointer entries
        const bool live = n0 + lane < n_nodes;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const int32_t* ptr_l = dir == 0 ? in_ptr_l : out_ptr_l;
            const int32_t* base = dir == 0 ? in_base : out_base;
            const ColT* col = dir == 0 ? in_col : out_col;
            int32_t* ptr = dir == 0 ? g.in_ptr : g.out_ptr;
            int32_t* eid = dir == 0 ? g.in_eid : g.out_eid;
            int32_t* endpoint = dir == 0 ? g.dst : g.src;
            const int end = __ldg(base + b) + __ldg(ptr_l + lp + 1);
            const int beg = live ? __ldg(base + b) + __ldg(ptr_l + lp) : end;       // spare lanes: an empty row behind the last node's
            if (live) {
                ptr[n] = beg;
                if (n == n_nodes - 1) ptr[n_nodes] = end;
            }
            const int r0 = __shfl_sync(FULL, beg, 0), r1 = __shfl_sync(FULL, end, 31);
            for (int k = r0 + lane; k < r1 + 31 - ((r1 - r0 + 31) & 31); k += 32) {      // whole warp trips
                int owner = 0;                                 // last lane whose row starts at or before k
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    const int cand = owner + s;
                    const int cb = __shfl_sync(FULL, beg, cand & 31);
                    if (cand < 32 && cb <= k) owner = cand;
                }
                const int ob = __shfl_sync(FULL, b, owner);
                if (k < r1) {
                    const int slot = ob * e_max + (int)col[k];
                    endpoint[slot] = n0 + owner;
                    eid[k] = slot;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
assemble_nbr_kernel(const int n_in, const int n_out, GnnsegGraphMut g) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < max(n_in, n_out); k += gridDim.x * blockDim.x) {
        if (k < n_in) g.in_nbr[k] = g.src[g.in_eid[k]];
        if (k < n_out) g.out_nbr[k] = g.dst[g.out_eid[k]];
    }
}

int assemble_batch(const int32_t* meta, int B, int n_nodes, int e_max, int n_in, int n_out, const int32_t* in_ptr_l,
                   const int32_t* out_ptr_l, const void* in_col, const void* out_col, int col_bytes,
                   const GnnsegGraphMut& g, cudaStream_t st, bool lean) {
    const size_t n_slots = (size_t)B * e_max;
    if (n_slots > 0) {
        // padding slots (and half edges) carry -1
        if (cudaMemsetAsync(g.src, 0xFF, n_slots * 4, st) != cudaSuccess || cudaMemsetAsync(g.dst, 0xFF, n_slots * 4, st) != cudaSuccess)
            return GNNSEG_ECUDA;
        if (!lean && (cudaMemsetAsync(g.in_pos, 0xFF, n_slots * 4, st) != cudaSuccess || cudaMemsetAsync(g.out_pos, 0xFF, n_slots * 4, st) != cudaSuccess))
            return GNNSEG_ECUDA;
    }
    if (n_nodes == 0) {
        fill_i32_kernel<<<1, 32, 0, st>>>(g.in_ptr, 0, 1);
        fill_i32_kernel<<<1, 32, 0, st>>>(g.out_ptr, 0, 1);
        return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
    }
    const int grid = grid_for(n_nodes, 256, 148 * 8);
    auto rows = [&](auto kern, auto* ic, auto* oc) { kern<<<grid, 256, 0, st>>>(meta, B, n_nodes, e_max, in_ptr_l, out_ptr_l, ic, oc, g); };
    if (col_bytes == 2) {
        const uint16_t* ic = static_cast<const uint16_t*>(in_col), *oc = static_cast<const uint16_t*>(out_col);
        if (lean) rows(assemble_rows_warp_kernel<uint16_t>, ic, oc); else rows(assemble_rows_kernel<uint16_t, false>, ic, oc);
    } else {
        const int32_t* ic = static_cast<const int32_t*>(in_col), *oc = static_cast<const int32_t*>(out_col);
        if (lean) rows(assemble_rows_warp_kernel<int32_t>, ic, oc); else rows(assemble_rows_kernel<int32_t, false>, ic, oc);
    }
    if (!lean && (n_in > 0 || n_out > 0))
        assemble_nbr_kernel<<<grid_for(n_in > n_out ? n_in : n_out, 256, 148 * 8), 256, 0, st>>>(n_in, n_out, g);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
