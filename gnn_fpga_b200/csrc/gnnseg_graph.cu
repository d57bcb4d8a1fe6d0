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
__global__ void histogram_kernel(const int32_t* __restrict__ key, const int n_slots,
                                 const int n_nodes, int32_t* __restrict__ count) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slots; j += gridDim.x * blockDim.x) {
        const int k = __ldg(key + j);
        if (k >= 0 && k < n_nodes) atomicAdd(count + k, 1);
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
scan_local_kernel(const int32_t* __restrict__ count, const int n, int32_t* __restrict__ ptr,
                  int32_t* __restrict__ block_sums) {
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        v[i] = (base + i < n) ? count[base + i] : 0;
        sum += v[i];
    }
    int total;
    int off = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        if (base + i < n) ptr[base + i] = off;
        off += v[i];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(int32_t* __restrict__ block_sums, const int n_blocks, int32_t* __restrict__ grand_total) {
    int carry = 0;
    for (int base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int32_t* __restrict__ ptr, const int n, const int32_t* __restrict__ block_sums,
                const int32_t* __restrict__ grand_total, int32_t* __restrict__ cursor) {
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
    const int off = block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i)
        if (base + i < n) {
            const int p = ptr[base + i] + off;
            ptr[base + i] = p;
            cursor[base + i] = p;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) ptr[n] = *grand_total;
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ key, const int n_slots,
                                const int n_nodes, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ eid) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slots; j += gridDim.x * blockDim.x) {
        const int k = __ldg(key + j);
        if (k >= 0 && k < n_nodes) eid[atomicAdd(cursor + k, 1)] = j;
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

__global__ void csr_sort_rows_kernel(const int32_t* __restrict__ ptr, const int n_nodes,
                                     int32_t* __restrict__ eid) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        const int beg = ptr[n], len = ptr[n + 1] - beg;
        int32_t* a = eid + beg;
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

__global__ void csr_nbr_kernel(const int32_t* __restrict__ eid, const int32_t* __restrict__ ptr,
                               const int n_nodes, const int32_t* __restrict__ other,
                               int32_t* __restrict__ nbr, int32_t* __restrict__ pos) {
    const int n_used = ptr[n_nodes];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_used; s += gridDim.x * blockDim.x) {
        const int j = eid[s];
        nbr[s] = __ldg(other + j);
        if (pos) pos[j] = s;       // inverse map (pos was filled with -1)
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

int build_csr(const int32_t* key, const int32_t* other, int n_slots, int n_nodes, int32_t* ptr,
              int32_t* eid, int32_t* nbr, int32_t* pos, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < csr_workspace_bytes(n_nodes, n_slots)) return GNNSEG_EWORKSPACE;
    int32_t* count = static_cast<int32_t*>(ws);
    int32_t* cursor = count + (n_nodes + 1);
    int32_t* block_sums = cursor + (n_nodes + 1);
    const int n_blocks = (n_nodes + SCAN_ITEMS - 1) / SCAN_ITEMS;
    int32_t* grand_total = block_sums + n_blocks + 1;

    fill_i32_kernel<<<grid_for(n_nodes + 1, 256, 4096), 256, 0, st>>>(count, 0, (size_t)n_nodes + 1);
    if (pos && n_slots > 0) fill_i32_kernel<<<grid_for(n_slots, 256, 4096), 256, 0, st>>>(pos, -1, (size_t)n_slots);
    if (n_slots > 0) histogram_kernel<<<grid_for(n_slots, 256, 4096), 256, 0, st>>>(key, n_slots, n_nodes, count);
    if (n_blocks > 0) {
        scan_local_kernel<<<n_blocks, SCAN_THREADS, 0, st>>>(count, n_nodes, ptr, block_sums);
        scan_block_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, n_blocks, grand_total);
        scan_add_kernel<<<n_blocks, SCAN_THREADS, 0, st>>>(ptr, n_nodes, block_sums, grand_total, cursor);
    } else {
        fill_i32_kernel<<<1, 32, 0, st>>>(ptr, 0, 1);
    }
    if (n_slots > 0 && n_nodes > 0) {
        csr_fill_kernel<<<grid_for(n_slots, 256, 4096), 256, 0, st>>>(key, n_slots, n_nodes, cursor, eid);
        csr_sort_rows_kernel<<<grid_for(n_nodes, 128, 8192), 128, 0, st>>>(ptr, n_nodes, eid);
        csr_nbr_kernel<<<grid_for(n_slots, 256, 4096), 256, 0, st>>>(eid, ptr, n_nodes, other, nbr, pos);
    }
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace gnnseg
