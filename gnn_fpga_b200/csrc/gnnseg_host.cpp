// Host-side batch packing: SparseGraph tuples -> flattened int32 endpoints, no densifying.
// Replaces graph_from_sparse (gnn/graph.py:28-35) + merge_graphs + np_to_torch
// (gnn/trainSegmentClassifier.py:35-44,66-95) in front of the device path.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "gnnseg.h"

extern "C" int gnnseg_pack_sparse_batch_host(
    int B, int F, int e_max, const float* const* X_host, const int64_t* n_nodes_host,
    const int64_t* const* Ri_rows_host, const int64_t* const* Ri_cols_host,
    const int64_t* const* Ro_rows_host, const int64_t* const* Ro_cols_host,
    const int64_t* n_in_host, const int64_t* n_out_host, float* X_out_host, int32_t* src_host,
    int32_t* dst_host, int n_threads) {
    if (B < 0 || F < 1 || e_max < 0) return GNNSEG_EINVAL;
    if (B == 0) return GNNSEG_OK;
    if (!X_host || !n_nodes_host || !Ri_rows_host || !Ri_cols_host || !Ro_rows_host || !Ro_cols_host ||
        !n_in_host || !n_out_host || !X_out_host)
        return GNNSEG_EINVAL;
    if ((int64_t)B * e_max > 0 && (!src_host || !dst_host)) return GNNSEG_EINVAL;

    std::vector<int64_t> node_off(B + 1, 0);
    for (int b = 0; b < B; ++b) {
        if (n_nodes_host[b] < 0 || n_in_host[b] < 0 || n_out_host[b] < 0) return GNNSEG_EINVAL;
        node_off[b + 1] = node_off[b] + n_nodes_host[b];
    }
    if (node_off[B] > 0x7fffffffLL || (int64_t)B * e_max > 0x7fffffffLL) return GNNSEG_EINVAL;

    std::atomic<int> bad(0);
    // default: all cores but two (the caller's launching thread and the CUDA driver need them;
    // an oversubscribed team doubles the per-batch time)
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency() - 2;
    nt = std::max(1, std::min(nt, std::min(B, 64)));
    // OpenMP keeps its worker team alive between calls (creating 16 std::threads per batch
    // cost a quarter of the packing time)
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int b = 0; b < B; ++b) {
        {
            const int64_t nn = n_nodes_host[b], off = node_off[b];
            if (nn > 0) std::memcpy(X_out_host + off * F, X_host[b], sizeof(float) * nn * F);
            int32_t* src = src_host + (int64_t)b * e_max;
            int32_t* dst = dst_host + (int64_t)b * e_max;
            std::fill(src, src + e_max, -1);
            std::fill(dst, dst + e_max, -1);
            const int64_t* rr = Ri_rows_host[b];
            const int64_t* rc = Ri_cols_host[b];
            for (int64_t i = 0; i < n_in_host[b]; ++i) {
                const int64_t r = rr[i], c = rc[i];
                if (r < 0 || r >= nn || c < 0 || c >= e_max) { bad.store(1); continue; }
                dst[c] = (int32_t)(off + r);
            }
            rr = Ro_rows_host[b];
            rc = Ro_cols_host[b];
            for (int64_t i = 0; i < n_out_host[b]; ++i) {
                const int64_t r = rr[i], c = rc[i];
                if (r < 0 || r >= nn || c < 0 || c >= e_max) { bad.store(1); continue; }
                src[c] = (int32_t)(off + r);
            }
        }
    }
    return bad.load() ? GNNSEG_EINVAL : GNNSEG_OK;
}
