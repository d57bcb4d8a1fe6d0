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

    std::atomic<int> bad(0), dup(0);
    // default: all cores but two (the caller's launching thread and the CUDA driver need them;
    // an oversubscribed team doubles the per-batch time)
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency() - 2;
    nt = std::max(1, std::min(nt, 64));
    // Work items are (event, range) chunks, not events, so that one big event (a mu200 event has 1 M
    // edges) is packed by the whole team too.  Two phases: fill + copy, then scatter (a slot is written
    // by at most one Ri entry and one Ro entry, so the chunks of the second phase never collide).
    constexpr int64_t CHUNK = 1 << 15;
    struct Task { int b; int kind; int64_t beg, end; };
    std::vector<Task> fill, scatter;
    for (int b = 0; b < B; ++b) {
        for (int64_t i = 0; i < n_nodes_host[b]; i += CHUNK) fill.push_back({b, 0, i, std::min(n_nodes_host[b], i + CHUNK)});
        for (int64_t i = 0; i < e_max; i += 2 * CHUNK) fill.push_back({b, 1, i, std::min((int64_t)e_max, i + 2 * CHUNK)});
        for (int64_t i = 0; i < n_in_host[b]; i += CHUNK) scatter.push_back({b, 2, i, std::min(n_in_host[b], i + CHUNK)});
        for (int64_t i = 0; i < n_out_host[b]; i += CHUNK) scatter.push_back({b, 3, i, std::min(n_out_host[b], i + CHUNK)});
    }
    nt = std::max(1, std::min<int>(nt, (int)std::max(fill.size(), scatter.size())));
    // OpenMP keeps its worker team alive between calls (creating 16 std::threads per batch
    // cost a quarter of the packing time)
#pragma omp parallel num_threads(nt)
    {
#pragma omp for schedule(dynamic, 1)
        for (int t = 0; t < (int)fill.size(); ++t) {
            const Task& k = fill[t];
            if (k.kind == 0) {
                std::memcpy(X_out_host + (node_off[k.b] + k.beg) * F, X_host[k.b] + k.beg * F, sizeof(float) * (k.end - k.beg) * F);
            } else {
                std::fill(src_host + (int64_t)k.b * e_max + k.beg, src_host + (int64_t)k.b * e_max + k.end, -1);
                std::fill(dst_host + (int64_t)k.b * e_max + k.beg, dst_host + (int64_t)k.b * e_max + k.end, -1);
            }
        }   // implicit barrier: every slot holds -1 before the first endpoint is written
#pragma omp for schedule(dynamic, 1)
        for (int t = 0; t < (int)scatter.size(); ++t) {
            const Task& k = scatter[t];
            const int64_t nn = n_nodes_host[k.b], off = node_off[k.b];
            const int64_t* rr = k.kind == 2 ? Ri_rows_host[k.b] : Ro_rows_host[k.b];
            const int64_t* rc = k.kind == 2 ? Ri_cols_host[k.b] : Ro_cols_host[k.b];
            int32_t* out = (k.kind == 2 ? dst_host : src_host) + (int64_t)k.b * e_max;
            for (int64_t i = k.beg; i < k.end; ++i) {
                const int64_t r = rr[i], c = rc[i];
                if (r < 0 || r >= nn || c < 0 || c >= e_max) { bad.store(1); continue; }
                // a column listed twice (hyper-edge: the dense reference would sum two rows, gnn/model.py:71-72) is an
                // error, not last-writer-wins; the exchange also makes the two writers' race well defined
                if (__atomic_exchange_n(&out[c], (int32_t)(off + r), __ATOMIC_RELAXED) != -1) dup.store(1);
            }
        }
    }
    if (bad.load()) return GNNSEG_EINVAL;
    return dup.load() ? GNNSEG_EHYPEREDGE : GNNSEG_OK;
}
