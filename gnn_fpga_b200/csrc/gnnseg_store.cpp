// Host-side event store: the graphs of a data set, narrowed and laid out ONCE at load time so that a
// batch of consecutive events reaches the GPU with a handful of contiguous copies and no per-batch host
// work at all.  Replaces, together with gnnseg_assemble_batch (gnnseg_graph.cu), the per-batch
// graph_from_sparse + merge_graphs + np_to_torch of the reference's generator
// (gnn/graph.py:28-35, gnn/trainSegmentClassifier.py:35-44,66-111): the reference keeps its list of
// SparseGraph tuples from load_graphs (gnn/trainSegmentClassifier.py:129) and densifies a slice of it
// for every batch; here the list becomes one arena in pinned host memory.
//
// What the arena holds per event (events back to back, every array 256-byte aligned as a whole):
//   X        float32 (n, F)      node features, in INTERNAL node order (see `reorder`)
//   in_ptr   int32   (n + 1)     destination-CSR row pointer of the event, local (starts at 0)
//   out_ptr  int32   (n + 1)     source-CSR row pointer
//   in_col   u16/i32 (n_in)      Ri_cols in CSR order (np.nonzero order of gnn/graph.py:23-26 when the
//   out_col  u16/i32 (n_out)     node order is kept): the edge column of every incidence entry
//   y        float32 (n_y)       edge labels by column
//   perm     int32   (n)         internal position -> original node of the event
// Validation happens here, once: indices in range, and no column listed twice in Ri or in Ro (a
// hyper-edge: the dense reference would sum two rows, gnn/model.py:71-72; unsupported, as for the dense
// entry point's GNNSEG_BAD_HYPEREDGE).
//
// reorder: nodes of an event may be renumbered internally so that the endpoints of an edge sit close
// together in memory (the gathers of the forward then hit in L1 instead of L2).  The order is by the
// feature column along which edges are most local (smallest mean |x_src - x_dst| relative to the
// column's spread), picked from the data; scores are per edge column and do not see the renumbering.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "gnnseg.h"

namespace {

inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// Column along which edges are most local, or -1 if none is local enough to be worth a renumbering.
int pick_order_column(int B, int F, const float* const* X, const int64_t* n_nodes, const int64_t* const* Ri_rows,
                      const int64_t* const* Ri_cols, const int64_t* const* Ro_rows, const int64_t* const* Ro_cols,
                      const int64_t* n_in, const int64_t* n_out) {
    std::vector<double> edge_dev(F, 0.0), node_dev(F, 0.0);
    double n_edges = 0.0, n_pts = 0.0;
    const int sample = std::min(B, 8);
    std::vector<int64_t> dst;
    for (int b = 0; b < sample; ++b) {
        const int64_t n = n_nodes[b], e = n_in[b];
        if (n == 0 || e == 0) continue;
        dst.assign((size_t)e, -1);
        for (int64_t k = 0; k < n_in[b]; ++k) {
            const int64_t c = Ri_cols[b][k], r = Ri_rows[b][k];
            if (c >= 0 && c < e && r >= 0 && r < n) dst[(size_t)c] = r;
        }
        for (int64_t k = 0; k < n_out[b]; ++k) {
            const int64_t c = Ro_cols[b][k], s = Ro_rows[b][k];
            if (c < 0 || c >= e || s < 0 || s >= n || dst[(size_t)c] < 0) continue;
            const float* xs = X[b] + s * F;
            const float* xd = X[b] + dst[(size_t)c] * F;
            for (int f = 0; f < F; ++f) edge_dev[f] += std::fabs((double)xs[f] - (double)xd[f]);
            n_edges += 1.0;
        }
        std::vector<double> mean(F, 0.0);
        for (int64_t i = 0; i < n; ++i)
            for (int f = 0; f < F; ++f) mean[f] += X[b][i * F + f];
        for (int f = 0; f < F; ++f) mean[f] /= (double)n;
        for (int64_t i = 0; i < n; ++i)
            for (int f = 0; f < F; ++f) node_dev[f] += std::fabs((double)X[b][i * F + f] - mean[f]);
        n_pts += (double)n;
    }
    if (n_edges == 0.0 || n_pts == 0.0) return -1;
    int best = -1;
    double best_ratio = 0.25;        // edges must span less than a quarter of the column's spread
    for (int f = 0; f < F; ++f) {
        const double spread = node_dev[f] / n_pts;
        if (spread <= 0.0) continue;
        const double ratio = (edge_dev[f] / n_edges) / spread;
        if (ratio < best_ratio) { best_ratio = ratio; best = f; }
    }
    return best;
}

template <typename ColT>
int fill_event(const GnnsegStoreLayout& L, int b, int F, const float* X, int64_t n, const int64_t* rows_in,
               const int64_t* cols_in, int64_t n_in, const int64_t* rows_out, const int64_t* cols_out, int64_t n_out,
               const float* y, int64_t n_y, int64_t n_edges, int order_col, char* arena) {
    const int64_t* node_off = reinterpret_cast<const int64_t*>(arena + L.o_node_off);
    const int64_t* in_off = reinterpret_cast<const int64_t*>(arena + L.o_in_off);
    const int64_t* out_off = reinterpret_cast<const int64_t*>(arena + L.o_out_off);
    const int64_t* y_off = reinterpret_cast<const int64_t*>(arena + L.o_y_off);
    float* Xo = reinterpret_cast<float*>(arena + L.o_X) + node_off[b] * F;
    int32_t* perm = reinterpret_cast<int32_t*>(arena + L.o_perm) + node_off[b];
    int32_t* ptr[2] = {reinterpret_cast<int32_t*>(arena + L.o_in_ptr) + node_off[b] + b,
                       reinterpret_cast<int32_t*>(arena + L.o_out_ptr) + node_off[b] + b};
    ColT* col[2] = {reinterpret_cast<ColT*>(arena + L.o_in_col) + in_off[b], reinterpret_cast<ColT*>(arena + L.o_out_col) + out_off[b]};
    const int64_t* rows[2] = {rows_in, rows_out};
    const int64_t* cols[2] = {cols_in, cols_out};
    const int64_t cnt[2] = {n_in, n_out};
    const int64_t e = n_edges;                                 // columns of the event (graph_from_sparse: len(Ri_rows))

    // internal order
    std::vector<int32_t> rank((size_t)n);
    std::iota(perm, perm + n, 0);
    if (order_col >= 0 && n > 1)
        std::stable_sort(perm, perm + n, [&](int32_t a, int32_t c) { return X[(int64_t)a * F + order_col] < X[(int64_t)c * F + order_col]; });
    for (int64_t i = 0; i < n; ++i) rank[(size_t)perm[i]] = (int32_t)i;
    for (int64_t i = 0; i < n; ++i) std::memcpy(Xo + i * F, X + (int64_t)perm[i] * F, sizeof(float) * F);
    if (n_y > 0) std::memcpy(reinterpret_cast<float*>(arena + L.o_y) + y_off[b], y, sizeof(float) * n_y);

    std::vector<uint8_t> seen((size_t)e);
    std::vector<int32_t> cursor((size_t)n + 1);
    for (int d = 0; d < 2; ++d) {
        std::fill(seen.begin(), seen.end(), 0);
        std::fill(ptr[d], ptr[d] + n + 1, 0);
        bool canonical = true;                                 // entries sorted by (internal row, column)?
        int64_t prev_r = -1, prev_c = -1;
        for (int64_t k = 0; k < cnt[d]; ++k) {
            const int64_t r = rows[d][k], c = cols[d][k];
            if (r < 0 || r >= n || c < 0 || c >= e) return GNNSEG_EINVAL;
            if (seen[(size_t)c]) return GNNSEG_EHYPEREDGE;
            seen[(size_t)c] = 1;
            const int64_t ri = rank[(size_t)r];
            if (ri < prev_r || (ri == prev_r && c < prev_c)) canonical = false;
            prev_r = ri; prev_c = c;
            ptr[d][ri + 1] += 1;
        }
        for (int64_t i = 0; i < n; ++i) ptr[d][i + 1] += ptr[d][i];
        if (canonical) {
            for (int64_t k = 0; k < cnt[d]; ++k) col[d][k] = (ColT)cols[d][k];
        } else {
            std::copy(ptr[d], ptr[d] + n + 1, cursor.begin());
            for (int64_t k = 0; k < cnt[d]; ++k) col[d][cursor[(size_t)rank[(size_t)rows[d][k]]]++] = (ColT)cols[d][k];
            for (int64_t i = 0; i < n; ++i) std::sort(col[d] + ptr[d][i], col[d] + ptr[d][i + 1]);   // ascending column = np.nonzero order
        }
    }
    return GNNSEG_OK;
}

}  // namespace

extern "C" int gnnseg_store_plan_host(int n_events, int F, const int64_t* n_nodes_host, const int64_t* n_in_host,
                                      const int64_t* n_out_host, const int64_t* n_y_host, const int64_t* n_edges_host,
                                      GnnsegStoreLayout* L) {
    if (n_events < 0 || F < 1 || !L || (n_events > 0 && (!n_nodes_host || !n_in_host || !n_out_host))) return GNNSEG_EINVAL;
    int64_t tn = 0, ti = 0, to = 0, ty = 0, emax = 0;
    for (int b = 0; b < n_events; ++b) {
        const int64_t ny = n_y_host ? n_y_host[b] : 0;
        if (n_nodes_host[b] < 0 || n_in_host[b] < 0 || n_out_host[b] < 0 || ny < 0) return GNNSEG_EINVAL;
        if (n_nodes_host[b] > 0x7ffffff0LL || n_in_host[b] > 0x7ffffff0LL || n_out_host[b] > 0x7ffffff0LL) return GNNSEG_EINVAL;
        tn += n_nodes_host[b]; ti += n_in_host[b]; to += n_out_host[b]; ty += ny;
        const int64_t ne = n_edges_host ? n_edges_host[b] : n_in_host[b];
        if (ne < n_in_host[b] || ne > 0x7ffffff0LL) return GNNSEG_EINVAL;
        emax = std::max(emax, ne);
    }
    std::memset(L, 0, sizeof(*L));
    L->n_events = n_events; L->n_features = F;
    L->total_nodes = tn; L->total_in = ti; L->total_out = to; L->total_y = ty;
    L->col_bytes = emax <= 65536 ? 2 : 4;                      // column ids are < e <= 65536
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += align256(bytes); return o; };
    L->o_node_off = take(8 * ((int64_t)n_events + 1));
    L->o_in_off = take(8 * ((int64_t)n_events + 1));
    L->o_out_off = take(8 * ((int64_t)n_events + 1));
    L->o_y_off = take(8 * ((int64_t)n_events + 1));
    L->o_X = take(4 * tn * F);
    L->o_in_ptr = take(4 * (tn + n_events));
    L->o_out_ptr = take(4 * (tn + n_events));
    L->o_in_col = take((int64_t)L->col_bytes * ti);
    L->o_out_col = take((int64_t)L->col_bytes * to);
    L->o_y = take(4 * ty);
    L->o_perm = take(4 * tn);
    L->o_n_edges = take(8 * (int64_t)n_events);
    L->bytes = off;
    return GNNSEG_OK;
}

extern "C" int gnnseg_store_fill_host(const GnnsegStoreLayout* L, const float* const* X_host, const int64_t* n_nodes_host,
                                      const int64_t* const* Ri_rows_host, const int64_t* const* Ri_cols_host,
                                      const int64_t* const* Ro_rows_host, const int64_t* const* Ro_cols_host,
                                      const int64_t* n_in_host, const int64_t* n_out_host, const float* const* y_host,
                                      const int64_t* n_y_host, const int64_t* n_edges_host, int reorder, int n_threads,
                                      void* arena_host, int32_t* info_host) {
    if (!L || !arena_host) return GNNSEG_EINVAL;
    const int B = (int)L->n_events, F = L->n_features;
    if (info_host) { info_host[0] = -1; info_host[1] = -1; }
    if (B == 0) return GNNSEG_OK;
    if (!X_host || !n_nodes_host || !Ri_rows_host || !Ri_cols_host || !Ro_rows_host || !Ro_cols_host || !n_in_host || !n_out_host)
        return GNNSEG_EINVAL;
    char* arena = static_cast<char*>(arena_host);
    int64_t* node_off = reinterpret_cast<int64_t*>(arena + L->o_node_off);
    int64_t* in_off = reinterpret_cast<int64_t*>(arena + L->o_in_off);
    int64_t* out_off = reinterpret_cast<int64_t*>(arena + L->o_out_off);
    int64_t* y_off = reinterpret_cast<int64_t*>(arena + L->o_y_off);
    int64_t* n_edges_arena = reinterpret_cast<int64_t*>(arena + L->o_n_edges);
    node_off[0] = in_off[0] = out_off[0] = y_off[0] = 0;
    for (int b = 0; b < B; ++b) {
        n_edges_arena[b] = n_edges_host ? n_edges_host[b] : n_in_host[b];
        node_off[b + 1] = node_off[b] + n_nodes_host[b];
        in_off[b + 1] = in_off[b] + n_in_host[b];
        out_off[b + 1] = out_off[b] + n_out_host[b];
        y_off[b + 1] = y_off[b] + (n_y_host ? n_y_host[b] : 0);
    }
    if (node_off[B] != L->total_nodes || in_off[B] != L->total_in || out_off[B] != L->total_out || y_off[B] != L->total_y)
        return GNNSEG_EINVAL;                                  // not the sizes the layout was planned for
    // reorder: 0 = keep the node order, 1 = always (if a local column exists), 2 = auto (events of >= 1024 nodes)
    int order_col = -1;
    if (reorder == 1 || (reorder == 2 && L->total_nodes >= 1024LL * B))
        order_col = pick_order_column(B, F, X_host, n_nodes_host, Ri_rows_host, Ri_cols_host, Ro_rows_host, Ro_cols_host,
                                      n_in_host, n_out_host);
    if (info_host) info_host[1] = order_col;

    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(B, 64)));
    std::atomic<int> next(0), rc(GNNSEG_OK), bad_event(-1);
    auto work = [&]() {
        for (int b = next.fetch_add(1); b < B; b = next.fetch_add(1)) {
            const float* y = (y_host && n_y_host && n_y_host[b] > 0) ? y_host[b] : nullptr;
            const int64_t ny = y ? n_y_host[b] : 0;
            const int64_t ne = n_edges_host ? n_edges_host[b] : n_in_host[b];
            const bool fits = L->col_bytes == 4 || ne <= 65536;             // the layout was planned for these edge counts
            const int r = !fits ? GNNSEG_EINVAL : L->col_bytes == 2
                ? fill_event<uint16_t>(*L, b, F, X_host[b], n_nodes_host[b], Ri_rows_host[b], Ri_cols_host[b], n_in_host[b],
                                       Ro_rows_host[b], Ro_cols_host[b], n_out_host[b], y, ny, ne, order_col, arena)
                : fill_event<int32_t>(*L, b, F, X_host[b], n_nodes_host[b], Ri_rows_host[b], Ri_cols_host[b], n_in_host[b],
                                      Ro_rows_host[b], Ro_cols_host[b], n_out_host[b], y, ny, ne, order_col, arena);
            if (r != GNNSEG_OK) {
                int expect = GNNSEG_OK;
                if (rc.compare_exchange_strong(expect, r)) bad_event.store(b);
            }
        }
    };
    std::vector<std::thread> team;
    for (int t = 1; t < nt; ++t) team.emplace_back(work);
    work();
    for (auto& t : team) t.join();
    if (info_host) info_host[0] = bad_event.load();
    return rc.load();
}
