// extern "C" entry points declared in include/gnnseg.h.  Argument checking happens here, on
// the host; the kernels live in gnnseg_forward.cu / gnnseg_graph.cu.
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {
int input_step(const float*, const float*, int, int, int, float*, float*, float*, float*, cudaStream_t);
int edge_step(const float*, const GnnsegGraph*, const float*, int, float*, float*, float*, cudaStream_t);
int node_step(const float*, const GnnsegGraph*, const float*, const float*, const float*, const float*, int, float*, float*, int, float*, float*, cudaStream_t);
int pack_weights(const GnnsegParams*, int, int, float*, cudaStream_t);
int pack_head_weights(const GnnsegParams*, const float*, const float*, int, int, float*, cudaStream_t);
int node_head(const float*, int, int, float*, cudaStream_t);
int node_gather_step(const GnnsegGraph*, const float*, const float*, const float*, int, float*, int, cudaStream_t);
int node_mlp_step(const float*, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int dense_to_edges(const float*, const float*, int, int, int, int32_t*, int32_t*, int32_t*, cudaStream_t);
size_t csr_workspace_bytes(int, int);
int build_csr(const int32_t*, const int32_t*, int, int, int32_t*, int32_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t);
int build_graph(const int32_t*, const int32_t*, int, int, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t);
int fused_gather_step(const float*, const GnnsegGraph*, const float*, int, float*, int, cudaStream_t);
int edge_final_step(const float*, const GnnsegGraph*, const float*, int, int, int, int, float*, cudaStream_t);
int build_adjacency(const GnnsegGraph*, int32_t*, int32_t*, int32_t*, cudaStream_t, bool from_endpoints = false);
size_t adjacency_entries(int, int);
int launch_input_tc32_ex(const float*, const float*, int, int, float*, const ProjOut&, float*, cudaStream_t);
int launch_input_tc64_ex(const float*, const float*, int, int, float*, const ProjOut&, float*, cudaStream_t);
int launch_node_mlp_tc32_ex(const float*, const float*, const float*, int, int, const ProjOut&, float*, bool, cudaStream_t);
int launch_node_mlp_tc64_ex(const float*, const float*, const float*, int, int, const ProjOut&, float*, bool, cudaStream_t);
bool use_pdl(int);
bool use_pdl_mlp();
int assemble_batch(const int32_t*, int, int, int, int, int, const int32_t*, const int32_t*, const void*, const void*, int,
                   const GnnsegGraphMut&, cudaStream_t, bool lean = false);
// gnnseg_backward.cu
struct GradOut {
    float* w_in; float* b_in; float* w_e1; float* b_e1; float* w_e2; float* b_e2;
    float* w_n1; float* b_n1; float* w_n2; float* b_n2;
    const float* m_e1; const float* m_e2; const float* m_n1; const float* m_n2;
};
struct TrainState {
    const float* x4;
    const float* const* Hs; const float* const* h1s; const float* const* Ps; const float* const* Qs;
    const float* const* e_in; const float* const* e_out;
    float* dg; float* dproj; float* ds_in; float* ds_out; float* partE; float* partN; float* dz;
};
size_t edge_part_floats(int h);
size_t node_part_floats(int h);
int backward(const float*, const GnnsegGraph*, int, int, int, const float*, const TrainState&, const GradOut&, const float*,
             float*, float*, cudaStream_t);
size_t segments_workspace_bytes(int, int);
size_t segments_batch_workspace_bytes(int, int, int);
int build_segments_batch(const int32_t*, const void*, const void*, const void*, int, const int64_t*, int, const int32_t*, int, int,
                         const int32_t*, int, int, double, double, double, int, int, int32_t*, int32_t*, float*, int32_t*, void*,
                         cudaStream_t);
int build_segments(const int32_t*, const void*, const void*, const void*, int, const int64_t*, int, const int32_t*, int, int,
                   double, double, double, int, int, int, int32_t*, int32_t*, float*, int32_t*, void*, cudaStream_t);
int scale_features(const void*, const void*, const void*, int, int, double, double, double, float*, cudaStream_t);
int bce_loss(const float*, const float*, const float*, int, float*, float*, float*, cudaStream_t);
int l1_penalty(const GnnsegParams*, int, int, float, float*, const GnnsegGrads*, cudaStream_t);
int adam_step(float*, const float*, float*, float*, int, int, float, float, float, float, float, cudaStream_t);
}  // namespace gnnseg

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct FwdWorkspace {
    float* x4;
    float* p;
    float* q[2];
    float* e_in;
    float* e_out;
    float* s[2];          // fused path: state rows (alias q / e_in / e_out)
    int32_t* status;      // fused path: range flag word
    size_t bytes;
};

// X4 | P | Q0 | Q1 | e_in | e_out, each 256-byte aligned.  The fused inference path (gnnseg_fused.cu) uses
// the same block as X4 | (P unused) | S0 | S1 | status word, with S = n_nodes + 1 state rows of 5h floats.
FwdWorkspace carve(void* ws, int n_nodes, int n_slots, int h) {
    FwdWorkspace w;
    const size_t x_b = align_up((size_t)n_nodes * 4 * 4, 256);
    const size_t p_b = align_up((size_t)n_nodes * 2 * h * 4, 256);
    const size_t q_b = align_up((size_t)n_nodes * 3 * h * 4, 256);
    const size_t e_b = align_up((size_t)n_slots * 4, 256);
    const size_t s_b = align_up(((size_t)n_nodes + 1) * 5 * h * 4, 256);      // + the extra row absent neighbours point at
    char* base = static_cast<char*>(ws);
    w.x4 = reinterpret_cast<float*>(base);
    w.p = reinterpret_cast<float*>(base + x_b);
    w.q[0] = reinterpret_cast<float*>(base + x_b + p_b);
    w.q[1] = reinterpret_cast<float*>(base + x_b + p_b + q_b);
    w.e_in = reinterpret_cast<float*>(base + x_b + p_b + 2 * q_b);
    w.e_out = reinterpret_cast<float*>(base + x_b + p_b + 2 * q_b + e_b);
    w.s[0] = reinterpret_cast<float*>(base + x_b + p_b);
    w.s[1] = reinterpret_cast<float*>(base + x_b + p_b + s_b);
    w.status = reinterpret_cast<int32_t*>(base + x_b + p_b + 2 * s_b);
    const size_t steps = 2 * q_b + 2 * e_b, fused = 2 * s_b + 256;
    w.bytes = x_b + p_b + (steps > fused ? steps : fused);
    return w;
}

inline bool fused_width(int h) { return h == 32 || h == 64; }
inline bool exact_env() {
    static const bool v = [] { const char* e = getenv("GNNSEG_EXACT"); return e && e[0] != '0'; }();
    return v;
}
inline int state_input(const float* blob, const float* X, int n, int F, int h, float* X4, const gnnseg::ProjOut& out, cudaStream_t st) {
    return h == 32 ? gnnseg::launch_input_tc32_ex(blob, X, n, F, X4, out, nullptr, st)
                   : gnnseg::launch_input_tc64_ex(blob, X, n, F, X4, out, nullptr, st);
}
inline int state_mlp(const float* blob, const float* X4, const float* h1, int ld_h1, int n, int h, const gnnseg::ProjOut& out,
                     bool pdl, cudaStream_t st) {
    return h == 32 ? gnnseg::launch_node_mlp_tc32_ex(blob, X4, h1, ld_h1, n, out, nullptr, pdl, st)
                   : gnnseg::launch_node_mlp_tc64_ex(blob, X4, h1, ld_h1, n, out, nullptr, pdl, st);
}

constexpr int MAX_ITERS = 64;
}  // namespace
namespace gnnseg {
// does gnnseg_forward_ex take the fused inference path for this width and these flags (given adjacency lists)?
bool forward_is_fused(int h, int flags) { return fused_width(h) && !(flags & GNNSEG_FWD_EXACT) && !exact_env(); }
}
namespace {

// Training workspace: everything the forward leaves behind for the backward pass, then the
// backward scratch.  X4 | H[0..T] | h1[0..T-1] | P[0..T] | Q[0..max(T,1)-1] | e_in[T] | e_out[T] |
// dg | dproj | ds_in | ds_out | partE | partN, each 256-byte aligned.
struct TrainWorkspace {
    float* x4;
    float* H[MAX_ITERS + 1];
    float* h1[MAX_ITERS];
    float* P[MAX_ITERS + 1];
    float* Q[MAX_ITERS];
    float* e_in[MAX_ITERS];
    float* e_out[MAX_ITERS];
    float* dg; float* dproj; float* ds_in; float* ds_out; float* partE; float* partN; float* dz;
    size_t bytes;
};

TrainWorkspace carve_train(void* ws, int n_nodes, int n_slots, int h, int T) {
    TrainWorkspace w;
    char* base = static_cast<char*>(ws);
    size_t off = 0;
    auto take = [&](size_t floats) {
        float* p = reinterpret_cast<float*>(base + off);
        off += align_up(floats * 4, 256);
        return p;
    };
    const size_t n = (size_t)n_nodes, m = (size_t)n_slots;
    w.x4 = take(n * 4);
    for (int t = 0; t <= T; ++t) w.H[t] = take(n * h);
    for (int t = 0; t < T; ++t) w.h1[t] = take(n * h);
    for (int t = 0; t <= T; ++t) w.P[t] = take(n * 2 * h);
    for (int t = 0; t < (T > 0 ? T : 1); ++t) w.Q[t] = take(n * 3 * h);
    for (int t = 0; t < T; ++t) w.e_in[t] = take(m);
    for (int t = 0; t < T; ++t) w.e_out[t] = take(m);
    w.dg = take(n * h);
    w.dproj = take(n * 5 * h);
    w.ds_in = take(m);
    w.ds_out = take(m);
    w.partE = take(gnnseg::edge_part_floats(h));
    w.partN = take(gnnseg::node_part_floats(h));
    w.dz = take(n * h);
    w.bytes = off;
    return w;
}

inline bool graph_ok(const GnnsegGraph* g) {
    if (!g || g->n_nodes < 0 || g->n_slots < 0) return false;
    if (g->n_slots > 0 && (!g->src || !g->dst)) return false;
    return true;
}
inline bool csr_ok(const GnnsegGraph* g) {
    if (!graph_ok(g)) return false;
    if (!g->in_ptr || !g->out_ptr) return false;
    if (g->n_slots > 0 && (!g->in_eid || !g->in_nbr || !g->out_eid || !g->out_nbr || !g->in_pos || !g->out_pos)) return false;
    return true;
}

}  // namespace

extern "C" {

int gnnseg_abi_version(void) { return GNNSEG_ABI_VERSION; }

const char* gnnseg_strerror(int code) {
    switch (code) {
        case GNNSEG_OK:           return "ok";
        case GNNSEG_EINVAL:       return "invalid argument";
        case GNNSEG_EUNSUPPORTED: return "unsupported (input_dim, hidden_dim): need input_dim in 1..4 and hidden_dim in {4,8,16,32,64}";
        case GNNSEG_EWORKSPACE:   return "workspace too small";
        case GNNSEG_ECUDA:        return "CUDA runtime error";
        case GNNSEG_ENODEVICE:    return "no CUDA device";
        case GNNSEG_EIO:          return "file cannot be opened or mapped";
        case GNNSEG_EFORMAT:      return "not an .npz graph file (np.savez of X, Ri_rows, Ri_cols, Ro_rows, Ro_cols, y)";
        case GNNSEG_EHYPEREDGE:   return "a column of Ri or Ro is listed more than once (hyper-edge)";
        default:                  return "unknown gnnseg error";
    }
}

int gnnseg_supported(int F, int h) {
    return (F >= 1 && F <= 4 && (h == 4 || h == 8 || h == 16 || h == 32 || h == 64)) ? 1 : 0;
}

int gnnseg_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GNNSEG_ENODEVICE;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return GNNSEG_ECUDA;
    return n;
}

size_t gnnseg_weights_floats(int F, int h) {
    return gnnseg_supported(F, h) ? (size_t)gnnseg::blob_total(h) : 0;
}

int gnnseg_pack_weights(const GnnsegParams* p, int F, int h, float* blob, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!p || !blob || !p->w_in || !p->b_in || !p->w_e1 || !p->b_e1 || !p->w_e2 || !p->b_e2 ||
        !p->w_n1 || !p->b_n1 || !p->w_n2 || !p->b_n2)
        return GNNSEG_EINVAL;
    return gnnseg::pack_weights(p, F, h, blob, static_cast<cudaStream_t>(stream));
}

int gnnseg_pack_node_head(const GnnsegParams* p, const float* w_out, const float* b_out, int F, int h,
                          float* head_blob, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!p || !head_blob || !w_out || !b_out || !p->w_in || !p->b_in || !p->w_n1 || !p->b_n1 || !p->w_n2 || !p->b_n2 ||
        !p->w_e2 || !p->b_e2)
        return GNNSEG_EINVAL;
    return gnnseg::pack_head_weights(p, w_out, b_out, F, h, head_blob, static_cast<cudaStream_t>(stream));
}

int gnnseg_dense_to_edges(const float* Ri, const float* Ro, int B, int N, int E, int32_t* src,
                          int32_t* dst, int32_t* err_flag, void* stream) {
    if (B < 0 || N < 0 || E < 0 || !err_flag) return GNNSEG_EINVAL;
    if ((long long)B * N > 0x7fffffffLL || (long long)B * E > 0x7fffffffLL) return GNNSEG_EINVAL;
    if ((long long)B * E > 0 && (!src || !dst)) return GNNSEG_EINVAL;
    if ((long long)B * N * E > 0 && (!Ri || !Ro)) return GNNSEG_EINVAL;
    return gnnseg::dense_to_edges(Ri, Ro, B, N, E, src, dst, err_flag, static_cast<cudaStream_t>(stream));
}

size_t gnnseg_csr_workspace_bytes(int n_nodes, int n_slots) {
    if (n_nodes < 0 || n_slots < 0) return 0;
    return gnnseg::csr_workspace_bytes(n_nodes, n_slots);
}

int gnnseg_build_csr(const int32_t* key, const int32_t* other, int n_slots, int n_nodes,
                     int32_t* ptr, int32_t* eid, int32_t* nbr, int32_t* pos, void* ws,
                     size_t ws_bytes, void* stream) {
    if (n_slots < 0 || n_nodes < 0 || !ptr || !ws) return GNNSEG_EINVAL;
    if (n_slots > 0 && (!key || !other || !eid || !nbr)) return GNNSEG_EINVAL;
    return gnnseg::build_csr(key, other, n_slots, n_nodes, ptr, eid, nbr, pos, ws, ws_bytes,
                             static_cast<cudaStream_t>(stream));
}

int gnnseg_build_graph(const int32_t* src, const int32_t* dst, int n_slots, int n_nodes,
                       int32_t* in_ptr, int32_t* in_eid, int32_t* in_nbr, int32_t* in_pos,
                       int32_t* out_ptr, int32_t* out_eid, int32_t* out_nbr, int32_t* out_pos,
                       void* ws, size_t ws_bytes, void* stream) {
    if (n_slots < 0 || n_nodes < 0 || !in_ptr || !out_ptr || !ws) return GNNSEG_EINVAL;
    if (n_slots > 0 && (!src || !dst || !in_eid || !in_nbr || !out_eid || !out_nbr)) return GNNSEG_EINVAL;
    return gnnseg::build_graph(src, dst, n_slots, n_nodes, in_ptr, in_eid, in_nbr, in_pos, out_ptr, out_eid,
                               out_nbr, out_pos, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int gnnseg_assemble_batch(const int32_t* meta, int B, int n_nodes, int e_max, int n_in, int n_out,
                          const int32_t* in_ptr_local, const int32_t* out_ptr_local, const void* in_col,
                          const void* out_col, int col_bytes, int32_t* src, int32_t* dst, int32_t* in_ptr,
                          int32_t* in_eid, int32_t* in_nbr, int32_t* in_pos, int32_t* out_ptr, int32_t* out_eid,
                          int32_t* out_nbr, int32_t* out_pos, void* stream) {
    if (B < 0 || n_nodes < 0 || e_max < 0 || n_in < 0 || n_out < 0 || (col_bytes != 2 && col_bytes != 4)) return GNNSEG_EINVAL;
    if ((long long)B * e_max > 0x7fffffffLL || (long long)n_nodes + B > 0x7fffffffLL) return GNNSEG_EINVAL;
    if (!in_ptr || !out_ptr) return GNNSEG_EINVAL;
    if (n_nodes > 0 && (!meta || !in_ptr_local || !out_ptr_local)) return GNNSEG_EINVAL;
    if ((long long)B * e_max > 0 && (!src || !dst || !in_pos || !out_pos)) return GNNSEG_EINVAL;
    if (n_in > 0 && (!in_col || !in_eid || !in_nbr)) return GNNSEG_EINVAL;
    if (n_out > 0 && (!out_col || !out_eid || !out_nbr)) return GNNSEG_EINVAL;
    const gnnseg::GnnsegGraphMut g{src, dst, in_ptr, in_eid, in_nbr, in_pos, out_ptr, out_eid, out_nbr, out_pos};
    return gnnseg::assemble_batch(meta, B, n_nodes, e_max, n_in, n_out, in_ptr_local, out_ptr_local, in_col, out_col,
                                  col_bytes, g, static_cast<cudaStream_t>(stream));
}

size_t gnnseg_forward_workspace_bytes(int n_nodes, int n_slots, int F, int h) {
    if (!gnnseg_supported(F, h) || n_nodes < 0 || n_slots < 0) return 0;
    return carve(nullptr, n_nodes, n_slots, h).bytes + 256;
}

int gnnseg_input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* X4,
                      float* P, float* Q, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || n_nodes < 0 || (n_nodes > 0 && (!X || !X4 || !P || !Q))) return GNNSEG_EINVAL;
    return gnnseg::input_step(blob, X, n_nodes, F, h, X4, P, Q, nullptr, static_cast<cudaStream_t>(stream));
}

int gnnseg_edge_step(const float* blob, const GnnsegGraph* g, const float* P, int h, float* e,
                     float* e_in, float* e_out, void* stream) {
    if (!gnnseg_supported(1, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !graph_ok(g) || (g->n_nodes > 0 && !P)) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && ((e_in && !g->in_pos) || (e_out && !g->out_pos))) return GNNSEG_EINVAL;
    return gnnseg::edge_step(blob, g, P, h, e, e_in, e_out, static_cast<cudaStream_t>(stream));
}

int gnnseg_node_step(const float* blob, const GnnsegGraph* g, const float* X4, const float* Q_in,
                     const float* e_in, const float* e_out, int h, float* P_out, float* Q_out,
                     void* stream) {
    if (!gnnseg_supported(1, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !csr_ok(g)) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && (!X4 || !Q_in || !P_out)) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && (!e_in || !e_out)) return GNNSEG_EINVAL;
    return gnnseg::node_step(blob, g, X4, Q_in, e_in, e_out, h, P_out, Q_out, Q_out != nullptr, nullptr, nullptr,
                             static_cast<cudaStream_t>(stream));
}

int gnnseg_node_gather_step(const GnnsegGraph* g, const float* Q_in, const float* e_in, const float* e_out, int h,
                            float* h1, int ld_h1, void* stream) {
    if (!gnnseg_supported(1, h)) return GNNSEG_EUNSUPPORTED;
    if (!csr_ok(g) || ld_h1 < h || (ld_h1 & 3)) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && (!Q_in || !h1)) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && (!e_in || !e_out)) return GNNSEG_EINVAL;
    return gnnseg::node_gather_step(g, Q_in, e_in, e_out, h, h1, ld_h1, static_cast<cudaStream_t>(stream));
}

int gnnseg_node_mlp_step(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, int h,
                         float* P_out, float* Q_out, void* stream) {
    if (!gnnseg_supported(1, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || n_nodes < 0 || ld_h1 < h || (ld_h1 & 3)) return GNNSEG_EINVAL;
    if (n_nodes > 0 && (!X4 || !h1 || !P_out)) return GNNSEG_EINVAL;
    return gnnseg::node_mlp_step(blob, X4, h1, ld_h1, n_nodes, h, P_out, Q_out, static_cast<cudaStream_t>(stream));
}

int gnnseg_forward_ex(const float* blob, const GnnsegGraph* g, const float* X, int F, int h, int n_iters, float* scores,
                      void* ws, size_t ws_bytes, int flags, int32_t* status, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !csr_ok(g) || n_iters < 0 || !ws) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && !X) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && !scores) return GNNSEG_EINVAL;
    // 256-byte align the caller's pointer ourselves
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const FwdWorkspace w = carve(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* flag = status ? status : w.status;
    if (cudaMemsetAsync(flag, 0, sizeof(int32_t), st) != cudaSuccess) return GNNSEG_ECUDA;

    const bool fused = gnnseg::forward_is_fused(h, flags) && g->adj_ptr && g->adj;
    if (fused) {
        // gnnseg_fused.cu: input -> n_iters x (edge step inside the node step's CSR walk, tensor-core MLP) -> final
        // edge step over the in-edges of every node.  2 * n_iters + 2 launches.
        const int n = g->n_nodes;
        const bool pdl = gnnseg::use_pdl_mlp();
        // row n of a state buffer is what an absent neighbour reads: zeros while the buffer feeds a node step
        if (cudaMemsetAsync(w.s[0] + (size_t)n * 5 * h, 0, sizeof(float) * 5 * h, st) != cudaSuccess ||
            cudaMemsetAsync(w.s[1] + (size_t)n * 5 * h, 0, sizeof(float) * 5 * h, st) != cudaSuccess)
            return GNNSEG_ECUDA;
        // the last producing step writes only the edge projections [SPs | SPd] of its rows (n_cols = 2h)
        int rc = state_input(blob, X, n, F, h, w.x4, gnnseg::ProjOut{w.s[0], nullptr, 1, n_iters == 0 ? 2 * h : 5 * h, flag}, st);
        int cur = 0;
        for (int it = 0; it < n_iters && rc == GNNSEG_OK; ++it) {
            const bool is_last = it + 1 == n_iters;
            float* rows = w.s[cur ^ 1];                          // h1 travels in the first h floats of the rows it becomes
            rc = gnnseg::fused_gather_step(blob, g, w.s[cur], h, rows, 5 * h, st);
            if (rc == GNNSEG_OK)
                rc = state_mlp(blob, w.x4, rows, 5 * h, n, h, gnnseg::ProjOut{rows, nullptr, 1, is_last ? 2 * h : 5 * h, flag}, pdl, st);
            cur ^= 1;
        }
        // ... and for the final edge step an absent start node reads 2^(log2e b1) there (Ps = b1, gnn/model.py:71-73)
        const int sb1 = gnnseg::blob_total(h) - h;
        if (rc == GNNSEG_OK &&
            cudaMemcpyAsync(w.s[cur] + (size_t)n * 5 * h, blob + sb1, sizeof(float) * h, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return GNNSEG_ECUDA;
        if (rc == GNNSEG_OK) rc = gnnseg::edge_final_step(blob, g, w.s[cur], 5 * h, 0, h, h, scores, st);
        return rc;
    }

    int rc = gnnseg::input_step(blob, X, g->n_nodes, F, h, w.x4, w.p, w.q[0], nullptr, st);
    int cur = 0;
    for (int it = 0; it < n_iters && rc == GNNSEG_OK; ++it) {
        // intermediate scores are only consumed by the node step: CSR order only
        rc = gnnseg::edge_step(blob, g, w.p, h, nullptr, w.e_in, w.e_out, st);
        // the last node step feeds only the final edge step: its Q' is never read
        if (rc == GNNSEG_OK)
            rc = gnnseg::node_step(blob, g, w.x4, w.q[cur], w.e_in, w.e_out, h, w.p, w.q[cur ^ 1], it + 1 < n_iters, nullptr, nullptr, st);
        cur ^= 1;
    }
    if (rc == GNNSEG_OK) rc = gnnseg::edge_step(blob, g, w.p, h, scores, nullptr, nullptr, st);
    return rc;
}

int gnnseg_forward(const float* blob, const GnnsegGraph* g, const float* X, int F, int h,
                   int n_iters, float* scores, void* ws, size_t ws_bytes, void* stream) {
    return gnnseg_forward_ex(blob, g, X, F, h, n_iters, scores, ws, ws_bytes, 0, nullptr, stream);
}

size_t gnnseg_adjacency_entries(int n_nodes, int n_slots) {
    if (n_nodes < 0 || n_slots < 0) return 0;
    return gnnseg::adjacency_entries(n_nodes, n_slots);
}

int gnnseg_build_adjacency(const GnnsegGraph* g, int32_t* adj_ptr, int32_t* adj, int32_t* node_order, void* stream) {
    if (!csr_ok(g) || !adj_ptr || !adj) return GNNSEG_EINVAL;
    return gnnseg::build_adjacency(g, adj_ptr, adj, node_order, static_cast<cudaStream_t>(stream));
}

int gnnseg_fused_gather_step(const float* blob, const GnnsegGraph* g, const float* S, int h, float* h1, int ld_h1, void* stream) {
    if (!fused_width(h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !graph_ok(g) || !g->adj_ptr || !g->adj || ld_h1 < h || (ld_h1 & 3)) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && (!S || !h1)) return GNNSEG_EINVAL;
    return gnnseg::fused_gather_step(blob, g, S, h, h1, ld_h1, static_cast<cudaStream_t>(stream));
}

int gnnseg_edge_final_step(const float* blob, const GnnsegGraph* g, const float* P, int ld, int off_s, int off_d, int h,
                           float* scores, void* stream) {
    if (!fused_width(h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !csr_ok(g) || !g->adj_ptr || !g->adj || ld < 2 * h || (ld & 3) || off_s < 0 || off_d < 0 || (off_s & 3) ||
        (off_d & 3) || off_s + h > ld || off_d + h > ld)
        return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && !P) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && !scores) return GNNSEG_EINVAL;
    return gnnseg::edge_final_step(blob, g, P, ld, off_s, off_d, h, scores, static_cast<cudaStream_t>(stream));
}

int gnnseg_state_input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* X4, float* S, int n_cols,
                            int32_t* status, void* stream) {
    if (!gnnseg_supported(F, h) || !fused_width(h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || n_nodes < 0 || (n_cols != 5 * h && n_cols != 2 * h) || (n_nodes > 0 && (!X || !X4 || !S))) return GNNSEG_EINVAL;
    return state_input(blob, X, n_nodes, F, h, X4, gnnseg::ProjOut{S, nullptr, 1, n_cols, status}, static_cast<cudaStream_t>(stream));
}

int gnnseg_state_mlp_step(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes, int h, float* S,
                          int n_cols, int32_t* status, void* stream) {
    if (!fused_width(h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || n_nodes < 0 || (n_cols != 5 * h && n_cols != 2 * h) || ld_h1 < h || (ld_h1 & 3)) return GNNSEG_EINVAL;
    if (n_nodes > 0 && (!X4 || !h1 || !S)) return GNNSEG_EINVAL;
    return state_mlp(blob, X4, h1, ld_h1, n_nodes, h, gnnseg::ProjOut{S, nullptr, 1, n_cols, status}, false,
                     static_cast<cudaStream_t>(stream));
}

int gnnseg_forward_nodes(const float* blob, const float* head_blob, const GnnsegGraph* g, const float* X, int F,
                         int h, int n_iters, float* node_scores, void* ws, size_t ws_bytes, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !head_blob || !csr_ok(g) || n_iters < 0 || !ws) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && (!X || !node_scores)) return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const FwdWorkspace w = carve(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the last producing step runs on the head blob: column 0 of its P is the node logit
    int rc = gnnseg::input_step(n_iters == 0 ? head_blob : blob, X, g->n_nodes, F, h, w.x4, w.p, w.q[0], nullptr, st);
    int cur = 0;
    for (int it = 0; it < n_iters && rc == GNNSEG_OK; ++it) {
        const bool last = it + 1 == n_iters;
        rc = gnnseg::edge_step(blob, g, w.p, h, nullptr, w.e_in, w.e_out, st);
        if (rc == GNNSEG_OK)
            rc = gnnseg::node_step(last ? head_blob : blob, g, w.x4, w.q[cur], w.e_in, w.e_out, h, w.p, w.q[cur ^ 1],
                                   !last, nullptr, nullptr, st);
        cur ^= 1;
    }
    if (rc == GNNSEG_OK) rc = gnnseg::node_head(w.p, 2 * h, g->n_nodes, node_scores, st);
    return rc;
}

size_t gnnseg_segments_workspace_bytes(int n_hits, int n_pairs) {
    if (n_hits < 0 || n_pairs < 0 || n_pairs > 32 || (long long)n_hits * n_pairs > 0x7ffffff0LL) return 0;
    return gnnseg::segments_workspace_bytes(n_hits, n_pairs) + 256;
}

int gnnseg_build_segments(const int32_t* layer, const void* r, const void* phi, const void* z, int dtype_bytes,
                          const int64_t* particle_id, int n_hits, const int32_t* layer_pairs_host, int n_pairs,
                          int n_layers, double phi_slope_max, double phi_slope_outer_max, double z0_max,
                          int outer_from_layer, int node_offset, int capacity, int32_t* src, int32_t* dst, float* y,
                          int32_t* n_edges, void* ws, size_t ws_bytes, void* stream) {
    if (dtype_bytes != 4 && dtype_bytes != 8) return GNNSEG_EUNSUPPORTED;
    if (n_hits < 0 || n_pairs < 0 || n_pairs > 32 || n_layers < 1 || n_layers > 32 || capacity < 0 || !n_edges || !ws)
        return GNNSEG_EINVAL;
    if (n_pairs > 0 && !layer_pairs_host) return GNNSEG_EINVAL;
    if (n_hits > 0 && (!layer || !r || !phi || !z)) return GNNSEG_EINVAL;
    if (capacity > 0 && (!src || !dst)) return GNNSEG_EINVAL;
    for (int p = 0; p < 2 * n_pairs; ++p)
        if (layer_pairs_host[p] < 0 || layer_pairs_host[p] >= n_layers) return GNNSEG_EINVAL;
    const size_t need = gnnseg_segments_workspace_bytes(n_hits, n_pairs);
    if (need == 0) return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    if (ws_bytes < need - 256 + (al - raw)) return GNNSEG_EWORKSPACE;
    return gnnseg::build_segments(layer, r, phi, z, dtype_bytes, particle_id, n_hits, layer_pairs_host, n_pairs, n_layers,
                                  phi_slope_max, phi_slope_outer_max, z0_max, outer_from_layer, node_offset, capacity,
                                  src, dst, y, n_edges, reinterpret_cast<void*>(al), static_cast<cudaStream_t>(stream));
}

size_t gnnseg_segments_batch_workspace_bytes(int n_events, int total_hits, int n_pairs) {
    if (n_events < 0 || total_hits < 0 || n_pairs < 0 || n_pairs > 32 || (long long)total_hits * n_pairs > 0x7ffffff0LL) return 0;
    return gnnseg::segments_batch_workspace_bytes(n_events, total_hits, n_pairs) + 256;
}

int gnnseg_build_segments_batch(const int32_t* layer, const void* r, const void* phi, const void* z, int dtype_bytes,
                                const int64_t* particle_id, int n_events, const int32_t* hit_off, int total_hits,
                                int max_hits_per_event, const int32_t* layer_pairs_host, int n_pairs, int n_layers,
                                double phi_slope_max, double phi_slope_outer_max, double z0_max, int outer_from_layer,
                                int e_max, int32_t* src, int32_t* dst, float* y, int32_t* n_edges, void* ws, size_t ws_bytes,
                                void* stream) {
    if (dtype_bytes != 4 && dtype_bytes != 8) return GNNSEG_EUNSUPPORTED;
    if (n_events < 0 || total_hits < 0 || max_hits_per_event < 0 || max_hits_per_event > total_hits || n_pairs < 0 ||
        n_pairs > 32 || n_layers < 1 || n_layers > 32 || e_max < 0 || !ws)
        return GNNSEG_EINVAL;
    if (n_events > 0 && (!hit_off || !n_edges)) return GNNSEG_EINVAL;
    if (n_pairs > 0 && !layer_pairs_host) return GNNSEG_EINVAL;
    if (total_hits > 0 && (!layer || !r || !phi || !z)) return GNNSEG_EINVAL;
    if (e_max > 0 && (!src || !dst)) return GNNSEG_EINVAL;
    if ((long long)n_events * e_max > 0x7fffffffLL || n_events > 65535) return GNNSEG_EINVAL;
    for (int p = 0; p < 2 * n_pairs; ++p)
        if (layer_pairs_host[p] < 0 || layer_pairs_host[p] >= n_layers) return GNNSEG_EINVAL;
    const size_t need = gnnseg_segments_batch_workspace_bytes(n_events, total_hits, n_pairs);
    if (need == 0) return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    if (ws_bytes < need - 256 + (al - raw)) return GNNSEG_EWORKSPACE;
    return gnnseg::build_segments_batch(layer, r, phi, z, dtype_bytes, particle_id, n_events, hit_off, total_hits,
                                        max_hits_per_event, layer_pairs_host, n_pairs, n_layers, phi_slope_max,
                                        phi_slope_outer_max, z0_max, outer_from_layer, e_max, src, dst, y, n_edges,
                                        reinterpret_cast<void*>(al), static_cast<cudaStream_t>(stream));
}

int gnnseg_scale_features(const void* a, const void* b, const void* c, int dtype_bytes, int n_hits, double scale_a,
                          double scale_b, double scale_c, float* X, void* stream) {
    if (dtype_bytes != 4 && dtype_bytes != 8) return GNNSEG_EUNSUPPORTED;
    if (n_hits < 0 || (n_hits > 0 && (!a || !b || !c || !X))) return GNNSEG_EINVAL;
    return gnnseg::scale_features(a, b, c, dtype_bytes, n_hits, scale_a, scale_b, scale_c, X, static_cast<cudaStream_t>(stream));
}

size_t gnnseg_train_workspace_bytes(int n_nodes, int n_slots, int F, int h, int n_iters) {
    if (!gnnseg_supported(F, h) || n_nodes < 0 || n_slots < 0 || n_iters < 0 || n_iters > MAX_ITERS) return 0;
    return carve_train(nullptr, n_nodes, n_slots, h, n_iters).bytes + 256;
}

int gnnseg_forward_train(const float* blob, const GnnsegGraph* g, const float* X, int F, int h, int n_iters,
                         float* scores, void* ws, size_t ws_bytes, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !csr_ok(g) || n_iters < 0 || n_iters > MAX_ITERS || !ws) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && !X) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && !scores) return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const TrainWorkspace w = carve_train(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h, n_iters);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = gnnseg::input_step(blob, X, g->n_nodes, F, h, w.x4, w.P[0], w.Q[0], w.H[0], st);
    for (int t = 0; t < n_iters && rc == GNNSEG_OK; ++t) {
        rc = gnnseg::edge_step(blob, g, w.P[t], h, nullptr, w.e_in[t], w.e_out[t], st);
        const bool more = t + 1 < n_iters;
        if (rc == GNNSEG_OK)
            rc = gnnseg::node_step(blob, g, w.x4, w.Q[t], w.e_in[t], w.e_out[t], h, w.P[t + 1],
                                   more ? w.Q[t + 1] : nullptr, more, w.h1[t], w.H[t + 1], st);
    }
    if (rc == GNNSEG_OK) rc = gnnseg::edge_step(blob, g, w.P[n_iters], h, scores, nullptr, nullptr, st);
    return rc;
}

int gnnseg_backward(const float* blob, const GnnsegParams* masks, const GnnsegGraph* g, int F, int h, int n_iters,
                    const float* dscores, const GnnsegGrads* grads, void* ws, size_t ws_bytes, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !csr_ok(g) || n_iters < 0 || n_iters > MAX_ITERS || !ws || !grads) return GNNSEG_EINVAL;
    if (g->n_slots > 0 && !dscores) return GNNSEG_EINVAL;
    if (!grads->w_in || !grads->b_in || !grads->w_e1 || !grads->b_e1 || !grads->w_e2 || !grads->b_e2 ||
        !grads->w_n1 || !grads->b_n1 || !grads->w_n2 || !grads->b_n2)
        return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const TrainWorkspace w = carve_train(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h, n_iters);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    gnnseg::TrainState s;
    s.x4 = w.x4;
    s.Hs = w.H; s.h1s = w.h1; s.Ps = w.P; s.Qs = w.Q; s.e_in = w.e_in; s.e_out = w.e_out;
    s.dg = w.dg; s.dproj = w.dproj; s.ds_in = w.ds_in; s.ds_out = w.ds_out; s.partE = w.partE; s.partN = w.partN; s.dz = w.dz;
    gnnseg::GradOut go;
    go.w_in = grads->w_in; go.b_in = grads->b_in; go.w_e1 = grads->w_e1; go.b_e1 = grads->b_e1;
    go.w_e2 = grads->w_e2; go.b_e2 = grads->b_e2; go.w_n1 = grads->w_n1; go.b_n1 = grads->b_n1;
    go.w_n2 = grads->w_n2; go.b_n2 = grads->b_n2;
    go.m_e1 = masks ? masks->m_e1 : nullptr; go.m_e2 = masks ? masks->m_e2 : nullptr;
    go.m_n1 = masks ? masks->m_n1 : nullptr; go.m_n2 = masks ? masks->m_n2 : nullptr;
    return gnnseg::backward(blob, g, F, h, n_iters, dscores, s, go, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int gnnseg_forward_nodes_train(const float* blob, const float* head_blob, const GnnsegGraph* g, const float* X,
                               int F, int h, int n_iters, float* node_scores, void* ws, size_t ws_bytes, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !head_blob || !csr_ok(g) || n_iters < 0 || n_iters > MAX_ITERS || !ws) return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && (!X || !node_scores)) return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const TrainWorkspace w = carve_train(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h, n_iters);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = gnnseg::input_step(n_iters == 0 ? head_blob : blob, X, g->n_nodes, F, h, w.x4, w.P[0], w.Q[0], w.H[0], st);
    for (int t = 0; t < n_iters && rc == GNNSEG_OK; ++t) {
        rc = gnnseg::edge_step(blob, g, w.P[t], h, nullptr, w.e_in[t], w.e_out[t], st);
        const bool more = t + 1 < n_iters;
        if (rc == GNNSEG_OK)
            rc = gnnseg::node_step(more ? blob : head_blob, g, w.x4, w.Q[t], w.e_in[t], w.e_out[t], h, w.P[t + 1],
                                   more ? w.Q[t + 1] : nullptr, more, w.h1[t], w.H[t + 1], st);
    }
    if (rc == GNNSEG_OK) rc = gnnseg::node_head(w.P[n_iters], 2 * h, g->n_nodes, node_scores, st);
    return rc;
}

int gnnseg_backward_nodes(const float* blob, const float* head_blob, const GnnsegParams* masks, const GnnsegGraph* g,
                          int F, int h, int n_iters, const float* dnode_scores, const GnnsegGrads* grads,
                          float* grad_w_out, float* grad_b_out, void* ws, size_t ws_bytes, void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!blob || !head_blob || !csr_ok(g) || n_iters < 0 || n_iters > MAX_ITERS || !ws || !grads || !grad_w_out ||
        !grad_b_out)
        return GNNSEG_EINVAL;
    if (g->n_nodes > 0 && !dnode_scores) return GNNSEG_EINVAL;
    if (!grads->w_in || !grads->b_in || !grads->w_e1 || !grads->b_e1 || !grads->w_e2 || !grads->b_e2 ||
        !grads->w_n1 || !grads->b_n1 || !grads->w_n2 || !grads->b_n2)
        return GNNSEG_EINVAL;
    const uintptr_t raw = reinterpret_cast<uintptr_t>(ws);
    const uintptr_t al = (raw + 255) & ~uintptr_t(255);
    const TrainWorkspace w = carve_train(reinterpret_cast<void*>(al), g->n_nodes, g->n_slots, h, n_iters);
    if (ws_bytes < w.bytes + (al - raw)) return GNNSEG_EWORKSPACE;
    gnnseg::TrainState s;
    s.x4 = w.x4;
    s.Hs = w.H; s.h1s = w.h1; s.Ps = w.P; s.Qs = w.Q; s.e_in = w.e_in; s.e_out = w.e_out;
    s.dg = w.dg; s.dproj = w.dproj; s.ds_in = w.ds_in; s.ds_out = w.ds_out; s.partE = w.partE; s.partN = w.partN; s.dz = w.dz;
    gnnseg::GradOut go;
    go.w_in = grads->w_in; go.b_in = grads->b_in; go.w_e1 = grads->w_e1; go.b_e1 = grads->b_e1;
    go.w_e2 = grads->w_e2; go.b_e2 = grads->b_e2; go.w_n1 = grads->w_n1; go.b_n1 = grads->b_n1;
    go.w_n2 = grads->w_n2; go.b_n2 = grads->b_n2;
    go.m_e1 = masks ? masks->m_e1 : nullptr; go.m_e2 = masks ? masks->m_e2 : nullptr;
    go.m_n1 = masks ? masks->m_n1 : nullptr; go.m_n2 = masks ? masks->m_n2 : nullptr;
    return gnnseg::backward(blob, g, F, h, n_iters, dnode_scores, s, go, head_blob, grad_w_out, grad_b_out,
                            static_cast<cudaStream_t>(stream));
}

int gnnseg_bce_loss(const float* scores, const float* targets, const float* weights, int n, float* loss,
                    float* dscores, void* ws, void* stream) {
    if (n < 0 || !loss || !ws || (n > 0 && (!scores || !targets))) return GNNSEG_EINVAL;
    return gnnseg::bce_loss(scores, targets, weights, n, loss, dscores, static_cast<float*>(ws),
                            static_cast<cudaStream_t>(stream));
}

int gnnseg_l1_penalty(const GnnsegParams* params, int F, int h, float l1, float* loss, const GnnsegGrads* grads,
                      void* stream) {
    if (!gnnseg_supported(F, h)) return GNNSEG_EUNSUPPORTED;
    if (!params || !params->w_e1 || !params->w_e2 || !params->w_n1 || !params->w_n2) return GNNSEG_EINVAL;
    if (grads && (!grads->w_e1 || !grads->w_e2 || !grads->w_n1 || !grads->w_n2)) return GNNSEG_EINVAL;
    return gnnseg::l1_penalty(params, F, h, l1, loss, grads, static_cast<cudaStream_t>(stream));
}

int gnnseg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n, int step,
                     float lr, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    if (n < 0 || step < 1 || (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) return GNNSEG_EINVAL;
    return gnnseg::adam_step(param, grad, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
