// One batch of a host event store through the device in ONE call: the per-batch driver of the pipelined
// inference path.  Replaces, per batch, the reference's generator + model call + .cpu()
// (gnn/trainSegmentClassifier.py:97-111, gnn/estimator.py:137-146) with a fixed sequence of asynchronous
// CUDA calls on three caller-owned streams:
//   copy stream     6 cudaMemcpyAsync out of the store's pinned arena (batch metadata + five contiguous slices), then the
//                   batch assembly and the adjacency lists (two small kernels: they run next to the previous batch's forward)
//   compute stream  waits for those; gnnseg_forward_ex
//   out stream      waits for the forward; scores and the range flag back into pinned host memory
// Nothing here synchronises with the host and the library keeps no state: the two events that order the
// streams are created and destroyed inside the call.  The caller keeps `depth` slots (GnnsegBatchBuffers)
// in flight and records its own event on the out stream after the call to know when a slot is done.
#include <algorithm>
#include <cstdlib>
#include "gnnseg_common.cuh"

namespace gnnseg {
int assemble_batch(const int32_t*, int, int, int, int, int, const int32_t*, const int32_t*, const void*, const void*, int,
                   const GnnsegGraphMut&, cudaStream_t, bool lean);                                   // gnnseg_graph.cu
int build_adjacency(const GnnsegGraph*, int32_t*, int32_t*, int32_t*, cudaStream_t, bool from_endpoints);   // gnnseg_fused.cu
bool forward_is_fused(int h, int flags);                                                              // gnnseg_abi.cu
}

namespace {

struct BatchShape { int B, n_nodes, e_max, n_in, n_out; };

// sizes of events [lo, hi) and the batch-relative metadata gnnseg_assemble_batch wants, from the arena
bool batch_shape(const GnnsegStoreLayout* L, const char* arena, int lo, int hi, int32_t* meta_host, BatchShape* s) {
    const int64_t* node_off = reinterpret_cast<const int64_t*>(arena + L->o_node_off);
    const int64_t* in_off = reinterpret_cast<const int64_t*>(arena + L->o_in_off);
    const int64_t* out_off = reinterpret_cast<const int64_t*>(arena + L->o_out_off);
    const int64_t* n_edges = reinterpret_cast<const int64_t*>(arena + L->o_n_edges);
    const int B = hi - lo;
    const int64_t n = node_off[hi] - node_off[lo], ni = in_off[hi] - in_off[lo], no = out_off[hi] - out_off[lo];
    int64_t e_max = 0;
    for (int b = lo; b < hi; ++b) e_max = std::max(e_max, n_edges[b]);
    if (n + B > 0x7fffffffLL || ni > 0x7fffffffLL || no > 0x7fffffffLL || (int64_t)B * e_max > 0x7fffffffLL) return false;
    s->B = B; s->n_nodes = (int)n; s->e_max = (int)e_max; s->n_in = (int)ni; s->n_out = (int)no;
    if (meta_host)
        for (int b = 0; b <= B; ++b) {
            meta_host[b] = (int32_t)(node_off[lo + b] - node_off[lo]);
            meta_host[(B + 1) + b] = (int32_t)(in_off[lo + b] - in_off[lo]);
            meta_host[2 * (B + 1) + b] = (int32_t)(out_off[lo + b] - out_off[lo]);
        }
    return true;
}

bool fits(const GnnsegBatchBuffers* b, const BatchShape& s) {
    return s.B <= b->cap_events && s.n_nodes <= b->cap_nodes && s.n_in <= b->cap_in && s.n_out <= b->cap_out &&
           (int64_t)s.B * s.e_max <= b->cap_slots;
}

struct ScopedEvent {
    cudaEvent_t ev = nullptr;
    bool ok;
    ScopedEvent() { ok = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess; }
    ~ScopedEvent() { if (ev) cudaEventDestroy(ev); }     // legal while a stream still waits on it: released when that completes
};

}  // namespace

extern "C" int gnnseg_store_batch_shape_host(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                                             int32_t* shape_host) {
    if (!layout || !arena_host || !shape_host || lo < 0 || hi <= lo || hi > layout->n_events) return GNNSEG_EINVAL;
    BatchShape s;
    if (!batch_shape(layout, static_cast<const char*>(arena_host), lo, hi, nullptr, &s)) return GNNSEG_EINVAL;
    shape_host[0] = s.n_nodes; shape_host[1] = s.e_max; shape_host[2] = s.n_in; shape_host[3] = s.n_out;
    return GNNSEG_OK;
}

// lean: the batch feeds the fused inference path only (no inverse maps, no neighbour arrays: gnnseg_graph.cu).
// on_copy_stream: the assembly kernels run on the copy stream behind the batch's copies, i.e. next to the PREVIOUS batch's
// forward on the compute stream instead of in front of this one's (they are small: they fill that forward's tails).
static int load_batch(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi, const GnnsegBatchBuffers* bufs,
                      void* copy_stream, void* compute_stream, int32_t* shape_host, bool lean, bool on_copy_stream) {
    if (!layout || !arena_host || !bufs || lo < 0 || hi <= lo || hi > layout->n_events) return GNNSEG_EINVAL;
    if (!bufs->meta || !bufs->meta_host || !bufs->in_ptr || !bufs->out_ptr || !bufs->adj_ptr || !bufs->adj) return GNNSEG_EINVAL;
    const char* arena = static_cast<const char*>(arena_host);
    BatchShape s;
    if (!batch_shape(layout, arena, lo, hi, nullptr, &s)) return GNNSEG_EINVAL;
    if (shape_host) { shape_host[0] = s.n_nodes; shape_host[1] = s.e_max; shape_host[2] = s.n_in; shape_host[3] = s.n_out; }
    if (!fits(bufs, s)) return GNNSEG_EWORKSPACE;
    batch_shape(layout, arena, lo, hi, bufs->meta_host, &s);
    const int64_t* node_off = reinterpret_cast<const int64_t*>(arena + layout->o_node_off);
    const int64_t* in_off = reinterpret_cast<const int64_t*>(arena + layout->o_in_off);
    const int64_t* out_off = reinterpret_cast<const int64_t*>(arena + layout->o_out_off);
    const int F = layout->n_features, cb = layout->col_bytes;
    cudaStream_t cs = static_cast<cudaStream_t>(copy_stream), ks = static_cast<cudaStream_t>(compute_stream);
    const int64_t n0 = node_off[lo], i0 = in_off[lo], o0 = out_off[lo];
    auto copy = [&](void* dst, const void* src, size_t bytes) {
        return bytes == 0 || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs) == cudaSuccess;
    };
    bool ok = copy(bufs->meta, bufs->meta_host, sizeof(int32_t) * 3 * (s.B + 1));
    ok = ok && copy(bufs->X, arena + layout->o_X + 4 * n0 * F, (size_t)4 * s.n_nodes * F);
    ok = ok && copy(bufs->in_ptr_local, arena + layout->o_in_ptr + 4 * (n0 + lo), (size_t)4 * (s.n_nodes + s.B));
    ok = ok && copy(bufs->out_ptr_local, arena + layout->o_out_ptr + 4 * (n0 + lo), (size_t)4 * (s.n_nodes + s.B));
    ok = ok && copy(bufs->in_col, arena + layout->o_in_col + (int64_t)cb * i0, (size_t)cb * s.n_in);
    ok = ok && copy(bufs->out_col, arena + layout->o_out_col + (int64_t)cb * o0, (size_t)cb * s.n_out);
    if (!ok) return GNNSEG_ECUDA;
    auto order_streams = [&](cudaStream_t from, cudaStream_t to) {
        if (from == to) return true;
        ScopedEvent e;
        return e.ok && cudaEventRecord(e.ev, from) == cudaSuccess && cudaStreamWaitEvent(to, e.ev, 0) == cudaSuccess;
    };
    // where the assembly runs: the compute stream, or (lean mode) the copy stream / the slot's own assembly stream
    const cudaStream_t as = !on_copy_stream ? ks : (bufs->assemble_stream ? static_cast<cudaStream_t>(bufs->assemble_stream) : cs);
    if (!order_streams(cs, as)) return GNNSEG_ECUDA;
    if ((int64_t)s.B * s.e_max > 0 && (!bufs->src || !bufs->dst || (!lean && (!bufs->in_pos || !bufs->out_pos)))) return GNNSEG_EINVAL;
    if ((s.n_in > 0 && (!bufs->in_col || !bufs->in_eid || (!lean && !bufs->in_nbr))) ||
        (s.n_out > 0 && (!bufs->out_col || !bufs->out_eid || (!lean && !bufs->out_nbr))))
        return GNNSEG_EINVAL;
    const gnnseg::GnnsegGraphMut gm{bufs->src, bufs->dst, bufs->in_ptr, bufs->in_eid, bufs->in_nbr, bufs->in_pos,
                                    bufs->out_ptr, bufs->out_eid, bufs->out_nbr, bufs->out_pos};
    int rc = gnnseg::assemble_batch(bufs->meta, s.B, s.n_nodes, s.e_max, s.n_in, s.n_out, bufs->in_ptr_local, bufs->out_ptr_local,
                                    bufs->in_col, bufs->out_col, cb, gm, as, lean);
    if (rc != GNNSEG_OK) return rc;
    const GnnsegGraph g{s.n_nodes, s.B * s.e_max, bufs->src, bufs->dst, bufs->in_ptr, bufs->in_eid, bufs->in_nbr, bufs->out_ptr,
                        bufs->out_eid, bufs->out_nbr, bufs->in_pos, bufs->out_pos, nullptr, nullptr, nullptr};
    rc = gnnseg::build_adjacency(&g, bufs->adj_ptr, bufs->adj, bufs->node_order, as, lean);
    if (rc != GNNSEG_OK) return rc;
    if (!order_streams(as, ks)) return GNNSEG_ECUDA;
    return GNNSEG_OK;
}

extern "C" int gnnseg_store_load_batch(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                                       const GnnsegBatchBuffers* bufs, void* copy_stream, void* compute_stream,
                                       int32_t* shape_host) {
    return load_batch(layout, arena_host, lo, hi, bufs, copy_stream, compute_stream, shape_host, false, false);
}

extern "C" int gnnseg_store_forward_batch(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                                          const float* blob, int h, int n_iters, int flags, const GnnsegBatchBuffers* bufs,
                                          void* copy_stream, void* compute_stream, void* out_stream, int32_t* shape_host) {
    if (!bufs || !bufs->scores || !bufs->status || !bufs->ws) return GNNSEG_EINVAL;
    int32_t shape[4] = {0, 0, 0, 0};
    // Where the batch assembly runs.  GNNSEG_STREAM_ASSEMBLE (A/B runs): 0 = full assembly on the compute stream, 1 = lean
    // assembly on the compute stream, 2 = lean assembly next to the previous batch's forward (copy / assembly stream).
    // Default: 2, except at hidden_dim 64, where the concurrent form measured slower (mu200 event: 1.36 against 1.23 ms
    // per batch; acts64, hidden_dim 32: 0.63 against 0.73 ms).
    static const int forced = [] { const char* v = std::getenv("GNNSEG_STREAM_ASSEMBLE"); return v ? std::atoi(v) : -1; }();
    const int mode = forced >= 0 ? forced : (h == 64 ? 1 : 2);
    const bool lean = mode >= 1 && gnnseg::forward_is_fused(h, flags);
    int rc = load_batch(layout, arena_host, lo, hi, bufs, copy_stream, compute_stream, shape, lean, lean && mode >= 2);
    if (shape_host) for (int i = 0; i < 4; ++i) shape_host[i] = shape[i];
    if (rc != GNNSEG_OK) return rc;
    const int B = hi - lo, n_nodes = shape[0], e_max = shape[1];
    const GnnsegGraph g{n_nodes, B * e_max, bufs->src, bufs->dst, bufs->in_ptr, bufs->in_eid, bufs->in_nbr, bufs->out_ptr,
                        bufs->out_eid, bufs->out_nbr, bufs->in_pos, bufs->out_pos, bufs->adj_ptr, bufs->adj, bufs->node_order};
    cudaStream_t ks = static_cast<cudaStream_t>(compute_stream), os = static_cast<cudaStream_t>(out_stream);
    rc = gnnseg_forward_ex(blob, &g, bufs->X, layout->n_features, h, n_iters, bufs->scores, bufs->ws, bufs->ws_bytes, flags,
                           bufs->status, ks);
    if (rc != GNNSEG_OK) return rc;
    if (bufs->scores_host || bufs->status_host) {
        if (os != ks) {
            ScopedEvent e;
            if (!e.ok || cudaEventRecord(e.ev, ks) != cudaSuccess || cudaStreamWaitEvent(os, e.ev, 0) != cudaSuccess) return GNNSEG_ECUDA;
        }
        const size_t n_slots = (size_t)B * e_max;
        if (bufs->scores_host && n_slots > 0 &&
            cudaMemcpyAsync(bufs->scores_host, bufs->scores, 4 * n_slots, cudaMemcpyDeviceToHost, os) != cudaSuccess)
            return GNNSEG_ECUDA;
        if (bufs->status_host && cudaMemcpyAsync(bufs->status_host, bufs->status, 4, cudaMemcpyDeviceToHost, os) != cudaSuccess)
            return GNNSEG_ECUDA;
    }
    return GNNSEG_OK;
}
