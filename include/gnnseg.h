/*
 * gnnseg.h — C ABI of libgnnseg_b200.so: the B200 (sm_100a) implementation of the
 * interaction-network segment classifier forward of jmduarte/gnn-fpga.
 *
 * The reference has no FFI: its boundary is the Python module API in gnn/model.py and
 * gnn/graph.py.  Every entry point below names the reference code it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - Every `stream` argument is a cudaStream_t passed as void*.
 *   - Pointers are DEVICE pointers unless the name ends in `_host`.
 *   - The caller owns every buffer, including workspaces.  The library never allocates or
 *     frees device memory, never synchronises, and keeps no mutable global state (only a per-device
 *     cache of immutable facts: SM count, kernel occupancy, shared-memory opt-in), so every
 *     device entry point can be captured in a CUDA graph.
 *   - Return value: GNNSEG_OK (0) or a negative GNNSEG_E* code; gnnseg_strerror() names it.
 *     Data errors that can only be seen on the device (bad incidence matrices) are written
 *     to a caller-supplied device `err_flag` word (bit set below).
 *   - All arithmetic is IEEE fp32 (the reference casts everything to float32,
 *     gnn/trainSegmentClassifier.py:38-44); node/edge indices are int32.
 *
 * Graph layout ("flattened batch"): the B graphs of a batch are one block-diagonal graph
 * with n_nodes nodes.  Edges live in `n_slots` slots; slot j of event b in a padded
 * (B, E_max) batch is b*E_max + j, so scores come out directly in the reference's (B, E)
 * layout.  A slot whose start or end is absent (a zero column of Ro or Ri, i.e. the
 * zero padding of merge_graphs, gnn/trainSegmentClassifier.py:66-95) carries -1.
 */
#ifndef GNNSEG_H_
#define GNNSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNNSEG_ABI_VERSION 5

/* error codes */
#define GNNSEG_OK            0
#define GNNSEG_EINVAL       -1   /* null pointer / negative size / bad argument        */
#define GNNSEG_EUNSUPPORTED -2   /* (input_dim, hidden_dim) outside the compiled set    */
#define GNNSEG_EWORKSPACE   -3   /* workspace smaller than gnnseg_*_workspace_bytes()   */
#define GNNSEG_ECUDA        -4   /* a CUDA runtime call / launch failed                 */
#define GNNSEG_ENODEVICE    -5   /* no sm_100 device visible                            */
#define GNNSEG_EIO          -6   /* a file could not be opened / mapped                 */
#define GNNSEG_EFORMAT      -7   /* not an .npz graph file as save_graph writes them     */
#define GNNSEG_EHYPEREDGE   -8   /* a column listed twice in Ri (or Ro) of a host graph  */

/* bits written to the device err_flag word by gnnseg_dense_to_edges */
#define GNNSEG_BAD_VALUE      1  /* an incidence entry is neither 0 nor 1                */
#define GNNSEG_BAD_HYPEREDGE  2  /* a column of Ri (or Ro) has more than one non-zero    */

/*
 * The ten parameter tensors of SegmentClassifier (gnn/model.py:128-138), row-major as
 * torch stores them, plus the four optional MaskedLinear masks (gnn/model.py:19-33;
 * masks_e = [m_e1, m_e2], masks_n = [m_n1, m_n2]).  D = input_dim + hidden_dim.
 */
typedef struct GnnsegParams {
    const float* w_in;  const float* b_in;   /* input_network.0      (h, F)  , (h) */
    const float* w_e1;  const float* b_e1;   /* edge_network.network.0 (h, 2D), (h) */
    const float* w_e2;  const float* b_e2;   /* edge_network.network.2 (1, h) , (1) */
    const float* w_n1;  const float* b_n1;   /* node_network.network.0 (h, 3D), (h) */
    const float* w_n2;  const float* b_n2;   /* node_network.network.2 (h, h) , (h) */
    const float* m_e1;  const float* m_e2;   /* nullable: masks, same shapes as w_e1, w_e2 */
    const float* m_n1;  const float* m_n2;   /* nullable: masks, same shapes as w_n1, w_n2 */
} GnnsegParams;

/*
 * Flattened batch graph in both CSR orders.  (in_ptr, in_eid) is exactly what
 * np.nonzero(Ri) yields in gnn/graph.py:23-26 (rows ascending, columns ascending within a
 * row): in_ptr = [0] + cumsum(bincount(Ri_rows)), in_eid = Ri_cols.  Same for Ro / out_*.
 * in_nbr[s] = src[in_eid[s]] and out_nbr[s] = dst[out_eid[s]] are the neighbour ids.
 */
typedef struct GnnsegGraph {
    int32_t        n_nodes;
    int32_t        n_slots;
    const int32_t* src;      /* [n_slots] start node (row of Ro holding the 1) or -1 */
    const int32_t* dst;      /* [n_slots] end node   (row of Ri holding the 1) or -1 */
    const int32_t* in_ptr;   /* [n_nodes+1] destination-CSR row pointer              */
    const int32_t* in_eid;   /* [in_ptr[n_nodes]] slot ids, ascending within a row   */
    const int32_t* in_nbr;   /* [in_ptr[n_nodes]] src[in_eid]                        */
    const int32_t* out_ptr;  /* [n_nodes+1] source-CSR row pointer                   */
    const int32_t* out_eid;  /* [out_ptr[n_nodes]]                                   */
    const int32_t* out_nbr;  /* [out_ptr[n_nodes]] dst[out_eid]                      */
    const int32_t* in_pos;   /* [n_slots] position of the slot in in_eid  (in_eid[in_pos[j]] == j), -1 if dst[j] < 0  */
    const int32_t* out_pos;  /* [n_slots] position of the slot in out_eid, -1 if src[j] < 0 */
    /* combined adjacency of the fused inference path (gnnseg_build_adjacency); both NULL: the forward runs
     * the edge step and the node step as separate kernels */
    const int32_t* adj_ptr;    /* [n_nodes+1] = in_ptr + out_ptr                                             */
    const int32_t* adj;        /* [gnnseg_adjacency_entries()] per node, starting at 4*ceil(adj_ptr[n]/4) + 4n: its
                                  in-edges (start node id), then its out-edges (end node id | 0x80000000), ascending
                                  slot order, padded to a multiple of four with n_nodes (= absent neighbour: the
                                  extra row every state buffer of the fused path carries)                          */
    const int32_t* node_order; /* [n_nodes] nullable: the order the fused kernels take the nodes in (inside every
                                  window of 4096 nodes by decreasing length of the adjacency list)                 */
} GnnsegGraph;

/* ---- library ------------------------------------------------------------------------ */

int         gnnseg_abi_version(void);
const char* gnnseg_strerror(int code);
/* 1 if (input_dim F, hidden_dim h) has compiled kernels: F in 1..4, h in {4,8,16,32,64}. */
int         gnnseg_supported(int F, int h);
/* number of SMs of the current device (grid sizing), or a negative error. */
int         gnnseg_device_sm_count(void);

/* ---- weights: replaces MaskedLinear.forward's per-call weight*mask, gnn/model.py:28-33 */

/* floats in the packed (transposed, 16-byte padded, mask-multiplied) weight blob. */
size_t gnnseg_weights_floats(int F, int h);
int    gnnseg_pack_weights(const GnnsegParams* params, int F, int h, float* blob,
                           void* stream);

/* ---- graph: replaces make_sparse_graph / graph_from_sparse, gnn/graph.py:23-35 -------- */

/*
 * Dense (B,N,E) fp32 incidence matrices, as batch_generator feeds them to the model
 * (gnn/trainSegmentClassifier.py:104-110), to per-slot endpoints: dst[b*E+e] = b*N+n where
 * Ri[b,n,e]==1, src likewise from Ro, -1 where the column is all zero.  *err_flag is
 * OR-ed with GNNSEG_BAD_* bits (caller zeroes it).
 */
int gnnseg_dense_to_edges(const float* Ri, const float* Ro, int B, int N, int E,
                          int32_t* src, int32_t* dst, int32_t* err_flag, void* stream);

/*
 * Per-slot keys to CSR.  ptr[n_nodes+1], eid/nbr[n_slots] (only the first ptr[n_nodes] are
 * meaningful).  Slots with key < 0 are left out.  Rows list slot ids in ascending order,
 * which is np.nonzero's order, so the result is bit-identical to the reference's sparse
 * tuples.  nbr[s] = other[eid[s]];  pos[n_slots] (nullable) is the inverse map:
 * pos[eid[s]] = s, -1 for slots that were left out.
 */
size_t gnnseg_csr_workspace_bytes(int n_nodes, int n_slots);
int    gnnseg_build_csr(const int32_t* key, const int32_t* other, int n_slots, int n_nodes,
                        int32_t* ptr, int32_t* eid, int32_t* nbr, int32_t* pos,
                        void* ws, size_t ws_bytes, void* stream);

/*
 * Both CSRs of a batch in one go (the same launches build the destination-CSR, key = dst, and
 * the source-CSR, key = src).  Workspace: 2 * gnnseg_csr_workspace_bytes(n_nodes, n_slots).
 * in_pos / out_pos may be NULL.
 */
int    gnnseg_build_graph(const int32_t* src, const int32_t* dst, int n_slots, int n_nodes,
                          int32_t* in_ptr, int32_t* in_eid, int32_t* in_nbr, int32_t* in_pos,
                          int32_t* out_ptr, int32_t* out_eid, int32_t* out_nbr, int32_t* out_pos,
                          void* ws, size_t ws_bytes, void* stream);

/*
 * The combined adjacency list the fused inference path walks: per node its in-edges then its out-edges
 * (see GnnsegGraph.adj).  adj_ptr[n_nodes+1], adj[gnnseg_adjacency_entries(n_nodes, n_slots)], node_order[n_nodes]
 * (nullable).  Two launches, reads the two CSRs of `graph` (its adj_ptr / adj / node_order members are ignored).
 */
size_t gnnseg_adjacency_entries(int n_nodes, int n_slots);
int    gnnseg_build_adjacency(const GnnsegGraph* graph, int32_t* adj_ptr, int32_t* adj, int32_t* node_order,
                              void* stream);

/* ---- forward: replaces SegmentClassifier.forward, gnn/model.py:140-156 ---------------- */

size_t gnnseg_forward_workspace_bytes(int n_nodes, int n_slots, int F, int h);
/*
 * X is (n_nodes, F) row-major fp32.  scores[n_slots] receives the final edge_network
 * output; a -1/-1 slot gets the reference's padding constant sigmoid(W2.tanh(b1)+b2).
 * Launches 2*n_iters+2 kernels on `stream` (hidden_dim 32 / 64: 3*n_iters+2, the node step being a
 * gather kernel and a tensor-core MLP kernel), no host synchronisation.
 */
int gnnseg_forward(const float* blob, const GnnsegGraph* graph, const float* X,
                   int F, int h, int n_iters, float* scores,
                   void* ws, size_t ws_bytes, void* stream);
/*
 * The same with two more arguments.  hidden_dim 32 / 64 and a graph with an adjacency list run the FUSED
 * inference path (2*n_iters+2 launches: the edge step is evaluated inside the node step's CSR walk, once from
 * each end of an edge, on exponentials 2^(log2e P) of the edge projections that the producing kernel takes
 * once per node; nothing is written per edge between the iterations).  That path represents |P| <= 43.6;
 * beyond, the exponent is clamped and *status (device word, nullable) gets GNNSEG_RANGE_CLAMPED: the scores
 * of that call may then be off and the caller reruns with GNNSEG_FWD_EXACT, which forces the step-by-step
 * kernels (as does the environment variable GNNSEG_EXACT=1).  *status is zeroed at the start of the call.
 */
#define GNNSEG_FWD_EXACT      1
#define GNNSEG_RANGE_CLAMPED  1
int gnnseg_forward_ex(const float* blob, const GnnsegGraph* graph, const float* X,
                      int F, int h, int n_iters, float* scores,
                      void* ws, size_t ws_bytes, int flags, int32_t* status, void* stream);

/*
 * The three steps on their own (used by the per-kernel parity tests).  State arrays, with
 * HX[n] = [hidden(h) | X(F)] and D = h + F (the reference's cat([H, X]), gnn/model.py:146,154):
 *   X4 (n_nodes, 4)   X zero padded to 4 columns
 *   P  (n_nodes, 2h)  [ W1[:, 0:D].HX + b1 | W1[:, D:2D].HX ]                 edge-step inputs
 *   Q  (n_nodes, 3h)  [ W3[:, 0:D].HX | W3[:, D:2D].HX | W3[:, 2D:3D].HX + b3 ] node-step inputs
 * (W1 = edge_network.network.0.weight, W3 = node_network.network.0.weight: the first linear
 * layer of each network is applied per node before the gather, which is the same algebra.)
 *   gnnseg_input_step : input_network + cat([H,X]) + projections      gnn/model.py:144-146
 *   gnnseg_edge_step  : EdgeNetwork.forward                            gnn/model.py:69-81
 *                       writes the edge scores in slot order to `e` and/or, for the node step,
 *                       in the two CSR orders to e_in[in_pos[j]] / e_out[out_pos[j]]
 *                       (each of the three outputs may be NULL)
 *   gnnseg_node_step  : NodeNetwork.forward + cat([H,X]) + projections gnn/model.py:113-125,154
 *                       reads the scores in CSR order; Q_out may be NULL when no further node
 *                       step follows.
 */
int gnnseg_input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* X4,
                      float* P, float* Q, void* stream);
int gnnseg_edge_step(const float* blob, const GnnsegGraph* graph, const float* P, int h,
                     float* e, float* e_in, float* e_out, void* stream);
int gnnseg_node_step(const float* blob, const GnnsegGraph* graph, const float* X4,
                     const float* Q_in, const float* e_in, const float* e_out, int h,
                     float* P_out, float* Q_out, void* stream);

/*
 * The two halves of gnnseg_node_step on their own.  hidden_dim = 32 and 64 run the node step as these two
 * kernels (a gather rate on B200 is set by the number of resident warps, which a kernel that also
 * hosts the tensor-core epilogue warps cannot provide):
 *   gnnseg_node_gather_step : h1[n] = tanh(Qs[n] + sum_in e Qi[src] + sum_out e Qo[dst])   gnn/model.py:114-122
 *                             own term, in-slots, out-slots in ascending slot order; no atomics.
 *                             h1 rows are written with row stride ld_h1 floats (>= h, multiple of 4).
 *   gnnseg_node_mlp_step    : H' = tanh(W4.h1 + b4), projections of [H' | X] -> P_out, Q_out (nullable)
 *                             gnn/model.py:122-125,154.  hidden_dim = 32 or 64 (GNNSEG_EUNSUPPORTED
 *                             otherwise: the narrower widths run gnnseg_node_step's fused kernel).  h1 may live inside P_out (row n of h1 in the first h floats
 *                             of row n of P_out, ld_h1 = 2h): a row of P' is written after its h1 was read.
 */
int gnnseg_node_gather_step(const GnnsegGraph* graph, const float* Q_in, const float* e_in,
                            const float* e_out, int h, float* h1, int ld_h1, void* stream);
int gnnseg_node_mlp_step(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes,
                         int h, float* P_out, float* Q_out, void* stream);

/*
 * The steps of the fused inference path on their own (per-kernel parity tests, per-kernel timing).
 * State rows S (n_nodes + 1, 5h) = [SPs | SPd | Qi | Qo | Qs] with SPs = 2^(log2e Ps), SPd = 2^(log2e Pd) and
 * Qi, Qo, Qs as in gnnseg_input_step.  hidden_dim 32 or 64.  Row n_nodes is the caller's: what an absent
 * neighbour reads.  All zeros while the buffer feeds gnnseg_fused_gather_step (such an entry then adds e * 0);
 * exp(b1) in its first h columns for gnnseg_edge_final_step (the start-node projection of an absent start node,
 * gnn/model.py:71-73 with a zero column of Ro).  The producing steps write rows 0 .. n_nodes-1 only.
 *   gnnseg_state_input_step : gnnseg_input_step writing state rows; n_cols = 5h, or 2h: only [SPs | SPd]
 *   gnnseg_fused_gather_step: h1[n] = tanh(Qs[n] + sum_in e Qi[src] + sum_out e Qo[dst]) with the edge scores e
 *                             computed on the fly from SPs / SPd (gnn/model.py:69-81 inside :113-122); walks
 *                             graph->adj in graph->node_order
 *   gnnseg_state_mlp_step   : gnnseg_node_mlp_step writing state rows (n_cols as above); h1 may live in the first h
 *                             floats of the rows it becomes (ld_h1 = 5h)
 *   gnnseg_edge_final_step  : scores per slot from rows of `ld` floats holding SPs at column off_s and SPd at
 *                             column off_d (the in-edges of every node, then the slots without an end node)
 */
int gnnseg_state_input_step(const float* blob, const float* X, int n_nodes, int F, int h, float* X4,
                            float* S, int n_cols, int32_t* status, void* stream);
int gnnseg_fused_gather_step(const float* blob, const GnnsegGraph* graph, const float* S, int h,
                             float* h1, int ld_h1, void* stream);
int gnnseg_state_mlp_step(const float* blob, const float* X4, const float* h1, int ld_h1, int n_nodes,
                          int h, float* S, int n_cols, int32_t* status, void* stream);
int gnnseg_edge_final_step(const float* blob, const GnnsegGraph* graph, const float* P, int ld,
                           int off_s, int off_d, int h, float* scores, void* stream);

/* ---- training: replaces loss.backward() / optimizer.step() of Estimator.training_step,
 *      gnn/estimator.py:49-60 ------------------------------------------------------------- */

/* Gradients of the ten parameter tensors, same shapes and row-major layout as GnnsegParams. */
typedef struct GnnsegGrads {
    float* w_in;  float* b_in;
    float* w_e1;  float* b_e1;
    float* w_e2;  float* b_e2;
    float* w_n1;  float* b_n1;
    float* w_n2;  float* b_n2;
} GnnsegGrads;

/*
 * gnnseg_forward_train is gnnseg_forward that keeps, in the caller's workspace, what the backward
 * pass needs (per iteration: hidden states, first-layer projections, edge scores in both CSR
 * orders).  n_iters <= 64.  gnnseg_backward must be given the same workspace, untouched, and the
 * same blob / graph / sizes.  It launches 3*(n_iters+1)+2 kernels on `stream` (hidden_dim 32 / 64:
 * 4*(n_iters+1)+2, the dense step being two tcgen05 kernels: gnnseg_dprop_tc.cu, gnnseg_wgrad_tc.cu;
 * 3xTF32 products, fp32 accumulation); the weight gradients are per-CTA partial sums added in a fixed
 * order (no atomics): bit-identical from run to run.
 *   dscores[n_slots] = dL/dscore for every slot of the padded batch, padding slots included (the
 *                      reference's BCELoss averages over them, gnn/estimator.py:57).
 *   masks            = nullable; only m_e1, m_e2, m_n1, m_n2 are read: dW = dW_eff * mask
 *                      (MaskedLinear, gnn/model.py:28-31).
 *   grads            = written, not accumulated.
 */
size_t gnnseg_train_workspace_bytes(int n_nodes, int n_slots, int F, int h, int n_iters);
int gnnseg_forward_train(const float* blob, const GnnsegGraph* graph, const float* X,
                         int F, int h, int n_iters, float* scores,
                         void* ws, size_t ws_bytes, void* stream);
int gnnseg_backward(const float* blob, const GnnsegParams* masks, const GnnsegGraph* graph,
                    int F, int h, int n_iters, const float* dscores, const GnnsegGrads* grads,
                    void* ws, size_t ws_bytes, void* stream);

/*
 * nn.BCELoss() as Estimator uses it (mean reduction, log clamped at -100): *loss (device) is
 * written; dscores (nullable) receives dL/dscore.  weights: nullable per-element weights.
 * ws: 1024 bytes of device scratch.
 */
int gnnseg_bce_loss(const float* scores, const float* targets, const float* weights, int n,
                    float* loss, float* dscores, void* ws, void* stream);
/*
 * The L1 penalty of gnn/estimator.py:54-56: *loss += l1 * sum|W| over w_n1, w_n2, w_e1, w_e2
 * (raw weights) and, when grads is not NULL, grads->w_* += l1 * sign(W).  loss may be NULL.
 */
int gnnseg_l1_penalty(const GnnsegParams* params, int F, int h, float l1, float* loss,
                      const GnnsegGrads* grads, void* stream);
/*
 * torch.optim.Adam step (amsgrad off) on one flat tensor, in place; step counts from 1.
 */
int gnnseg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n,
                     int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                     void* stream);

/* ---- NodeClassifier: the per-hit sibling of the segment classifier,
 *      gnn/MPNN_HitClassifier.ipynb cell 21 (class NodeClassifier): the same input / edge / node
 *      steps, n_iters times, then output_network = Linear(F+h, 1) + Sigmoid on [H | X] per node
 *      and no final edge step ------------------------------------------------------------- */

/*
 * Packs the blob of the model's LAST producing step.  The head's logit Wo.[H|X] + bo is one more
 * linear map of [H | X], so it rides in column 0 of that step's start-node projection (which has
 * no reader in this model); params' w_e1 / b_e1 are not read.  head_blob: gnnseg_weights_floats
 * floats.  w_out [F+h] = output_network.0.weight (1, F+h), b_out [1] = its bias.
 */
int gnnseg_pack_node_head(const GnnsegParams* params, const float* w_out, const float* b_out,
                          int F, int h, float* head_blob, void* stream);
/*
 * node_scores[n_nodes] = NodeClassifier.forward.  blob from gnnseg_pack_weights, head_blob from
 * gnnseg_pack_node_head; workspace of gnnseg_forward_workspace_bytes.  2*n_iters + 2 launches
 * (h = 32 / 64: 3*n_iters + 2).
 */
int gnnseg_forward_nodes(const float* blob, const float* head_blob, const GnnsegGraph* graph,
                         const float* X, int F, int h, int n_iters, float* node_scores,
                         void* ws, size_t ws_bytes, void* stream);
/*
 * Training pair, as gnnseg_forward_train / gnnseg_backward (workspace of
 * gnnseg_train_workspace_bytes, kept untouched in between).  dnode_scores[n_nodes] = dL/dscore
 * per node; grad_w_out [F+h] and grad_b_out [1] receive the head's gradients, grads the ten
 * tensors of the body (written, not accumulated; bit-identical from run to run).
 */
int gnnseg_forward_nodes_train(const float* blob, const float* head_blob, const GnnsegGraph* graph,
                               const float* X, int F, int h, int n_iters, float* node_scores,
                               void* ws, size_t ws_bytes, void* stream);
int gnnseg_backward_nodes(const float* blob, const float* head_blob, const GnnsegParams* masks,
                          const GnnsegGraph* graph, int F, int h, int n_iters,
                          const float* dnode_scores, const GnnsegGrads* grads, float* grad_w_out,
                          float* grad_b_out, void* ws, size_t ws_bytes, void* stream);

/* ---- host side: replaces graph_from_sparse + merge_graphs + np_to_torch --------------- */

/*
 * Pack B host SparseGraph tuples (gnn/graph.py:20-21; int64 indices as np.nonzero and
 * np.load give them) into flattened int32 endpoint arrays and one X array, without
 * densifying.  Event b owns nodes [node_off[b], node_off[b+1]) and slots
 * [b*e_max, (b+1)*e_max).  src_host/dst_host are filled with -1 first.  Returns
 * GNNSEG_EINVAL if an index is out of range.  Pure CPU; n_threads <= 0 picks a default.
 */
int gnnseg_pack_sparse_batch_host(int B, int F, int e_max,
                                  const float* const* X_host, const int64_t* n_nodes_host,
                                  const int64_t* const* Ri_rows_host,
                                  const int64_t* const* Ri_cols_host,
                                  const int64_t* const* Ro_rows_host,
                                  const int64_t* const* Ro_cols_host,
                                  const int64_t* n_in_host, const int64_t* n_out_host,
                                  float* X_out_host, int32_t* src_host, int32_t* dst_host,
                                  int n_threads);

/*
 * The event store: the graphs of a data set narrowed, validated and laid out ONCE at load time in one
 * caller-allocated (pinned) host arena, so that a batch of consecutive events needs no per-batch host
 * work: five contiguous slices go to the device and gnnseg_assemble_batch builds the batch graph there.
 * This is the reference's `graphs = load_graphs(filenames, SparseGraph)` list
 * (gnn/trainSegmentClassifier.py:129) plus what its batch_generator does per batch
 * (graph_from_sparse + merge_graphs + np_to_torch, gnn/trainSegmentClassifier.py:97-111), split into a
 * load-time part (here) and a device part (below).
 *
 * Arena arrays, events back to back (offsets in GnnsegStoreLayout, all 256-byte aligned):
 *   node_off, in_off, out_off, y_off   int64 [n_events+1]  first node / Ri entry / Ro entry / label of an event
 *   X        float32 [total_nodes, F]   features in internal node order
 *   in_ptr   int32 [total_nodes + n_events]   per event n+1 local row pointers of the destination-CSR;
 *   out_ptr  int32 [total_nodes + n_events]   event b's block starts at node_off[b] + b
 *   in_col   uint16|int32 [total_in]   Ri_cols in CSR order (col_bytes = 2 when every event has <= 65536 edges)
 *   out_col  uint16|int32 [total_out]  Ro_cols in CSR order
 *   y        float32 [total_y]         labels by edge column
 *   perm     int32 [total_nodes]       internal position -> node id within the event
 *   n_edges  int64 [n_events]          edge columns of every event (see n_edges_host)
 * Index arrays may come in any order; np.nonzero order (gnn/graph.py:23-26) is the fast case.
 * reorder: 0 keeps the node order; 1 renumbers the nodes of every event internally along the feature
 * column in which edges are most local (the forward's gathers then hit in L1; scores, being per edge
 * column, do not see it; `perm` undoes it for per-node outputs); 2 = 1 for events of >= 1024 nodes.
 * n_edges_host: the number of edge columns of every event; NULL = len(Ri_rows), which is what
 * graph_from_sparse takes (gnn/graph.py:30).  A caller whose tuples have zero columns in Ri (an edge
 * without an end node: a dense graph can hold one, the reference's sparse form cannot) passes the true count.
 * Errors: GNNSEG_EINVAL for an index out of range (row >= n_nodes, column >= n_edges),
 * GNNSEG_EHYPEREDGE for a column listed twice in Ri or in Ro (the dense reference would sum two rows,
 * gnn/model.py:71-72).  info_host (nullable, 2 words): [0] = the first offending event or -1,
 * [1] = the column chosen for the renumbering or -1.  Pure CPU; n_threads <= 0 picks a default.
 */
typedef struct GnnsegStoreLayout {
    int64_t n_events;    int32_t n_features;  int32_t col_bytes;
    int64_t total_nodes; int64_t total_in;    int64_t total_out;  int64_t total_y;
    int64_t o_node_off;  int64_t o_in_off;    int64_t o_out_off;  int64_t o_y_off;
    int64_t o_X;         int64_t o_in_ptr;    int64_t o_out_ptr;
    int64_t o_in_col;    int64_t o_out_col;   int64_t o_y;        int64_t o_perm;
    int64_t o_n_edges;   /* int64 [n_events]: edge columns of every event */
    int64_t bytes;       /* size of the arena */
} GnnsegStoreLayout;
int gnnseg_store_plan_host(int n_events, int F, const int64_t* n_nodes_host, const int64_t* n_in_host,
                           const int64_t* n_out_host, const int64_t* n_y_host /* nullable */,
                           const int64_t* n_edges_host /* nullable */, GnnsegStoreLayout* layout);
int gnnseg_store_fill_host(const GnnsegStoreLayout* layout,
                           const float* const* X_host, const int64_t* n_nodes_host,
                           const int64_t* const* Ri_rows_host, const int64_t* const* Ri_cols_host,
                           const int64_t* const* Ro_rows_host, const int64_t* const* Ro_cols_host,
                           const int64_t* n_in_host, const int64_t* n_out_host,
                           const float* const* y_host /* nullable */, const int64_t* n_y_host /* nullable */,
                           const int64_t* n_edges_host /* nullable */,
                           int reorder, int n_threads, void* arena_host, int32_t* info_host);
/*
 * Device part: the flattened batch graph of B consecutive events of a store from their slices
 * (all DEVICE pointers; copied from the arena with one cudaMemcpyAsync each):
 *   meta      int32 [3*(B+1)]  node_off | in_base | out_base: first node, first Ri entry, first Ro
 *                              entry of every event RELATIVE to the batch (entry B = the totals)
 *   in_ptr_local / out_ptr_local  int32 [n_nodes + B]  the store's in_ptr / out_ptr slices
 *   in_col / out_col              [n_in] / [n_out]     the store's in_col / out_col slices
 * Writes every array of `graph` (src, dst, both CSRs, in_pos, out_pos; sizes as in GnnsegGraph with
 * n_slots = B*e_max; padding slots get -1).  No sort, no atomics, 4 memsets + 2 kernels on `stream`.
 */
int gnnseg_assemble_batch(const int32_t* meta, int B, int n_nodes, int e_max, int n_in, int n_out,
                          const int32_t* in_ptr_local, const int32_t* out_ptr_local,
                          const void* in_col, const void* out_col, int col_bytes,
                          int32_t* src, int32_t* dst,
                          int32_t* in_ptr, int32_t* in_eid, int32_t* in_nbr, int32_t* in_pos,
                          int32_t* out_ptr, int32_t* out_eid, int32_t* out_nbr, int32_t* out_pos,
                          void* stream);

/*
 * One batch of a store through the device in one call: the per-batch driver of pipelined inference
 * (replaces per batch the reference's generator + model call + .cpu(), gnn/trainSegmentClassifier.py:97-111,
 * gnn/estimator.py:137-146).  GnnsegBatchBuffers holds one pipeline slot: device buffers of the given capacities
 * (elements) and the pinned host words the results come back to; everything caller-allocated.
 *   gnnseg_store_batch_shape_host  shape_host[4] = n_nodes, e_max, n_in, n_out of events [lo, hi)   (pure CPU)
 *   gnnseg_store_load_batch        copy stream: six cudaMemcpyAsync out of the arena (batch metadata + five contiguous
 *                                  slices); compute stream: waits for them, gnnseg_assemble_batch + gnnseg_build_adjacency
 *   gnnseg_store_forward_batch     the same, then gnnseg_forward_ex on the compute stream (blob from gnnseg_pack_weights,
 *                                  flags as gnnseg_forward_ex) and, on the out stream after it, scores -> scores_host and
 *                                  the range flag -> status_host (either may be NULL: the result stays on the device).
 *                                  When the forward takes the fused inference path the assembly is the lean one (src, dst,
 *                                  row pointers, in_eid, out_eid and the adjacency lists only: in_pos / out_pos / in_nbr /
 *                                  out_nbr of the slot are NOT written) and runs on the copy stream behind the copies,
 *                                  i.e. next to the previous batch's forward; the compute stream waits for it
 * GNNSEG_EWORKSPACE (with shape_host filled) when the batch exceeds a capacity.  No host synchronisation, no library
 * state: the events ordering the three streams live inside the call; the caller records its own event on the out
 * stream afterwards to learn when the slot may be reused.
 */
typedef struct GnnsegBatchBuffers {
    int32_t  cap_nodes, cap_in, cap_out, cap_slots, cap_events, reserved_;
    int32_t* meta;            /* [3 * (cap_events + 1)]                                  */
    float*   X;               /* [cap_nodes * F]                                         */
    int32_t* in_ptr_local;    /* [cap_nodes + cap_events]                                */
    int32_t* out_ptr_local;   /* [cap_nodes + cap_events]                                */
    void*    in_col;          /* [cap_in]  uint16 / int32 (layout->col_bytes)            */
    void*    out_col;         /* [cap_out]                                               */
    int32_t* src;  int32_t* dst;  int32_t* in_pos;  int32_t* out_pos;     /* [cap_slots]      */
    int32_t* in_ptr;  int32_t* out_ptr;  int32_t* adj_ptr;                /* [cap_nodes + 1]  */
    int32_t* in_eid;  int32_t* in_nbr;                                    /* [cap_in]         */
    int32_t* out_eid; int32_t* out_nbr;                                   /* [cap_out]        */
    int32_t* adj;             /* [gnnseg_adjacency_entries(cap_nodes, cap_slots)]        */
    int32_t* node_order;      /* [cap_nodes]                                             */
    float*   scores;          /* [cap_slots]                                             */
    int32_t* status;          /* [1] range flag of gnnseg_forward_ex                     */
    void*    ws;  size_t ws_bytes;   /* gnnseg_forward_workspace_bytes(cap_nodes, cap_slots, F, h) */
    int32_t* meta_host;       /* pinned [3 * (cap_events + 1)]                           */
    float*   scores_host;     /* pinned [cap_slots] or NULL                              */
    int32_t* status_host;     /* pinned [1] or NULL                                      */
    void*    assemble_stream; /* nullable: a fourth stream for the lean assembly kernels of gnnseg_store_forward_batch
                                 (default: the copy stream).  With its own stream the next batch's copies do not queue
                                 behind assembly kernels that wait for a free SM (eight ranks sharing a host)      */
} GnnsegBatchBuffers;
int gnnseg_store_batch_shape_host(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                                  int32_t* shape_host);
int gnnseg_store_load_batch(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                            const GnnsegBatchBuffers* bufs, void* copy_stream, void* compute_stream,
                            int32_t* shape_host);
int gnnseg_store_forward_batch(const GnnsegStoreLayout* layout, const void* arena_host, int lo, int hi,
                               const float* blob, int h, int n_iters, int flags,
                               const GnnsegBatchBuffers* bufs, void* copy_stream, void* compute_stream,
                               void* out_stream, int32_t* shape_host);

/* ---- segment construction: replaces construct_graph / select_segments, gnn/graph.py:44-142 --- */

/*
 * Edge candidates of one event from its hits, on the device.  Hit columns as the reference's
 * DataFrame holds them, in row order: layer (int32, 0 <= layer < n_layers <= 32), r, phi, z (all
 * float32 or all float64: dtype_bytes 4 / 8), particle_id (int64, nullable).  For every layer pair
 * (layer_pairs_host[2p], [2p+1]; HOST array, n_pairs <= 32), hit i on the first layer and hit j on
 * the second, in row order:  dphi = phi_j - phi_i wrapped once into [-pi, pi];  phi_slope = dphi / dr;
 * z0 = z_i - r_i * dz / dr;  kept iff |phi_slope| < (first layer < outer_from_layer ? phi_slope_max
 * : phi_slope_outer_max) and |z0| < z0_max  (select_segments, gnn/graph.py:58-66; the reference passes
 * outer_from_layer = 5).  All arithmetic in the columns' dtype, IEEE round-to-nearest, in the reference's
 * order of evaluation: the selection is bit-identical to pandas / numpy.
 * Output, in the reference's edge order (layer pair, i, j):  src[k] = node_offset + i (Ro row),
 * dst[k] = node_offset + j (Ri row), y[k] = (particle_id_i == particle_id_j) (nullable), for
 * k < capacity;  n_edges[0] = number of kept pairs, n_edges[1] = 1 if that exceeds capacity (device
 * words).  capacity = 0 with src = dst = NULL only counts.  Two passes, no atomics.
 */
size_t gnnseg_segments_workspace_bytes(int n_hits, int n_pairs);
int gnnseg_build_segments(const int32_t* layer, const void* r, const void* phi, const void* z,
                          int dtype_bytes, const int64_t* particle_id, int n_hits,
                          const int32_t* layer_pairs_host, int n_pairs, int n_layers,
                          double phi_slope_max, double phi_slope_outer_max, double z0_max,
                          int outer_from_layer, int node_offset, int capacity,
                          int32_t* src, int32_t* dst, float* y, int32_t* n_edges,
                          void* ws, size_t ws_bytes, void* stream);

/*
 * The same for all events of a batch in one set of launches (the event loop of construct_graphs,
 * gnn/graph.py:145-175).  The hit columns hold the events back to back; event b owns rows
 * [hit_off[b], hit_off[b+1]) (hit_off: DEVICE array of n_events + 1 words; total_hits = hit_off[n_events];
 * max_hits_per_event sizes the grid).  Node ids are row numbers of the concatenated table; event b's
 * edges go to slots [b*e_max, b*e_max + count_b) of src / dst / y (the caller pre-fills the padding
 * with -1 / 0).  Two calls:
 *   e_max == 0  counting pass: n_edges[2b] = number of kept pairs of event b (device words);
 *   e_max  > 0  filling pass alone; it needs the workspace exactly as the counting pass on the same
 *               arguments left it.  e_max >= every count.  n_events <= 65535.
 */
size_t gnnseg_segments_batch_workspace_bytes(int n_events, int total_hits, int n_pairs);
int gnnseg_build_segments_batch(const int32_t* layer, const void* r, const void* phi, const void* z,
                                int dtype_bytes, const int64_t* particle_id, int n_events,
                                const int32_t* hit_off, int total_hits, int max_hits_per_event,
                                const int32_t* layer_pairs_host, int n_pairs, int n_layers,
                                double phi_slope_max, double phi_slope_outer_max, double z0_max,
                                int outer_from_layer, int e_max, int32_t* src, int32_t* dst, float* y,
                                int32_t* n_edges, void* ws, size_t ws_bytes, void* stream);
/* X (n_hits, 3) float32 = (column / scale) computed in float64 then rounded, as
 * (hits[feature_names].values / feature_scale).astype(np.float32), gnn/graph.py:118. */
int gnnseg_scale_features(const void* a, const void* b, const void* c, int dtype_bytes, int n_hits,
                          double scale_a, double scale_b, double scale_c, float* X, void* stream);

/* ---- on-disk format: replaces load_graph, gnn/graph.py:188-191 -------------------------- */

/*
 * One graph file as save_graph writes it (gnn/graph.py:179-181: np.savez of the six SparseGraph
 * members), memory mapped.  The pointers point into the mapping (or into an aligned copy when a
 * member is not aligned for its type inside the archive) and stay valid until
 * gnnseg_npz_close_graph_host; they are what gnnseg_pack_sparse_batch_host takes.  X must be
 * float32 (N, F), the four index arrays int64 (E,), y float32 (E,) (optional member), little
 * endian, C order, stored uncompressed: GNNSEG_EUNSUPPORTED for np.savez_compressed files or other
 * dtypes, GNNSEG_EFORMAT for anything that is not such an archive, GNNSEG_EIO if the file cannot
 * be mapped.  Pure CPU.
 */
typedef struct GnnsegNpzGraph {
    void*          handle;       /* owned by the library */
    const float*   X;            int64_t n_nodes;  int32_t n_features;
    const int64_t* Ri_rows;      const int64_t* Ri_cols;   int64_t n_in;
    const int64_t* Ro_rows;      const int64_t* Ro_cols;   int64_t n_out;
    const float*   y;            int64_t n_y;      /* y == NULL, n_y == 0 when the file has no y */
} GnnsegNpzGraph;
int gnnseg_npz_open_graph_host(const char* path, GnnsegNpzGraph* graph);
int gnnseg_npz_close_graph_host(GnnsegNpzGraph* graph);
/* B files at once, opened by n_threads threads (<= 0: a default); all or nothing: on the first
 * failure every graph is closed again and that file's error code is returned. */
int gnnseg_npz_open_batch_host(int B, const char* const* paths, GnnsegNpzGraph* graphs, int n_threads);
int gnnseg_npz_close_batch_host(int B, GnnsegNpzGraph* graphs);

#ifdef __cplusplus
}
#endif
#endif /* GNNSEG_H_ */
