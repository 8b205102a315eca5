/*
 * kgat_b200.h -- C ABI of the B200-native KGAT hot path (libkgat_b200.so).
 *
 * The reference (jennyzhang0215/DGL-KGAT) is pure Python and has no FFI layer of
 * its own: its GPU work happens inside DGL 0.4.x and ATen.  Each entry point
 * below replaces one group of those library calls; the reference call site it
 * stands in for is cited as file:line.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions (all entry points)
 *   - return 0 on success, a negative kgat_status on failure; nothing throws,
 *     nothing is printed.  kgat_error_string() maps a code to text;
 *     kgat_last_cuda_error() returns the cudaError_t seen by the failing call.
 *   - every pointer is a DEVICE pointer to contiguous row-major data owned by the
 *     caller (fp32 values, int32 ids), unless its name ends in _host;
 *   - the library allocates nothing and keeps no state between calls; scratch
 *     comes in through an explicit workspace argument sized by a *_workspace_bytes
 *     query (only exception: the multi-GPU arenas of section 8, which have explicit
 *     create / destroy calls);
 *   - work is enqueued on the given stream and the call returns without
 *     synchronising (CUDA-graph capturable), except kgat_graph_build which
 *     performs no host read-back either;
 *   - 16-byte alignment is required for every float array; all feature widths
 *     must be multiples of 4.
 */
#ifndef KGAT_B200_H
#define KGAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KGAT_API __attribute__((visibility("default")))
#else
#define KGAT_API
#endif

typedef void* kgat_stream_t;          /* a cudaStream_t */

enum kgat_status {
    KGAT_OK = 0,
    KGAT_ERR_INVALID_ARG = -1,        /* null pointer, negative size, misaligned buffer */
    KGAT_ERR_UNSUPPORTED_DIM = -2,    /* feature width not a multiple of 4 or above the compiled maximum */
    KGAT_ERR_WORKSPACE = -3,          /* workspace smaller than *_workspace_bytes says */
    KGAT_ERR_CUDA = -4,               /* a CUDA runtime call failed; see kgat_last_cuda_error() */
    KGAT_ERR_NO_DEVICE = -5           /* no sm_100 device is current */
};

KGAT_API int         kgat_version(void);                   /* ABI version, currently 1 */
KGAT_API const char* kgat_error_string(int status);
KGAT_API int         kgat_last_cuda_error(void);
/* 0 when the current device can run the kernels (compute capability 10.x). */
KGAT_API int         kgat_device_check(void);

/*
 * Collaborative knowledge graph in the three layouts the kernels read.
 * Built from the (E,3) int32 (h, r, t) triplet array of dataset.py:84-93; edge i
 * runs src = t[i] (tail) -> dst = h[i] (head) and has edge id i, exactly as
 * dataset.py:114-119 hands it to DGL (add_edges(t, h); edata['type'] = r).
 *
 *   CSR, destination-sorted: position c = rank of the edge under (head, rel, eid)
 *   CSC, source-sorted:      position k = rank of CSR position under (tail, c)
 *   relation buckets:        position p = rank of CSR position under (rel, c);
 *                            bucket r = the edges filter_edges(type == r) selects
 *                            (models.py:150)
 */
typedef struct kgat_graph {
    int64_t  n_nodes, n_edges, n_rel;
    int32_t *csr_ptr;   /* [n_nodes+1] */
    int32_t *csr_src;   /* [E] tail  */
    int32_t *csr_dst;   /* [E] head  */
    int32_t *csr_rel;   /* [E]       */
    int32_t *csr_eid;   /* [E] triplet row */
    int32_t *csc_ptr;   /* [n_nodes+1] */
    int32_t *csc_dst;   /* [E] head of the k-th out-edge */
    int32_t *csc_pos;   /* [E] its CSR position */
    int32_t *csc_eid;   /* [E] its triplet row */
    int32_t *rel_ptr;   /* [n_rel+1] */
    int32_t *bk_pos;    /* [E] CSR position */
    int32_t *bk_head;   /* [E] */
    int32_t *bk_tail;   /* [E] */
    /* Row splitting for power-law graphs (filled by kgat_graph_split_build; all NULL / 0 = never split).
     * A hot entity of a real KG has 10^4..10^6 edges (dataset.py:84-93 stacks every interaction of an
     * item onto the item's row); such a row must not be walked by one warp or one CTA.  Rows with more
     * than KGAT_SPLIT_ROW in-edges (CSR side) / out-edges (CSC side) are cut into segments of at most
     * KGAT_SPLIT_SEG edges; the aggregation and softmax entry points handle them in a pre-pass with one
     * CTA per SEGMENT and add the partial results in segment order (deterministic).  The threshold is low
     * on purpose: hot rows come in clusters (popular items / entities have neighbouring ids), and a tile of
     * 32 consecutive rows with a few hundred edges each kept one CTA busy long after the rest of the grid
     * had finished (measured on the Zipf graph of bench.py: 0.46 ms instead of 0.15 for the first layer). */
    int32_t *split;          /* [kgat_graph_split_ints(E)] segment tables of both sides */
    float   *split_scratch;  /* [max(n_csr_seg, n_csc_seg) * KGAT_SPLIT_SCRATCH_WIDTH] partial sums */
    int32_t  n_csr_seg, n_csr_long, n_csc_seg, n_csc_long;   /* host copies of the table sizes */
} kgat_graph_t;

#define KGAT_SPLIT_ROW 256
#define KGAT_SPLIT_SEG 512
#define KGAT_SPLIT_SCRATCH_WIDTH 128

/* (1) graph construction -- replaces DGLGraph()/add_nodes/add_edges/readonly
 *     (dataset.py:112-130, :203-211) and the per-relation filter_edges scans
 *     (models.py:150).  head/rel/tail: device int32 columns of the triplet array.
 *     Fills every array of *g (the caller allocated them).  Deterministic and
 *     bit-exact against oracle/graph_ref.py. */
KGAT_API size_t kgat_graph_build_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t n_rel);
KGAT_API int    kgat_graph_build(const kgat_graph_t* g, const int32_t* head, const int32_t* rel,
                        const int32_t* tail, void* workspace, size_t workspace_bytes,
                        kgat_stream_t stream);

/* (1') segment tables of the rows that are split (see kgat_graph_t).  Reads g->csr_ptr / g->csc_ptr,
 *      fills g->split (caller-allocated, kgat_graph_split_ints(E) int32) and returns the four table
 *      sizes {n_csr_seg, n_csr_long, n_csc_seg, n_csc_long} in counts_out_host - the caller stores them in
 *      its kgat_graph_t and allocates split_scratch when any is non-zero.  Synchronises the stream
 *      (one-off per graph).  Deterministic: rows ascending, segments of a row in edge order. */
KGAT_API size_t kgat_graph_split_ints(int64_t n_edges);
KGAT_API int    kgat_graph_split_build(const kgat_graph_t* g, int32_t* counts_out_host, kgat_stream_t stream);

/* (2) TransR attention logits pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r) for every
 *     edge, relation-grouped -- replaces the loop of apply_edges(_att_score)
 *     (models.py:135-152).  entity [n_nodes,d], rel_emb [n_rel,rd], W_R [n_rel,d,rd].
 *     logits_csr[c] is the logit of the edge at CSR position c.
 *     d = rd = 64 runs on the tensor cores (tcgen05 split-fp16, fp32 accumulate in TMEM); other widths use
 *     fp32 FFMA kernels.  The reference computes in fp32 (models.py:140-143); the split-fp16 kernel is only
 *     valid for |W_R| < 32 and |entity| < 65504, so the call checks the range ON THE DEVICE (no host round
 *     trip, CUDA-graph capturable) and re-runs the logits with the fp32-range 3xTF32 kernel when an operand
 *     is outside it; kgat_attention_fallbacks() reports how often that happened.
 *     Environment: KGAT_ATT_IMPL=tf32 / ffma force those kernels (A/B measurements). */
KGAT_API int kgat_attention_logits(const kgat_graph_t* g, const float* entity, const float* rel_emb,
                          const float* W_R, int d, int rd, float* logits_csr,
                          kgat_stream_t stream);
/* number of kgat_attention_logits calls on the current device that took the fp32-range fallback since the
 * last call of this function (then cleared).  Synchronises the device. */
KGAT_API int kgat_attention_fallbacks(void);

/* (2') the same logits through the (head, relation)-run shortcut:
 *      pi(h,r,t) = e_t . (W_r tanh(e_h W_r + e_r)), and the bracket depends on (h, r) only.  Inside a relation
 *      bucket the edges are head-sorted, so the edges of one (h, r) pair - a run - are contiguous.
 *      kgat_runs_build derives the runs from a built graph, once per graph (deterministic; bit-exact against
 *      oracle/graph_ref.build_runs): bk_run[p] = run of bucket position p, run_ptr[k] = first bucket position of
 *      run k (run_ptr[n_runs] = E), run_head[k] = its head node, rel_run_ptr[r] = first run of relation r,
 *      n_runs[0] = the run count (device scalar).  The caller allocates bk_run[E], run_ptr[E + 1], run_head[E],
 *      rel_run_ptr[n_rel + 1], n_runs[1].
 *      kgat_attention_logits_runs computes two 64 x 64 mat-vecs per RUN on the tensor cores (tcgen05 split-fp16
 *      grouped GEMMs over 128-run tiles) and a 64-wide dot per edge; d = rd = 64 only (KGAT_ERR_UNSUPPORTED_DIM
 *      otherwise), same operand-range protocol and fallback as kgat_attention_logits.
 *      v_rows: scratch for the per-run vectors, fp32 [n_runs <= E][64], 16-byte aligned - the run kernel writes
 *      them and a second, full-occupancy kernel computes the edge dots (the right split when runs are much
 *      fewer than edges: the rows stay in L2); NULL fuses the edge walk into the run kernel (no scratch, no
 *      extra traffic; latency-bound on the read-out warps, see profiles/README.md). */
typedef struct kgat_runs_t {
    int32_t *bk_run;        /* [E] */
    int32_t *run_ptr;       /* [E+1] */
    int32_t *run_head;      /* [E] */
    int32_t *rel_run_ptr;   /* [n_rel+1] */
    int32_t *n_runs;        /* [1] */
    int32_t *rel_mode;      /* [n_rel] or NULL: 1 = the relation has at most a third as many runs as edges */
} kgat_runs_t;
KGAT_API size_t kgat_runs_build_workspace_bytes(int64_t n_edges);
KGAT_API int kgat_runs_build(const kgat_graph_t* g, const kgat_runs_t* runs, void* workspace, size_t workspace_bytes,
                    kgat_stream_t stream);
KGAT_API int kgat_attention_logits_runs(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                               const float* rel_emb, const float* W_R, int d, int rd, float* logits_csr,
                               float* v_rows, kgat_stream_t stream);
/* (2h) the choice per RELATION: relations with runs->rel_mode[r] == 1 go through the run shortcut, the others
 *      through the per-edge kernel of kgat_attention_logits (a synthetic KG with random heads next to the two
 *      interaction relations, whose rows are one run each).  The caller reads rel_mode / rel_ptr / rel_run_ptr
 *      once per graph and passes the spans of the marked relations: v_rows holds the runs
 *      [v_base, v_base + v_count) (64 floats each; NULL = fused edge walk), [p_begin, p_end) are the bucket
 *      positions from the first to the last marked relation. */
KGAT_API int kgat_attention_logits_hybrid(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                                 const float* rel_emb, const float* W_R, int d, int rd, float* logits_csr,
                                 float* v_rows, int64_t v_base, int64_t v_count, int64_t p_begin, int64_t p_end,
                                 kgat_stream_t stream);

/* (3a) softmax over the in-edges of every destination row -- replaces
 *      edge_softmax (models.py:153).  Rows [row_begin,row_end).  Writes the weights
 *      in CSR order and, when non-null, scattered to original edge-id order
 *      (what compute_attention returns). */
KGAT_API int kgat_segment_softmax(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                         const float* logits_csr, float* w_csr, float* w_eid,
                         kgat_stream_t stream);

/* (3b) edge-weight re-layout: w_csr[c] = w_eid[csr_eid[c]] and
 *      w_csc[k] = w_eid[csc_eid[k]] -- what `g.edata['w'] = A_w` (kgat.py:144, :58)
 *      means for the CSR/CSC kernels.  Either output may be null. */
KGAT_API int kgat_edge_weights_from_eid(const kgat_graph_t* g, const float* w_eid, float* w_csr,
                               float* w_csc, kgat_stream_t stream);

/* (3c)+(4) one KGATConv forward over rows [row_begin,row_end) -- replaces
 *      update_all(u_mul_e, sum) + mul + Linear + leaky_relu + Dropout
 *      (models.py:60-70) and the F.normalize / cat of Model.gnn (models.py:163-167):
 *        h_nb[v]  = sum_{c in CSR row v} w_csr[c] * h_in[csr_src[c]]
 *        act[v]   = LeakyReLU_slope((h_in[v] (.) h_nb[v]) W2^T)
 *        h_out[v] = act[v] (.) drop_mask[v]            (drop_mask null: no dropout)
 *        concat[v, out_off : out_off+d_out] = h_out[v] / max(|h_out[v]|_2, 1e-12)
 *        concat[v, in_off  : in_off +d_in ] = h_in[v]  (only when in_off >= 0)
 *      W2 is [d_out, d_in] (nn.Linear weight).  act may be null when drop_mask is
 *      null (then act == h_out).  concat may be null. */
KGAT_API int kgat_layer_fwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                   const float* h_in, const float* w_csr, const float* W2,
                   const float* drop_mask, int d_in, int d_out, float slope,
                   float* h_nb, float* act, float* h_out,
                   float* concat, int64_t concat_stride, int64_t out_off, int64_t in_off,
                   kgat_stream_t stream);

/* Backward of one KGATConv for rows [row_begin,row_end) -- replaces the autograd
 * of the ops above incl. DGL's BackwardBinaryReduce (transposed aggregation).
 * Gradient w.r.t. this layer's output row v is assembled in-kernel from
 *   - the normalised copy:   d_concat[v, out_off : out_off+d_out]   (d_concat may be null).
 *     When dc_row_slot is non-null the gradient of the concat buffer is COMPACT (what
 *     kgat_bpr_loss_bwd produces): row v's gradient is d_concat[dc_row_slot[v], ...] and is zero
 *     when dc_row_slot[v] < 0;
 *   - the next layer (both null for the last layer; each term is optional on its own):
 *       nxt_dx_direct[v] + sum_{k in CSC row v} w_csc[k] * nxt_dh_nb[csc_dst[k]]
 * then pushed through dropout / LeakyReLU / Linear / (.):
 *   dW2       = sum_v dpre[v]^T (h_in[v] (.) h_nb[v])      (summed over the given rows)
 *   dx_direct = (dpre W2) (.) h_nb        dh_nb = (dpre W2) (.) h_in
 * dW2_partials: scratch of kgat_layer_bwd_partials() * d_out * d_in floats.
 * dW2 may be NULL: the call then leaves one partial per CTA in dW2_partials and kgat_layer_bwd_reduce (same
 * row count and widths) adds them in fixed order later - e.g. on a second stream, next to the following
 * kernels, since nothing in the rest of the backward pass reads dW2. */
KGAT_API int kgat_layer_bwd_partials(void);
KGAT_API int kgat_layer_bwd_reduce(int64_t rows, int d_in, int d_out, const float* dW2_partials, float* dW2,
                          kgat_stream_t stream);
KGAT_API int kgat_layer_bwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                   const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride, int64_t out_off,
                   const float* nxt_dx_direct, const float* nxt_dh_nb, const float* w_csc,
                   const float* h_in, const float* h_nb, const float* act, const float* h_out,
                   const float* drop_mask, const float* W2, int d_in, int d_out, float slope,
                   float* dx_direct, float* dh_nb, float* dW2, float* dW2_partials,
                   kgat_stream_t stream);

/* Gradient of the entity table (input of layer 0) for rows [row_begin,row_end):
 *   d_entity[v] = d_concat[v, in_off : in_off+d] + dx_direct[v]
 *               + sum_{k in CSC row v} w_csc[k] * dh_nb[csc_dst[k]] */
/* (3c') transposed aggregation on its own: out[v] = (add ? add[v] : 0) + sum over the CSC row of v of
 *       w_csc[k] * feat[csc_dst[k]], rows [row_begin, row_end), width d in {8, 16, 32, 64, 128} -- autograd of
 *       update_all(u_mul_e, sum) (models.py:63; DGL's BackwardBinaryReduce with atomics), gather-only and
 *       deterministic; rows above 512 edges are gathered by a whole CTA.  The host layer runs it in front of
 *       kgat_layer_bwd (add = the next layer's dx_direct, feat = its dh_nb) and hands the sum to that call as
 *       nxt_dx_direct with nxt_dh_nb = NULL. */
KGAT_API int kgat_csc_gather(const kgat_graph_t* g, int64_t row_begin, int64_t row_end, const float* add,
                    const float* feat, const float* w_csc, int d, float* out, kgat_stream_t stream);
KGAT_API int kgat_input_bwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                   const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride, int64_t in_off,
                   const float* dx_direct, const float* dh_nb, const float* w_csc, int d,
                   float* d_entity, kgat_stream_t stream);

/* BPR loss head of the CF step -- replaces Model.get_loss (models.py:170-178):
 *   loss = -mean(logsigmoid(<s,p> - <s,n>)) + reg_lambda * (mean|s|^2/2 + mean|p|^2/2 + mean|n|^2/2),
 *   s,p,n = emb[src], emb[pos], emb[neg]  (int64 ids, as kgat.py:153-158 produces them).
 * fwd writes the scalar loss and sig[b] = sigmoid(-(<s,p>-<s,n>)) for the backward.
 * bwd writes the gradient w.r.t. emb COMPACT: rows[i,:] for the sorted positions i that start a run of
 * equal node ids (rows_node[i] = node id, -1 for the other positions) and row_slot[node] = i, -1 for
 * untouched nodes; rows is (3*batch, dim).  grad_loss: device scalar, the upstream gradient.
 * Multi-GPU: emb may be the (3*batch, dim) table of rows gathered from their owners, indexed by
 * src/pos/neg = positions in that table, while node_src/node_pos/node_neg (null = same as src/pos/neg)
 * give the graph node of every slot; only nodes in [own_begin, own_end) receive a gradient row.
 * Deterministic (sorted, fixed-order accumulation; no atomics). */
KGAT_API size_t kgat_bpr_workspace_bytes(int64_t batch);
KGAT_API int kgat_bpr_loss_fwd(const float* emb, int64_t emb_stride, int dim, const int64_t* src,
                      const int64_t* pos, const int64_t* neg, int64_t batch, float reg_lambda,
                      float* loss, float* sig, void* workspace, size_t workspace_bytes,
                      kgat_stream_t stream);
KGAT_API int kgat_bpr_loss_bwd(const float* emb, int64_t emb_stride, int dim, const int64_t* src,
                      const int64_t* pos, const int64_t* neg, int64_t batch, float reg_lambda,
                      const float* sig, const float* grad_loss,
                      const int64_t* node_src, const int64_t* node_pos, const int64_t* node_neg,
                      int64_t own_begin, int64_t own_end, int64_t n_nodes,
                      float* rows, int32_t* rows_node, int32_t* row_slot,
                      void* workspace, size_t workspace_bytes, kgat_stream_t stream);

/* (6) TransR step of the KG phase -- replaces Model.transR (models.py:114-133, :46-47) and its
 *     autograd:
 *       v_x  = normalize(e_x W_R[r]) for x in {h, pos_t, neg_t};  r_v = normalize(rel_emb[r])
 *       loss = mean(-logsigmoid(|v_h + r_v - v_n|^2 - |v_h + r_v - v_p|^2))
 *            + reg_lambda * mean_b((|v_h|^2 + |r_v|^2 + |v_p|^2 + |v_n|^2) / 2)
 *     fwd writes the scalar loss and leaves the per-sample / per-chunk gradient pieces in the
 *     workspace; bwd (same workspace, untouched in between) scales them by the upstream gradient
 *     (device scalar, null = 1) and writes the DENSE gradients the reference's optimizer sees:
 *     d_entity (n_nodes, d) in full (zero rows for entities not in the batch), d_rel (n_rel, rd),
 *     d_W (n_rel, d, rd).  int64 ids as kgat.py:117-126 produces them, in range like nn.Embedding
 *     requires.  The batch is grouped by relation so W_R[r] is staged once per 32 samples instead of
 *     gathered per sample.  Deterministic (sorted, fixed-order sums).
 *     d = rd = 64, n_rel <= 1024 and 3 * batch <= 16384, else KGAT_ERR_UNSUPPORTED_DIM. */
KGAT_API size_t kgat_transr_workspace_bytes(int64_t batch, int64_t n_rel);
KGAT_API int kgat_transr_fwd(const float* entity, int64_t n_nodes, const float* rel_emb, const float* W_R,
                    int64_t n_rel, int d, int rd, const int64_t* h, const int64_t* r,
                    const int64_t* pos_t, const int64_t* neg_t, int64_t batch, float reg_lambda,
                    float* loss, void* workspace, size_t workspace_bytes, kgat_stream_t stream);
KGAT_API int kgat_transr_bwd(const float* grad_loss, int64_t n_nodes, int64_t n_rel, int d, int rd, int64_t batch,
                    float* d_entity, float* d_rel, float* d_W, const void* workspace,
                    size_t workspace_bytes, kgat_stream_t stream);

/* (7) Adam step over a list of fp32 tensors in one launch -- replaces optimizer.step() of
 *     torch.optim.Adam as the reference configures it (kgat.py:101; no amsgrad):
 *       m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *       p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)      (g += weight_decay * p first)
 *     params/grads/exp_avg/exp_avg_sq: HOST arrays of n_tensors device pointers (16-byte aligned),
 *     numel their element counts; step counts from 1.  At most 16 tensors per call. */
KGAT_API int kgat_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int64_t step, kgat_stream_t stream);

/* (5) full-rank scoring with fused per-row top-K -- replaces the per-user loop of
 *     metric.calc_recall_ndcg (metric.py:44-51): for eval user j,
 *       score[i] = <emb[user_ids[j]], emb[item_ids[i]]>,
 *       score[mask_items[mask_ptr[j] .. mask_ptr[j+1])] = 0.0   (train items, metric.py:50;
 *                                  positions into item_ids, ascending per user),
 *       topk_idx[j, :] = the K best positions, descending score, ties -> lower position.
 *     The U x I score matrix is never materialised.  topk_score may be null.
 *     dim <= 256 and K <= 32 run on the tensor cores (tcgen05 3xTF32, scores stay in TMEM); otherwise an
 *     fp32 FFMA kernel is used.  Environment: KGAT_EVAL_IMPL=ffma forces the FFMA kernel (A/B runs). */
KGAT_API int kgat_score_topk(const float* emb, int64_t emb_stride, int dim,
                    const int32_t* user_ids, int64_t n_users,
                    const int32_t* item_ids, int64_t n_items,
                    const int32_t* mask_ptr, const int32_t* mask_items,
                    int K, int32_t* topk_idx, float* topk_score, kgat_stream_t stream);

/* (5m) Recall@K / NDCG@K of the top-K lists -- replaces the per-user hit matching and the helpers of
 *      metric.py:5-31, :56-63 (recall = hits / len(pos); NDCG method 1 with the IDCG of the HIT vector), fp64:
 *        pos_items[pos_ptr[j] .. pos_ptr[j+1])  the held-out items of eval user j as positions into item_ids,
 *                                               ascending and unique;  pos_len[j] = len(test_user_dict[u_j])
 *        out_recall_ndcg[0..1] = the two means over the n_users eval users (device doubles).
 *      Deterministic (fixed-order sums).  workspace: kgat_recall_ndcg_workspace_bytes(n_users) bytes. */
KGAT_API size_t kgat_recall_ndcg_workspace_bytes(int64_t n_users);
KGAT_API int kgat_recall_ndcg(const int32_t* topk_idx, int64_t n_users, int K, const int32_t* pos_ptr,
                     const int32_t* pos_items, const int32_t* pos_len, double* out_recall_ndcg,
                     void* workspace, size_t workspace_bytes, kgat_stream_t stream);

/* (5') the same contract with a caller-provided workspace: the fast path of the evaluation.  The user rows and
 *      the item rows are split ONCE into fp16 (hi, 2^11-scaled residual) operand tiles laid out for the tensor
 *      cores inside the workspace, every CTA keeps its 128-user tile resident and streams item tiles with bulk
 *      asynchronous copies (tcgen05 kind::f16, fp32 accumulate in TMEM), the work is cut into
 *      (user tile x item range) units so that any user count fills the GPU, and the per-range top-K lists are
 *      merged (ties -> lower position).  Same results as kgat_score_topk to fp32 rounding of the scores.
 *      Covers dim <= 192 and K <= 22 (K <= 54 for dim <= 128); other shapes, KGAT_EVAL_IMPL=tf32|ffma, and
 *      operands outside the fp16 range (|x| >= 65504, detected on the device, counted by
 *      kgat_eval_fallbacks()) run the kgat_score_topk kernels.  workspace: device memory of at least
 *      kgat_score_topk_workspace_bytes(...) bytes (0: the shape is not covered and no workspace is needed),
 *      16-byte aligned, only used during the call (stream order). */
KGAT_API size_t kgat_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int dim, int K);
KGAT_API int kgat_score_topk_ws(const float* emb, int64_t emb_stride, int dim,
                       const int32_t* user_ids, int64_t n_users,
                       const int32_t* item_ids, int64_t n_items,
                       const int32_t* mask_ptr, const int32_t* mask_items,
                       int K, int32_t* topk_idx, float* topk_score,
                       void* workspace, size_t workspace_bytes, kgat_stream_t stream);
/* number of kgat_score_topk_ws calls on the current device that fell back to the fp32-range kernel since the
 * last call of this function (then cleared).  Synchronises the device. */
KGAT_API int kgat_eval_fallbacks(void);

/* ------------------------------------------------------------------------------------------------
 * (8) Multi-GPU: peer-memory exchange for the destination-partitioned step (one process per GPU of one
 *     NVLink / NVSwitch box).  New capability -- the reference is single-GPU (kgat.py:14,69-71).
 *
 *     Every rank owns an ARENA, exported with cudaIpc and mapped by every other rank.  A buffer that all
 *     ranks need in full (a layer output, dh_nb, the entity gradient, the attention weights, the batch
 *     rows of the loss) lives at the same offset of every arena; the *_push entry points below are the
 *     kernels of the single-GPU path with ONE addition: every owned row they produce is stored into the
 *     own arena AND into every peer's arena (NVLink stores issued next to the local ones), so the
 *     all-gather overlaps the kernel row by row.  kgat_peer_barrier() is what remains of the collective.
 *     The caller allocates all exchanged buffers identically on every rank (same offsets).
 *
 *     The first KGAT_PEER_HEADER_BYTES of an arena hold the barrier state and are not for payload.
 *     These are the only entry points that own memory (explicit create / destroy). */
#define KGAT_MAX_RANKS 8
#define KGAT_PEER_HEADER_BYTES 1024
#define KGAT_IPC_HANDLE_BYTES 64

typedef struct kgat_peer_group {
    int32_t  world, rank;
    uint64_t arena_bytes;                 /* the same on every rank */
    void*    arena[KGAT_MAX_RANKS];       /* arena[k]: address, in THIS process, of rank k's arena */
    void*    multicast;                   /* NVSwitch multicast mapping of the arenas of ALL ranks (one store
                                           * reaches every rank: multimem.st), or NULL: the *_push kernels then
                                           * store to every peer separately.  How the mapping is created is the
                                           * host's business (cuMulticastCreate / cuMulticastBindMem; the Python
                                           * host takes it from torch's symmetric-memory rendezvous). */
} kgat_peer_group_t;

/* cudaMalloc + zero-fill an arena on the current device and export it: handle_out_host receives the
 * KGAT_IPC_HANDLE_BYTES of its cudaIpcMemHandle_t (exchange them over torch.distributed / MPI / a file). */
KGAT_API int kgat_peer_arena_create(size_t bytes, void** arena_out_host, void* handle_out_host);
/* Map a peer's arena from its handle (peer access is enabled on demand). */
KGAT_API int kgat_peer_arena_open(const void* handle_host, void** arena_out_host);
KGAT_API int kgat_peer_arena_close(void* mapped_arena);        /* for arenas obtained from _open    */
KGAT_API int kgat_peer_arena_destroy(void* own_arena);         /* for the arena obtained from _create */

/* All ranks enqueue the same sequence of barriers.  On the device: publish everything this rank's stream
 * has written so far (system-scope release), signal every peer, wait until every peer has signalled the
 * same barrier count (acquire).  One 32-thread CTA, no host involvement, CUDA-graph capturable.
 * A peer that does not arrive within timeout_ms (<= 0: 20 000) sets the arena's error flag instead of
 * hanging the GPU; later barriers then return at once.  kgat_peer_error() reads the flag (synchronises). */
KGAT_API int kgat_peer_barrier(const kgat_peer_group_t* pg, int timeout_ms, kgat_stream_t stream);
KGAT_API int kgat_peer_error(const kgat_peer_group_t* pg);

/* kgat_segment_softmax whose w_csr lies in the own arena and is stored to every peer as well (the
 * attention weights of ALL edges in global CSR order = the concatenation of the ranks' slices). */
KGAT_API int kgat_segment_softmax_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                              const float* logits_csr, float* w_csr, float* w_eid,
                              const kgat_peer_group_t* pg, kgat_stream_t stream);
/* kgat_layer_fwd whose h_out (full (n_nodes, d_out) buffer in the own arena) is stored to every peer. */
KGAT_API int kgat_layer_fwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                        const float* h_in, const float* w_csr, const float* W2,
                        const float* drop_mask, int d_in, int d_out, float slope,
                        float* h_nb, float* act, float* h_out,
                        float* concat, int64_t concat_stride, int64_t out_off, int64_t in_off,
                        const kgat_peer_group_t* pg, kgat_stream_t stream);
/* kgat_layer_bwd whose dh_nb (full (n_nodes, d_in) buffer in the own arena) is stored to every peer. */
KGAT_API int kgat_layer_bwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                        const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride, int64_t out_off,
                        const float* nxt_dx_direct, const float* nxt_dh_nb, const float* w_csc,
                        const float* h_in, const float* h_nb, const float* act, const float* h_out,
                        const float* drop_mask, const float* W2, int d_in, int d_out, float slope,
                        float* dx_direct, float* dh_nb, float* dW2, float* dW2_partials,
                        const kgat_peer_group_t* pg, kgat_stream_t stream);
/* kgat_input_bwd whose d_entity (full (n_nodes, d) buffer in the own arena) is stored to every peer. */
KGAT_API int kgat_input_bwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                        const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride, int64_t in_off,
                        const float* dx_direct, const float* dh_nb, const float* w_csc, int d,
                        float* d_entity, const kgat_peer_group_t* pg, kgat_stream_t stream);
/* Batch rows of the loss head: table[s, :] = emb[ids[s], :] for every slot s whose node
 * ids[s] lies in [own_begin, own_end), written into the (n_slots, dim) table of EVERY rank (own arena
 * included); after the barrier every rank holds all rows.  Replaces gather + mask + all-reduce. */
KGAT_API int kgat_peer_gather_rows_push(const float* emb, int64_t emb_stride, int dim, const int64_t* ids,
                               int64_t n_slots, int64_t own_begin, int64_t own_end, float* table,
                               const kgat_peer_group_t* pg, kgat_stream_t stream);
/* Small replicated sums (the dW2 of all layers): slots is a (world, n) buffer in the arena.
 * _push_slot copies src[0..n) into slot `rank` of every rank's buffer; after the barrier _sum_slots
 * adds the world slots in rank order (identical bits on every rank) into out[0..n). */
KGAT_API int kgat_peer_push_slot(const float* src, int64_t n, float* slots, const kgat_peer_group_t* pg,
                        kgat_stream_t stream);
/* dst[0..n) = src[0..n) in the own arena and in every peer's (dst lies in the own arena, n % 4 == 0, 16-byte
 * aligned): the all-gather push of one contiguous buffer.  Through the multicast mapping the NVSwitch writes
 * every copy, the own one included. */
KGAT_API int kgat_peer_push_buffer(const float* src, int64_t n, float* dst, const kgat_peer_group_t* pg,
                          kgat_stream_t stream);
KGAT_API int kgat_peer_sum_slots(const float* slots, int64_t n, int world, float* out, kgat_stream_t stream);

/* Hardware self-test of the tcgen05 path (diagnostics; not on the hot path):
 * C[128,N] = A[128,K] * B[N,K]^T on one CTA, mode 0 = one TF32 pass, mode 1 = 3xTF32.
 * *err_flag (device int) is set to 1 if the tensor-core completion barrier never fires. */
/* 1 if a tensor-core completion barrier ever timed out since the last call (then cleared); 0 otherwise.
 * Synchronises the device. */
KGAT_API int kgat_umma_error_flag(void);
KGAT_API int kgat_selftest_umma(const float* A, const float* B, float* C, int N, int K, int variant, int mode,
                       int* err_flag, kgat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KGAT_B200_H */
