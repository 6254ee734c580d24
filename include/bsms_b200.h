/*
 * bsms_b200.h — C-ABI of libbsms_b200.so, the B200 (sm_100a) implementation of the BSMS-GNN
 * processor hot path.
 *
 * The reference (Eydcao/BSMS-GNN) is pure Python and has no FFI of its own; its "operator
 * interface" for this path is the set of torch.nn.Modules in src/ops/basic.py and src/ops/BSMS.py.
 * Every entry point below replaces the arithmetic of one of those (file:line cited per function)
 * and is bound from Python with ctypes by bsms_gnn_b200/_lib.py (see INTEGRATION.md for the
 * reference-side stub).
 *
 * Conventions
 *   - plain C: pointers + sizes only.  Unless a parameter says "host", every pointer is a DEVICE
 *     pointer on the current CUDA device; `stream` is a cudaStream_t passed as void*.
 *   - features are fp32 row-major [B, N, C] (B = 1 for un-batched tensors), C = 128 for latent
 *     features; edge lists are int64 [2, E] exactly as the reference's m_gs (src/ops/BSMS.py:39-57).
 *   - every function returns 0 on success, a negative BSMS_E* code on failure; the message is
 *     available from bsms_last_error() (thread-local).  Nothing falls back to the CPU.
 *   - all kernels are asynchronous on `stream` and graph-capturable unless stated otherwise.
 */
#ifndef BSMS_B200_H
#define BSMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSMS_OK 0
#define BSMS_EINVAL (-1)   /* bad argument (shape, null pointer, unsupported width)            */
#define BSMS_ECUDA (-2)    /* CUDA runtime error; bsms_last_error() has cudaGetErrorString     */
#define BSMS_EINDEX (-3)   /* edge / node index out of range (reference: ATen index error)     */
#define BSMS_EWORKSPACE (-4) /* workspace too small                                            */

#define BSMS_LATENT 128    /* the only latent width built (configs/model/(any).yaml: latent_dim 128) */

/* arithmetic of the MLP contractions */
#define BSMS_MODE_FP32 0      /* fp32 FFMA, exact fp32 products                                  */
#define BSMS_MODE_FP16X3 1    /* tcgen05 kind::f16, 2-way fp16 split (3 MMAs), fp32 accumulate   */
#define BSMS_MODE_BF16 2      /* tcgen05 kind::f16 bf16 operands (1 MMA), fp32 accumulate        */

const char* bsms_last_error(void);
int bsms_version(void);
/* Device facts the host uses to size grids / report rooflines. out[0]=SM count, out[1]=L2 bytes,
 * out[2]=max dynamic smem per block, out[3]=compute capability major*10+minor.  (host pointer) */
int bsms_device_info(int64_t* out4_host);

/* ---------------------------------------------------------------------------------------------
 * Level plan: the per-graph index structures every kernel consumes.  Built once per mesh level
 * from the reference's int64 edge list g = m_gs[l]  (g[0] = sender i, g[1] = receiver j;
 * src/ops/basic.py:66).  All arrays are int32 device buffers owned by the caller.
 *   dst-sorted order ("_d", stable): position k holds edge perm_d[k]; rowptr_d is the CSR of
 *     receivers — GMP aggregation and the down transfer reduce over rows of it.
 *   src-sorted order ("_s", stable): CSR of senders — the up transfer (aggragating=False) and
 *     every backward gather reduce over rows of it; s2d[k] is the dst-sorted position of the edge
 *     at src-sorted position k.
 * ------------------------------------------------------------------------------------------- */
typedef struct bsms_level_plan {
  int32_t n_nodes;
  int32_t n_edges;
  const int32_t* src_d;    /* [E] */
  const int32_t* dst_d;    /* [E] */
  const int32_t* rowptr_d; /* [N+1] */
  const int32_t* perm_d;   /* [E] */
  const int32_t* src_s;    /* [E] */
  const int32_t* dst_s;    /* [E] */
  const int32_t* rowptr_s; /* [N+1] */
  const int32_t* s2d;      /* [E] */
} bsms_level_plan;

size_t bsms_plan_workspace_bytes(int64_t n_edges, int64_t n_nodes);
/* Fills the eight arrays of `out` (pre-allocated by the caller with the sizes above) from g.
 * status_dev: int32[4] device scratch; after the call (which synchronises the stream — plan
 * building is one-time per mesh, off the hot path) an out-of-range index yields BSMS_EINDEX and
 * status_dev[1] holds max(g[0]) (the reference's degree() sizes itself by it, src/utils/basic.py:305-307). */
int bsms_plan_build(const int64_t* g, int64_t n_edges, int64_t n_nodes, const bsms_level_plan* out,
                    int32_t* status_dev, void* workspace, size_t workspace_bytes, void* stream);

/* Content fingerprints of n <= 32 device buffers (sizes multiples of 8 bytes) in one launch:
 * out_dev[k] = position-dependent 64-bit hash of buffer k.  The reference re-uploads m_gs / m_ids
 * every step (src/trainer/trainer.py:100-117, src/models/model.py:189-200), so the host caches its
 * plans by CONTENT: one launch + one 8n-byte read-back decides whether a mesh was seen before.
 * ptrs / nbytes are HOST arrays.  Asynchronous on `stream`. */
int bsms_fingerprint(const void* const* ptrs_host, const int64_t* nbytes_host, int32_t n,
                     uint64_t* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * WeightedEdgeConv.cal_ew (src/ops/basic.py:142-167): deg_i = out-degree, s_e = w_i/deg_i,
 * aggr_w_j = sum_{e: dst=j} s_e + 1e-12, ew_e = s_e / aggr_w_j.   w: [N].
 * Outputs: ew_orig [E] in the caller's edge order (what the reference returns), ew_d / ew_s the
 * same weights in the plan's two orders (what the transfer kernels read), aggr_w [N].
 * Any of ew_orig may be NULL.  The reference raises when max(g[0])+1 != N; the host checks that.
 * ------------------------------------------------------------------------------------------- */
int bsms_cal_ew(const bsms_level_plan* plan, const float* w, float* ew_orig, float* ew_d,
                float* ew_s, float* aggr_w, void* stream);
/* ew given in the caller's edge order -> the plan's two orders (for WeightedEdgeConv.forward
 * called with arbitrary user weights, src/ops/basic.py:107). */
int bsms_permute_ew(const bsms_level_plan* plan, const float* ew_orig, float* ew_d, float* ew_s,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * WeightedEdgeConv.forward (src/ops/basic.py:107-140), x: [B,N,C] -> out: [B,N,C], any C >= 1.
 *   up == 0 (aggragating=True):  out[j] = sum_{e: dst=j} ew_e x[i_e]    (reads ew_d)
 *   up == 1 (aggragating=False): out[i] = sum_{e: src=i} ew_e x[j_e]    (reads ew_s)
 * The two are adjoint, so each is the other's backward.  Atomic-free, deterministic.
 * ------------------------------------------------------------------------------------------- */
int bsms_edge_conv(const bsms_level_plan* plan, const float* ew, const float* x, float* out,
                   int32_t B, int32_t C, int32_t up, void* stream);
/* Down transfer fused with pooling (src/ops/BSMS.py:74-89): out[b,k,:] = conv_down(x)[b,ids[k],:],
 * only the kept rows are formed.  ids: int32 [n_keep]. */
int bsms_conv_down_pool(const bsms_level_plan* plan, const float* ew_d, const int32_t* ids,
                        int32_t n_keep, const float* x, float* out, int32_t B, int32_t C,
                        void* stream);
/* Unpool fused with the up transfer (src/ops/BSMS.py:98-100): out = conv_up(unpool(hc)) without
 * materialising the zero-filled tensor.  inv: int32 [N], coarse row of a fine node or -1. */
int bsms_unpool_conv_up(const bsms_level_plan* plan, const float* ew_s, const int32_t* inv,
                        int32_t n_keep, const float* hc, float* out, int32_t B, int32_t C,
                        void* stream);
/* Pooling gather out[b,k,:] = x[b,ids[k],:] (src/ops/BSMS.py:79-89) and Unpool.forward
 * (src/ops/basic.py:176-201; zero-fill + row injection). */
int bsms_gather_rows(const float* x, const int32_t* ids, int32_t n_keep, int32_t n_rows, float* out,
                     int32_t B, int32_t C, void* stream);
int bsms_unpool_rows(const float* h, const int32_t* ids, int32_t n_keep, int32_t n_rows, float* out,
                     int32_t B, int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GMP block (src/ops/basic.py:26-98), latent 128, hidden_layer 3, pos_dim P in {1,2,3}.
 * Weights are the reference's own tensors, untouched: w_edge[l] = mlp_edge.seq.{2l}.weight
 * ([128, 2*128+P+1] for l=0, [128,128] after), b_edge[l] its bias, same for mlp_node
 * ([128,256] for l=0).  Column order of layer 0: edge [dir(P), norm, x_src(128), x_dst(128)]
 * (basic.py:85-90), node [x(128), aggr(128)] (basic.py:97).
 *   out = LN(mlp_node([x, sum_{e->j} LN(mlp_edge([fiber_e, x_i, x_j]))])) + x  (+ skip if given)
 * x: [B,N,128]; pos: [N,P] (pos_batched=0) or [B,N,P]; out: [B,N,128].
 * ------------------------------------------------------------------------------------------- */
typedef struct bsms_gmp_weights {
  const float* w_edge[4];
  const float* b_edge[4];
  const float* w_node[4];
  const float* b_node[4];
} bsms_gmp_weights;

typedef struct bsms_gmp_grads { /* accumulated into (+=); all required in backward */
  float* w_edge[4];
  float* b_edge[4];
  float* w_node[4];
  float* b_node[4];
} bsms_gmp_grads;

/* bytes of scratch bsms_gmp_forward / bsms_gmp_backward need for this problem size */
size_t bsms_gmp_workspace_bytes(int32_t B, int32_t n_nodes, int32_t n_edges, int32_t mode,
                                int32_t backward);
/* `saved` (may be NULL): bsms_gmp_saved_bytes(B, n_nodes) bytes that receive the node-level
 * intermediates (projected rows, aggregated messages, node-MLP activations; no per-edge tensor)
 * and the packed 16-bit weight images of this call, so that backward neither recomputes nor
 * re-packs them.  The weights must not change between the forward and its backward. */
size_t bsms_gmp_saved_bytes(int32_t B, int32_t n_nodes);
int bsms_gmp_forward(const bsms_level_plan* plan, const bsms_gmp_weights* w, const float* x,
                     const float* pos, int32_t pos_batched, const float* skip /* may be NULL */,
                     float* out, float* saved /* may be NULL */, int32_t B, int32_t P, int32_t mode,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Inference / rollout form: the 16-bit operand images of a GMP's weights can be made ONCE
 * (bsms_gmp_pack into bsms_gmp_packed_bytes() bytes, tensor-core modes only) and handed to every
 * forward while the weights do not change — the reference's rollout (src/utils/rollout_utils.py:48-62)
 * calls the same 13 blocks 599 times.  `packed` may be NULL (then it is bsms_gmp_forward). */
size_t bsms_gmp_packed_bytes(void);
int bsms_gmp_pack(const bsms_gmp_weights* w, int32_t P, int32_t mode, void* packed, void* stream);
int bsms_gmp_forward_packed(const bsms_level_plan* plan, const bsms_gmp_weights* w, const void* packed,
                            const float* x, const float* pos, int32_t pos_batched, const float* skip,
                            float* out, float* saved, int32_t B, int32_t P, int32_t mode,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the block above by recomputation (nothing is saved by forward): given g_out
 * [B,N,128] writes g_x [B,N,128] (gradient w.r.t. x INCLUDING the residual path; the gradient
 * w.r.t. skip is g_out itself) and accumulates the 16 parameter gradients.  No gradient flows to
 * pos (it is an input, src/ops/BSMS.py:75 runs under the same autograd but pos never requires
 * grad in the reference's use). */
int bsms_gmp_backward(const bsms_level_plan* plan, const bsms_gmp_weights* w, const float* x,
                      const float* pos, int32_t pos_batched,
                      const float* saved /* from forward with the same inputs, or NULL to recompute */,
                      const float* g_out, float* g_x,
                      const bsms_gmp_grads* grads, int32_t B, int32_t P, int32_t mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Test hook for the fused tcgen05 edge stage (mode FP16X3 or BF16): runs the per-node
 * pre-projection and the fused kernel and dumps one intermediate per edge row, dst-sorted order,
 * into dbg [B*E,128]: stage 0 = a0 after gather+ReLU, 1/2 = activations after edge layers 1/2,
 * 3 = layer-3 output before LayerNorm.  aggr [B*N,128] receives the aggregated messages. */
int bsms_debug_edge_stage(const bsms_level_plan* plan, const bsms_gmp_weights* w, const float* x,
                          const float* pos, int32_t pos_batched, int32_t B, int32_t P, int32_t mode,
                          int32_t stage, float* dbg, float* aggr, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The rest of the training step (src/trainer/trainer.py:79-98,134-156).  Device-resident scalars
 * only: no host synchronisation, CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------- */
/* Masked RMSE (trainer.py:96-98): loss = sqrt(sum((pred-tar)^2 * mask) / sum(mask) / C) with
 * pred, tar [rows, C], mask [rows] (the reference's [B,N,1] mask).  acc2_dev: double[2] scratch.
 * loss_dev (float[1], may be NULL) receives the loss; grad_pred ([rows, C], may be NULL) receives
 * d loss / d pred times *g_loss_dev (1 if g_loss_dev is NULL). */
int bsms_masked_rmse(const float* pred, const float* tar, const float* mask, int64_t rows, int32_t C,
                     double* acc2_dev, const float* g_loss_dev, float* loss_dev, float* grad_pred,
                     void* stream);
/* One optimiser update over flat fp32 buffers of n values (16-byte aligned): global-norm clipping
 * (torch.nn.utils.clip_grad_norm_, trainer.py:151; max_norm <= 0 disables it) followed by AdamW
 * (torch.optim.AdamW as configured at trainer.py:24-28) at the learning rate of the reference's
 * WarmupCosineDecayScheduler (src/utils/basic.py:168-184; warmup_steps = decay_steps = 0 keeps the
 * rate constant).  state_dev: double[2] = {number of updates done so far, scratch}; zero it once.
 * hyper_dev: float[4] scratch that receives {lr, 1-beta1^t, sqrt(1-beta2^t), clip coefficient} of
 * this update (readable afterwards).  zero_grad != 0 clears the gradient buffer in the same pass. */
int bsms_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                         double* state_dev, float* hyper_dev, double peak_lr, double warmup_steps,
                         double decay_steps, double beta1, double beta2, double eps,
                         double weight_decay, double max_norm, int32_t zero_grad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The caller side of the processor for inference / rollouts, fused around it
 * (src/models/model.py:83-106,127-164; src/utils/normalizer.py:39-52,80-83;
 * src/utils/rollout_utils.py:48-62).  node_in rows are [state(C) | mesh_pos(P) | node_type(1)],
 * Cin = C + P + 1, C = out_dim <= 4.  mean / std are HOST arrays of doubles (the reference's
 * normalisers are fp64: mean() and std_with_epsilon()).
 * ------------------------------------------------------------------------------------------- */
/* a1 = relu(W0 * normalise([state, type]) + b0) [rows,128] (first encoder layer, model.py:137-147,
 * W0 = encode.seq.0.weight [128, C+1]) and pos [rows,P] (model.py:62). */
int bsms_encode_in(const float* node_in, int64_t rows, int32_t Cin, int32_t C, int32_t P,
                   const double* mean_host, const double* std_host, const float* W0, const float* b0,
                   float* a1, float* pos, void* stream);
/* n_layers (1..3) Linear 128->128 layers with ReLU where bit l of relu_mask is set, optionally
 * followed by LayerNorm without affine (src/ops/basic.py:6-23), in arithmetic `mode`.  W / b are
 * HOST arrays of device pointers; `packed` from bsms_dense128_pack (unused in BSMS_MODE_FP32);
 * scratch: 2 * rows * 128 floats. */
size_t bsms_dense128_packed_bytes(int32_t mode);
int bsms_dense128_pack(const float* const* W_host, int32_t n_layers, int32_t mode, void* packed, void* stream);
int bsms_dense128_stack(const float* x, int64_t rows, const float* const* W_host, const float* const* b_host,
                        int32_t n_layers, int32_t relu_mask, int32_t layer_norm, int32_t mode,
                        const void* packed, float* out, float* scratch, void* stream);
/* pred = state + mask * denormalise(W3 y + b3)  (last decoder layer decode.seq.6 [C,128],
 * model.py:150-163) and, when next_in != NULL, the next rollout input
 * next_in = mask == 0 ? ic : [pred | pos | type]  (rollout_utils.py:57-62; ic may be NULL). */
int bsms_decode_out(const float* y, int64_t rows, int32_t Cin, int32_t C, const float* W3, const float* b3,
                    const double* mean_host, const double* std_host, const float* node_in,
                    const float* mask, const float* ic, float* pred, float* next_in,
                    int32_t pos_feedback /* deforming meshes (pos_dim == out_dim): next_in's position channels
                                            receive the new state instead of passing through */,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6: halo exchange of the node-partitioned processor (one process per GPU of one NVLink box) as
 * ONE kernel per exchange over peer-mapped memory.  The reference has no partitioning; this is the
 * exchange step SURVEY.md §8e derives: before every operator that reads neighbours the ghost rows
 * of a level are refreshed from their owners, and the adjoint returns ghost gradients in backward.
 * Arenas come from bsms_ipc_alloc (cudaMalloc, zero-filled) and are mapped into the peers with
 * bsms_ipc_export / bsms_ipc_open (CUDA IPC, 64-byte handles, host pointers).
 * ------------------------------------------------------------------------------------------- */
int bsms_ipc_alloc(size_t bytes, void** ptr_out_host);
int bsms_ipc_free(void* ptr);
int bsms_ipc_export(void* ptr, uint8_t* handle64_host);
int bsms_ipc_open(const uint8_t* handle64_host, void** ptr_out_host);
int bsms_ipc_close(void* ptr);

typedef struct bsms_halo_args {
  int32_t world, rank, channels /* floats per row, multiple of 4 */, backward;
  int64_t n_own, n_ghost, n_send;
  const float* src;         /* forward: x_own [n_own, C];  backward: g_local [n_own + n_ghost, C]        */
  float* dst;               /* forward: out [n_own + n_ghost, C] IN THIS RANK'S ARENA (peers write its ghost
                               rows); backward: g_own [n_own, C]                                          */
  const int64_t* send_idx;  /* [n_send] owned rows this rank sends, grouped by destination peer (device)  */
  int32_t send_off[9];      /* [world + 1] row offsets of the groups of send_idx                          */
  int32_t recv_off[9];      /* [world + 1] row offsets of this rank's ghost rows, grouped by owner        */
  void* peer_dst[8];        /* forward: where my rows start in peer q's out buffer; backward: where my ghost
                               gradients start in owner q's `back` region (peer-mapped device pointers)    */
  const float* back;        /* backward: this rank's own `back` region [n_send, C] (peers write it)       */
  void* my_flags;           /* uint32[world] in this rank's arena: slot q is written by peer q            */
  void* peer_flag[8];       /* address of slot [rank] inside peer q's flags of the same call site         */
  void* ctrl;               /* uint32[4] in this rank's arena, zero-initialised: epoch + arrive counters  */
} bsms_halo_args;
/* One exchange (see above).  Asynchronous on `stream`, CUDA-graph capturable; returns when launched.
 * Every rank of the group must call it for the same call site in the same order. */
int bsms_halo_exchange(const bsms_halo_args* args, void* stream);

/* Training noise of the reference's datapipe on the device (src/datasets/base.py:274-289): per node
 * and output channel c, n ~ N(0, noise_level[c]) (zero where mask == 0); node_in[:, c] += n,
 * node_tar[:, c] += (1 - noise_gamma) n.  node_in [rows, Cin], node_tar [rows, C], C <= 4.
 * Counter-based Philox stream keyed by (seed, row, offset): reproducible, NOT torch's CPU stream. */
int bsms_inject_noise(float* node_in, int32_t Cin, float* node_tar, int32_t C, const float* mask,
                      int64_t rows, const float* noise_level_host, float noise_gamma, uint64_t seed,
                      uint64_t offset, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Native bi-stride hierarchy builder, HOST code (OpenMP): the integer part of the reference's
 * BistrideMultiLayerGraph (src/graph_wrappers/bsms_graph_wrapper.py:58-154,
 * src/graph_wrappers/graph_wrapper.py:67-134).  All pointers are HOST pointers; flat_edge is the
 * reference's int64 [2, E] list.  Exact (integer) results.
 * ------------------------------------------------------------------------------------------- */
/* labels_out[n]: weakly connected cluster of every node, clusters numbered by their smallest node. */
int bsms_components_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes,
                         int64_t* labels_out, int64_t* n_comp_out);
/* One pooling level: BFS parity from one seed per cluster (seeds chosen by the caller: the node
 * nearest the cluster centroid, bsms_graph_wrapper.py:107-126), keep the smaller parity class,
 * new edges = pattern of (A+I)^2 minus the diagonal on the kept nodes, re-indexed, row-major with
 * sorted columns.  keep_out: capacity n_nodes; *edges_out: malloc'ed int64 [2, *n_edges_out],
 * release with bsms_host_free. */
int bsms_bistride_level_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes,
                             const int64_t* labels, int64_t n_comp, const int64_t* seeds,
                             int64_t* keep_out, int64_t* n_keep_out, int64_t** edges_out,
                             int64_t* n_edges_out);
void bsms_host_free(void* p);

/* The whole hierarchy in one call (replaces BistrideMultiLayerGraph(...).get_multi_layer_graphs(),
 * src/graph_wrappers/bsms_graph_wrapper.py:9-28, including its seed choice :107-126 in the positions' own
 * floating-point type and numpy's operation order).  flat_edge int64 [2, n_edges]; pos [n_nodes, pos_dim] float32
 * (pos_is_f64 = 0) or float64; depth = number of pooling levels.  *handle_out owns the result. */
int bsms_hierarchy_build_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes, const void* pos,
                              int32_t pos_dim, int32_t pos_is_f64, int32_t depth, void** handle_out);
/* level in [1, depth]: its edge list int64 [2, *n_edges_out] (row-major, sorted columns) and the ids of its
 * *n_nodes_out nodes in level - 1 (ascending).  The pointers stay valid until bsms_hierarchy_free_host. */
int bsms_hierarchy_level_host(void* handle, int32_t level, int64_t* n_nodes_out, int64_t* n_edges_out,
                              const int64_t** edges_out, const int64_t** ids_out);
void bsms_hierarchy_free_host(void* handle);

/* Test hook for the split-operand tensor-core layer of the fp32-parity backward: Y[rows,128] = X W^T (b_mn = 0) or
 * X W (b_mn = 1), optionally masked by (mask > 0); a_is_grad != 0 scales X by its own max (gradient operands), else
 * by the static activation scale.  scratch: 64 KB + 64 bytes of device memory. */
int bsms_debug_lin_split(const float* X, int64_t rows, const float* W, int32_t b_mn, const float* mask,
                         int32_t a_is_grad, float* Y, void* scratch, void* stream);

/* Deterministic option (process-wide switch, default off): bitwise run-to-run reproducible results.
 *  - BSMS_MODE_FP32: segment sums walk CSR rows in order without atomics; with the switch on the split-over-rows
 *    weight-gradient kernels commit their partial sums in a fixed (ticket) order.
 *  - BSMS_MODE_BF16: the fused tcgen05 edge kernels write their per-edge-row results as [B*E,128] rows and order-fixed
 *    CSR segment sums replace the red.add reductions into node rows; every per-CTA atomic flush of a weight / bias
 *    gradient becomes a per-CTA partial-sum block + one ordered reduction (bsms_gmp_workspace_bytes grows accordingly —
 *    set the switch BEFORE sizing workspaces).
 *  - BSMS_MODE_FP16X3: bsms_gmp_forward / bsms_gmp_backward return BSMS_EINVAL while the switch is on (use
 *    BSMS_MODE_FP32, the same 1e-5 grade).
 * The transfer operators are order-fixed CSR sums in every mode.  The reference offers
 * torch.use_deterministic_algorithms for the same purpose (its scatter_add_ is otherwise order-dependent on CUDA,
 * src/utils/basic.py:287-343). */
void bsms_set_deterministic(int32_t on);
int32_t bsms_get_deterministic(void);

/* Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline pass).
 * Kinds: 0 edge-MLP forward GEMM/chain, 1 node-level forward GEMMs, 2 edge gather+combine,
 * 3 LayerNorm+segment-sum, 4 dgrad, 5 wgrad, 6 LayerNorm backward, 7 edge-gradient segment sums,
 * 8 transfer (restriction/prolongation/conv), 9 other, 10 fused tcgen05 edge stage, 11 its fused backward.
 * Host pointers; collect synchronises. */
#define BSMS_PROF_KINDS 12
int bsms_prof_enable(int on);
int bsms_prof_collect(double* ms_by_kind_host, int64_t* launches_by_kind_host, int n_kinds);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t bsms_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BSMS_B200_H */
