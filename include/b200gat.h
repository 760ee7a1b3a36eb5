/* b200gat.h -- C ABI of libb200gat.so: the B200 (sm_100a) GAT training hot path.
 *
 * The reference (Axionis47/PlotPointe-GAT-Recommendation) is pure Python and has no FFI of its own;
 * the interface it offers for this path is two torch modules and five loss lines.  Each entry point
 * below names the reference lines it replaces.  A Python/ctypes binding (the one this repo ships in
 * plotpointe-gat-recommendation_b200/_lib.py, and the one a reference maintainer would add -- see
 * INTEGRATION.md) passes raw device pointers, sizes and a cudaStream_t; there are no torch types in
 * any signature.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer unless it says "host"; buffers are caller-owned;
 *   - all float tensors are fp32, row-major, 16-byte aligned; index outputs are int32;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); every call is asynchronous;
 *   - return value: 0 = ok, <0 = error (B200GAT_ERR_*); b200gat_last_error() returns the message of
 *     the last failure on the calling host thread;
 *   - there is no CPU path: with no CUDA device every launch fails with B200GAT_ERR_CUDA.
 */
#ifndef B200GAT_H_
#define B200GAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GAT_ABI_VERSION 1

#define B200GAT_ERR_ARG (-1)
#define B200GAT_ERR_CUDA (-2)
#define B200GAT_ERR_UNSUPPORTED (-3)
#define B200GAT_ERR_WORKSPACE (-4)

/* softmax dialect of the edge kernels */
#define B200GAT_POLICY_CUSTOM 0 /* SimpleGATLayer: clamp[-10,10], no max, denom+1e-9 (train_gat_custom.py:80-88) */
#define B200GAT_POLICY_PYG 1    /* PyG GATConv: max-subtracted, denom+1e-16, head mean + bias (train_gat_pyg.py:77,87) */

#define B200GAT_LOSS_BPR 0 /* train_gat_custom.py:354-355 */
#define B200GAT_LOSS_BCE 1 /* train_gat_custom.py:356-359 */

const char* b200gat_last_error(void);
int b200gat_abi_version(void);
/* number of kernels launched through this library by the calling process (monotonic) */
int64_t b200gat_launch_count(void);

/* ---- (1) COO -> CSR/CSC -------------------------------------------------------------------------
 * Replaces the edge ordering that the reference's per-edge scatter ops imply
 * (denom.scatter_add_ / out.index_add_ over edge_index, scripts/train_gat_custom.py:85-92; the list
 * itself comes from build_edge_index, :166-175).  CSR = stable sort by destination (edge_index row
 * 1), CSC = stable sort by source (row 0); duplicates kept; bit-exact with
 * `edge_index[1].argsort(stable=True)`.
 *   edge_index : int64 [2, n_edges] (row 0 = src, row 1 = dst)
 *   rowptr,colptr : int32 [n_nodes+1]     col : src ids in CSR order     row : dst ids in CSC order
 *   perm, perm_csc : original edge ids in CSR / CSC order
 *   csr2csc : CSC position of the edge at each CSR position
 *   n_bad : device int32, number of edges with an endpoint outside [0, n_nodes) (caller checks == 0)
 */
int b200gat_graph_workspace_bytes(int64_t n_nodes, int64_t n_edges, size_t* bytes /*host*/);
int b200gat_build_graph(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int32_t* rowptr, int32_t* col,
                        int32_t* perm, int32_t* colptr, int32_t* row, int32_t* perm_csc, int32_t* csr2csc,
                        int32_t* n_bad, void* workspace, size_t workspace_bytes, void* stream);

/* Row schedule for the edge kernels: sched[k] = (row, ptr[row], ptr[row+1], 0) (int32 x4) for the k-th row of
 * `ptr` (a CSR rowptr or CSC colptr slice, n_rows+1 entries) in descending-degree order.  The persistent warps
 * of (3)/(4) walk it round-robin: hubs start first and every warp gets the same degree mix.
 * degree_bound: any value > the largest degree (e.g. n_edges+1); <= 0 keeps the natural row order. */
int b200gat_schedule_workspace_bytes(int64_t n_rows, size_t* bytes /*host*/);
int b200gat_build_schedule(const int32_t* ptr, int64_t n_rows, int64_t degree_bound, int32_t* sched, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- (2) projection + attention logits ----------------------------------------------------------
 * h = x W^T (self.lin(x), train_gat_custom.py:77; GATConv.lin) and the per-node halves of the edge
 * logit, s[:, 0:H] = (h * a_src).sum(-1), s[:, H:2H] = (h * a_dst).sum(-1) (:79).
 *   x [n_rows, in_features]; W [heads*channels, in_features]; a_src, a_dst [heads, channels]
 *   h [n_rows, heads*channels]; s [n_rows, 2*heads]
 * precision: B200GAT_GEMM_FP32 = CUDA-core FFMA; B200GAT_GEMM_TF32X3 = tcgen05 tensor cores, every
 * fp32 operand split into tf32 hi + lo and four UMMAs per K step (element error ~1e-7 relative).
 * The tensor-core path needs in_features == channels == 128; other shapes use the FP32 kernels.
 */
#define B200GAT_GEMM_FP32 0
#define B200GAT_GEMM_TF32X3 1
int b200gat_set_gemm_mode(int mode); /* process-wide; default B200GAT_GEMM_TF32X3 where the shape allows it */
int b200gat_get_gemm_mode(void);
int b200gat_project_f32(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                        int in_features, int heads, int channels, float* h, float* s, void* workspace,
                        size_t workspace_bytes, void* stream);

/* "bf16 projection" (BASELINE configs 3 and 5): operands rounded to bf16 on the way into shared memory, one UMMA
 * kind::f16 per K step of 16, fp32 accumulation in TMEM; h is STORED as bf16 [n_rows, heads*channels] (the edge kernels
 * then gather half the bytes), s is taken from the fp32 accumulator.  in_features and channels in {128, 256},
 * heads * channels <= 1024; workspace >= b200gat_dense_workspace_bytes(heads, channels, in_features).
 * project_bwd_bf16: the matching backward -- dx = dh_full W as ONE bf16 tensor-core launch contracting over all heads
 * (K = heads * channels), dW = dh_full^T x in fp32-accurate TF32-split tiles, da_src / da_dst; same arguments as
 * b200gat_project_bwd_f32 (dh is NOT modified). */
int b200gat_project_bf16(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                         int in_features, int heads, int channels, void* h_bf16, float* s, void* workspace,
                         size_t workspace_bytes, void* stream);
int b200gat_project_bwd_bf16(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh,
                             const float* ds, int64_t n_rows, int in_features, int heads, int channels, float* dx, float* dW,
                             float* da_src, float* da_dst, void* workspace, size_t workspace_bytes, void* stream);
/* _ex forms for the sharded / streaming path: x may already be stored as bf16 (x_is_bf16; the exchanged layer inputs of the
 * bf16 tier), and dx may accumulate (one call per head into the same dx). */
int b200gat_project_bf16_ex(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                            int in_features, int heads, int channels, void* h_bf16, float* s, void* workspace,
                            size_t workspace_bytes, void* stream);
int b200gat_project_bwd_bf16_ex(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst,
                                const float* dh, const float* ds, int64_t n_rows, int in_features, int heads, int channels,
                                float* dx, int accumulate_dx, float* dW, float* da_src, float* da_dst, void* workspace,
                                size_t workspace_bytes, void* stream);

/* Backward of (2).  `dh` holds the aggregation part on entry and may be overwritten with the full
 * gradient of h (adds ds_src*a_src + ds_dst*a_dst); ds = [ds_src | ds_dst] [n_rows, 2*heads].
 * workspace (both directions): b200gat_dense_workspace_bytes.
 * Outputs: dx [n_rows, in_features] (may be NULL), dW [heads*channels, in_features], da_src, da_dst. */
int b200gat_dense_workspace_bytes(int heads, int channels, int in_features, size_t* bytes /*host*/);
int b200gat_project_bwd_f32(const float* x, const float* W, const float* a_src, const float* a_dst, float* dh,
                            const float* ds, int64_t n_rows, int in_features, int heads, int channels, float* dx,
                            float* dW, float* da_src, float* da_dst, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Item-feature projection of the node-feature assembly (item_proj, train_gat_custom.py:100,105-109):
 * y[n, 0:out_features] (row pitch ldy) = x W^T + bias, so the result can be written straight into the tail rows of the
 * [N, C] node-feature buffer (no torch.cat copy); backward: dW = dy^T x, dbias = column sums (x needs no gradient). */
int b200gat_linear_f32(const float* x, const float* W, const float* bias, int64_t n_rows, int in_features,
                       int out_features, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);
/* The same projection on the tensor cores (TF32 hi/lo split): used by the bf16 tier, where the next step rounds x to bf16.
 * b200gat_linear_f32 itself accumulates with fp32 FFMA (round to nearest): the tensor core's truncating accumulation shrinks
 * the item rows by ~2e-6 relative to the user rows of the same matrix, which config 1's gradients amplify. */
int b200gat_linear_tc_f32(const float* x, const float* W, const float* bias, int64_t n_rows, int in_features,
                          int out_features, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);
int b200gat_linear_bwd_f32(const float* x, const float* dy, int64_t ldy, int64_t n_rows, int in_features,
                           int out_features, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                           void* stream);
/* out[c] = sum_n a[n, c] in a fixed order (GATConv bias gradient). workspace >= 296*channels floats. */
int b200gat_colsum_f32(const float* a, int64_t n_rows, int channels, float* out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- (3) fused edge forward ---------------------------------------------------------------------
 * Replaces train_gat_custom.py:79-92 (logit, LeakyReLU, clamp, exp, scatter_add denominator,
 * normalise, dropout, index_add aggregate) / GATConv's propagate at train_gat_pyg.py:87.
 * Persistent warps, one destination row at a time; local row r is global node `row_offset + r`.
 *   h [*, heads*channels], s [*, 2*heads] : indexed by global node id
 *   sched [n_sched, 4] : (row, beg, end, slot+1 | 0).  Entries with a non-zero 4th field are SEGMENTS of a long
 *     row that the host split (a hub row processed by one warp would be the tail of the launch): the warp
 *     parks its un-normalised softmax state in partial[slot] (heads*(channels+4) floats per slot) and a second
 *     small kernel merges the segments of each long row in order.  long_table [n_long, 4] = (row, first slot,
 *     n segments, degree); n_long may be 0 (then long_table/partial may be NULL).  The backward uses the same
 *     scheme with heads*channels+4 floats per slot.
 *   col/perm : CSR arrays
 *   bias [channels] or NULL; out [n_rows, channels] (head mean + bias)
 *   out_heads [n_rows, heads, channels] or NULL (per-head outputs, needed by the backward if heads>1)
 *   rowstat [n_rows, heads, 2] = (running max m (0 for CUSTOM), 1/(denominator+eps)) or NULL
 *   p_drop/seed : attention dropout (0 = off); the mask is a function of (seed, original edge id,
 *   head) only.
 */
int b200gat_edge_fwd_f32(const float* h, const float* s, const int32_t* sched, int64_t n_sched,
                         const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                         const int32_t* perm, int64_t row_offset, int heads, int channels, int policy,
                         float negative_slope, const float* bias, float* out, float* out_heads, float* rowstat,
                         float p_drop, uint64_t seed, void* stream);
/* same, h stored as bf16 (b200gat_project_bf16); accumulation and outputs stay fp32 */
int b200gat_edge_fwd_bf16(const void* h_bf16, const float* s, const int32_t* sched, int64_t n_sched,
                          const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                          const int32_t* perm, int64_t row_offset, int heads, int channels, int policy,
                          float negative_slope, const float* bias, float* out, float* out_heads, float* rowstat,
                          float p_drop, uint64_t seed, void* stream);

/* ---- (4) fused edge backward --------------------------------------------------------------------
 * Replaces autograd's replay of the lines above (loss.backward(), train_gat_custom.py:361).
 * node_prep : nodestat[r,h] = (s_dst, m, 1/D, t), t = (1/H) dout[r,:].out_heads[r,h,:]
 *             (for heads==1 pass out as out_heads, and bias to subtract it if it was added);
 *             dbias (nullable) = column sums of dout in a fixed order (GATConv bias gradient),
 *             workspace >= 1184*channels floats when dbias != NULL;
 *             dout_bf16 (nullable) = bf16 copy of dout [n_rows, channels] for b200gat_edge_bwd_bf16
 * edge_bwd  : warp per SOURCE row over the CSC schedule: dh[r,:,:] = sum_i alpha_ij dout_i / H (no atomics),
 *             de[q,h] = d loss / d logit for CSC position q, ds_src[r*ld_ds + h] = sum_q de
 * ds_dst    : ds_dst[r*ld_ds + h] = sum over the in-edges of r (CSR order) of de[csr2csc[e], h]
 */
int b200gat_node_prep_f32(const float* dout, const float* out_heads, const float* bias, const float* s,
                          const float* rowstat, int64_t n_rows, int64_t row_offset, int heads, int channels,
                          float* nodestat, float* dbias, void* dout_bf16, void* workspace, size_t workspace_bytes,
                          void* stream);
int b200gat_edge_bwd_f32(const float* h, const float* s, const float* dout, const float* nodestat,
                         const int32_t* sched, int64_t n_sched, const int32_t* long_table, int64_t n_long,
                         float* partial, const int32_t* row, const int32_t* perm_csc, int64_t row_offset, int heads,
                         int channels, int policy, float negative_slope, float* dh, float* de, float* ds_src,
                         int ld_ds, float p_drop, uint64_t seed, void* stream);
/* same, with h and the gathered dout stored as bf16 */
int b200gat_edge_bwd_bf16(const void* h_bf16, const float* s, const void* dout_bf16, const float* nodestat,
                          const int32_t* sched, int64_t n_sched, const int32_t* long_table, int64_t n_long,
                          float* partial, const int32_t* row, const int32_t* perm_csc, int64_t row_offset, int heads,
                          int channels, int policy, float negative_slope, float* dh, float* de, float* ds_src,
                          int ld_ds, float p_drop, uint64_t seed, void* stream);
int b200gat_ds_dst_f32(const float* de, const int32_t* rowptr, const int32_t* csr2csc, int64_t n_rows, int64_t n_edges,
                       int heads, float* ds_dst, int ld_ds, void* stream);

/* ---- (5) fused ranking loss ---------------------------------------------------------------------
 * Replaces train_gat_custom.py:350-359: pos/neg dot products of gathered rows of Z, BPR
 * -log(sigmoid(pos-neg)+1e-8).mean() or BCE-with-logits, and (bwd) the duplicate-row index_put.
 *   z [n_users+n_items, channels]; u, i, j int64 [n_triples] (i, j are item ids, not node ids)
 *   loss : device float[1]; an out-of-range triple makes it NaN.
 *   The forward leaves per-triple coefficients and node-sorted incidence lists in `workspace`; pass the
 *   same workspace to the backward.  grad_out : device float[1].  The backward writes the gradient
 *   rows of nodes [node_begin, node_begin+node_count) -- or of the nodes listed in node_list
 *   [node_count] when it is not NULL -- into dz [node_count, channels] (a row shard; 0 / n_users+n_items
 *   = everything; node_list entries < 0 are padding rows and get zeros); every row written (rows without
 *   triples = 0); no atomics, bitwise reproducible.  dz_bf16 (nullable): a bf16 copy of dz for the bf16 tier.
 *   node_map (nullable, int32 [n_users+n_items]): row of z that holds node v (a row-sharded z is stored
 *   block-permuted); NULL = identity.
 */
int b200gat_loss_workspace_bytes(int64_t n_nodes, int64_t n_triples, size_t* bytes /*host*/);
int b200gat_rank_loss_fwd_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* u,
                              const int64_t* i, const int64_t* j, int64_t n_triples, const int32_t* node_map,
                              int loss_kind, int need_backward, float* loss, void* workspace, size_t workspace_bytes,
                              void* stream);
int b200gat_rank_loss_bwd_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* u,
                              const int64_t* i, const int64_t* j, int64_t n_triples, const int32_t* node_map,
                              int loss_kind, const float* grad_out, const int32_t* node_list, int64_t node_begin,
                              int64_t node_count, float* dz, void* dz_bf16, void* workspace, size_t workspace_bytes,
                              void* stream);
/* Row-sharded forms of the same two lines (:350-359) for the multi-GPU path.  z_blocks: HOST array of n_blocks device
 * pointers, z_blocks[b] = rank b's [n_max, channels] block of Z (its own or peer-mapped memory); node_map (required) gives
 * the row of node v in the gathered row space (block = row / n_max).  The forward evaluates triples
 * [t_begin, t_begin + t_count) only -- their three rows are read out of the owners' memory over NVLink, no all-gather of Z --
 * writes those slices of coef [2 * n_triples] (d/dpos | d/dneg) and loss_partial[0] = this rank's share of the mean; with
 * need_backward it also sorts all 3 * n_triples incidences by node into `workspace`.  The backward needs the complete coef
 * (the caller gathers the ranks' slices) and writes dz [node_count, channels] for the nodes in node_list (this rank's rows). */
int b200gat_rank_loss_fwd_peer_f32(const void* const* z_blocks /*host*/, int n_blocks, int64_t n_max, int64_t n_users,
                                   int64_t n_items, int channels, const int64_t* u, const int64_t* i, const int64_t* j,
                                   int64_t n_triples, int64_t t_begin, int64_t t_count, const int32_t* node_map, int loss_kind,
                                   int need_backward, float* coef, float* loss_partial, void* workspace, size_t workspace_bytes,
                                   void* stream);
int b200gat_rank_loss_bwd_peer_f32(const void* const* z_blocks /*host*/, int n_blocks, int64_t n_max, int64_t n_users,
                                   int64_t n_items, int channels, const int64_t* u, const int64_t* i, const int64_t* j,
                                   int64_t n_triples, const int32_t* node_map, int loss_kind, const float* coef,
                                   const float* grad_out, const int32_t* node_list, int64_t node_count, float* dz, void* dz_bf16,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-head streaming of a heads > 1 layer (BASELINE config 5: 30 M nodes x 4 heads x 256 channels) ----------------------
 * When [N, heads*channels] does not fit, the layer (scripts/train_gat_pyg.py:77,87) runs one head at a time with heads = 1
 * arguments: h_head [N, channels], s_head [N, 2] (b200gat_project_bf16 on that head's rows of W), rowstat_head [n_rows, 2].
 *   edge_fwd_stream : the forward of one head; out = (accumulate ? out : 0) + out_scale * result (+ bias): out_scale = 1/heads,
 *                     accumulate = head > 0, bias with the first head -> `out` ends up as the head mean + bias.
 * The backward does not keep per-head outputs (t_i = sum_k alpha_ik dalpha_ik is recovered from the edges instead):
 *   node_stat        : nodestat[r] = (s_dst, m, 1/D, 0) from s and rowstat
 *   edge_bwd_phase1  : arguments of b200gat_edge_bwd_* with heads = 1; dout must already carry the 1/heads of the head mean
 *                      (b200gat_cast_bf16 with scale); writes dh (aggregation part) and w_ij = alpha_ij * dalpha_ij into `de`
 *   (caller)         : t = per-destination sums of w (b200gat_ds_dst_f32 on `de`; summed over ranks when sharded),
 *                      node_stat_set_t stores them into nodestat[.].w (and the blocks are exchanged again when sharded)
 *   edge_bwd_phase2  : de = slope * (w - alpha * t) in place, ds_src[j] = sum over j's out-edges; colptr/row = CSC arrays,
 *                      rows [row_offset, row_offset + n_rows)
 *   (caller)         : ds_dst = b200gat_ds_dst_f32 on `de`, then b200gat_project_bwd_bf16 / b200gat_project_dx_bf16 per head.
 *   cast_bf16        : dst = bf16(scale * src) (n a multiple of 4). */
int b200gat_edge_fwd_stream_f32(const float* h, const float* s, const int32_t* sched, int64_t n_sched,
                                const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                const int32_t* perm, int64_t row_offset, int channels, int policy, float negative_slope,
                                const float* bias, float* out, float* rowstat, float p_drop, uint64_t seed, float out_scale,
                                int accumulate, void* stream);
int b200gat_edge_fwd_stream_bf16(const void* h_bf16, const float* s, const int32_t* sched, int64_t n_sched,
                                 const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                 const int32_t* perm, int64_t row_offset, int channels, int policy, float negative_slope,
                                 const float* bias, float* out, float* rowstat, float p_drop, uint64_t seed, float out_scale,
                                 int accumulate, void* stream);
int b200gat_node_stat_f32(const float* s, const float* rowstat, int64_t n_rows, int64_t row_offset, int heads, float* nodestat,
                          void* stream);
int b200gat_node_stat_set_t_f32(float* nodestat, const float* t, int64_t n, void* stream);
int b200gat_edge_bwd_phase1_f32(const float* h, const float* s, const float* dout, const float* nodestat, const int32_t* sched,
                                int64_t n_sched, const int32_t* long_table, int64_t n_long, float* partial, const int32_t* row,
                                const int32_t* perm_csc, int64_t row_offset, int channels, int policy, float negative_slope,
                                float* dh, float* de, float* ds_src, int ld_ds, float p_drop, uint64_t seed, void* stream);
int b200gat_edge_bwd_phase1_bf16(const void* h_bf16, const float* s, const void* dout_bf16, const float* nodestat,
                                 const int32_t* sched, int64_t n_sched, const int32_t* long_table, int64_t n_long, float* partial,
                                 const int32_t* row, const int32_t* perm_csc, int64_t row_offset, int channels, int policy,
                                 float negative_slope, float* dh, float* de, float* ds_src, int ld_ds, float p_drop, uint64_t seed,
                                 void* stream);
int b200gat_edge_bwd_phase2_f32(float* de, const int32_t* colptr, const int32_t* row, const float* s, const float* nodestat,
                                int64_t n_rows, int64_t row_offset, int policy, float negative_slope, float* ds_src, int ld_ds,
                                void* stream);
int b200gat_cast_bf16(const float* src, void* dst_bf16, int64_t n, float scale, void* stream);

/* ---- callers either side of the path ---------------------------------
 * adam_step : one step of torch.optim.Adam(lr, weight_decay=l2) exactly as the reference builds it
 *   (scripts/train_gat_custom.py:335,362): g += wd*p; m, v moments; bias-corrected update.  step counts from 1.
 * eval_ranks: inner loop of eval_sampled (scripts/train_gat_custom.py:200-206).  candidates [n_eval, n_candidates]
 *   int64 item ids, column 0 = the held-out positive, the rest the sampled negatives; ranks[q] =
 *   (scores > scores[0]).sum() + 1 with scores = I[candidates[q]] @ U[users[q]].  n_bad: device int32 count of
 *   out-of-range ids (those are scored against row 0).  Recall@k / NDCG@k follow on the host as at :207-209.
 */
int b200gat_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);
/* sample_bpr (row f4): the per-epoch triple sampler, scripts/train_gat_custom.py:213-224 -- uniform user among those
 * with positives, uniform positive, uniform non-positive item by rejection.  The user's positives are the out-edges of
 * the user node in the CSC of b200gat_build_graph (colptr/row; row holds n_users + item).  Counter-based RNG keyed on
 * (seed, sample index): distribution-level parity with the reference's Python `random` stream.  n_fail: device int32,
 * samples for which no user / negative was found within the attempt budget (0 unless the graph is degenerate). */
int b200gat_sample_bpr(const int32_t* colptr, const int32_t* row, int64_t n_users, int64_t n_items, int64_t n_samples,
                       uint64_t seed, int64_t* u, int64_t* i, int64_t* j, int32_t* n_fail, void* stream);
/* _ex: `active_users` (device int32 [n_active], may be NULL) lists the users that have positives, so the user draw needs no
 * rejection on graphs where most user ids are cold (the reference draws from the keys of its user -> positives dict).  A user
 * whose positives cover almost the whole catalogue gets its negative by a forward walk from a random item after 1024
 * rejections; n_fail counts only samples for which no negative exists at all. */
int b200gat_sample_bpr_ex(const int32_t* colptr, const int32_t* row, int64_t n_users, int64_t n_items, int64_t n_samples,
                          uint64_t seed, const int32_t* active_users, int64_t n_active, int64_t* u, int64_t* i, int64_t* j,
                          int32_t* n_fail, void* stream);
int b200gat_eval_ranks_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* users,
                           const int64_t* candidates, int64_t n_eval, int n_candidates, int32_t* ranks, int32_t* n_bad,
                           void* stream);

/* ---- next row f1: Item-Item cosine kNN (graphs/build_ii_knn.py:56-99) ------------------------------------------
 * emb [n_items, dim] fp32, dim = 128 (fused features) or 384 (text embeddings).  Normalisation as the reference
 * (x/(|x|+1e-8), then sklearn's row normalisation), bf16 tcgen05 pass that selects 48 candidate columns per row, exact
 * fp32 re-rank.  Outputs: nbr_idx / nbr_sim
 * [n_items, k] descending (self excluded; unused slots -1 / 0), counts[n_items] = how many pass >= min_similarity
 * (a prefix of the row).  n_unsafe: device int32, rows whose bf16 candidate margin could not prove the selection exact
 * and were therefore recomputed with exact fp32 dots against all columns (informational).
 * The COO triple of the reference (:103-111) is (row = item, col = nbr_idx[item, :counts[item]], sim).
 * Workspace: ~1.2 KB per item (128-d), plus 8 KB per item from 8,192 items on (the appended candidate lists of the three-sweep
 * pipeline, csrc/knn.cu); environment: B200GAT_KNN_MODE=lists forces the register-list kernel at every size (A/B, tests). */
int b200gat_knn_workspace_bytes(int64_t n_items, int dim, size_t* bytes /*host*/);
int b200gat_knn_cosine_f32(const float* emb, int64_t n_items, int dim, int k, float min_similarity, int32_t* nbr_idx,
                           float* nbr_sim, int32_t* counts, int32_t* n_unsafe, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---- multi-GPU exchange over peer memory (the reference itself is single-GPU) -------------
 * The row-sharded path all-gathers a layer's rows once per layer and direction.  Instead of a collective kernel, every
 * rank keeps its block in a buffer from b200gat_peer_alloc, exports it (64-byte CUDA IPC handle, exchanged by the host
 * code), opens its peers' buffers once, and per exchange PULLS each peer's block with b200gat_peer_pull (copy engine over
 * NVLink; one stream per peer).  The host code orders producers and pulls with a stream-ordered barrier. */
#define B200GAT_PEER_HANDLE_BYTES 64
int b200gat_peer_alloc(size_t bytes, void** ptr /*host out*/);
int b200gat_peer_free(void* ptr);
int b200gat_peer_export(const void* ptr, void* handle /*host, >= 64 bytes*/, size_t handle_bytes);
int b200gat_peer_open(const void* handle /*host*/, void** ptr /*host out: the peer's buffer, mapped here*/);
int b200gat_peer_close(void* ptr);
int b200gat_peer_pull(void* dst, const void* src, size_t bytes, void* stream);

/* Device-side exchange ("fabric"): every rank's exported buffer has the same layout -- b200gat_peer_flag_bytes(n_channels)
 * bytes of flags (zeroed once, before the first signal), then regions at offsets the host code chooses (the same on every
 * rank).  bases: HOST array of `world` device pointers, bases[r] = rank r's buffer as mapped on this rank (bases[rank] = the
 * local one).  epoch: the caller's step counter (grows by one per use of a channel; never reset).
 *   signal    : ordered after the producer kernels on `stream`, stores epoch into flags[channel][rank] of every peer.
 *   wait      : a kernel that returns once flags[channel][p] >= epoch for all p (bounded: traps after 20 s).
 *   allgather : wait, then pull, with all SMs, block p of each part from rank p's buffer into the same place of the local
 *               buffer (part k = world blocks of block_bytes[k] at offsets[k] + p * block_bytes[k]; the local block is produced
 *               in place).  1 or 2 parts, 16-byte aligned.
 *   reduce    : wait, then out[i] = sum over p = 0..world-1, in that order on every rank, of
 *               ((float*)(bases[p] + offset))[first + i], i < n  (the all-reduce / reduce-scatter of the small tensors:
 *               ds_dst partial sums, parameter gradients, the loss). */
int b200gat_peer_flag_bytes(int n_channels, size_t* bytes /*host*/);
int b200gat_peer_signal(const void* const* bases /*host*/, int world, int rank, int channel, uint32_t epoch, void* stream);
int b200gat_peer_wait(const void* const* bases /*host*/, int world, int rank, int channel, uint32_t epoch, void* stream);
int b200gat_peer_allgather(const void* const* bases /*host*/, int world, int rank, int channel, uint32_t epoch, int n_parts,
                           const uint64_t* offsets /*host*/, const uint64_t* block_bytes /*host*/, void* stream);
/* Fused projection + exchange (row-sharded path; heads == 1, in_features == channels == 128, tensor-core mode; otherwise
 * B200GAT_ERR_UNSUPPORTED and the caller projects, then pushes): the tcgen05 GEMM's epilogue stores every output tile -- and
 * the logits -- into this rank's block of each peer's mapped exchange buffer as well as locally, so the NVLink transfer runs
 * under the GEMM.  peer_h[q] / peer_s[q] / peer_dx[q]: the address of row 0 of this call's h / s / dx in peer q's buffer.
 * Follow with b200gat_peer_signal on a channel of its own; the consumer waits on that channel (as after b200gat_peer_push). */
int b200gat_project_push_f32(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                             int in_features, int heads, int channels, float* h, float* s, void* const* peer_h,
                             void* const* peer_s, int n_peers, void* workspace, size_t workspace_bytes, void* stream);
int b200gat_project_bwd_push_f32(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh,
                                 const float* ds, int64_t n_rows, int in_features, int heads, int channels, float* dx,
                                 void* const* peer_dx, int n_peers, float* dW, float* da_src, float* da_dst, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* Push variant of the gather: stores this rank's block of every part into the same place of every peer's buffer (reads local
 * memory once, posted NVLink writes).  Follow it with b200gat_peer_signal on a channel of its own and have the consumer
 * b200gat_peer_wait on that channel. */
int b200gat_peer_push(const void* const* bases, int world, int rank, int n_parts, const uint64_t* offsets,
                      const uint64_t* block_bytes, void* stream);
int b200gat_peer_reduce_f32(const void* const* bases /*host*/, int world, int rank, int channel, uint32_t epoch, uint64_t offset,
                            int64_t first, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GAT_H_ */
