/*
 * feta_b200.h -- C ABI of libfeta_b200.so: B200 (sm_100a) kernels for the FeTA
 * spectral hot path.
 *
 * Conventions (all entry points):
 *   - plain C, no C++/torch types; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;
 *   - the caller owns every buffer, including workspaces; the library keeps no
 *     state and allocates nothing;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     hidden synchronisation, so every call is CUDA-graph capturable;
 *   - returns 0 on success, a negative FETA_E* code otherwise; never throws,
 *     never exits; feta_last_error_string() describes the last failure on the
 *     calling thread;
 *   - floating point is fp32; index inputs that come from the reference's
 *     callers are int64 as PyTorch/PyG produce them, library-built index
 *     arrays (CSR, graph_ptr) are int32.
 *
 * Each declaration cites the reference interface it replaces
 * (paths relative to the ansonb/FeTA_TMLR tree).
 */
#ifndef FETA_B200_H_
#define FETA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FETA_OK 0
#define FETA_EINVAL (-1)     /* bad argument (null pointer, unsupported size) */
#define FETA_ECUDA (-2)      /* a CUDA runtime call / launch failed */
#define FETA_EWORKSPACE (-3) /* workspace too small */
#define FETA_EUNSUPPORTED (-4)

/* dtype tags for `batch` (models.py:179-182 hands ChebConvDynamic a FLOAT batch) */
#define FETA_DT_I64 0
#define FETA_DT_I32 1
#define FETA_DT_F32 2
#define FETA_DT_F64 3

/* slots of the int32 `meta` array written by feta_cheb_plan_build */
#define FETA_META_NNZ 0        /* entries of L_hat after self-loop removal */
#define FETA_META_NUM_GRAPHS 1 /* runs of equal values found in `batch` */
#define FETA_META_SORTED 2     /* 1 iff batch is non-decreasing */
#define FETA_META_BLOCKDIAG 3  /* 1 iff every edge stays inside its graph's row range */
#define FETA_META_MAX_NODES 4  /* largest graph (rows) */
#define FETA_META_MAX_DEG 5    /* longest CSR row (either orientation) */
#define FETA_META_BAD_INDEX 6  /* 1 iff an edge endpoint was outside [0, R) (edge dropped) */
#define FETA_META_GUARD 7      /* set to 1 by feta_cheb_fwd/bwd when they refused to run (see below) */
#define FETA_META_WORDS 8

int feta_version(void);
const char* feta_last_error_string(void);
/* number of kernels this library has launched from the calling process (bench.py's
 * gpu_launches claim is read from here) */
int64_t feta_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * A2  ChebConvDynamic.__norm__  (transformer/ChebNetDynamic.py:108-130; PyG-1.7
 *     remove_self_loops -> get_laplacian('sym') -> 2w/lambda_max -> add_self_loops(-1))
 *     + the `torch.unique(batch, return_counts=True)` segmenting of :148.
 *
 * Builds, once per mini-batch, the CSR of L_hat = -(2/lambda_max) D^-1/2 A D^-1/2 grouped by
 * TARGET (rowptr/colidx/vals: row t lists its sources, in input edge order) and the same
 * matrix grouped by SOURCE (the *_t arrays, used by the backward pass).  The +1/-1 self-loop
 * pair of the reference cancels and is not stored.  graph_ptr[g]..graph_ptr[g+1] is the row
 * range of the g-th run of equal `batch` values; row_graph[r] is that run index.
 * Pass batch == NULL for a single graph covering all rows.  A column (-1, -1) of edge_index is
 * padding (fixed-width edge lists of CUDA-graph replays) and is ignored; any other endpoint
 * outside [0, R) drops the edge and sets meta[FETA_META_BAD_INDEX].
 * --------------------------------------------------------------------------------------- */
size_t feta_cheb_plan_workspace_bytes(int64_t num_rows, int64_t num_edges);
int feta_cheb_plan_build(const int64_t* edge_index /* [2, E] */, int64_t num_edges,
                         const void* batch /* [R] or NULL */, int batch_dtype, int64_t num_rows,
                         int64_t num_graphs /* capacity of graph_ptr - 1 */, float lambda_max,
                         int32_t* rowptr /* [R+1] */, int32_t* colidx /* [E] */, float* vals /* [E] */,
                         int32_t* rowptr_t /* [R+1] */, int32_t* colidx_t /* [E] */, float* vals_t /* [E] */,
                         int32_t* graph_ptr /* [G+1] */, int32_t* row_graph /* [R] */,
                         int32_t* meta /* [FETA_META_WORDS] */, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---------------------------------------------------------------------------------------
 * A1/A3  ChebConvDynamic.forward  (transformer/ChebNetDynamic.py:132-189, message :192-193)
 *     out[r] = sum_k T_k[r] . Theta_k[g(r)] + bias,  T_0 = x, T_1 = L x, T_k = 2 L T_{k-1} - T_{k-2}
 * theta is addressed as theta[k*theta_stride_k + g*theta_stride_g + i*fout + j] (elements), so
 * the reference's permuted view of a contiguous [G, K*Fin*Fout] tensor (models.py:357) is
 * consumed without a copy.  One fused launch when `block_diagonal` != 0, Fin == Fout in
 * {4, 8, 16, 32} and the largest graph fits shared memory (`max_nodes`: an upper bound, from
 * meta[FETA_META_MAX_NODES] or the caller's own knowledge); otherwise an un-fused per-order
 * path through `workspace` (feta_cheb_workspace_bytes).
 * `plan_meta` (may be NULL) is the meta array of feta_cheb_plan_build: when given, the fused
 * kernels verify ON THE DEVICE that the plan is sorted, block diagonal, has exactly
 * `num_graphs` graphs, none larger than `max_nodes`, and no bad index; if not they write
 * nothing, set plan_meta[FETA_META_GUARD] = 1 and return -- so a caller that passed host-side
 * hints instead of synchronising on the meta words cannot fault the GPU, and finds out at
 * its next synchronisation point.
 * --------------------------------------------------------------------------------------- */
/* The same builder with the normalisation chosen by `norm_mode`:
 *   FETA_NORM_CHEB_SYM  the scaled Laplacian above (feta_cheb_plan_build == this mode);
 *   FETA_NORM_GCN       ARMAConvDynamic.forward's gcn_norm(add_self_loops=False)
 *                       (transformer/ChebNetDynamic.py:301-305): vals = d_s^-1/2 d_t^-1/2 with the degree
 *                       counted over the TARGET index, input self-loops kept, lambda_max ignored. */
#define FETA_NORM_CHEB_SYM 0
#define FETA_NORM_GCN 1
int feta_graph_plan_build(const int64_t* edge_index /* [2, E] */, int64_t num_edges,
                          const void* batch /* [R] or NULL */, int batch_dtype, int64_t num_rows,
                          int64_t num_graphs, int norm_mode, float lambda_max,
                          int32_t* rowptr, int32_t* colidx, float* vals,
                          int32_t* rowptr_t, int32_t* colidx_t, float* vals_t,
                          int32_t* graph_ptr, int32_t* row_graph, int32_t* meta,
                          void* workspace, size_t workspace_bytes, void* stream);

size_t feta_cheb_workspace_bytes(int64_t num_rows, int fin, int fout, int K);
int feta_cheb_fwd(const float* x /* [R, Fin] */, const int32_t* rowptr, const int32_t* colidx,
                  const float* vals, const int32_t* graph_ptr, const int32_t* row_graph,
                  int32_t* plan_meta /* [FETA_META_WORDS] or NULL */,
                  const float* theta, int64_t theta_stride_k, int64_t theta_stride_g,
                  const float* bias /* [Fout] or NULL */, float* out /* [R, Fout] */,
                  int64_t num_rows, int64_t num_graphs, int K, int fin, int fout, int max_nodes,
                  int block_diagonal, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above.  dx (may be NULL) needs the SOURCE-grouped CSR (*_t); dtheta (may be
 * NULL) recomputes T_k with the TARGET-grouped CSR.  dtheta uses the same strides as theta and
 * is fully overwritten.  dbias [Fout] (may be NULL) is overwritten with sum_r dout[r]. */
int feta_cheb_bwd(const float* dout /* [R, Fout] */, const float* x, const int32_t* rowptr,
                  const int32_t* colidx, const float* vals, const int32_t* rowptr_t,
                  const int32_t* colidx_t, const float* vals_t, const int32_t* graph_ptr,
                  const int32_t* row_graph, int32_t* plan_meta, const float* theta, int64_t theta_stride_k,
                  int64_t theta_stride_g, float* dx, float* dtheta, float* dbias, int64_t num_rows,
                  int64_t num_graphs, int K, int fin, int fout, int max_nodes, int block_diagonal,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * N4  ARMAConvDynamic.forward  (transformer/ChebNetDynamic.py:297-346; constructed with
 *     num_layers = 1 at transformer/models.py:139; _batch_multiply_coeff :274-295)
 *     out[r] = 1/K sum_k relu( a[g,k] (A_hat x)[r] W_k + b[g,k] x_root[r] V_k + bias_k )
 * with coeff[g] = (a[g,0..K-1], b[g,0..K-1]) the per-graph row of the reference's filter_coeff
 * (:313-316), W = init_weight [K,F,F], V = root_weight[0] [K,F,F], bias = bias[0] [K,F] (may be NULL)
 * and A_hat the FETA_NORM_GCN plan (feta_graph_plan_build).  x_root is the dropout-ed copy of x
 * (:335; NULL = x itself).  `prop` (may be NULL) receives A_hat x for the backward pass.
 * One fused launch; requires a block-diagonal plan, F in {4, 8, 16} (the reference's reshape at
 * :284 already forces in_channels == out_channels) and the largest graph to fit one CTA
 * (FETA_EUNSUPPORTED otherwise).  plan_meta: device-side guard as for feta_cheb_fwd.
 * --------------------------------------------------------------------------------------- */
int feta_arma_fwd(const float* x /* [R, F] */, const float* x_root /* [R, F] or NULL */, const int32_t* rowptr,
                  const int32_t* colidx, const float* vals, const int32_t* graph_ptr, const int32_t* row_graph,
                  int32_t* plan_meta, const float* coeff /* [G, 2K] */, const float* init_weight,
                  const float* root_weight, const float* bias, float* out /* [R, F] */,
                  float* prop /* [R, F] or NULL */, int64_t num_rows, int64_t num_graphs, int K, int F,
                  int max_nodes, void* stream);
/* autograd of the above.  dx = A_hat^T d(A_hat x) (+ dx_root when `dx_root` is NULL, i.e. x_root was x);
 * dcoeff [G, 2K] (every graph with at least one row is written; zero-fill for empty graphs is the
 * caller's); dz / dza / dzb [R, K*F] = dz_k, a_k dz_k, b_k dz_k, from which
 *   d init_weight[k] = prop^T dza[:, k],  d root_weight[k] = x_root^T dzb[:, k],  d bias[k] = colsum dz[:, k]
 * (feta_linear_wgrad). */
int feta_arma_bwd(const float* dout /* [R, F] */, const float* prop, const float* x_root,
                  const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                  const int32_t* graph_ptr, const int32_t* row_graph, int32_t* plan_meta, const float* coeff,
                  const float* init_weight, const float* root_weight, const float* bias, float* dx,
                  float* dx_root /* or NULL */, float* dz, float* dza, float* dzb, float* dcoeff,
                  int64_t num_rows, int64_t num_graphs, int K, int F, int max_nodes, void* stream);

/* ---------------------------------------------------------------------------------------
 * A6  kernel-biased attention core of DiffTransformerEncoderLayer (the layer transformer/models.py:4
 *     imports; contract from models.py:166-167):  per (graph b, head h)
 *         S = scale * Q K^T ; masked keys -> -inf ; E = exp(S - rowmax) * pe ;
 *         P = E / max(rowsum(E), 1e-6) ; O = P V
 * q/k/v are addressed as ptr[n*stride_n + b*stride_b + h*dh + c] (the [Nmax, B, 3d] in_proj output
 * is consumed in place).  mask [B, Nmax] bytes, nonzero = padding.  pe [B, Nmax, Nmax] or NULL.
 * Writes attn [B, H, Nmax, Nmax] (rows of padded queries are written as 0), o_heads
 * (= `out_each_head`, models.py:179; addressed o[n*o_stride_n + b*o_stride_b + h*dh + c], so the
 * caller picks [B, Nmax, H, dh] or the seq-first [Nmax, B, H*dh] the out-projection consumes without a
 * copy) and rowflag [B, H, Nmax] for the backward pass: 1 where rowsum > 1e-6; -(1 + argmax_j S_ij) where the clamp
 * was active (the denominator is then a constant but the row maximum still carries a gradient to its argmax --
 * the reference does not detach it; the tcgen05 variant writes 0 there and drops that term); 0 for a padded query.
 * use_tensor_cores != 0: QK^T and PV run as tcgen05.mma kind::tf32 with a 3xTF32 split (fp32-grade
 * accuracy), accumulators and P in TMEM (csrc/attention_tc.cu), when dh in {8,16,32} and
 * Nmax <= 224; otherwise, or with 0, the fp32 CUDA-core kernel (csrc/attention.cu).
 * --------------------------------------------------------------------------------------- */
int feta_attn_fwd(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                  const float* pe, const uint8_t* mask, float* attn, float* o_heads, int64_t o_stride_n,
                  int64_t o_stride_b, float* rowflag, int B, int H, int nmax, int dh, float scale,
                  int use_tensor_cores, void* stream);
/* Backward: d_o_heads (same addressing as o_heads) and optional d_attn [B, H, Nmax, Nmax] in; dq/dk/dv out with
 * their own strides (dq_ptr[n*dstride_n + b*dstride_b + h*dh + c]); every real (n, b) row is
 * written, padded rows are written as 0. */
int feta_attn_bwd(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                  const uint8_t* mask, const float* attn, const float* rowflag, const float* d_o_heads,
                  int64_t o_stride_n, int64_t o_stride_b, const float* d_attn, float* dq, float* dk, float* dv,
                  int64_t dstride_n,
                  int64_t dstride_b, int B, int H, int nmax, int dh, float scale, void* stream);

/* Attention-weight dropout (--dropout > 0; the layer applies it to P before P V, and the returned attention matrix
 * is the dropped one -- GraphiT semantics restated in oracle/layers.py): `drop` [B, H, Nmax, Nmax] holds the
 * multiplier of every weight (0 or 1/(1-p)), generated by the caller.  Forward writes BOTH attn (P, kept for
 * the backward pass) and attn_post (P * drop, what the caller returns); O = (P * drop) V.  Backward takes the
 * gradient of attn_post (or NULL) and applies dP = dP_post * drop, dV = (P * drop)^T dO. */
int feta_attn_fwd_dropout(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                          const float* pe, const uint8_t* mask, const float* drop, float* attn, float* attn_post,
                          float* o_heads, int64_t o_stride_n, int64_t o_stride_b, float* rowflag, int B, int H,
                          int nmax, int dh, float scale, void* stream);
int feta_attn_bwd_dropout(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                          const uint8_t* mask, const float* attn, const float* rowflag, const float* drop,
                          const float* d_o_heads, int64_t o_stride_n, int64_t o_stride_b, const float* d_attn_post,
                          float* dq, float* dk, float* dv, int64_t dstride_n, int64_t dstride_b, int B, int H,
                          int nmax, int dh, float scale, void* stream);

/* A7 / N2, static-shape batches (the collate of transformer/data.py:161-225 padded to dataset-wide capacities, see
 * DESIGN.md section 5): what the padded-domain forward derives from the padding mask [B, Nmax] and the packed edge
 * list [2, Ecap] (int32 or int64 node ids of the reference's packed numbering, (-1, -1) in unused columns):
 *   lens[b], node_end[b] (inclusive prefix), slot_ptr[b] = b*Nmax, seg_lo/seg_hi [H*B] (rows g*Nmax .. g*Nmax+lens of
 *   the coefficient pooling, models.py:283), real [Nmax, B] (1 = real row), and ei_out [2, Ecap (* H)]: every endpoint
 *   re-numbered to its padded slot b*Nmax + i (tile_heads != 0: one copy per head, offset h*B*Nmax -- the per-head
 *   edge tiling the reference omits).  Two launches; replaces ~15 tensor ops of the step's chain. */
int feta_static_context(const uint8_t* mask, const void* edge_index, int edge_dtype, int64_t ecap, int B, int nmax,
                        int H, int tile_heads, int64_t* ei_out, int32_t* node_end, int32_t* lens, int32_t* slot_ptr,
                        int32_t* seg_lo, int32_t* seg_hi, float* real, void* stream);

/* out[c] = sum_r x[r, c], x [R, C] row-major: the `gcn.weight.sum(dim=0)` of the collapsed coefficient path
 * (transformer/models.py:252-282 with an all-ones feature matrix) and the bias gradient of ChebConvDynamic
 * (ChebNetDynamic.py:187).  Deterministic two-stage sum; `partial` holds feta_colsum_partial_floats(C) floats. */
int64_t feta_colsum_partial_floats(int C);
int feta_colsum(const float* x, int64_t R, int C, float* out, float* partial, void* stream);

/* A6 / N1: the same attention core for the layers whose attention matrix nobody reads (transformer/models.py:169-173:
 * under `last_layer_filter` only the LAST layer's matrix feeds get_filter_coefficients; the others use O alone).
 * No [B, H, Nmax, Nmax] tensor is written or read: the forward pass keeps `stats` [B, H, Nmax, 4] = (row maximum of
 * log2(e)*scale*q.k, 1/max(sum, 1e-6), sum > 1e-6, row is real) and the backward pass recomputes P from q, k and
 * `pe` (csrc/attention_rows.cu: one thread per query / key row, K/V/Q/dO of the (graph, head) broadcast from shared
 * memory).  Same argument meaning as feta_attn_fwd / feta_attn_bwd; `o_heads` in the backward pass is the forward
 * output (delta_i = dO_i . O_i).  Needs nmax <= 256, dh in {4, 8, 16, 32}, 16-byte aligned head slices
 * (feta_attn_rows_supported), else FETA_EUNSUPPORTED.  No dropout, no gradient of the attention matrix. */
int feta_attn_rows_supported(int nmax, int dh);
int feta_attn_rows_fwd(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                       const float* pe, const uint8_t* mask, float* o_heads, int64_t o_stride_n, int64_t o_stride_b,
                       float* stats, int B, int H, int nmax, int dh, float scale, void* stream);
/* N1: the per-node scalar of the collapsed filter-coefficient path (see feta_coeff_scalar) WITHOUT the attention matrix:
 * a_ij is recomputed from q, k, `pe` and the `stats` feta_attn_rows_fwd left behind (one thread per key row, two sweeps
 * over the query rows: degrees, then the normalised sums).  s_out [H * N] is indexed h*N + node_ptr[b] + rank, rank =
 * packed index of the real position, like feta_coeff_scalar.  Not differentiable (the reference detaches the matrix,
 * transformer/models.py:282). */
int feta_attn_rows_coeff(const float* q, const float* k, int64_t stride_n, int64_t stride_b, const float* pe,
                         const uint8_t* mask, const float* stats, const int32_t* node_ptr, float* s_out, int B, int H,
                         int nmax, int dh, float scale, int64_t N, void* stream);
int feta_attn_rows_bwd(const float* q, const float* k, const float* v, int64_t stride_n, int64_t stride_b,
                       const float* pe, const uint8_t* mask, const float* stats, const float* o_heads,
                       const float* d_o_heads, int64_t o_stride_n, int64_t o_stride_b, float* dq, float* dk, float* dv,
                       int64_t dstride_n, int64_t dstride_b, int B, int H, int nmax, int dh, float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * A6 (layer glue)  token-axis reductions of the layer's backward (the residual + norm1 / FFN +
 * norm2 part of the layer transformer/models.py:4 imports; torch F.linear / nn.LayerNorm in the
 * reference's upstream).  T = Nmax*B tokens.
 *   feta_linear_wgrad:  dW[out,in] = sum_t dY[t,out] X[t,in],  db[out] = sum_t dY[t,out] (db may be
 *     NULL).  dY [T,out], X [T,in] contiguous, 16-byte aligned, out/in multiples of 4.  `partial`
 *     needs feta_linear_wgrad_slices(T) * (out*in + out) floats.  The token axis is split over CTAs
 *     (each contracts its slice on the tensor cores, 3xTF32: fp32-grade), a second pass folds the slices
 *     in slice order (deterministic).  No shared scratch: calls on different streams may overlap.
 *     `counters` is reserved (may be NULL).
 *   feta_add_layernorm_fwd:  z = a + bscale[row] * b (b, bscale may be NULL), y = LayerNorm(z)*gamma + beta;
 *     saves z, mean, rstd [T] for the backward.  D <= 256.  bscale is the per-node `degree` factor the
 *     layer applies to the attention branch before the residual.
 *   feta_add_layernorm_bwd:  dz (= gradient of a), db_scaled (= bscale * dz, gradient of b; may be
 *     NULL), dgamma, dbeta; `partial` needs feta_add_layernorm_bwd_blocks(T) * 2 * D floats;
 *     `counter`: one int32, ZERO on entry, left zero on exit (the last CTA folds the per-block partials
 *     and re-arms it; reusable by the next call on the same stream, not by concurrent streams).
 *     With dgamma == NULL (dbeta, counter ignored) only dz / db_scaled / `partial` are produced and the
 *     parameter gradients come from a later feta_add_layernorm_bwd_fold(partial, T, D, dgamma, dbeta) --
 *     which a caller can put on another stream, off the critical path of the backward pass.
 * --------------------------------------------------------------------------------------- */
int feta_linear_wgrad_slices(int64_t T);
int feta_linear_wgrad(const float* dY, const float* X, float* dW, float* db, float* partial, size_t partial_floats,
                      int32_t* counters, int64_t T, int out, int in, void* stream);
int feta_add_layernorm_fwd(const float* a, const float* b, const float* bscale, const float* gamma, const float* beta,
                           float* y, float* z, float* mean, float* rstd, int64_t T, int D, float eps, void* stream);
int feta_add_layernorm_bwd_blocks(int64_t T);
int feta_add_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                           const float* bscale, float* dz, float* db_scaled, float* dgamma, float* dbeta,
                           float* partial, int32_t* counter, int64_t T, int D, void* stream);
int feta_add_layernorm_bwd_fold(const float* partial, int64_t T, int D, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------
 * A6 (layer glue)  the layer's projections on the tensor cores (csrc/dense_tc.cu): F.linear of the
 * in/out projections and the FFN of the layer transformer/models.py:4 imports, and its input gradient.
 * m16n8k8 TF32 MMAs with both operands split hi + lo (three MMAs per product): fp32-grade accuracy.
 *   feta_linear_fwd:  Y[T,out] = act(X[T,in] . W[out,in]^T + bias)   (bias may be NULL; relu != 0: ReLU)
 *   feta_linear_dx:   dX[T,in] = (dY[T,out] . W[out,in]) * [mask_src > 0] + dres
 *                     mask_src [T,in] (may be NULL): the ReLU output this layer consumed;
 *                     dres [T,in] (may be NULL): gradient of the residual connection that branches off X.
 * Shapes: in, out multiples of 8, <= 256 (feta_linear_tc_supported); other shapes: use a library GEMM.
 * --------------------------------------------------------------------------------------- */
/* Linear + degree scale + residual + LayerNorm in ONE launch on the tcgen05 path (out = 64 = d_model, in a multiple
 * of 64):  z = res + bscale[row] * (X . W^T + b)  (bscale, b may be NULL);  y = LayerNorm(z) * gamma + beta.
 * z, mean, rstd are what feta_add_layernorm_bwd takes; the Linear's own backward is feta_linear_dx / feta_linear_wgrad. */
int feta_linear_layernorm_supported(int in, int out);
/* The same fused launch on the fp32 CUDA-core kernel (csrc/linear_simt.cu: a CTA owns whole 64-wide output rows, the
 * LayerNorm is two half-warp shuffle folds in the projection's epilogue; exact fp32): out = 64, in a multiple of 64
 * up to 256.  feta_linear_layernorm_fwd = impl FETA_LINEAR_AUTO (this kernel when eligible, else tcgen05). */
int feta_linear_layernorm_simt_supported(int in, int out);
/* Backward of that launch, input-gradient half, in ONE launch (csrc/linear_simt.cu): LayerNorm backward over the
 * 64-wide rows (dz = gradient of `res`; dlin = bscale * dz = gradient of the projection's output, written only when
 * bscale != NULL -- otherwise dlin == dz) as the prologue of dX = (dlin . W) * [mask_src > 0]; `partial` receives
 * feta_lnbwd_linear_dx_blocks(T) * 2 * 64 floats (per-CTA sums of dgamma / dbeta) for feta_ln_fold.  out must be 64, in a
 * multiple of 64.  W is the projection's weight [out = 64, in]. */
int feta_lnbwd_linear_dx_blocks(int64_t T);
int feta_lnbwd_linear_dx(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                         const float* bscale, const float* W, const float* mask_src, float* dz, float* dlin, float* dX,
                         float* partial, int64_t T, int in, int out, void* stream);
/* dgamma[c] = sum_k partial[k, 0, c], dbeta[c] = sum_k partial[k, 1, c] over nblk per-CTA partials (deterministic). */
int feta_ln_fold(const float* partial, int nblk, int D, float* dgamma, float* dbeta, void* stream);
int feta_linear_layernorm_fwd_ex(const float* X, const float* W, const float* bias, const float* res,
                                 const float* bscale, const float* gamma, const float* beta, float* y, float* z,
                                 float* mean, float* rstd, int64_t T, int in, int out, float eps, int impl, void* stream);
int feta_linear_layernorm_fwd(const float* X, const float* W, const float* bias, const float* res, const float* bscale,
                              const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd,
                              int64_t T, int in, int out, float eps, void* stream);
/* The tail of an encoder layer in ONE launch (d_model 64, dim_feedforward 128): three chained tcgen05 GEMMs per
 * 128-token tile, activations stay on the SM between them:
 *   z1 = res + bscale*(o Wo^T + bo); y1 = LN1(z1); h = relu(y1 W1^T + b1); z2 = y1 + h W2^T + b2; y2 = LN2(z2).
 * o, res [T, 64]; bscale [T] or NULL; Wo [64,64], W1 [128,64], W2 [64,128] as nn.Linear stores them; biases may be
 * NULL.  Everything the backward pass needs is written: z1, mean1, rstd1, y1, h, z2, mean2, rstd2 (the backward is
 * feta_add_layernorm_bwd / feta_linear_dx / feta_linear_wgrad per Linear). */
int feta_layer_tail_supported(int d_model, int dff);
int feta_layer_tail_fwd(const float* o, const float* res, const float* bscale, const float* Wo, const float* bo,
                        const float* g1, const float* be1, const float* W1, const float* b1, const float* W2,
                        const float* b2, const float* g2, const float* be2, float* z1, float* mean1, float* rstd1,
                        float* y1, float* h, float* z2, float* mean2, float* rstd2, float* y2, int64_t T, int d_model,
                        int dff, float eps1, float eps2, void* stream);
/* Which kernel family runs a projection.  AUTO: the fp32 CUDA-core latency kernel (csrc/linear_simt.cu: whole
 * reduction dimension staged by one wave of cp.async, exact fp32) when the shape is eligible (in, out multiples of 64,
 * <= 256), else tcgen05 (3xTF32), else legacy mma.sync (3xTF32).  feta_linear_fwd / feta_linear_dx = AUTO. */
#define FETA_LINEAR_AUTO 0
#define FETA_LINEAR_SIMT 1
#define FETA_LINEAR_TC5 2
#define FETA_LINEAR_MMA 3
int feta_linear_simt_supported(int in, int out);
int feta_linear_fwd_ex(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out,
                       int relu, int impl, void* stream);
int feta_linear_dx_ex(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                      int in, int out, int impl, void* stream);
/* 1 when feta_linear_fwd / feta_linear_dx run this (in, out) pair on the tcgen05 path (csrc/linear_tc5.cu:
 * 128 x 64 x 64 tiles, 3xTF32 in TMEM): both multiples of 64. */
int feta_linear_tc5_supported(int in, int out);
int feta_linear_tc_supported(int in, int out);
int feta_linear_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out, int relu,
                    void* stream);
int feta_linear_dx(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                   int in, int out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimizer step of the measured training step: `optim.Adam(model.parameters(), lr=args.lr)`
 * (experiments/run_transformer_gengcn.py:302) / `optim.AdamW(..., weight_decay=)`
 * (experiments/run_transformer_gengcn_SBM_cv.py:371) over ONE flat fp32 buffer each for parameters,
 * gradients and the two moments (n elements, 16-byte aligned):
 *     step += 1;  g' = grad_scale * g;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *     p = p (1 - lr wd) - lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
 * `step` (one float, starts at 0) and `lr` (one float) live on the device: graph-replay safe, and a
 * scheduler changes the rate without re-capture.  weight_decay is decoupled (AdamW); 0 = Adam.
 * grad_scale folds the 1/world_size of the gradient mean.  Two launches (tick + update).
 * --------------------------------------------------------------------------------------- */
int feta_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const float* lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   float* step, void* stream);

/* ---------------------------------------------------------------------------------------
 * A4  DiffTransformerEncoderGenGCN.get_filter_coefficients (transformer/models.py:240-287).
 * With x == 1 the all-pairs GCNConv of :280-282 is  s_j * colsum(W) + b  with a per-node scalar
 *     loop_j = a_jj != 0 ? a_jj : 1;  deg_j = sum_{i != j} a_ij + loop_j;
 *     s_j = deg_j^-1/2 * ( sum_{i != j} deg_i^-1/2 a_ij + deg_j^-1/2 loop_j )
 * (a = attn[b, h] restricted to real nodes; PyG-1.7 gcn_norm / add_remaining_self_loops).
 * feta_coeff_scalar writes s in stacked-row order  s[h*N + node_ptr[b] + j].
 * feta_coeff_pool_fwd:  pooled[g, c] = mean_{j in [seg_lo[g], seg_hi[g])} tanh(s_j * wbar[c] + gbias[c])
 * (:282-283), g = h*B + b; seg_lo/seg_hi are two int32 arrays of G entries (for a packed plan pass
 * graph_ptr and graph_ptr + 1; a padded layout passes g*Nmax and g*Nmax + len_b).
 * feta_coeff_pool_bwd returns d wbar, d gbias (attention is detached, :282).
 * --------------------------------------------------------------------------------------- */
int feta_coeff_scalar(const float* attn /* [B,H,Nmax,Nmax] */, const uint8_t* mask /* [B,Nmax] */,
                      const int32_t* node_ptr /* [B+1] packed offsets of real nodes */,
                      float* s /* [H*N] */, int B, int H, int nmax, int64_t num_nodes, void* stream);
int feta_coeff_pool_fwd(const float* s /* [R] */, const int32_t* seg_lo /* [G] */, const int32_t* seg_hi /* [G] */,
                        const float* wbar /* [C] */, const float* gbias /* [C] */,
                        float* pooled /* [G, C] */, int64_t num_graphs, int C, void* stream);
int feta_coeff_pool_bwd(const float* s, const int32_t* seg_lo, const int32_t* seg_hi, const float* wbar,
                        const float* gbias, const float* d_pooled /* [G, C] */,
                        float* d_wbar /* [C] */, float* d_gbias /* [C] */, float* partial /* [nblk, 2, C] */,
                        int nblk, int64_t num_graphs, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * A5  scatter / pool / pack ops (torch_scatter + advanced indexing in the reference).
 * feature_indices [N, 2] int64 = (graph, node) rows as data.py:218 builds them.
 * --------------------------------------------------------------------------------------- */
/* models.py:177-185 + :347: x[h*N + i, :] = o_heads[fi[i,0], fi[i,1], h, :] */
int feta_pack_heads(const float* o_heads /* [B,Nmax,H,dh] */, const int64_t* feature_indices,
                    float* x /* [H*N, dh] */, int64_t N, int B, int nmax, int H, int dh, void* stream);
/* adjoint of pack_heads: d_o_heads zero-filled then scattered */
int feta_pack_heads_bwd(const float* dx, const int64_t* feature_indices, float* d_o_heads,
                        int64_t N, int B, int nmax, int H, int dh, void* stream);
/* models.py:200-202: out[fi[i,1], fi[i,0], h*dh + c] = y[h*N + i, c], zero elsewhere; out [Nmax,B,H*dh] */
int feta_unpack_heads(const float* y /* [H*N, dh] */, const int64_t* feature_indices,
                      float* out /* [Nmax,B,H*dh] */, int64_t N, int B, int nmax, int H, int dh,
                      void* stream);
int feta_unpack_heads_bwd(const float* d_out, const int64_t* feature_indices, float* dy, int64_t N,
                          int B, int nmax, int H, int dh, void* stream);
/* PyG global_mean_pool over sorted segments (models.py:283):  out[g] = mean_{r in g} x[r] */
int feta_segment_mean_fwd(const float* x /* [R, C] */, const int32_t* graph_ptr, float* out /* [G, C] */,
                          int64_t num_graphs, int C, void* stream);
int feta_segment_mean_bwd(const float* d_out, const int32_t* graph_ptr, float* dx, int64_t num_graphs,
                          int C, void* stream);
/* GlobalAvg1D (models.py:586-595): out[b] = sum_{n real} x[b,n] / count;  x addressed
 * x[b*stride_b + n*stride_n + c] */
int feta_masked_mean_fwd(const float* x, int64_t stride_b, int64_t stride_n, const uint8_t* mask,
                         float* out /* [B, C] */, int B, int nmax, int C, void* stream);
int feta_masked_mean_bwd(const float* d_out /* [B, C] */, const uint8_t* mask, float* dx /* [B,Nmax,C] contiguous */,
                         int B, int nmax, int C, void* stream);
/* cls_output[~masks] (models.py:1070-1071) and the packed<->padded moves of :347 / :201-202 for
 * one feature block:  packed[i, :] = padded[fi[i,0]*stride_b + fi[i,1]*stride_n + :] */
int feta_gather_rows(const float* padded, int64_t stride_b, int64_t stride_n,
                     const int64_t* feature_indices, float* packed /* [N, C] */, int64_t N, int C,
                     void* stream);
int feta_scatter_rows(const float* packed, const int64_t* feature_indices, float* padded,
                      int64_t stride_b, int64_t stride_n, int64_t N, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * A6 layer glue, BatchNorm variant (the reference's ZINC default: experiments/run_transformer_gengcn.py:57,64;
 * the layer flattens [Nmax, B, d] to rows -- padding included -- before nn.BatchNorm1d).
 *   z = a + bscale[row] * b;  y = (z - mean_c) * rstd_c * gamma_c + beta_c,  batch statistics (biased variance)
 *   over the rows with roww[row] != 0 (NULL: all rows);  running_mean / running_var (unbiased) / num_batches
 *   updated like torch.nn.BatchNorm1d in training mode (NULL: skipped).  D must divide 256.
 *   partial: feta_add_batchnorm_blocks(T) * 3 * D floats of scratch.  Backward returns dz (= d a), dbs (= d b,
 *   NULL when b was absent), dgamma, dbeta.
 * --------------------------------------------------------------------------------------- */
int feta_add_batchnorm_blocks(int64_t T);
int feta_add_batchnorm_fwd(const float* a, const float* b, const float* bscale, const float* roww,
                           const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd,
                           float* running_mean, float* running_var, int64_t* num_batches, float* partial,
                           float momentum, float eps, int64_t T, int D, void* stream);
int feta_add_batchnorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                           const float* gamma, const float* bscale, const float* roww, float* dz, float* dbs,
                           float* dgamma, float* dbeta, float* partial, int64_t T, int D, void* stream);

/* ---------------------------------------------------------------------------------------
 * A7  collate index builders (transformer/data.py:161-225, :394-460): GPU batch builder over a
 * dataset pre-packed on the device.  Given the ids of the B graphs of a mini-batch and the
 * packed dataset (node_ptr/edge_ptr prefix sums, edge_index local to each graph), writes the
 * reference's integer outputs bit-exactly: mask [B,Nmax] (1 = pad), edge_indices [2,E]
 * (= local + node offset), batch_indices [N], feature_indices [N,2].
 * --------------------------------------------------------------------------------------- */
int feta_collate_indices(const int64_t* graph_ids /* [B] */, const int64_t* ds_node_ptr,
                         const int64_t* ds_edge_ptr, const int64_t* ds_edge_index /* [2, E_ds] */,
                         int64_t ds_num_edges, const int64_t* out_node_ptr /* [B+1] */,
                         const int64_t* out_edge_ptr /* [B+1] */, uint8_t* mask, int64_t* edge_indices,
                         int64_t* batch_indices, int64_t* feature_indices, int B, int nmax,
                         int64_t N, int64_t E, void* stream);
/* static-shape variant of the edge list (engine.GraphedTrainStep): edge_indices is [2, e_cap], columns >= E
 * are written as (-1, -1), which feta_graph_plan_build ignores; everything else as feta_collate_indices with
 * nmax = the static node capacity. */
int feta_collate_edges_static(const int64_t* graph_ids, const int64_t* ds_edge_ptr,
                              const int64_t* ds_edge_index /* [2, E_ds] */, int64_t ds_num_edges,
                              const int64_t* out_node_ptr, const int64_t* out_edge_ptr,
                              int64_t* edge_indices /* [2, e_cap] */, int B, int64_t E, int64_t e_cap,
                              void* stream);
/* padded feature / PE / degree fill for the same batch:
 *   dst[b, n, :] = src[ds_node_ptr[graph_ids[b]] + n, :]  (n < len_b), 0 elsewhere;
 *   pe_dst[b, i, j] = pe_src[ds_pe_ptr[gid] + i*len + j]   (i, j < len_b), 0 elsewhere. */
int feta_collate_pad_rows(const int64_t* graph_ids, const int64_t* ds_node_ptr, const float* src,
                          float* dst, int B, int nmax, int C, void* stream);
int feta_collate_pad_pe(const int64_t* graph_ids, const int64_t* ds_node_ptr, const int64_t* ds_pe_ptr,
                        const float* pe_src, float* pe_dst, int B, int nmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FETA_B200_H_ */
