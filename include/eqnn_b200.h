/*
 * eqnn_b200 -- C ABI of the B200-native (sm_100a) kernels behind eqnn-jax's
 * equivariant message-passing hot path.
 *
 * The reference (smsharma/eqnn-jax) is pure Python on jax/e3nn-jax/jraph and
 * exposes no FFI of its own; each entry point below therefore cites the Python
 * call site(s) whose arithmetic it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the jax.ffi binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - all arrays are dense fp32 / int32, row-major unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library
 *     never allocates or frees device memory and keeps no pointer after return;
 *   - kernels are enqueued asynchronously on `stream`; no call synchronises the
 *     device, so every call is CUDA-graph capturable;
 *   - return value 0 = ok, non-zero = EQNN_ERR_*; eqnn_last_error_string()
 *     gives the thread-local message.  No exception crosses this boundary.
 *   - a batch of B graphs with N nodes each is addressed with GLOBAL node ids
 *     n = b*N + i and GLOBAL edge ids e = b*E + j.
 */
#ifndef EQNN_B200_H_
#define EQNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* eqnn_stream_t; /* cudaStream_t */

enum {
  EQNN_OK = 0,
  EQNN_ERR_ARG = 1,        /* invalid argument */
  EQNN_ERR_WORKSPACE = 2,  /* workspace too small */
  EQNN_ERR_SIGNATURE = 3,  /* tensor-product signature not compiled in */
  EQNN_ERR_CUDA = 4        /* CUDA runtime error at launch */
};

const char* eqnn_last_error_string(void);
int eqnn_version(void);
long long eqnn_launch_count(void); /* kernels enqueued by this library since load (process-wide) */

/* ---- tensor-product signatures compiled into the library (csrc/gen/tp_gen.cuh) ---- */
int eqnn_tp_signature_count(void);
const char* eqnn_tp_signature(int id);       /* host string */
int eqnn_tp_lookup(const char* signature);   /* -> id or -1 */
int eqnn_tp_dims(int id, int* dxp, int* ny, int* n_live, int* d_live);

/* ---- graph construction -------------------------------------------------------------
 * eqnn_knn_pbc replaces models/utils/graph_utils.py:40-72 `nearest_neighbors` vmapped
 * over the batch (:280-282) with `apply_pbc` of :13-37 fused in.  Bit-exact contract
 * (SURVEY.md C.3): every fp32 operation individually rounded (no FMA contraction)
 *     d_c  = (x_i[c] - x_j[c]) + 1e-7f
 *     d_c -= rint(d . cell_inv)[.] . cell          (only if EQNN_KNN_PBC)
 *     D_ij = (d_0^2 + d_1^2) + d_2^2
 * and the k smallest D_ij per row in stable order (ties -> smaller j); j == i included.
 *   x          [B, N, ld_x] fp32, the first three columns are the coordinates
 *   cell, cell_inv   HOST pointers to 9 floats (row-major 3x3); may be NULL without PBC
 *   senders    [B, N*k] int32  = repeat(arange(N), k)          (graph_utils.py:70)
 *   receivers  [B, N*k] int32  = neighbour indices, local ids  (graph_utils.py:71)
 *   dr         [B, N*k, 3] fp32 = wrapped displacement of each selected pair (:72); may be NULL
 * flags: EQNN_KNN_PBC; EQNN_KNN_FMA contracts D_ij as fma(d2,d2,fma(d1,d1,d0*d0))
 *        (what an FMA-fusing CPU backend would produce; off by default);
 *        EQNN_KNN_DENSE forces the dense N x N sweep.
 * ws / ws_bytes: optional workspace of eqnn_knn_workspace_bytes(B, N, k) bytes.  With it, periodic diagonal cells
 * and N >= 2048, the search runs on a cell list (points binned into G^3 cells, (2R+1)^3 cells swept per query, R
 * grown until the k-th distance lies inside the searched region): same distance expression, list ordered by
 * (distance, index) -- output bit-identical to the dense sweep.  Without it (NULL / too small) the dense sweep runs. */
enum { EQNN_ACT_GELU = 0, EQNN_ACT_SILU = 1 };   /* models/nequip.py:61 getattr(nn, activation); gelu = tanh form */
enum { EQNN_KNN_PBC = 1, EQNN_KNN_FMA = 2, EQNN_KNN_DENSE = 4 };
size_t eqnn_knn_workspace_bytes(int B, int N, int k);
int eqnn_knn_pbc(const float* x, int ld_x, const float* cell, const float* cell_inv, int B, int N, int k,
                 int flags, int32_t* senders, int32_t* receivers, float* dr, void* ws, size_t ws_bytes,
                 eqnn_stream_t stream);

/* Radius graph: the `radius is not None` branch of build_graph (models/utils/graph_utils.py:266-277) as it evidently
 * means to work (upstream it cannot run: `x` is undefined at :277 and np.where of the batched mask returns three arrays).
 * Edges i -> j of every graph with 0 < |apply_pbc(x_i - x_j)| < radius (compute_distances, :209-226: no EPS), in np.where
 * order (i, then j ascending).  Two passes over the same kernel:
 *   rowptr == NULL: count[b * N + i] = number of edges leaving i;
 *   rowptr given (eqnn_exclusive_scan_i32 of count, B * N + 1 entries): senders / receivers (local ids) and dist are
 *   written at rowptr[b * N + i] ...  -- a ragged edge list, graph b owning [rowptr[b N], rowptr[(b + 1) N]).     */
int eqnn_radius_graph(const float* x, int ld_x, const float* cell, const float* cell_inv, int B, int N, float radius,
                      int flags, int32_t* count, const int32_t* rowptr, int32_t* senders, int32_t* receivers, float* dist,
                      eqnn_stream_t stream);
int eqnn_exclusive_scan_i32(const int32_t* count, int64_t n, int32_t* out, eqnn_stream_t stream);

/* GraphsTuple.edges of build_graph (models/utils/graph_utils.py:298-302):
 *   n_rb == 0: out [n_edges, 1] = |dr|;  n_rb > 0: out [n_edges, n_rb] = e3nn.bessel(|dr|, n_rb, r_max) */
int eqnn_edge_features(const float* dr, int64_t n_edges, int n_rb, double r_max, float* out, eqnn_stream_t stream);

/* Receiver-sorted CSR of the batch (replaces the unsorted scatter of
 * jraph segment_sum/segment_mean at models/nequip.py:118-120,154-159).
 *   senders/receivers [B, E] local ids  ->
 *   rowptr [B*N + 1]  global CSR offsets per global receiver id
 *   perm   [B*E]      global edge id stored at each CSR position (ascending inside a row,
 *                     i.e. the sequential order of a CPU scatter-add)
 *   src    [B*E]      global sender id at each CSR position                          */
size_t eqnn_csr_workspace_bytes(int B, int N, int E);
int eqnn_csr_build(const int32_t* senders, const int32_t* receivers, int B, int N, int E, int32_t* rowptr,
                   int32_t* perm, int32_t* src, void* ws, size_t ws_bytes, eqnn_stream_t stream);

/* Per-edge geometry in CSR order (models/nequip.py:129-134 r_ij = x_s - x_r, no PBC, no EPS;
 * :55 d = |r_ij|; :57 e3nn.bessel):
 *   geom   [B*E] float4 = (r_hat.x, r_hat.y, r_hat.z, d); r_hat = 0 where d == 0
 *   radial [B*E, n_rb]  = sqrt(2/r_cut) sin(j pi d / r_cut)/d, j=1..n_rb (limit at d=0);
 *                         if n_rb == 0: radial [B*E, 1] = d (nequip.py:59)
 * Every operation is rounded to fp32 where the reference's XLA:CPU graph rounds (the Bessel
 * phase j*pi/r_cut*d is ill-conditioned; reproducing its rounding points is what keeps parity). */
int eqnn_edge_geom(const float* pos, int ld_pos, int n_nodes, const int32_t* rowptr, const int32_t* src,
                   int n_rb, double r_cut, float* geom, float* radial, eqnn_stream_t stream);

/* ---- NequIP interaction block -------------------------------------------------------
 * eqnn_tpconv_fwd: for every receiver r
 *     A[r, :] = agg_{e -> r}  w_t[:, e] (x)  CG( xt[src(e)],  Y_lm(r_hat_e) )
 * = models/nequip.py:46-52 (spherical harmonics, tensor_product), :68 (per-irrep
 * radial weights) and the jraph aggregation :118-120, restricted to live columns.
 *   tp_id  signature id (x irreps, lmax, live set)
 *   xt     [n_nodes, DXP] regrouped + padded sender record (Linear of :43 applied)
 *   w_t    [NLIVE, n_edges] column-major radial weights in CSR order
 *   agg    0 = sum, 1 = mean (divide by max(in-degree, 1)), 2 = max (jraph.segment_max)
 *   A      [n_nodes, ld_a], first DLIVE columns written
 *   argmax [n_nodes, DLIVE] int32, agg == 2 only (else NULL): CSR position of the first edge attaining the
 *          maximum of each component; the backward pass routes dL/dA to that edge alone
 * Harmonics are unit-norm ("norm"); the caller folds the e3nn normalisation factor
 * of each l into the weights that produce w_t.                                       */
int eqnn_tpconv_fwd(int tp_id, const int32_t* rowptr, const int32_t* src, const float* geom, const float* w_t,
                    int64_t n_edges, const float* xt, int n_nodes, int agg, float* A, int ld_a, int32_t* argmax,
                    eqnn_stream_t stream);
/* Backward of the above w.r.t. w_t and xt (the custom_vjp rule):
 *   gA    [n_nodes, ld_ga]  dL/dA
 *   dw_t  [NLIVE, n_edges]  dL/dw_t
 *   dxe   [n_edges, DXP]    per-edge dL/dxt contribution, stored at the ORIGINAL edge id
 *                           (perm) so that senders' k edges are contiguous            */
int eqnn_tpconv_bwd(int tp_id, const int32_t* rowptr, const int32_t* src, const int32_t* perm, const float* geom,
                    const float* w_t, int64_t n_edges, const float* xt, int n_nodes, int agg, const float* gA,
                    int ld_ga, float* dw_t, float* dxe, const int32_t* argmax, eqnn_stream_t stream);

/* Per-edge messages of one interaction step, un-reduced (the `edges` field of the node task's output,
 * models/nequip.py:154-171: jraph.GraphNetwork stores the update_edge_fn result):
 *     out[perm[p], :] = (w_rows[p, col[c]] * scale[c])_c (x) CG( xt[src(p)], Y_lm(r_hat_p) )
 *   tp_id   a signature whose live set is every message column (NLIVE <= 32)
 *   w_rows  [n_edges, ld_w] radial-MLP outputs per CSR position, row-major; col / scale: HOST arrays of NLIVE entries
 *   out     [n_edges, ld_out] in the reference's (sender-major) edge order                                       */
int eqnn_tp_messages(int tp_id, const int32_t* src, const int32_t* perm, const float* geom, const float* w_rows, int ld_w,
                     const int* col, const float* scale, const float* xt, int64_t n_edges, float* out, int ld_out,
                     eqnn_stream_t stream);
/* dxt[s, :] = sum_{j<k} dxe[s*k + j, :]   (sender-major fixed out-degree k, graph_utils.py:70) */
int eqnn_sender_reduce(const float* dxe, int n_nodes, int k, int dxp, float* dxt, eqnn_stream_t stream);
/* general out-degree: dxt[s, :] = sum over row s of a SENDER-sorted CSR (eqnn_csr_build with the roles of
 * senders and receivers swapped) of dxe[perm_s[q], :]; ascending edge id inside a row => deterministic */
int eqnn_segment_gather_sum(const float* dxe, const int32_t* rowptr_s, const int32_t* perm_s, int n_nodes, int dxp,
                            float* dxt, eqnn_stream_t stream);

/* node_attrs = scatter_sum(Y(r_ij), receivers) * scale   (models/nequip.py:180-191), unit-norm Y */
int eqnn_sh_scatter(int lmax, const int32_t* rowptr, const float* geom, int n_nodes, float scale, float* attrs,
                    eqnn_stream_t stream);
/* T = tensor_product(x, y) restricted to live columns, per node (models/nequip.py:197) */
int eqnn_node_tp_fwd(int tp_id, const float* x, const float* y, int64_t n, float* T, eqnn_stream_t stream);
int eqnn_node_tp_bwd(int tp_id, const float* x, const float* y, const float* gT, int64_t n, float* dx,
                     eqnn_stream_t stream);

/* ---- dense building block ------------------------------------------------------------
 * C[m,n] (+)= epi( alpha * sum_k pro(A[m,k]) * B[k,n] ), arbitrary element strides.
 * Used for the radial MLP (e3nn.flax.MultiLayerPerceptron, nequip.py:62-66), the
 * irreps Linear layers lowered to dense matrices (:43,85,162,197) and models/mlp.py.
 *   pro: 0 none | 1 A := gelu_tanh(A) * pro_scale
 *   epi: 0 none | 1 += bias[n] (aux = bias) | 2 *= gelu_tanh'(aux[m*ld_aux+n]) * epi_scale
 *        | 3 C = (alpha*acc + C) * gelu_tanh'(aux[m*ld_aux+n]) * epi_scale
 *   flags: EQNN_GEMM_ACCUMULATE adds to C.  Large-K problems are split over CTAs
 *   through `ws` (eqnn_gemm_workspace_bytes) and reduced in a fixed order.          */
/*   EQNN_GEMM_TF32X3: large-M problems with K, N <= 128 run on the tensor cores (tcgen05, TMEM) with every
 *   operand split into two TF32 halves and three MMAs per product (fp32-class accuracy, error ~2^-21);
 *   shapes the tensor-core kernels do not cover silently use the fp32 CUDA-core kernel.           */
/*   EQNN_GEMM_TF32X3_ANY (with EQNN_GEMM_TF32X3): every other shape with pro = epi = 0 and M, N >= 32 also runs
 *   on the tensor cores, through a general 128 x 128/256-tile 3xTF32 kernel (any strides and alignment, K split
 *   over CTAs for short-and-wide outputs) -- the SEGNN dense-lowered layers (models/segnn.py:30-64).            */
enum { EQNN_GEMM_ACCUMULATE = 1, EQNN_GEMM_TF32X3 = 2, EQNN_GEMM_TF32X3_ANY = 4 };
size_t eqnn_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int eqnn_gemm_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                  const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int flags,
                  int pro, float pro_scale, int epi, const float* aux, int64_t ld_aux, float epi_scale, void* ws,
                  size_t ws_bytes, eqnn_stream_t stream);

/* e3nn.gate with default activations (models/nequip.py:88-90):
 *   z   [n, n_extra + n_gates + sum_c mul_c (2 l_c + 1)] = [scalars | gates | gated chunks]
 *   out [n, n_extra + gated dims] = [gelu(z_extra)/c_act | sigmoid(gate)/c_gate * gated]
 *   mul, l: HOST arrays of the n_chunks gated chunks.                                */
int eqnn_gate_fwd(const float* z, int64_t n, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                  float inv_c_act, float inv_c_gate, float* out, eqnn_stream_t stream);
int eqnn_gate_bwd(const float* z, const float* gout, int64_t n, int n_extra, int n_gates, int n_chunks,
                  const int* mul, const int* l, float inv_c_act, float inv_c_gate, float* gz,
                  eqnn_stream_t stream);
/* the same with the scalar activation chosen by the caller (models/segnn.py:27,56 scalar_activation: EQNN_ACT_GELU or
 * EQNN_ACT_SILU, inv_c_act = 1 / e3nn.normalize_function constant of that activation); gates stay sigmoid */
int eqnn_gate_act_fwd(const float* z, int64_t n, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                      int activation, float inv_c_act, float inv_c_gate, float* out, eqnn_stream_t stream);
int eqnn_gate_act_bwd(const float* z, const float* gout, int64_t n, int n_extra, int n_gates, int n_chunks,
                      const int* mul, const int* l, int activation, float inv_c_act, float inv_c_gate, float* gz,
                      eqnn_stream_t stream);

/* Fused node update of one interaction step (models/nequip.py:80-90 and :162):
 *     z = alpha1 * A @ W1 + h @ W2 ;  h_out = gate(z) @ W3 ;  xt_out = h_out @ W0n (optional, next step's :43)
 *   A [n, ld_a] (first d_a columns), h [n, d_h]; W1 [d_a, d_z], W2 [d_h, d_z] or NULL, W3 [d_g, d_h],
 *   W0n [d_h, dxp] or NULL: dense lowerings of the irreps Linear layers;  gate layout and constants as in
 *   eqnn_gate_fwd;  z [n, d_z] is saved for the backward pass, the gate output never leaves the SM.
 * Backward: dA [n, d_a], dh_in [n, d_h] (self-connection part), dW1/dW2/dW3 (dense, overwritten; dW2 may be
 * NULL iff W2 is); weight gradients are reduced per thread -> per CTA -> fixed-order sum (deterministic).     */
size_t eqnn_node_update_workspace_bytes(int d_a, int d_h, int d_z, int d_g);
int eqnn_node_update_fwd(const float* A, int ld_a, int d_a, const float* h, int d_h, const float* W1, float alpha1,
                         const float* W2, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                         float inv_c_act, float inv_c_gate, const float* W3, const float* W0n, int dxp, int64_t n,
                         float* z, float* h_out, float* xt_out, eqnn_stream_t stream);
int eqnn_node_update_bwd(const float* A, int ld_a, int d_a, const float* h, int d_h, const float* W1, float alpha1,
                         const float* W2, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                         float inv_c_act, float inv_c_gate, const float* W3, const float* z, const float* dh_out,
                         int64_t n, float* dA, float* dh_in, float* dW1, float* dW2, float* dW3, void* ws,
                         size_t ws_bytes, eqnn_stream_t stream);

/* out[b, :] = reduce_n x[b, n, :]   (mode 0 sum, 1 mean; nequip.py:196-199) and its backward */
int eqnn_pool_fwd(const float* x, int B, int N, int D, int mode, float* out, eqnn_stream_t stream);
int eqnn_pool_bwd(const float* gout, int B, int N, int D, int mode, float* gx, eqnn_stream_t stream);
/* readout_agg = "max" (jnp.max over the nodes, models/nequip.py:195-199): out [B, D], argmax [B, D] = first node
 * attaining the maximum; the backward pass routes gout[b, c] to gx[b, argmax[b, c], c] and writes zeros elsewhere */
int eqnn_pool_max_fwd(const float* x, int B, int N, int D, float* out, int32_t* argmax, eqnn_stream_t stream);
int eqnn_pool_max_bwd(const float* gout, const int32_t* argmax, int B, int N, int D, float* gx, eqnn_stream_t stream);
/* LayerNorm over x[b, :] = concat(a[b, :da], g[b, :dg]) (pooled node scalars | graph globals), nequip.py:201-205:
 * y = (x - mean) * rsqrt(var + eps) * scale + bias, flax.linen.LayerNorm defaults (eps = 1e-6).  stats[b] = (mean,
 * rstd) is kept for the backward, which returns d_a (B x da; the globals are data), dscale and dbias (da + dg). */
int eqnn_layernorm_fwd(const float* a, int da, const float* g, int dg, int B, const float* scale, const float* bias,
                       float eps, float* y, float* stats, eqnn_stream_t stream);
int eqnn_layernorm_bwd(const float* a, int da, const float* g, int dg, int B, const float* scale, const float* stats,
                       const float* dy, float* d_a, float* dscale, float* dbias, eqnn_stream_t stream);
/* y[n] = sum_m x[m, n]  (bias gradients) */
int eqnn_colsum(const float* x, int64_t M, int N, float* y, eqnn_stream_t stream);

/* ---- steerable graph attributes (models/utils/equivariant_graph_utils.py:27-102, `get_equivariant_graph`) ----
 * Edges in the reference's order (senders / receivers: [B, E] local ids; rowptr / perm: receiver-sorted CSR of
 * eqnn_csr_build).  r_ij[e] = x[sender] - x[receiver], minimum image when cell != NULL, no EPS (:46-52);
 * edge_attrs[e, (lmax+1)^2] = normalisation * Y_lm(r_ij / |r_ij|) (:56-61; sh_norm 0 "norm", 1 "component",
 * 2 "integral"); node_attrs[n] = e3nn.scatter_mean of edge_attrs over the receivers (:64-74) + Y_lm(vel[n]) when
 * vel != NULL (steerable_velocities, :76-83).  additional_messages (:85-90) = eqnn_edge_features(r_ij).
 * cell / cell_inv: HOST pointers to 9 floats or NULL. */
int eqnn_steerable_attrs(const float* pos, int ld_pos, const float* vel, int ld_vel, int B, int N, int E,
                         const int32_t* senders, const int32_t* receivers, const int32_t* rowptr, const int32_t* perm,
                         const float* cell, const float* cell_inv, int lmax, int sh_norm, float* r_ij,
                         float* edge_attrs, float* node_attrs, eqnn_stream_t stream);

/* ---- SEGNN TensorProductLinearGate (models/segnn.py:30-64), dense-lowered ---------------------------------
 * Linear(TP(x, y)) = sum_j y_j (x @ M_j) + b: the host lowers the e3nn Linear weights and the Clebsch-Gordan
 * coefficients into M_cat (Dx x Do*Dy, column o*Dy + j = output o, attribute component j; eqnn_weights_expand),
 * eqnn_gemm_f32 forms Xbig = x @ M_cat, and these kernels contract / expand the attribute axis:
 *   z[r, o] = bias[o - bias_col] (for bias_col <= o < bias_col + n_bias: the 0e block) + sum_j y[r, j] Xbig[r, o*Dy + j]
 *                                                                          (y NULL: Dy = 1, y = 1; segnn.py:35-36)
 *   dXbig[r, o*Dy + j] = y[r, j] dz[r, o]
 * The gate (segnn.py:55-63) is eqnn_gate_fwd/bwd. */
/* Forward of the dense-lowered layer with the attribute contraction fused into the tensor-core GEMM's epilogue
 * (tc_gemm.cu): z[r, o] = bias + sum_j y[r, j] (x @ M_cat)[r, o*Dy + j]; the (R x Do*Dy) product is reduced in
 * registers straight out of tensor memory and never written.  Dy in {1, 2, 4, 8} (attributes l <= 1: Dy = 4), y != NULL,
 * and a shape the kernel covers (R >= 32, Do*Dy >= 32, R*Do*Dy*Dx >= 2^24); anything else is an error and the caller
 * keeps eqnn_gemm_f32 + eqnn_segnn_ycontract. */
int eqnn_segnn_tplg_fwd(const float* x, int ld_x, int Dx, const float* Mcat, const float* y, int ld_y, int Dy, int Do,
                        const float* bias, int bias_col, int n_bias, int64_t R, float* z, int ld_z, eqnn_stream_t stream);
int eqnn_segnn_ycontract(const float* Xbig, const float* y, int ld_y, const float* bias, int bias_col, int n_bias,
                         int64_t R, int Dy, int Do, float* z, eqnn_stream_t stream);
int eqnn_segnn_yexpand(const float* dz, const float* y, int ld_y, int64_t R, int Dy, int Do, float* dXbig,
                       eqnn_stream_t stream);
/* Irrep-blocked form of the same layer (the default of eqnn_jax_b200.segnn): the multiplicity contraction runs
 * first, T[r, (path, w, m1)] = sum_u x[r, (u, m1)] W_path[u, w] (eqnn_gemm_f32 per stored irrep block of x), then
 * the Clebsch-Gordan terms are applied from a CSR table (entries: idx = source column | attribute component << 24,
 * coef = sqrt(2 l3 + 1) C[m1, m2, m3]):
 *   out[r, o] = bias[o - bias_col] (0e block) + sum_{t in ptr[o] .. ptr[o+1]} coef[t] y[r, idx[t] >> 24] in[r, idx[t] & 0xFFFFFF]
 * forward: in = T, rows of the table = Linear outputs; backward: in = dz, rows = T columns (bias NULL).
 * y NULL: Dy = 1, y = 1.  Dy <= 32, n_in < 2^24. */
int eqnn_segnn_cg_apply(const float* in, int ld_in, int n_in, const float* y, int ld_y, int Dy, const int32_t* ptr,
                        const int32_t* idx, const float* coef, const float* bias, int bias_col, int n_bias, int64_t R,
                        int n_out, float* out, int ld_out, eqnn_stream_t stream);
/* First edge layer of a message-passing step, evaluated at the nodes (segnn.py:111-113: its input is
 * concat(add[e], h[sender], h[receiver]), so x @ M_cat = add @ M_a + (h @ M_s)[sender] + (h @ M_r)[receiver]).
 * The caller forms P_s = h @ M_s and P_r = h @ M_r with eqnn_gemm_f32 on the n_nodes rows (1/k of the per-edge
 * work); edges are in receiver-sorted order (rowptr / src of eqnn_csr_build):
 *   fwd : z[p, o] = bias + sum_j y[p, j] (P_r[row(p), o*Dy+j] + P_s[src[p], o*Dy+j] + sum_a add[p, a] M_a[a, o*Dy+j])
 *   bwd : with G[p, o*Dy+j] = y[p, j] dz[p, o]:  dP_r[n, :] = sum of G over the edges received by n,
 *         dP_s[n, :] = sum of G over the edges sent by n (rowptr_s: sender-sorted CSR; spos[q] = receiver-sorted
 *         position of the q-th sender-sorted edge, from eqnn_perm_compose(perm, perm_s))
 *   wadd: gM_a[a, o*Dy+j] = sum_p add[p, a] G[p, o*Dy+j]   (1 <= n_add <= 8; two-stage, fixed order)
 * The remaining gradients (h, M_s, M_r) are node-level GEMMs on dP_s / dP_r. */
int eqnn_segnn_edge0_fwd(const float* Ps, const float* Pr, const float* add, int n_add, const float* Ma, const float* y,
                         int Dy, int Do, const float* bias, int bias_col, int n_bias, const int32_t* rowptr,
                         const int32_t* src, int n_nodes, float* z, eqnn_stream_t stream);
int eqnn_segnn_edge0_bwd(const float* dz, const float* y, int Dy, int Do, const int32_t* rowptr, const int32_t* rowptr_s,
                         const int32_t* spos, int n_nodes, float* dPs, float* dPr, eqnn_stream_t stream);
size_t eqnn_segnn_edge0_wadd_workspace_bytes(int n_add, int Dy, int Do);
int eqnn_segnn_edge0_wadd(const float* add, int n_add, const float* y, int Dy, const float* dz, int Do, int64_t E,
                          float* gMa, void* ws, size_t ws_bytes, eqnn_stream_t stream);
int eqnn_perm_compose(const int32_t* perm, const int32_t* perm_s, int64_t n, int32_t* inv, int32_t* spos,
                      eqnn_stream_t stream);
/* Last edge layer of a message-passing step, evaluated at the nodes: it has no gate (segnn.py:66-77) and its output
 * is only aggregated over the receivers (:218), so with f = 1 (sum) or 1/deg (mean), w(n) = f deg:
 *   agg[n, o] = sum_j (S_j[n] @ M_j)[o] + b[o] w(n),   S_j[n, k] = f sum_{p in n} y[p, j] m[p, k],  M_j[k, o] = M_cat[k, o*Dy+j]
 *   yagg_fwd: S [Dy][n_nodes][Dx] from the per-edge inputs m [E][Dx] (receiver-sorted, rowptr);  the caller runs the
 *             Dy node-level GEMMs (eqnn_gemm_f32 with strides Do*Dy, Dy on M_cat);  bias_deg_fwd adds b w(n) in place
 *   yagg_bwd: dm[p, k] = f sum_j y[p, j] dS_j[n(p), k];   bias_deg_bwd: gb[c] = sum_n w(n) d[n, c]
 * agg: 0 sum | 1 mean.  Dy <= 16. */
int eqnn_segnn_yagg_fwd(const float* m, int Dx, const float* y, int Dy, const int32_t* rowptr, int n_nodes, int agg,
                        float* S, eqnn_stream_t stream);
int eqnn_segnn_yagg_bwd(const float* dS, int Dx, const float* y, int Dy, const int32_t* rowptr, int n_nodes, int agg,
                        float* dm, eqnn_stream_t stream);
int eqnn_segnn_bias_deg_fwd(float* out, int ld, const float* bias, int n_bias, const int32_t* rowptr, int n_nodes, int agg,
                            eqnn_stream_t stream);
int eqnn_segnn_bias_deg_bwd(const float* d, int ld, const int32_t* rowptr, int n_nodes, int agg, int n_bias, float* gb,
                            eqnn_stream_t stream);
/* dst[p, :] = src[perm[p], :]  (per-edge arrays into receiver-sorted order);  y += a * x */
int eqnn_gather_rows(const float* src, const int32_t* perm, int64_t n, int width, float* dst, eqnn_stream_t stream);
int eqnn_axpy(int64_t n, float a, const float* x, float* y, eqnn_stream_t stream);

/* ---- EGNN coordinate layer (models/egnn.py, positions-only) ---------------------------------------------
 * geometry (:82-90):  rij4[p] = (r_ij, |r_ij|), r_ij = x_s - x_r + 1e-7 then minimum image (cell may be NULL);
 *                     feats[p, :] = e3nn.bessel(|r_ij|, n_rb, r_max)  ([n_edges, 1] = |r_ij| if n_rb == 0)
 * coordinates (:107-114, jraph segment reduce, :146-150):
 *                     pos_out[r] = wrap( pos[r] + agg_{e->r} clip(r_ij * [tanh](t_e), +-100) )
 * backward, two phases around the edge-MLP backward:  dfeats == NULL -> dt[p];  dfeats given -> dre (per-edge
 * dL/dr_ij stored at the ORIGINAL edge id) and dpos_in (identity + receiver end); eqnn_egnn_sender_acc adds the
 * sender end from the sender-sorted CSR.  cell / cell_inv: HOST pointers to 9 floats or NULL.
 * agg: 0 sum, 1 mean, 2 max (jraph.segment_max, egnn.py:282-284): argmax [n_nodes, 3] = CSR position of the first edge
 * attaining each component's maximum, written by the forward pass and read by the backward pass (NULL otherwise). */
int eqnn_egnn_geom_fwd(const float* pos, int ld_pos, int n_nodes, const int32_t* rowptr, const int32_t* src,
                       const float* cell, const float* cell_inv, int n_rb, double r_max, float* rij4, float* feats,
                       eqnn_stream_t stream);
int eqnn_egnn_coord_fwd(const float* rij4, const float* t, int use_tanh, int agg, const int32_t* rowptr, int n_nodes,
                        const float* pos, int ld_pos, const float* cell, const float* cell_inv, float* pos_out,
                        int ld_out, int32_t* argmax, eqnn_stream_t stream);
int eqnn_egnn_coord_bwd(const float* rij4, const float* t, int use_tanh, int agg, const int32_t* rowptr,
                        const int32_t* perm, int n_nodes, const float* dpos_out, int n_rb, double r_max,
                        const float* dfeats, float* dt, float* dre, float* dpos_in, const int32_t* argmax,
                        eqnn_stream_t stream);
int eqnn_egnn_sender_acc(const float* dre, const int32_t* rowptr_s, const int32_t* perm_s, int n_nodes, float* dpos,
                         eqnn_stream_t stream);
/* node layout (x, v, h) with scalar attributes (egnn.py:47-60, 198-229):
 * edge-MLP input U[p] = [feats[p] (nf) | h[sender] (dh) | h[receiver] (dh) | globals[graph of p] (n_glob)] in
 * receiver-sorted order (jraph broadcasts the graph globals to every edge, egnn.py:41,46,53,60), and its backward
 * (dfeats may be NULL; dh_out[n, :dh] += receiver row sums + sender-sorted gather; dhs_e: [n_edges, dh] scratch). */
int eqnn_egnn_edge_input_fwd(const float* feats, int nf, const float* h, int ld_h, int dh, const float* glob, int n_glob,
                             int nodes_per_graph, const int32_t* rowptr, const int32_t* src, int n_nodes, float* U,
                             eqnn_stream_t stream);
int eqnn_egnn_edge_input_bwd(const float* dU, int nf, int dh, int n_glob, const int32_t* rowptr, const int32_t* perm,
                             const int32_t* rowptr_s, const int32_t* perm_s, int n_nodes, float* dfeats, float* dhs_e,
                             float* dh_out, int ld_dh, eqnn_stream_t stream);
/* base[n] = sx[n] x[n] + sv[n] v[n] (sx NULL = 1): the coordinate layer then forms x' = wrap(base + agg)  (:213-227) */
int eqnn_egnn_xv_base(const float* nodes, int F, const float* sx, const float* sv, int64_t n, float* base,
                      eqnn_stream_t stream);
/* nodes_next[n, 3:] = (columns [3, h_off) unchanged (the OLD velocity, :229), h + ph); x' is written by
 * eqnn_egnn_coord_fwd */
int eqnn_egnn_node_assemble(const float* nodes, int F, int h_off, const float* ph, int64_t n, float* nodes_next,
                            eqnn_stream_t stream);
/* aggregated messages (jraph segment_sum / segment_mean of m_ij; read by the (x, h) and (x, v) node functions,
 * egnn.py:161-166,180-196): out[r, :d] = agg_{e->r} act(rows[e, :]) with act 1 = gelu (messages kept as
 * pre-activations) and its backward target[e, :] += dmi[r(e), :] / deg * act'(z[e, :]);  agg = 2 (segment_max):
 * argmax [n_nodes, d] = CSR position of the winning edge, which alone receives the gradient (NULL for sum / mean) */
int eqnn_egnn_segment_rows_fwd(const float* rows, int d, int act, const int32_t* rowptr, int n_nodes, int agg, float* out,
                               int ld_out, int32_t* argmax, eqnn_stream_t stream);
int eqnn_egnn_segment_rows_bwd(const float* dmi, int ld_dmi, int d, const float* z, int act, const int32_t* rowptr,
                               int n_nodes, int agg, float* target, const int32_t* argmax, eqnn_stream_t stream);
/* backward of the node update: dsx = g.x, dsv = g.v, din = (sx g + (mix - g), sv g + dnext_v, dnext_h); mix = g +
 * geometry gradient (NULL: none); din may be NULL (first step); dnext has row stride ld_dnext >= F (the (x, v) layout
 * returns 6 + d_hidden columns) */
int eqnn_egnn_xv_bwd(const float* nodes, int F, const float* sx, const float* sv, const float* dnext, int ld_dnext,
                     const float* mix, int64_t n, float* din, float* dsx, float* dsv, eqnn_stream_t stream);
/* dst[n, col_dst + j] (+)= src[n / group, col_src + j], j < width  (group > 1: one source row per `group` destination
 * rows, e.g. graph globals broadcast to the nodes) */
int eqnn_copy_cols(const float* src, int ld_src, int col_src, int width, int64_t n, int group, int accumulate, float* dst,
                   int ld_dst, int col_dst, eqnn_stream_t stream);
/* soft edges (egnn.py:101-105): m = gelu(z) * sigmoid(za); backward: dza = dm * gelu(z) * sigmoid'(za),
 * g1 = dm * sigmoid(za)  (g1 may alias dm) */
int eqnn_softgate_fwd(const float* z, const float* za, int64_t n, float* m, eqnn_stream_t stream);
int eqnn_softgate_bwd(const float* z, const float* za, const float* dm, int64_t n, float* dza, float* g1,
                      eqnn_stream_t stream);
/* Parameter gradients of a Dense layer (models/mlp.py, egnn.py phi_e / phi_x / phi_h) over a tall batch, one pass
 * over G:  gW[M x N] = alpha * pro(A)^T G,  gb[n] = sum_k G[k, n]   (A: [K x M] rows lda, G: [K x N] rows ldg).
 * With EQNN_GEMM_TF32X3 and a shape the tcgen05 weight-gradient kernel covers (M in {32, 64, 128}, N = 128,
 * K >= 4096) the column sums are accumulated by the warps that already stream G into the tensor-core operands;
 * other shapes run eqnn_gemm_f32 followed by eqnn_colsum_tall (G dense and 16-byte aligned).  Fixed-order sums. */
size_t eqnn_dense_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int eqnn_dense_wgrad_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int pro,
                         float pro_scale, const float* G, int64_t ldg, float* gW, int64_t ldw, float* gb, int flags,
                         void* ws, size_t ws_bytes, eqnn_stream_t stream);
/* y[n] = sum_m x[m, n] for tall x (bias gradients of the edge MLPs): two-stage, fixed order */
size_t eqnn_colsum_workspace_bytes(int N);
int eqnn_colsum_tall(const float* x, int64_t M, int N, float* y, void* ws, size_t ws_bytes, eqnn_stream_t stream);

/* ---- fused radial MLP (models/nequip.py:55-68: e3nn.bessel -> e3nn.flax.MultiLayerPerceptron) -------------
 * w_t[c, e] = ( act(...act(bessel(d_e) @ W_0 / sqrt(n_rb)) ... @ W_{n_hidden-1} / sqrt(128)) @ W_out / sqrt(128) )[c]
 * with act(x) = activation(x) * act_scale (e3nn.normalize_function; activation EQNN_ACT_GELU or EQNN_ACT_SILU).  ONE kernel from the edge length to the radial
 * weights: the Bessel basis is evaluated in-kernel (fp32 rounding points of the reference: phase fl(fl(j*pi/r_cut)*d)),
 * every weight matrix is resident in shared memory, activations stay in tensor memory between layers (tcgen05.mma with
 * the A operand in TMEM); per edge the kernel reads d (4 bytes) and writes n_out floats.  Arithmetic: operands split
 * into fp16 (hi, lo) pairs, three tensor-core products each, fp32 accumulation (error ~2^-21, like EQNN_GEMM_TF32X3).
 *   dist[e * dist_stride]  edge length (e.g. geom + 3 with stride 4)
 *   weights  HOST array of n_hidden + 1 device pointers; weights[l][k * ldw[l] + n]; ldw HOST array
 *   the last matrix has n_out <= 16 columns (the live message columns, SH normalisation folded in)
 *   w_t [n_out][ld_wt] column-major (the layout eqnn_tpconv_fwd reads)
 * eqnn_radial_mlp_supported: 1 if the shape is covered (n_rb in {16,32,64}, d_hidden = 128, 1 <= n_hidden <= 3). */
int eqnn_radial_mlp_supported(int n_rb, int d_hidden, int n_hidden, int n_out);
/* Backward of the same MLP w.r.t. its weights (the only gradient the reference takes: jax.value_and_grad w.r.t.
 * params, train_cosmology.py:245; edge lengths carry no gradient).  dw_t [n_out][ld_dwt] = dL/dw_t.
 * grad_weights: HOST array of n_hidden + 1 device pointers with the shapes / leading dimensions of `weights`
 * (overwritten).  A chain kernel recomputes the activations from d in tensor memory and runs the data-gradient
 * chain dL/dz_l = (dL/dz_{l+1} @ W_{l+1}^T) * gelu'(z_l) without leaving the SM; it emits gelu(z_l) and dL/dz_l once
 * (the operands of the weight-gradient GEMMs, which run on the tcgen05 weight-gradient kernel with the edge axis as K).
 * Two variants.  With z_stash (written by eqnn_radial_mlp_fwd): every hidden layer l >= 2 runs as ONE kernel that forms
 * dL/dz_{l-1} and dL/dW_l in a single pass over the edges (both weight-gradient operands edge-major in shared memory; the
 * layer fed by dw forms dL/dz_H itself), the first and the output layer's weight gradients run on the tcgen05
 * weight-gradient kernel; radial = the Bessel basis [n_edges][n_rb] of eqnn_edge_geom (may be NULL for n_rb 32 / 64,
 * n_edges >= 4096, dist_stride 4: the basis is then regenerated inside the weight-gradient kernel, slower).
 * With z_stash = NULL nothing is kept from the forward pass: the activations are recomputed from d (radial unused).
 * ws: eqnn_radial_mlp_bwd_workspace_bytes. */
size_t eqnn_radial_mlp_bwd_workspace_bytes(int64_t n_edges, int n_rb, int n_hidden);
int eqnn_radial_mlp_bwd(const float* dist, int64_t dist_stride, int64_t n_edges, int n_rb, double r_cut, int n_hidden,
                        int d_hidden, const float* const* weights, const int64_t* ldw, int n_out, int activation,
                        float act_scale, const float* dw_t, int64_t ld_dwt, const float* z_stash, const float* radial,
                        float* const* grad_weights, void* ws, size_t ws_bytes, eqnn_stream_t stream);
int eqnn_radial_mlp_fwd(const float* dist, int64_t dist_stride, int64_t n_edges, int n_rb, double r_cut, int n_hidden,
                        int d_hidden, const float* const* weights, const int64_t* ldw, int n_out, int activation,
                        float act_scale, float* w_t, int64_t ld_wt, float* z_stash, eqnn_stream_t stream);
/* z_stash (optional, eqnn_radial_mlp_stash_bytes, 16-byte aligned): the forward kernel also writes what the backward
 * pass reads instead of recomputing -- the pre-activations z_1 .. z_{H-1} and, for the last hidden layer, gelu(z_H) and
 * gelu'(z_H) (opaque blocked layout; 512 bytes per edge and matrix, n_hidden + 1 matrices). */
size_t eqnn_radial_mlp_stash_bytes(int64_t n_edges, int n_hidden);

/* ---- shared radial evaluations ----------------------------------------------------------
 * The radial MLP (models/nequip.py:55-68) sees an edge only through its length d, and the two directions of a
 * reciprocated edge have bit-identical lengths in fp32 (x_s - x_r is exactly antisymmetric; 85 % of the edges of a
 * kNN-20 graph are reciprocated).  eqnn_edge_pairs finds those pairs on the receiver-sorted CSR and numbers the unique
 * evaluations (a pair's smaller CSR slot represents it; unpaired slots -- self-loops, one-directional edges, reverse
 * edges whose lengths differ in any bit -- represent themselves).  The MLP is then evaluated once per unique row
 * (eqnn_radial_mlp_fwd_rows), the interaction kernels read the rows through the map (eqnn_tpconv_*_shared), and the
 * cotangents of a pair are summed before the MLP's backward pass (eqnn_pair_combine; that pass is linear in dL/dw for a
 * fixed d).  Same arithmetic on the same numbers per edge: the forward result is bit-identical to the per-edge evaluation.
 * The unique count stays in device memory (n_unique): consumers take the capacity n_edges as their launch bound and read
 * the live count from that word -- no host synchronisation, CUDA-graph capturable.
 * Self-loops of length +0.0 (every node of a kNN graph has one: graph_utils.py:67) all share ONE row, the "zero class";
 * its cotangent is the sum over all of them (eqnn_pair_combine reduces it in a fixed order).
 *   uidx        [n_edges] unique row of every CSR slot          rep_slot    [n_edges cap] CSR slot of row u's representative
 *   rep_partner [n_edges cap] its partner slot, -1 (none) or -2 (zero class)   d_u  [n_edges cap] edge length of row u
 *   zero_slot   [n_nodes] the row's zero-class slot or -1       zero_row    [1] unique row of the zero class or -1
 *   n_unique    [1] int32                                       ws: eqnn_edge_pairs_workspace_bytes                       */
size_t eqnn_edge_pairs_workspace_bytes(int n_nodes, int64_t n_edges);
int eqnn_edge_pairs(const int32_t* rowptr, const int32_t* src, const float* geom, int n_nodes, int64_t n_edges, int32_t* uidx,
                    int32_t* rep_slot, int32_t* rep_partner, float* d_u, int32_t* zero_slot, int32_t* n_unique,
                    int32_t* zero_row, void* ws, size_t ws_bytes, eqnn_stream_t stream);
/* dw_t[c][u] = dw_rows[rep_slot[u]][c] + dw_rows[rep_partner[u]][c] for u < *n_unique, and the sum over all zero-class slots
 * for u = *zero_row (dw_rows: [n_edges][ld_rows] records written by eqnn_tpconv_bwd_shared; dw_t: [n_out][ld_dwt]
 * column-major, the layout eqnn_radial_mlp_bwd_rows reads); zero_slot / zero_row may be NULL (no zero class); ws:
 * eqnn_pair_combine_workspace_bytes */
size_t eqnn_pair_combine_workspace_bytes(void);
int eqnn_pair_combine(const float* dw_rows, int ld_rows, int n_out, const int32_t* rep_slot, const int32_t* rep_partner,
                      const int32_t* zero_slot, const int32_t* zero_row, int n_nodes, int64_t cap, const int32_t* n_unique,
                      float* dw_t, int64_t ld_dwt, void* ws, size_t ws_bytes, eqnn_stream_t stream);
/* out [cap][n_rb] = e3nn.bessel(d_rows, n_rb, r_cut) for the first *n_rows rows (reference rounding points, as eqnn_edge_geom) */
int eqnn_bessel_rows(const float* d_rows, int64_t cap, const int32_t* n_rows, int n_rb, double r_cut, float* out,
                     eqnn_stream_t stream);
/* eqnn_radial_mlp_fwd / _bwd over a compact row array whose live count *n_rows (<= n_cap) sits in device memory; n_cap sizes
 * the buffers, the stash and the launch.  Forward: row-major records w_rows [n_cap][ld_rows] (ld_rows = n_out rounded up to
 * 4) and / or column-major w_t (either may be NULL).  Backward: the z-stash variant on the fused layer kernels (n_hidden >= 2,
 * n_cap >= 4096); radial = eqnn_bessel_rows of the same rows. */
int eqnn_radial_mlp_fwd_rows(const float* dist, int64_t dist_stride, int64_t n_cap, const int32_t* n_rows, int n_rb,
                             double r_cut, int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw,
                             int n_out, int activation, float act_scale, float* w_t, int64_t ld_wt, float* w_rows,
                             int ld_rows, float* z_stash, eqnn_stream_t stream);
int eqnn_radial_mlp_bwd_rows(const float* dist, int64_t dist_stride, int64_t n_cap, const int32_t* n_rows, int n_rb,
                             double r_cut, int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw,
                             int n_out, int activation, float act_scale, const float* dw_t, int64_t ld_dwt,
                             const float* z_stash, const float* radial, float* const* grad_weights, void* ws,
                             size_t ws_bytes, eqnn_stream_t stream);
/* eqnn_tpconv_fwd / _bwd with the radial weights of CSR slot p read from w_rows[uidx[p]] (ld_rows = NLIVE rounded up to 4);
 * the backward writes dL/dw per slot as row-major records dw_rows [n_edges][ld_rows] (eqnn_pair_combine sums the pairs) */
int eqnn_tpconv_fwd_shared(int tp_id, const int32_t* rowptr, const int32_t* src, const float* geom, const int32_t* uidx,
                           const float* w_rows, int ld_rows, const float* xt, int n_nodes, int agg, float* A, int ld_a,
                           int32_t* argmax, eqnn_stream_t stream);
int eqnn_tpconv_bwd_shared(int tp_id, const int32_t* rowptr, const int32_t* src, const int32_t* perm, const float* geom,
                           const int32_t* uidx, const float* w_rows, int ld_rows, const float* xt, int n_nodes, int agg,
                           const float* gA, int ld_ga, float* dw_rows, float* dxe, const int32_t* argmax,
                           eqnn_stream_t stream);

/* ---- parameter plumbing ----------------------------------------------------------------
 * dst[t] = idx[t] < 0 ? 0 : scale[t] * params[idx[t]]     (Linear -> dense matrix, live columns)
 * grads[pidx[j]] = sum_{t in [ptr[j], ptr[j+1])} scale_t[t] * ddst[tidx[t]]               */
int eqnn_weights_expand(const float* params, const int32_t* idx, const float* scale, int64_t n_dst, float* dst,
                        eqnn_stream_t stream);
int eqnn_weights_contract(const float* ddst, const int32_t* pidx, const int32_t* ptr, const int32_t* tidx,
                          const float* scale_t, int64_t n_params, float* grads, eqnn_stream_t stream);

/* ---- training step tail (benchmarks/galaxies/train_cosmology.py:236-247) --------------- */
/* loss = mean_b mean_o (pred - target)^2 ; gpred = dloss/dpred * gscale                    */
int eqnn_mse_loss(const float* pred, const float* target, int B, int O, float gscale, float* loss, float* gpred,
                  eqnn_stream_t stream);
/* optax.adamw (decoupled weight decay); lr read from the host argument                     */
int eqnn_adamw(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float b1, float b2,
               float eps, float wd, int step, eqnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EQNN_B200_H_ */
