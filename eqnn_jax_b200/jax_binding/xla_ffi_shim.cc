// XLA typed-FFI handlers (jax.ffi custom calls) over the C ABI of include/eqnn_b200.h.
//
// One handler per op of the NequIP hot path: kNN, CSR, edge geometry, fused radial MLP (fwd / bwd), interaction
// kernel (fwd / bwd), fused node update (fwd / bwd).  A handler only unpacks ffi::Buffers and forwards pointers, sizes
// and the platform stream to the C function; XLA owns every buffer (inputs, results and the scratch results that
// serve as workspaces), nothing is allocated here, nothing synchronises -- the calls are command-buffer capturable.
// Non-zero status -> ffi::Error carrying eqnn_last_error_string().
//
// Build (only where jaxlib's headers exist; eqnn_jax_b200/build.py does it when it finds them):
//   g++ -std=c++17 -shared -fPIC -I<jaxlib>/include -I<cuda>/include -Iinclude xla_ffi_shim.cc \
//       -L eqnn_jax_b200 -leqnn_b200 -Wl,-rpath,'$ORIGIN/..' -o eqnn_jax_b200/jax_binding/libeqnn_xla.so
// The Python side (custom_vjp rules, registration) is eqnn_jax_b200/jax_binding/ffi.py.
// NOTE: this image has no jax / jaxlib, so the file is compiled and exercised only on a machine that has them.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <vector>

#include "eqnn_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using U8 = ffi::Buffer<ffi::U8>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

ffi::Error status(int rc) { return rc ? ffi::Error::Internal(eqnn_last_error_string()) : ffi::Error::Success(); }

// ---- graph_utils.nearest_neighbors (vmapped over the batch), models/utils/graph_utils.py:40-72 ----
// x [B, N, F]; cell / cell_inv: 9 floats each as attributes-by-value are not supported for arrays, so they arrive as
// tiny HOST-resident attribute spans (row-major 3x3); pbc = 0 -> no minimum image.
ffi::Error KnnPbc(cudaStream_t stream, F32 x, ffi::Span<const float> cell, ffi::Span<const float> cell_inv, int32_t k,
                  int32_t flags, RS32 senders, RS32 receivers, RF32 dr, RU8 ws) {
  const auto d = x.dimensions();
  const bool pbc = cell.size() == 9 && cell_inv.size() == 9;
  return status(eqnn_knn_pbc(x.typed_data(), (int)d[2], pbc ? cell.begin() : nullptr, pbc ? cell_inv.begin() : nullptr,
                             (int)d[0], (int)d[1], k, flags | (pbc ? EQNN_KNN_PBC : 0), senders->typed_data(),
                             receivers->typed_data(), dr->typed_data(), ws->untyped_data(), ws->size_bytes(), stream));
}

// ---- receiver-sorted CSR (replaces jraph's unsorted segment reduction, models/nequip.py:118-120) ----
ffi::Error CsrBuild(cudaStream_t stream, S32 senders, S32 receivers, int32_t n_nodes, RS32 rowptr, RS32 perm, RS32 src,
                    RU8 ws) {
  const auto d = receivers.dimensions();
  return status(eqnn_csr_build(senders.typed_data(), receivers.typed_data(), (int)d[0], n_nodes, (int)d[1],
                               rowptr->typed_data(), perm->typed_data(), src->typed_data(), ws->untyped_data(),
                               ws->size_bytes(), stream));
}

// ---- r_ij, |r_ij|, Bessel basis in CSR order, models/nequip.py:129-134,55-57 ----
ffi::Error EdgeGeom(cudaStream_t stream, F32 pos, S32 rowptr, S32 src, int32_t n_rb, double r_cut, RF32 geom, RF32 radial) {
  const auto d = pos.dimensions();
  return status(eqnn_edge_geom(pos.typed_data(), (int)d[d.size() - 1], (int)(pos.element_count() / d[d.size() - 1]),
                               rowptr.typed_data(), src.typed_data(), n_rb, r_cut, geom->typed_data(), radial->typed_data(),
                               stream));
}

// ---- fused radial MLP, models/nequip.py:55-68 (hidden kernels W0..W2 [fan_in, 128], output kernel [128, n_out]) ----
// Hidden kernels arrive as w0, w1, w2 (the unused trailing ones as zero-size arrays when n_hidden < 3).
ffi::Error RadialMlpFwd(cudaStream_t stream, F32 geom, F32 w0, F32 w1, F32 w2, F32 wout, int32_t n_hidden, int32_t n_rb,
                        double r_cut, int32_t activation, float act_scale, RF32 w_t, RF32 z_stash) {
  const float* hidden[3] = {w0.typed_data(), w1.typed_data(), w2.typed_data()};
  const float* W[4];
  int64_t ldw[4];
  for (int l = 0; l < n_hidden && l < 3; ++l) { W[l] = hidden[l]; ldw[l] = 128; }
  W[n_hidden] = wout.typed_data();
  ldw[n_hidden] = (int64_t)wout.dimensions()[1];
  const int64_t n_edges = geom.dimensions()[0];
  return status(eqnn_radial_mlp_fwd(geom.typed_data() + 3, 4, n_edges, n_rb, r_cut, n_hidden, 128, W, ldw,
                                    (int)wout.dimensions()[1], activation, act_scale, w_t->typed_data(), n_edges,
                                    z_stash->element_count() ? z_stash->typed_data() : nullptr, stream));
}
ffi::Error RadialMlpBwd(cudaStream_t stream, F32 geom, F32 radial, F32 z_stash, F32 w0, F32 w1, F32 w2, F32 wout, F32 dw_t,
                        int32_t n_hidden, int32_t n_rb, double r_cut, int32_t activation, float act_scale, RF32 g0, RF32 g1,
                        RF32 g2, RF32 gout, RU8 ws) {
  const float* hidden[3] = {w0.typed_data(), w1.typed_data(), w2.typed_data()};
  float* ghidden[3] = {g0->typed_data(), g1->typed_data(), g2->typed_data()};
  const float* W[4];
  float* G[4];
  int64_t ldw[4];
  for (int l = 0; l < n_hidden && l < 3; ++l) { W[l] = hidden[l]; G[l] = ghidden[l]; ldw[l] = 128; }
  W[n_hidden] = wout.typed_data();
  G[n_hidden] = gout->typed_data();
  ldw[n_hidden] = (int64_t)wout.dimensions()[1];
  const int64_t n_edges = geom.dimensions()[0];
  return status(eqnn_radial_mlp_bwd(geom.typed_data() + 3, 4, n_edges, n_rb, r_cut, n_hidden, 128, W, ldw,
                                    (int)wout.dimensions()[1], activation, act_scale, dw_t.typed_data(), n_edges,
                                    z_stash.element_count() ? z_stash.typed_data() : nullptr, radial.typed_data(), G,
                                    ws->untyped_data(), ws->size_bytes(), stream));
}

// ---- interaction kernel: SH + Clebsch-Gordan TP + radial weights + receiver reduction, models/nequip.py:46-52,68,118-120 ----
// argmax: [n_nodes, d_live] for agg == 2 (segment_max), a zero-size array otherwise
ffi::Error TpconvFwd(cudaStream_t stream, S32 rowptr, S32 src, F32 geom, F32 w_t, F32 xt, int32_t tp_id, int32_t agg, RF32 A,
                     RS32 argmax) {
  return status(eqnn_tpconv_fwd(tp_id, rowptr.typed_data(), src.typed_data(), geom.typed_data(), w_t.typed_data(),
                                (int64_t)src.element_count(), xt.typed_data(), (int)xt.dimensions()[0], agg, A->typed_data(),
                                (int)A->dimensions()[1], argmax->element_count() ? argmax->typed_data() : nullptr, stream));
}
ffi::Error TpconvBwd(cudaStream_t stream, S32 rowptr, S32 src, S32 perm, F32 geom, F32 w_t, F32 xt, F32 gA, S32 argmax,
                     int32_t tp_id, int32_t agg, RF32 dw_t, RF32 dxe) {
  return status(eqnn_tpconv_bwd(tp_id, rowptr.typed_data(), src.typed_data(), perm.typed_data(), geom.typed_data(),
                                w_t.typed_data(), (int64_t)src.element_count(), xt.typed_data(), (int)xt.dimensions()[0], agg,
                                gA.typed_data(), (int)gA.dimensions()[1], dw_t->typed_data(), dxe->typed_data(),
                                argmax.element_count() ? argmax.typed_data() : nullptr, stream));
}

// ---- shared radial evaluations: one radial-MLP evaluation per unique edge length ----
// counts = [n_unique, zero_row] (two int32 DEVICE words); every per-row result has the edge count as its static capacity
ffi::Error EdgePairs(cudaStream_t stream, S32 rowptr, S32 src, F32 geom, int32_t n_rb, double r_cut, RS32 uidx, RS32 rep_slot,
                     RS32 rep_partner, RF32 d_u, RS32 zero_slot, RS32 counts, RF32 radial_u, RU8 ws) {
  const int n_nodes = (int)rowptr.element_count() - 1;
  const int64_t n_edges = (int64_t)src.element_count();
  int rc = eqnn_edge_pairs(rowptr.typed_data(), src.typed_data(), geom.typed_data(), n_nodes, n_edges, uidx->typed_data(),
                           rep_slot->typed_data(), rep_partner->typed_data(), d_u->typed_data(), zero_slot->typed_data(),
                           counts->typed_data(), counts->typed_data() + 1, ws->untyped_data(), ws->size_bytes(), stream);
  if (!rc && n_rb > 0)
    rc = eqnn_bessel_rows(d_u->typed_data(), n_edges, counts->typed_data(), n_rb, r_cut, radial_u->typed_data(), stream);
  return status(rc);
}
ffi::Error RadialMlpFwdRows(cudaStream_t stream, F32 d_u, S32 counts, F32 w0, F32 w1, F32 w2, F32 wout, int32_t n_hidden,
                            int32_t n_rb, double r_cut, int32_t activation, float act_scale, RF32 w_rows, RF32 z_stash) {
  const float* hidden[3] = {w0.typed_data(), w1.typed_data(), w2.typed_data()};
  const float* W[4];
  int64_t ldw[4];
  for (int l = 0; l < n_hidden && l < 3; ++l) { W[l] = hidden[l]; ldw[l] = 128; }
  W[n_hidden] = wout.typed_data();
  ldw[n_hidden] = (int64_t)wout.dimensions()[1];
  return status(eqnn_radial_mlp_fwd_rows(d_u.typed_data(), 1, (int64_t)d_u.element_count(), counts.typed_data(), n_rb, r_cut,
                                         n_hidden, 128, W, ldw, (int)wout.dimensions()[1], activation, act_scale, nullptr, 0,
                                         w_rows->typed_data(), (int)w_rows->dimensions()[1],
                                         z_stash->element_count() ? z_stash->typed_data() : nullptr, stream));
}
// dw_rows [n_edges, ld_rows] (from TpconvBwdShared) -> per-row cotangents (scratch result dw_t) -> weight gradients
ffi::Error RadialMlpBwdRows(cudaStream_t stream, F32 d_u, S32 counts, S32 rep_slot, S32 rep_partner, S32 zero_slot, F32 radial_u,
                            F32 z_stash, F32 w0, F32 w1, F32 w2, F32 wout, F32 dw_rows, int32_t n_hidden, int32_t n_rb,
                            double r_cut, int32_t activation, float act_scale, RF32 g0, RF32 g1, RF32 g2, RF32 gout, RF32 dw_t,
                            RU8 ws_pair, RU8 ws) {
  const float* hidden[3] = {w0.typed_data(), w1.typed_data(), w2.typed_data()};
  float* ghidden[3] = {g0->typed_data(), g1->typed_data(), g2->typed_data()};
  const float* W[4];
  float* G[4];
  int64_t ldw[4];
  for (int l = 0; l < n_hidden && l < 3; ++l) { W[l] = hidden[l]; G[l] = ghidden[l]; ldw[l] = 128; }
  W[n_hidden] = wout.typed_data();
  G[n_hidden] = gout->typed_data();
  const int n_out = (int)wout.dimensions()[1];
  ldw[n_hidden] = n_out;
  const int64_t cap = (int64_t)d_u.element_count();
  int rc = eqnn_pair_combine(dw_rows.typed_data(), (int)dw_rows.dimensions()[1], n_out, rep_slot.typed_data(),
                             rep_partner.typed_data(), zero_slot.typed_data(), counts.typed_data() + 1,
                             (int)zero_slot.element_count(), cap, counts.typed_data(), dw_t->typed_data(), cap,
                             ws_pair->untyped_data(), ws_pair->size_bytes(), stream);
  if (!rc)
    rc = eqnn_radial_mlp_bwd_rows(d_u.typed_data(), 1, cap, counts.typed_data(), n_rb, r_cut, n_hidden, 128, W, ldw, n_out,
                                  activation, act_scale, dw_t->typed_data(), cap, z_stash.typed_data(), radial_u.typed_data(), G,
                                  ws->untyped_data(), ws->size_bytes(), stream);
  return status(rc);
}
ffi::Error TpconvFwdShared(cudaStream_t stream, S32 rowptr, S32 src, F32 geom, S32 uidx, F32 w_rows, F32 xt, int32_t tp_id,
                           int32_t agg, RF32 A, RS32 argmax) {
  return status(eqnn_tpconv_fwd_shared(tp_id, rowptr.typed_data(), src.typed_data(), geom.typed_data(), uidx.typed_data(),
                                       w_rows.typed_data(), (int)w_rows.dimensions()[1], xt.typed_data(), (int)xt.dimensions()[0],
                                       agg, A->typed_data(), (int)A->dimensions()[1],
                                       argmax->element_count() ? argmax->typed_data() : nullptr, stream));
}
ffi::Error TpconvBwdShared(cudaStream_t stream, S32 rowptr, S32 src, S32 perm, F32 geom, S32 uidx, F32 w_rows, F32 xt, F32 gA,
                           S32 argmax, int32_t tp_id, int32_t agg, RF32 dw_rows, RF32 dxe) {
  return status(eqnn_tpconv_bwd_shared(tp_id, rowptr.typed_data(), src.typed_data(), perm.typed_data(), geom.typed_data(),
                                       uidx.typed_data(), w_rows.typed_data(), (int)w_rows.dimensions()[1], xt.typed_data(),
                                       (int)xt.dimensions()[0], agg, gA.typed_data(), (int)gA.dimensions()[1],
                                       dw_rows->typed_data(), dxe->typed_data(),
                                       argmax.element_count() ? argmax.typed_data() : nullptr, stream));
}

// ---- fused node update: Linear + self-connection + gate + re-projection, models/nequip.py:80-90,162 ----
ffi::Error NodeUpdateFwd(cudaStream_t stream, F32 A, F32 h, F32 W1, F32 W2, F32 W3, float alpha1, int32_t n_extra,
                         int32_t n_gates, ffi::Span<const int32_t> mul, ffi::Span<const int32_t> l, float inv_c_act,
                         float inv_c_gate, RF32 z, RF32 h_out) {
  std::vector<int> m(mul.begin(), mul.end()), ll(l.begin(), l.end());
  return status(eqnn_node_update_fwd(A.typed_data(), (int)A.dimensions()[1], (int)W1.dimensions()[0], h.typed_data(),
                                     (int)h.dimensions()[1], W1.typed_data(), alpha1, W2.element_count() ? W2.typed_data() : nullptr,
                                     n_extra, n_gates, (int)m.size(), m.data(), ll.data(), inv_c_act, inv_c_gate, W3.typed_data(),
                                     nullptr, 0, (int64_t)A.dimensions()[0], z->typed_data(), h_out->typed_data(), nullptr, stream));
}
ffi::Error NodeUpdateBwd(cudaStream_t stream, F32 A, F32 h, F32 W1, F32 W2, F32 W3, F32 z, F32 dh_out, float alpha1,
                         int32_t n_extra, int32_t n_gates, ffi::Span<const int32_t> mul, ffi::Span<const int32_t> l,
                         float inv_c_act, float inv_c_gate, RF32 dA, RF32 dh_in, RF32 dW1, RF32 dW2, RF32 dW3, RU8 ws) {
  std::vector<int> m(mul.begin(), mul.end()), ll(l.begin(), l.end());
  const bool res = W2.element_count() != 0;
  return status(eqnn_node_update_bwd(A.typed_data(), (int)A.dimensions()[1], (int)W1.dimensions()[0], h.typed_data(),
                                     (int)h.dimensions()[1], W1.typed_data(), alpha1, res ? W2.typed_data() : nullptr, n_extra,
                                     n_gates, (int)m.size(), m.data(), ll.data(), inv_c_act, inv_c_gate, W3.typed_data(),
                                     z.typed_data(), dh_out.typed_data(), (int64_t)A.dimensions()[0], dA->typed_data(),
                                     dh_in->typed_data(), dW1->typed_data(), res ? dW2->typed_data() : nullptr, dW3->typed_data(),
                                     ws->untyped_data(), ws->size_bytes(), stream));
}

}  // namespace

#define EQNN_STREAM Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_knn_pbc_ffi, KnnPbc,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Attr<ffi::Span<const float>>("cell")
                                  .Attr<ffi::Span<const float>>("cell_inv").Attr<int32_t>("k").Attr<int32_t>("flags")
                                  .Ret<S32>().Ret<S32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_csr_build_ffi, CsrBuild,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Attr<int32_t>("n_nodes")
                                  .Ret<S32>().Ret<S32>().Ret<S32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_edge_geom_ffi, EdgeGeom,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<S32>().Arg<S32>().Attr<int32_t>("n_rb")
                                  .Attr<double>("r_cut").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_radial_mlp_fwd_ffi, RadialMlpFwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("n_hidden").Attr<int32_t>("n_rb").Attr<double>("r_cut").Attr<int32_t>("activation")
                                  .Attr<float>("act_scale").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_radial_mlp_bwd_ffi, RadialMlpBwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Attr<int32_t>("n_hidden").Attr<int32_t>("n_rb").Attr<double>("r_cut")
                                  .Attr<int32_t>("activation").Attr<float>("act_scale").Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_tpconv_fwd_ffi, TpconvFwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("tp_id").Attr<int32_t>("agg").Ret<F32>().Ret<S32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_tpconv_bwd_ffi, TpconvBwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<S32>().Attr<int32_t>("tp_id").Attr<int32_t>("agg").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_node_update_fwd_ffi, NodeUpdateFwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<float>("alpha1").Attr<int32_t>("n_extra").Attr<int32_t>("n_gates")
                                  .Attr<ffi::Span<const int32_t>>("mul").Attr<ffi::Span<const int32_t>>("l")
                                  .Attr<float>("inv_c_act").Attr<float>("inv_c_gate").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_node_update_bwd_ffi, NodeUpdateBwd,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Attr<float>("alpha1").Attr<int32_t>("n_extra").Attr<int32_t>("n_gates")
                                  .Attr<ffi::Span<const int32_t>>("mul").Attr<ffi::Span<const int32_t>>("l")
                                  .Attr<float>("inv_c_act").Attr<float>("inv_c_gate").Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Ret<U8>());
// shared radial evaluations
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_edge_pairs_ffi, EdgePairs,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Arg<F32>().Attr<int32_t>("n_rb")
                                  .Attr<double>("r_cut").Ret<S32>().Ret<S32>().Ret<S32>().Ret<F32>().Ret<S32>().Ret<S32>()
                                  .Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_radial_mlp_fwd_rows_ffi, RadialMlpFwdRows,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("n_hidden").Attr<int32_t>("n_rb").Attr<double>("r_cut").Attr<int32_t>("activation")
                                  .Attr<float>("act_scale").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_radial_mlp_bwd_rows_ffi, RadialMlpBwdRows,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<F32>().Arg<S32>().Arg<S32>().Arg<S32>().Arg<S32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Attr<int32_t>("n_hidden")
                                  .Attr<int32_t>("n_rb").Attr<double>("r_cut").Attr<int32_t>("activation").Attr<float>("act_scale")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U8>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_tpconv_fwd_shared_ffi, TpconvFwdShared,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("tp_id").Attr<int32_t>("agg").Ret<F32>().Ret<S32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eqnn_tpconv_bwd_shared_ffi, TpconvBwdShared,
                              ffi::Ffi::Bind().EQNN_STREAM.Arg<S32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<S32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<S32>().Attr<int32_t>("tp_id").Attr<int32_t>("agg").Ret<F32>()
                                  .Ret<F32>());
