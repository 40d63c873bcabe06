// fesr_nnconv_forward: one block-diagonal batch of subdomains through KernelNN / TEECNet.
#include <stdlib.h>

#include "kernels.cuh"
#include "workspace.cuh"

using namespace fesr;

namespace fesr {

ForwardWs carve_forward(void* base, const fesr_model_dims& d, int64_t n, int64_t E, int keep, int z_half) {
  Carver c(base);
  ForwardWs ws;
  ws.prep = carve_prepared(c, d);
  ws.g = c.take<float>((size_t)(E > 0 ? E : 1) * d.kp);
  const int nh = keep ? d.layers + 1 : 2;
  const int nz = keep ? d.layers : 1;
  ws.n_h = nh;
  ws.n_z = nz;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  for (int i = 0; i < nh; ++i) ws.h[i] = c.take<float>(nn * d.wp);
  for (int i = 0; i < nz; ++i)
    ws.Z[i] = (keep && z_half) ? reinterpret_cast<float*>(c.take<uint16_t>(nn * d.zk)) : c.take<float>(nn * d.zk);
  ws.bytes = c.used();
  return ws;
}

}  // namespace fesr

extern "C" {

size_t fesr_forward_workspace_bytes(const fesr_model_dims* dims, int64_t n, int64_t E, int keep_for_backward) {
  if (!dims || n < 0 || E < 0 || dims->layers > FESR_MAX_LAYERS) return 0;
  return carve_forward(nullptr, *dims, n, E, keep_for_backward & FESR_FWD_KEEP, (keep_for_backward & FESR_FWD_KEEP_Z16) != 0).bytes;
}

size_t fesr_forward_overflow_offset(const fesr_model_dims* dims) {
  if (!dims) return 0;
  ForwardWs ws = carve_forward(nullptr, *dims, 1, 1, 0, 0);
  return (size_t)reinterpret_cast<uintptr_t>(ws.prep.ovf);      // carved from a NULL base: the address IS the offset
}

int fesr_nnconv_forward(const fesr_model_dims* dims, const fesr_params* params, const float* x,
                        const int32_t* rowptr, const int32_t* src_sorted, const int32_t* perm,
                        const float* edge_attr, int64_t n, int64_t E, int precision, int fwd_flags,
                        float* y, void* workspace, size_t workspace_bytes, void* stream_) {
  const int keep_for_backward = fwd_flags & FESR_FWD_KEEP;
  FESR_CHECK_ARG(dims && params, "dims/params NULL");
  FESR_CHECK_ARG(n >= 0 && E >= 0 && n < (1ll << 31) && E < (1ll << 31), "n/E out of range");
  FESR_CHECK_ARG(dims->layers <= FESR_MAX_LAYERS, "too many layers");
  FESR_CHECK_ARG(precision == FESR_PREC_FP32 || precision == FESR_PREC_TF32 || precision == FESR_PREC_F16,
                 "unsupported precision %d (fp32 | tf32 | f16)", precision);
  if (n == 0) return FESR_OK;
  const bool edge_only = (fwd_flags & FESR_FWD_EDGE_ONLY) != 0, edge_done = (fwd_flags & FESR_FWD_EDGE_DONE) != 0;
  FESR_CHECK_ARG(!(edge_only && edge_done), "FESR_FWD_EDGE_ONLY and FESR_FWD_EDGE_DONE exclude each other");
  FESR_CHECK_ARG((edge_only || (x && y)) && rowptr && (E == 0 || (src_sorted && edge_attr)), "NULL pointer");
  const fesr_model_dims& d = *dims;
  // (the layout of a kept workspace follows the precision, not the caller's size-query flag: fesr_nnconv_backward
  // carves it the same way)
  ForwardWs ws = carve_forward(workspace, d, n, E, keep_for_backward, keep_for_backward && z_stash_half(precision));
  if (!workspace || workspace_bytes < ws.bytes) {
    set_error("forward workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  cudaStream_t s = as_stream(stream_);
  int rc;
  // fp16 range guard (common.cuh F16Guard): the flag starts at 0 with the pass (its edge phase when it is issued in
  // two calls), every kernel that packs fp16 raises it on overflow, fc_out turns a raised flag into a NaN output
  if (!edge_done) FESR_CUDA(cudaMemsetAsync(ws.prep.ovf, 0, sizeof(int), s));
  OvfScope ovf_scope(ws.prep.ovf);
  if (!(fwd_flags & FESR_FWD_WEIGHTS_PREPARED) && !edge_done && (rc = launch_prepare_weights(d, *params, ws.prep, s, !keep_for_backward))) return rc;
  static const bool ffma_only = getenv("FESR_ZBUILD_FFMA") != nullptr;   // A/B switch for profiling
  // reduced-precision arms: g and h are rounded to tf32 by their producers, so the gather kernel
  // feeds them to the tensor cores without converting
  const int rnd_in = precision == FESR_PREC_FP32 ? 0 : (precision == FESR_PREC_F16 ? 2 : 1);   // 2: fp16 g and h
  // FESR_PREC_F16 predict of the KernelNN shape: one fused kernel per layer (layer_fused.cu), Z stays on chip.
  // FESR_FUSE=0 selects the two-kernel path (A/B switch for profiling); 1..3 = parts per launch
  const char* fuse_env = getenv("FESR_FUSE");
  const int fuse_mode = fuse_env ? atoi(fuse_env) : 3;
  const bool fused = precision == FESR_PREC_F16 && !keep_for_backward && fuse_mode > 0 && layer_fused_supported(d) &&
                     ws.prep.tfused_h != nullptr && E > 0;
  {
    // fused arm: centred edge features (g - g(0), the lo slot constant) against two-term weights -- DESIGN.md 4.2
    GcenterScope gc_scope(fused ? ws.prep.gcenter : nullptr);
    if (!edge_done && (rc = launch_edge_hidden(d, *params, edge_attr, perm, E, ws.g, s, fused ? 3 : rnd_in))) return rc;
  }
  if (edge_only) return FESR_OK;
  if ((rc = launch_fc_in(d, ws.prep, x, n, ws.h[0], s, rnd_in))) return rc;
  const float* h_last = ws.h[0];
  for (int l = 0; l < d.layers; ++l) {
    const float* h_in = keep_for_backward ? ws.h[l] : ws.h[l & 1];
    float* h_out = keep_for_backward ? ws.h[l + 1] : ws.h[(l + 1) & 1];
    float* Z = keep_for_backward ? ws.Z[l] : ws.Z[0];
    if (fused) {
      if ((rc = launch_layer_fused_f16(d, rowptr, src_sorted, ws.g, E, h_in, n, ws.prep.tfused_h, ws.prep.bias_p, Z, h_out,
                                       d.w <= 43 ? fuse_mode : (fuse_mode > 2 ? 2 : fuse_mode), s, l == d.layers - 1)))
        return rc;
      h_last = h_out;
      continue;
    }
    if (keep_for_backward && z_stash_half(precision)) {
      // training forward of the tf32 arm: fp16 Z stash (workspace.cuh), fp32 g / h as everywhere in this arm
      if ((rc = launch_zbuild_mma(d, rowptr, src_sorted, ws.g, h_in, n, Z, 2, s))) return rc;
      const int epi16 = d.kind == FESR_TEECNET ? EPI_BIAS_CONST1 : EPI_BIAS_RELU;
      if ((rc = launch_node_gemm_f16(d, ws.prep.tprime_t_h, ws.prep.bias_p, epi16, Z, n, h_out, s, 1))) return rc;
      h_last = h_out;
      continue;
    }
    const int zmode = precision == FESR_PREC_FP32 ? 0 : (precision == FESR_PREC_F16 ? 2 : 1);
    if (zmode == 2)
      rc = launch_zbuild_f16(d, rowptr, src_sorted, ws.g, h_in, n, Z, s);
    else if (zmode == 0 || ffma_only)
      rc = launch_zbuild(d, rowptr, src_sorted, ws.g, h_in, n, Z, zmode, s);
    else
      rc = launch_zbuild_mma(d, rowptr, src_sorted, ws.g, h_in, n, Z, zmode, s);
    if (rc) return rc;
    const int epi = d.kind == FESR_TEECNET ? EPI_BIAS_CONST1 : EPI_BIAS_RELU;
    // predict (nothing kept for a backward): multi-term operands, see gemm_tc.cu.  A kept forward runs the plain
    // one-term product the backward differentiates.  FESR_FP32_SIMT=1: the CUDA-core fp32 GEMM (test cross-check)
    static const bool fp32_simt = getenv("FESR_FP32_SIMT") && atoi(getenv("FESR_FP32_SIMT")) != 0;
    const bool multi = !keep_for_backward;
    if (precision == FESR_PREC_FP32 && (fp32_simt || d.zk > 4096 || !(d.wp % 16 == 0 && d.wp <= 64)))
      rc = launch_node_gemm_fp32(d, ws.prep.tprime, ws.prep.bias_p, epi, Z, n, h_out, s);
    else if (precision == FESR_PREC_FP32)
      rc = launch_node_gemm_tf32(d, ws.prep.tprime_t, ws.prep.bias_p, epi, Z, n, h_out, s, 0, ws.prep.tprime_t_lo, 3);
    else if (precision == FESR_PREC_TF32)
      rc = launch_node_gemm_tf32(d, ws.prep.tprime_t, ws.prep.bias_p, epi, Z, n, h_out, s, 1, ws.prep.tprime_t_lo, multi ? 2 : 1);
    else
      rc = launch_node_gemm_f16(d, ws.prep.tprime_t_h, ws.prep.bias_p, epi, Z, n, h_out, s, 2,
                                multi ? ws.prep.tprime_t_h_lo : nullptr);
    if (rc) return rc;
    h_last = h_out;
  }
  return launch_fc_out(d, *params, h_last, n, y, s, precision == FESR_PREC_F16 && !fused);
}

}  // extern "C"
