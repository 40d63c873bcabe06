// Internal launch wrappers shared between the translation units of libfesr.so.
#pragma once
#include "common.cuh"

namespace fesr {

// mlp.cu ------------------------------------------------------------------------------
size_t prepared_bytes(const fesr_model_dims& d);
Prepared carve_prepared(Carver& c, const fesr_model_dims& d);
// with_fused: also the fused f16 predict arm's copies (g(0) centring, T' in the fused K order) -- not needed by a forward
// that keeps its intermediates for a backward (one single-block kernel of 0.1 ms per train step)
int launch_prepare_weights(const fesr_model_dims& d, const fesr_params& p, const Prepared& w, cudaStream_t s,
                           bool with_fused = true);
// g[E, kp]: hidden activations of the edge MLP in CSR edge order, channel-permuted layout
int launch_edge_hidden(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr,
                       const int32_t* perm, int64_t E, float* g, cudaStream_t s, int round_tf32 = 0);
// edge_mlp_mma.cu: the two-hidden-layer (KernelNN) case on mma.sync 3xTF32 (fp32-class accuracy)
int launch_edge_hidden2_mma(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr,
                            const int32_t* perm, int64_t E, float* g, cudaStream_t s, int round_tf32);
// edge_mlp_mma.cu: TEECNet's three hidden layers (1 -> 32 -> 64 -> 128, LeakyReLU) on mma.sync split-fp16 (3 terms);
// omode 1: tf32-rounded fp32 rows, 2: fp16 rows
int launch_edge_hidden3_mma(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr,
                            const int32_t* perm, int64_t E, float* g, cudaStream_t s, int omode);
int launch_fc_in(const fesr_model_dims& d, const Prepared& w, const float* x, int64_t n, float* h, cudaStream_t s,
                 int round_tf32 = 0);
int launch_fc_out(const fesr_model_dims& d, const fesr_params& p, const void* h, int64_t n, float* y, cudaStream_t s,
                  int h_half = 0);

// zbuild.cu ---------------------------------------------------------------------------
// Z[i, :] = (1/max(deg,1)) * sum_{e -> i} g_e (x) h[src_e]   ++  h[i]  (root block)
int launch_zbuild(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted,
                  const float* g, const float* h, int64_t n, float* Z, int round_tf32, cudaStream_t s,
                  int mean = 1, const float* gather_scale = nullptr);

// zbuild_mma.cu: same contract, outer products on mma.sync tf32 (reduced-precision arms only);
// zmode 1 = fp32 Z rounded to tf32, 2 = fp16 Z
int launch_zbuild_mma(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted,
                      const float* g, const float* h, int64_t n, void* Z, int zmode, cudaStream_t s,
                      int mean = 1, const float* gather_scale = nullptr);

// zbuild_f16.cu: FESR_PREC_F16 arm -- g [E,kp], h [n,wp] and Z [n,zk] are all fp16
int launch_zbuild_f16(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted,
                      const void* g_half, const void* h_half, int64_t n, void* Z_half, cudaStream_t s);

// layer_fused.cu: zbuild + node contraction + epilogue of one layer in one kernel (FESR_PREC_F16)
bool layer_fused_supported(const fesr_model_dims& d);
size_t layer_fused_tf_elems(const fesr_model_dims& d);
int launch_prepare_tfused(const fesr_model_dims& d, const float* tprime, const float* mfull, void* tf, cudaStream_t s);
// g3: fp16 planar [kp/16][E][16] (launch_edge_hidden with round mode 3); P: fp32 [n, 48] scratch;
// mode: parts per launch (0 = as many as the TMEM lanes hold)
int launch_layer_fused_f16(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g3,
                           int64_t E, const void* h_in, int64_t n, const void* tf, const float* bias_p, float* P,
                           void* h_out, int mode, cudaStream_t s, int out_f32 = 0, int sum_mode = 0, const void* h_own = nullptr);
// (sum_mode / h_own: the backward's reversed-graph pass, single-launch KernelNN shape only -- backward.cu)

// gemm_simt.cu ------------------------------------------------------------------------
// h_out[n, wp] = epilogue(Z[n, zk] x tprime[zk, wp] + bias)
// epilogue modes
enum { EPI_BIAS_RELU = 0, EPI_BIAS_CONST1 = 1, EPI_NONE = 2 };
// B_rowmajor: [zk, wp]
int launch_node_gemm_fp32(const fesr_model_dims& d, const float* B_rowmajor, const float* bias_p, int epi,
                          const float* Z, int64_t n, float* h_out, cudaStream_t s);

// gemm_tc.cu --------------------------------------------------------------------------
// fp16 variant: Z [n, zk] fp16, B_kmajor [wp, zk] fp16 (tcgen05 kind::f16, fp32 accumulate)
// B_lo_h: [wp, zk] fp16 low-order term of T' (predict: two-term weights) or NULL (one term)
int launch_node_gemm_f16(const fesr_model_dims& d, const void* B_kmajor_h, const float* bias_p, int epi,
                         const void* Z_h, int64_t n, float* h_out, cudaStream_t s, int round_out = 0,
                         const void* B_lo_h = nullptr);
// B_kmajor: [wp, zk] tf32-rounded; terms 1 | 2 (+ Z x B_lo) | 3 (3xTF32 on fp32 Z: the fp32 arm); B_lo: tf32 low-order term
int launch_node_gemm_tf32(const fesr_model_dims& d, const float* B_kmajor, const float* bias_p, int epi,
                          const float* Z, int64_t n, float* h_out, cudaStream_t s, int round_out = 0,
                          const float* B_lo = nullptr, int terms = 1);

}  // namespace fesr
