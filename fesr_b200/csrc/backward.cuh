// Helpers shared by the backward translation units.
#pragma once
#include "kernels.cuh"

namespace fesr {

struct GemmArgs {
  const float* A;
  int64_t sAm, sAk;
  const float* B;
  int64_t sBk, sBn;
  float* C;
  int64_t sCm, sCn;
  int64_t M, N, K;
  int accumulate;
  int tf32;          // 1: products on mma.sync tf32 (fp32 accumulate) -- the reduced-precision arm
};

size_t gemm_ws_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemm(const GemmArgs& a, float* ws, size_t ws_bytes, cudaStream_t s);

constexpr int COLSUM_BLOCKS = 1184;      // 8 blocks of 256 threads on each of the 148 SMs: one full wave
size_t colsum_ws_bytes(int cols);
int launch_colsum(const float* X, int64_t rows, int cols, int64_t ld, int accumulate, float* out, float* ws,
                  cudaStream_t s);
int launch_colsum_final(const float* partial, int nb, int cols, int accumulate, float* out, cudaStream_t s);

// dg[e, :] (+)= inv_deg[dst_e] * sum_a h[src_e, a] * dZ[dst_e, k, a]      (forward CSR order)
int launch_edge_grad(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* dZ,
                     const float* h, int64_t n, int use_mma, float* dg, cudaStream_t s, int dz_bf16 = 0);

// the same sum over all layers at once (tf32 arm): dZ_bf16[l] / h[l] per layer, dg written (not accumulated, no zero fill)
#define FESR_EG_MAX_LAYERS 8
bool edge_grad_layers_supported(const fesr_model_dims& d, int n_layers);
int launch_edge_grad_layers(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* const* dZ_bf16,
                            const float* const* h, int n_layers, int64_t n, float* dg, cudaStream_t s);

// bwd_gemm_mma.cu (tf32 arm): dT' += Z^T dpre (ws: wgrad_mma_ws_bytes) and dZ = dpre T'^T
size_t wgrad_mma_ws_bytes(const fesr_model_dims& d);
// terms 3: fp32 Z and dpre split into tf32 hi + lo in registers (the fp32 arm)
int launch_wgrad_mma(const fesr_model_dims& d, const void* Z, int z_half, const float* dpre, int64_t n, float* dT, float* ws,
                     cudaStream_t s, int terms = 1);
int launch_dz_mma(const fesr_model_dims& d, const float* dpre, const float* tprime, int64_t n, float* dZ, cudaStream_t s);
// gemm_tc.cu: the same product on tcgen05 (TMA-fed, TMEM accumulators); operands tf32-rounded by the caller
bool dz_tc_supported(const fesr_model_dims& d);
// out_bf16: dZ rows as bf16 (half the bytes written here and read by the edge-gradient kernel)
// dpre_lo / tprime_r_lo != NULL: the fp32 arm's three-term product (operands split into tf32 hi + lo by the caller)
int launch_dz_tc(const fesr_model_dims& d, const float* dpre, const float* tprime_r, int64_t n, void* dZ, int out_bf16,
                 cudaStream_t s, const float* dpre_lo = nullptr, const float* tprime_r_lo = nullptr);

// gemm_tc.cu: dT' += Z^T dpre on tcgen05 (both operands MN-major); Z_half: the fp16 Z stash, dpre_half: fp16 rows of dpre S
// (S the layer's power-of-two gradient scale), inv_scale -> 1 / S on the device; ws as for launch_wgrad_mma
bool wgrad_tc_supported(const fesr_model_dims& d);
int launch_wgrad_tc(const fesr_model_dims& d, const void* Z_half, const void* dpre_half, const float* inv_scale, int64_t n,
                    float* dT, float* ws, cudaStream_t s);

// edge_mlp_bwd.cu (tf32 arm, KernelNN shape): the whole backward of the edge-MLP hidden layers in one kernel
bool edge_mlp_bwd_supported(const fesr_model_dims& d);
size_t edge_mlp_bwd_ws_bytes(const fesr_model_dims& d, int64_t E);
int launch_edge_mlp_bwd(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                        const float* dg, const float* g, int64_t E, fesr_param_grads* grads, float* ws, cudaStream_t s);

}  // namespace fesr
