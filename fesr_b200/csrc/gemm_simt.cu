// fp32 CUDA-core version of the node contraction  h' = act(Z x T' + bias).
// This is the FESR_PREC_FP32 arm (rel-L2 <= 1e-5 against the reference's fp32 CPU result);
// the tensor-core arm lives in gemm_tc.cu.  Applies the last Linear layer of the edge MLP and
// the root weight in one GEMM (reference models/model.py:528-529 + :533-535, reordered).
#include "kernels.cuh"

namespace fesr {

constexpr int SG_BM = 128;
constexpr int SG_BK = 32;
constexpr int SG_THREADS = 256;

template <int WP>
__global__ void __launch_bounds__(SG_THREADS, 2)
node_gemm_fp32_kernel(const float* __restrict__ Z, const float* __restrict__ tprime, const float* __restrict__ bias_p,
                      int64_t n, int zk, int w, int epi, float* __restrict__ h_out) {
  constexpr int TN = WP / 8;
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][WP];
  const int tid = threadIdx.x;
  const int rg = tid >> 3, cg = tid & 7;      // 32 row groups x 8 column groups
  const int64_t row0 = (int64_t)blockIdx.x * SG_BM;
  float acc[4][TN];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < TN; ++y) acc[x][y] = 0.f;

  for (int k0 = 0; k0 < zk; k0 += SG_BK) {
    // A tile: 128 x 32 floats, transposed into As[k][row]
#pragma unroll
    for (int it = 0; it < (SG_BM * SG_BK / 4) / SG_THREADS; ++it) {
      const int idx = tid + it * SG_THREADS;
      const int r = idx >> 3, k4 = (idx & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < n) v = __ldg(reinterpret_cast<const float4*>(Z + (row0 + r) * zk + k0 + k4));
      As[k4 + 0][r] = v.x;
      As[k4 + 1][r] = v.y;
      As[k4 + 2][r] = v.z;
      As[k4 + 3][r] = v.w;
    }
    for (int idx = tid; idx < SG_BK * WP / 4; idx += SG_THREADS) {
      const int kk = idx / (WP / 4), c4 = (idx % (WP / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[kk][c4]) =
          __ldg(reinterpret_cast<const float4*>(tprime + (size_t)(k0 + kk) * WP + c4));
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][rg * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      float b[TN];
#pragma unroll
      for (int y = 0; y < TN; y += 2) {
        const float2 t = *reinterpret_cast<const float2*>(&Bs[kk][cg * TN + y]);
        b[y] = t.x;
        b[y + 1] = t.y;
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < TN; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int64_t r = row0 + rg * 4 + x;
    if (r >= n) continue;
#pragma unroll
    for (int y = 0; y < TN; y += 2) {
      const int c = cg * TN + y;
      float v0 = acc[x][y], v1 = acc[x][y + 1];
      if (epi != EPI_NONE) {
        v0 += bias_p[c];
        v1 += bias_p[c + 1];
      }
      if (epi == EPI_BIAS_CONST1) {  // no activation between layers (models/model.py:280-282); keep h[:, w] = 1
        if (c == w) v0 = 1.f;
        if (c + 1 == w) v1 = 1.f;
      } else if (epi == EPI_BIAS_RELU) {  // F.relu(conv1(...))   models/model.py:559
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      *reinterpret_cast<float2*>(h_out + r * WP + c) = make_float2(v0, v1);
    }
  }
}

int launch_node_gemm_fp32(const fesr_model_dims& d, const float* B, const float* bias_p, int epi, const float* Z,
                          int64_t n, float* h_out, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  const unsigned grid = (unsigned)ceil_div(n, SG_BM);
  ProfScope prof(PROF_NODE_GEMM, s);
  switch (d.wp) {
    case 16: node_gemm_fp32_kernel<16><<<grid, SG_THREADS, 0, s>>>(Z, B, bias_p, n, d.zk, d.w, epi, h_out); break;
    case 32: node_gemm_fp32_kernel<32><<<grid, SG_THREADS, 0, s>>>(Z, B, bias_p, n, d.zk, d.w, epi, h_out); break;
    case 48: node_gemm_fp32_kernel<48><<<grid, SG_THREADS, 0, s>>>(Z, B, bias_p, n, d.zk, d.w, epi, h_out); break;
    case 64: node_gemm_fp32_kernel<64><<<grid, SG_THREADS, 0, s>>>(Z, B, bias_p, n, d.zk, d.w, epi, h_out); break;
    default: set_error("unsupported padded width %d", d.wp); return FESR_EINVAL;
  }
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // namespace fesr
