// Backward of the outer-product aggregation with respect to the edge features g:
//   dg[e, k] += 1/deg(dst_e) * sum_a h[src_e, a] * dZ[dst_e, k*wp + a]
// One warp per destination node holds dZ_i in registers in exactly the lane tiling zbuild
// writes Z with (lane = 8*q + ag), loops over the node's incoming edges in CSR order, reduces the
// partial dot products over the 8 column groups with shuffles and the ag == 0 lanes accumulate
// into dg (one writer per element => deterministic).
#include "backward.cuh"

namespace fesr {

constexpr int EG_WARPS = 8;

template <int KT, int WP>
__global__ void __launch_bounds__(EG_WARPS * 32, 2)
edge_grad_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                 const float* __restrict__ dZ, const float* __restrict__ h, int64_t n, int passes, int kp, int zk,
                 float* __restrict__ dg) {
  constexpr int KTP = (KT + 3) / 4 * 4;
  constexpr int AT = WP / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 3, ag = lane & 7;
  const int64_t warp_global = (int64_t)blockIdx.x * EG_WARPS + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * EG_WARPS;
  for (int64_t i = warp_global; i < n; i += warp_stride) {
    const int eb = rowptr[i], ee = rowptr[i + 1];
    if (ee == eb) continue;
    const float inv = 1.0f / (float)(ee - eb);
    const float* zrow = dZ + i * (int64_t)zk;
    for (int p = 0; p < passes; ++p) {
      float dz[KT][AT];
#pragma unroll
      for (int k = 0; k < KT; ++k)
#pragma unroll
        for (int t = 0; t < AT; ++t) dz[k][t] = zrow[((p * 4 + q) * KT + k) * WP + ag * AT + t] * inv;
      for (int e = eb; e < ee; ++e) {
        const int64_t s = src_sorted[e];
        float ha[AT];
        if (AT % 2 == 0) {            // the lane's AT consecutive floats as 8-byte loads (ag * AT floats is 8-byte aligned)
          const float2* hp = reinterpret_cast<const float2*>(h + s * WP + ag * AT);
#pragma unroll
          for (int t = 0; t < AT / 2; ++t) {
            const float2 v = __ldg(hp + t);
            ha[2 * t] = v.x;
            ha[2 * t + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int t = 0; t < AT; ++t) ha[t] = __ldg(h + s * WP + ag * AT + t);
        }
        float part[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          float v = 0.f;
#pragma unroll
          for (int t = 0; t < AT; ++t) v = fmaf(dz[k][t], ha[t], v);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          part[k] = v;
        }
        if (ag == 0) {
          // the lane group's KTP slots of the dg row are one 16-byte-aligned run: read-modify-write as float4
          // (the pad slots get + 0); 2 * KTP / 4 transactions per edge and group instead of 2 * KT scalar ones
          float4* out = reinterpret_cast<float4*>(dg + (int64_t)e * kp + (p * 4 + q) * KTP);
#pragma unroll
          for (int k4 = 0; k4 < KTP / 4; ++k4) {
            float4 v = out[k4];
            v.x += part[4 * k4];
            if (4 * k4 + 1 < KT) v.y += part[4 * k4 + 1];
            if (4 * k4 + 2 < KT) v.z += part[4 * k4 + 2];
            if (4 * k4 + 3 < KT) v.w += part[4 * k4 + 3];
            out[k4] = v;
          }
        }
      }
    }
  }
}

template <int KT>
static int eg_dispatch(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src, const float* dZ,
                       const float* h, int64_t n, float* dg, cudaStream_t s) {
  const int64_t blocks = ceil_div(n, EG_WARPS);
  const int grid = (int)(blocks < 16ll * num_sms() ? blocks : 16ll * num_sms());
  switch (d.wp) {
    case 16: edge_grad_kernel<KT, 16><<<grid, EG_WARPS * 32, 0, s>>>(rowptr, src, dZ, h, n, d.passes, d.kp, d.zk, dg); break;
    case 32: edge_grad_kernel<KT, 32><<<grid, EG_WARPS * 32, 0, s>>>(rowptr, src, dZ, h, n, d.passes, d.kp, d.zk, dg); break;
    case 48: edge_grad_kernel<KT, 48><<<grid, EG_WARPS * 32, 0, s>>>(rowptr, src, dZ, h, n, d.passes, d.kp, d.zk, dg); break;
    case 64: edge_grad_kernel<KT, 64><<<grid, EG_WARPS * 32, 0, s>>>(rowptr, src, dZ, h, n, d.passes, d.kp, d.zk, dg); break;
    default: set_error("unsupported padded width %d", d.wp); return FESR_EINVAL;
  }
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

int launch_edge_grad(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* dZ,
                     const float* h, int64_t n, int accumulate, float* dg, cudaStream_t s) {
  (void)accumulate;   // dg is zero-initialised by the caller and always accumulated into
  if (n == 0) return FESR_OK;
  ProfScope prof(PROF_BACKWARD, s);
  switch (d.kt) {
    case 4: return eg_dispatch<4>(d, rowptr, src_sorted, dZ, h, n, dg, s);
    case 8: return eg_dispatch<8>(d, rowptr, src_sorted, dZ, h, n, dg, s);
    case 11: return eg_dispatch<11>(d, rowptr, src_sorted, dZ, h, n, dg, s);
    case 13: return eg_dispatch<13>(d, rowptr, src_sorted, dZ, h, n, dg, s);
  }
  set_error("unsupported kt %d", d.kt);
  return FESR_EINVAL;
}

}  // namespace fesr
