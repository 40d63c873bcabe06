// Gather + deterministic segmented mean over the destination CSR.
//
//   Z[i, k*wp + a] = 1/max(deg_i,1) * sum_{e: dst_e = i} g_e[k] * h[src_e, a]
//   Z[i, zk_main + a] = h[i, a]                                  (root block)
//
// This replaces PyG's index_select(x, edge_index[0]) + per-edge [w,w] mat-vec + atomic
// scatter_add_ mean (reference models/model.py:525-529 via MessagePassing.propagate): the
// edge-conditioned matrix A_e = reshape(W3 g_e + b3) is never formed; because the message is
// bilinear in (g_e, h_src) the sum over a node's incoming edges commutes with W3, so the
// per-edge work is one outer product and W3 is applied once per NODE by the Z x T' GEMM.
//
// One warp owns one destination node.  The node's g rows are contiguous in CSR order
// (streamed, 128-bit loads); the h[src] rows are gathered with 128-bit loads into a per-warp
// shared-memory slab; each lane then accumulates a KT x AT register tile of the outer product
// (lane = 8*q + ag: channel group q of 4, column group ag of 8) in a fixed edge order, so the
// fp32 result is reproducible run to run.
#include "kernels.cuh"

namespace fesr {

constexpr int ZB_WARPS = 8;
constexpr int ZB_DEGC = 16;   // edges staged per chunk

template <int AT>
__device__ __forceinline__ void load_cols(const float* p, float (&v)[AT]) {
  if constexpr (AT % 4 == 0) {
#pragma unroll
    for (int i = 0; i < AT / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < AT / 2; ++i) {
      const float2 t = *reinterpret_cast<const float2*>(p + 2 * i);
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  }
}

template <int AT>
__device__ __forceinline__ void store_cols(float* p, const float (&v)[AT]) {
  if constexpr (AT % 4 == 0) {
#pragma unroll
    for (int i = 0; i < AT / 4; ++i)
      *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < AT / 2; ++i) *reinterpret_cast<float2*>(p + 2 * i) = make_float2(v[2 * i], v[2 * i + 1]);
  }
}

template <int KT, int WP>
__global__ void __launch_bounds__(ZB_WARPS * 32, 2)
zbuild_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
              const float* __restrict__ g, const float* __restrict__ h, int64_t n, int passes, int kp,
              int zk_main, int zk, float* __restrict__ Z) {
  constexpr int KTP = (KT + 3) / 4 * 4;
  constexpr int AT = WP / 8;
  constexpr int GROW = 4 * KTP;              // floats of g staged per edge per pass
  constexpr int SLAB = ZB_DEGC * (GROW + WP);
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sg = smem + warp * SLAB;            // [DEGC][GROW]
  float* sh = sg + ZB_DEGC * GROW;           // [DEGC][WP]
  const int q = lane >> 3, ag = lane & 7;

  const int64_t warp_global = (int64_t)blockIdx.x * ZB_WARPS + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * ZB_WARPS;
  for (int64_t i = warp_global; i < n; i += warp_stride) {
    const int e_begin = rowptr[i], e_end = rowptr[i + 1];
    const int deg = e_end - e_begin;
    const float inv = 1.0f / (float)(deg > 0 ? deg : 1);
    float* zrow = Z + i * (int64_t)zk;
    for (int p = 0; p < passes; ++p) {
      float acc[KT][AT];
#pragma unroll
      for (int k = 0; k < KT; ++k)
#pragma unroll
        for (int t = 0; t < AT; ++t) acc[k][t] = 0.f;
      for (int c0 = e_begin; c0 < e_end; c0 += ZB_DEGC) {
        const int m = min(ZB_DEGC, e_end - c0);
        // stage g rows (contiguous in CSR order) and gathered h rows, 128-bit
        for (int t = lane; t < m * (GROW / 4); t += 32) {
          const int j = t / (GROW / 4), c = t % (GROW / 4);
          const float4 v = __ldg(reinterpret_cast<const float4*>(g + (int64_t)(c0 + j) * kp + p * GROW) + c);
          *reinterpret_cast<float4*>(sg + j * GROW + 4 * c) = v;
        }
        for (int t = lane; t < m * (WP / 4); t += 32) {
          const int j = t / (WP / 4), c = t % (WP / 4);
          const int s = __ldg(src_sorted + c0 + j);
          const float4 v = __ldg(reinterpret_cast<const float4*>(h + (int64_t)s * WP) + c);
          *reinterpret_cast<float4*>(sh + j * WP + 4 * c) = v;
        }
        __syncwarp();
        for (int j = 0; j < m; ++j) {
          float gq[KTP];
          load_cols<KTP>(sg + j * GROW + q * KTP, gq);
          float ha[AT];
          load_cols<AT>(sh + j * WP + ag * AT, ha);
#pragma unroll
          for (int k = 0; k < KT; ++k)
#pragma unroll
            for (int t = 0; t < AT; ++t) acc[k][t] = fmaf(gq[k], ha[t], acc[k][t]);
        }
        __syncwarp();
      }
      // channel k = (p*4 + q)*KT + kt  ->  columns [k*WP + ag*AT, +AT)
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        float o[AT];
#pragma unroll
        for (int t = 0; t < AT; ++t) o[t] = acc[k][t] * inv;
        store_cols<AT>(zrow + ((p * 4 + q) * KT + k) * WP + ag * AT, o);
      }
    }
    // root block + zero tail
    for (int c = lane; c < zk - zk_main; c += 32) zrow[zk_main + c] = (c < WP) ? h[i * WP + c] : 0.f;
  }
}

template <int KT, int WP>
static int launch_zb(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                     const float* h, int64_t n, float* Z, cudaStream_t s) {
  constexpr int KTP = (KT + 3) / 4 * 4;
  constexpr size_t smem = (size_t)ZB_WARPS * ZB_DEGC * (4 * KTP + WP) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(zbuild_kernel<KT, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t blocks_needed = ceil_div(n, ZB_WARPS);
  const int64_t cap = (int64_t)num_sms() * 2 * 8;      // a few waves of persistent-ish CTAs
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  ProfScope prof(PROF_ZBUILD, s);
  zbuild_kernel<KT, WP><<<grid, ZB_WARPS * 32, smem, s>>>(rowptr, src_sorted, g, h, n, d.passes, d.kp, d.zk_main,
                                                         d.zk, Z);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <int KT>
static int dispatch_wp(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                       const float* h, int64_t n, float* Z, cudaStream_t s) {
  switch (d.wp) {
    case 16: return launch_zb<KT, 16>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 32: return launch_zb<KT, 32>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 48: return launch_zb<KT, 48>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 64: return launch_zb<KT, 64>(d, rowptr, src_sorted, g, h, n, Z, s);
  }
  set_error("unsupported padded width %d", d.wp);
  return FESR_EINVAL;
}

int launch_zbuild(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                  const float* h, int64_t n, float* Z, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  switch (d.kt) {
    case 4: return dispatch_wp<4>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 8: return dispatch_wp<8>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 11: return dispatch_wp<11>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 13: return dispatch_wp<13>(d, rowptr, src_sorted, g, h, n, Z, s);
  }
  set_error("unsupported kt %d", d.kt);
  return FESR_EINVAL;
}

}  // namespace fesr
