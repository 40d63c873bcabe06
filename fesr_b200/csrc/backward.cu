// fesr_nnconv_backward: gradients of one block-diagonal batch through KernelNN / TEECNet.
//
// Mirrors the forward factorisation (zbuild.cu / gemm_*.cu).  Per layer l (last to first), with
// dpre = dL/d(pre-activation of layer l):
//   dT'    += Z_l^T dpre                         (reduction over nodes, split-K, fixed order)
//   dZ      = dpre T'^T                          -> dg += edge_grad(dZ, h_l)     (forward CSR)
//   Z~      = sum_{e: src=j} g_e (x) dpre[dst_e]/deg[dst_e]  ++ dpre[j]          (zbuild on the
//             REVERSED CSR: the same gather/segmented-sum kernel as the forward)
//   dh_l    = Z~ T~                              (the same node GEMM as the forward, fp32 or
//             tcgen05 tf32, with every wp x wp block of T' transposed)
// then the edge-MLP hidden layers, fc1 / fc2.  Replaces loss.backward() of the reference's train
// step (models/scheduler_gnn.py:407) -- where autograd materialises the [E, w*w] edge matrices
// and their gradients -- for the same parameters, named as in the state_dict.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "backward.cuh"
#include "workspace.cuh"

namespace fesr {

// ------------------------------------------------------------------------------- small kernels
__global__ void inv_deg_kernel(const int32_t* __restrict__ rowptr, int64_t n, float* __restrict__ inv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int deg = rowptr[i + 1] - rowptr[i];
  inv[i] = 1.0f / (float)(deg > 0 ? deg : 1);
}

// dpre = dh * act'(h_next) on the first w columns, 0 on the padding (and on TEECNet's constant column)
// (tf32 arm: rounded to tf32 here -- every consumer rounds it anyway, and the tcgen05 dZ product truncates what it is given)
__global__ void mask_kernel(const float* __restrict__ dh, const float* __restrict__ h_next, int64_t n, int wp, int w,
                            int relu, int round_tf32, float* __restrict__ dpre, float* __restrict__ dpre_hi,
                            float* __restrict__ dpre_lo, const float* __restrict__ in_scale) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * wp) return;
  const int c = (int)(idx % wp);
  float v = 0.f;
  if (c < w) {
    v = dh[idx];
    if (in_scale != nullptr) v *= __ldg(in_scale);      // dh of the scaled fp16 Z~ path: exact power-of-two unscaling
    if (relu && !(h_next[idx] > 0.f)) v = 0.f;
  }
  if (round_tf32) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    v = __uint_as_float(u);
  }
  dpre[idx] = v;
  if (dpre_hi != nullptr) {      // fp32 arm: tf32 hi + lo pair for the three-term tcgen05 dZ product
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    const float hi = __uint_as_float(u);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - hi));
    dpre_hi[idx] = hi;
    dpre_lo[idx] = __uint_as_float(u);
  }
}

// Scaled fp16 Z~ (tf32 arm).  The reversed-graph outer products Z~ = sum g (x) dpre[dst]/deg[dst] ++ dpre and the dh = Z~ T~
// product move 2.7 GB per layer as fp32; the forward of this arm already keeps its Z in fp16 (same 11-bit mantissa as
// tf32).  Gradients live far below fp16's range, so dpre is multiplied by a per-layer power of two S that brings its
// largest magnitude to 2^6 (found on the device: absmax -> S, no host round trip), Z~ and T~ go through the fp16 kernels of
// the forward, and the next layer's mask kernel multiplies dh by 1/S -- both scalings are exact.
__global__ void absmax_kernel(const float* __restrict__ x, int64_t count, unsigned* __restrict__ out) {
  unsigned m = 0u;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(x[i]) & 0x7fffffffu);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(out, m);      // non-negative floats order like their bit patterns
}

// out = x * S with S = 2^floor(log2(target / absmax)); scale[0] = S, scale[1] = 1 / S
__global__ void scale_copy_kernel(const float* __restrict__ x, int64_t count, const unsigned* __restrict__ amax_bits,
                                  float target, float* __restrict__ out, float* __restrict__ scale) {
  const float amax = __uint_as_float(*amax_bits);
  float S = 1.f;
  if (amax > 0.f && amax < 3.0e38f) S = exp2f(fminf(fmaxf(floorf(log2f(target / amax)), -100.f), 100.f));
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx == 0) {
    scale[0] = S;
    scale[1] = 1.0f / S;
  }
  if (idx < count) out[idx] = x[idx] * S;
}

// Fused reversed-graph pass (KernelNN shape): the gathered rows q = dpre S / deg[node] and the own rows dpre S as fp16 rows
// for layer_fused16_kernel in sum mode (Z~ stays on chip, as Z does in the predict arm)
__global__ void scale_rows_f16_kernel(const float* __restrict__ x, int64_t n, int wp, const unsigned* __restrict__ amax_bits,
                                      float target, const float* __restrict__ inv_deg, __half* __restrict__ own,
                                      __half* __restrict__ q, float* __restrict__ scale) {
  const float amax = __uint_as_float(*amax_bits);
  float S = 1.f;
  if (amax > 0.f && amax < 3.0e38f) S = exp2f(fminf(fmaxf(floorf(log2f(target / amax)), -100.f), 100.f));
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // four consecutive elements of one row (wp % 4 == 0)
  if (t == 0) {
    scale[0] = S;
    scale[1] = 1.0f / S;
  }
  const int64_t idx = t * 4;
  if (idx >= n * wp) return;
  const float4 v = *reinterpret_cast<const float4*>(x + idx);
  const float d = __ldg(inv_deg + idx / wp);
  const float a0 = v.x * S, a1 = v.y * S, a2 = v.z * S, a3 = v.w * S;
  __half2 o[2] = {__floats2half2_rn(a0, a1), __floats2half2_rn(a2, a3)};
  __half2 g[2] = {__floats2half2_rn(a0 * d, a1 * d), __floats2half2_rn(a2 * d, a3 * d)};
  *reinterpret_cast<uint2*>(own + idx) = *reinterpret_cast<const uint2*>(o);
  *reinterpret_cast<uint2*>(q + idx) = *reinterpret_cast<const uint2*>(g);
}

// g rows [E, kp] fp32 (reversed-CSR order, slot layout) -> planar fp16 [kp / 16][E][16]; the lo slot (first padding slot)
// holds the constant FESR_LO_SCALE that the fused kernel's two-term weights expect (layer_fused.cu)
// (index != NULL: row e is g[index[e]] -- the forward-order g gathered into reversed order on the way)
__global__ void g3_planar_kernel(const float* __restrict__ g, const int32_t* __restrict__ index, int64_t E, int kp, int lo_slot,
                                 __half* __restrict__ g3) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // one 8-slot chunk of one edge
  const int c8 = kp / 8;
  if (t >= E * c8) return;
  const int64_t e = t / c8;
  const int c = (int)(t - e * c8);
  const float* row = g + (index != nullptr ? (int64_t)__ldg(index + e) : e) * kp + c * 8;
  const float4 v0 = *reinterpret_cast<const float4*>(row), v1 = *reinterpret_cast<const float4*>(row + 4);
  float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  if (lo_slot >= c * 8 && lo_slot < c * 8 + 8) f[lo_slot - c * 8] = FESR_LO_SCALE;
  __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  const int part = (c * 8) / 16, off = (c * 8) % 16;
  *reinterpret_cast<uint4*>(g3 + ((size_t)part * E + e) * 16 + off) = *reinterpret_cast<const uint4*>(h);
}

__global__ void scale_inplace_kernel(float* __restrict__ x, int64_t count, const float* __restrict__ factor) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < count) x[idx] *= __ldg(factor);
}

// mask_kernel + the column sums of dpre (dbias) + its largest magnitude in ONE pass over dh (tf32 arm: the three used to be
// three kernels reading the same 30 MB).  Blocks own row ranges like colsum_partial_kernel; a thread owns four consecutive
// columns (16-byte loads and stores; wp / 4 threads per row, 256 / (wp / 4) row groups) and walks its rows with two
// independent accumulator sets; the row groups are combined in a fixed order, so the partial sums -- finished by
// colsum_final_kernel -- are deterministic; the absmax is order-independent.
__global__ void __launch_bounds__(256)
mask_colsum_kernel(const float* __restrict__ dh, const float* __restrict__ h_next, int64_t n, int wp, int w, int relu,
                   const float* __restrict__ in_scale, int64_t rchunk, float* __restrict__ dpre, float* __restrict__ partial,
                   unsigned* __restrict__ amax) {
  __shared__ float4 red[256];
  __shared__ unsigned mred[8];
  const int cq = wp >> 2, ng = 256 / cq;                 // threads per row, row groups
  const int jq = threadIdx.x % cq, rg = threadIdx.x / cq;
  const float sc = in_scale != nullptr ? __ldg(in_scale) : 1.f;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  unsigned m = 0u;
  if (rg < ng) {
    const int64_t r0 = (int64_t)blockIdx.x * rchunk, r1 = min(n, r0 + rchunk);
    const int c = 4 * jq;
    auto one = [&](int64_t i, float4& acc) {
      const float4 d = *reinterpret_cast<const float4*>(dh + i * wp + c);
      float v[4] = {d.x * sc, d.y * sc, d.z * sc, d.w * sc};
      if (relu) {
        const float4 hn = *reinterpret_cast<const float4*>(h_next + i * wp + c);
        if (!(hn.x > 0.f)) v[0] = 0.f;
        if (!(hn.y > 0.f)) v[1] = 0.f;
        if (!(hn.z > 0.f)) v[2] = 0.f;
        if (!(hn.w > 0.f)) v[3] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (c + q >= w) v[q] = 0.f;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v[q]));
        v[q] = __uint_as_float(u);
        m = max(m, u & 0x7fffffffu);
      }
      *reinterpret_cast<float4*>(dpre + i * wp + c) = make_float4(v[0], v[1], v[2], v[3]);
      acc.x += v[0];
      acc.y += v[1];
      acc.z += v[2];
      acc.w += v[3];
    };
    int64_t i = r0 + rg;
    for (; i + ng < r1; i += 2 * ng) {
      one(i, s0);
      one(i + ng, s1);
    }
    if (i < r1) one(i, s0);
  }
  red[threadIdx.x] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < wp) {
    // column j = 4 jq' + q of every row group, groups in index order (two interleaved chains)
    const int jq2 = threadIdx.x >> 2, q = threadIdx.x & 3;
    float a0 = 0.f, a1 = 0.f;
    int g = 0;
    for (; g + 1 < ng; g += 2) {
      a0 += reinterpret_cast<const float*>(&red[g * cq + jq2])[q];
      a1 += reinterpret_cast<const float*>(&red[(g + 1) * cq + jq2])[q];
    }
    if (g < ng) a0 += reinterpret_cast<const float*>(&red[g * cq + jq2])[q];
    partial[(int64_t)blockIdx.x * wp + threadIdx.x] = a0 + a1;
  }
  if (threadIdx.x == 0 && amax != nullptr) {
    unsigned mm = 0u;
    for (int q = 0; q < 8; ++q) mm = max(mm, mred[q]);
    if (mm != 0u) atomicMax(amax, mm);
  }
}

// out = tf32(in) (and out_lo = tf32(in - out) when asked for)
__global__ void round_tf32_kernel(const float* __restrict__ in, int64_t count, float* __restrict__ out,
                                  float* __restrict__ out_lo) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  uint32_t u;
  const float v = in[idx];
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  const float hi = __uint_as_float(u);
  out[idx] = hi;
  if (out_lo != nullptr) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - hi));
    out_lo[idx] = __uint_as_float(u);
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ index, int64_t rows,
                                   int row_f4, float* __restrict__ dst) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= rows * row_f4) return;
  const int64_t r = idx / row_f4;
  const int c = (int)(idx % row_f4);
  reinterpret_cast<float4*>(dst)[idx] = reinterpret_cast<const float4*>(src)[(int64_t)index[r] * row_f4 + c];
}

__global__ void gather_scalar_kernel(const float* __restrict__ src, const int32_t* __restrict__ index, int64_t m,
                                     float* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < m) dst[i] = src[index ? index[i] : i];
}

__global__ void add_first_cols_kernel(const float* __restrict__ srcv, int w, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w) dst[i] += srcv[i];
}

// dT' [zk, wp] -> gradients of the last edge-MLP layer, kernel.linear and root
__global__ void scatter_tgrad_kernelnn(fesr_model_dims d, const float* __restrict__ dT, float* __restrict__ gw_last,
                                       float* __restrict__ gb_last, float* __restrict__ groot) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int w = d.w, K = d.k1 - 1;
  const int64_t total = (int64_t)(d.k1 + 1) * w * w;     // k in [0,k1): mlp ; k == k1: root
  if (idx >= total) return;
  const int k = (int)(idx / (w * w)), a = (int)((idx / w) % w), b = (int)(idx % w);
  if (k < K) gw_last[(size_t)(a * w + b) * K + k] += dT[(size_t)(k * d.wp + a) * d.wp + b];
  else if (k == K) gb_last[a * w + b] += dT[(size_t)(K * d.wp + a) * d.wp + b];
  else groot[a * w + b] += dT[(size_t)(d.zk_main + a) * d.wp + b];
}

// TEECNet: T''[k,a,b] = sum_a' Laug[a',a] T0[k,a',b]  =>
//   dT0[k,a',b] = sum_a Laug[a',a] dT''[k,a,b] ; dLaug[a',a] = sum_{k,b} T0[k,a',b] dT''[k,a,b]
__global__ void scatter_tgrad_teecnet_t0(fesr_model_dims d, const float* __restrict__ dT, const float* __restrict__ lin_w,
                                         const float* __restrict__ lin_b, float* __restrict__ gw_last,
                                         float* __restrict__ gb_last, float* __restrict__ groot) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int w = d.w, K = d.k1 - 1;
  const int64_t total = (int64_t)(d.k1 + 1) * w * w;
  if (idx >= total) return;
  const int k = (int)(idx / (w * w)), ap = (int)((idx / w) % w), b = (int)(idx % w);
  if (k > K) {
    groot[ap * w + b] += dT[(size_t)(d.zk_main + ap) * d.wp + b];
    return;
  }
  float acc = 0.f;
  for (int a = 0; a <= w; ++a) {
    const float l = (a < w) ? lin_w[ap * w + a] : lin_b[ap];
    acc = fmaf(l, dT[(size_t)(k * d.wp + a) * d.wp + b], acc);
  }
  if (k < K) gw_last[(size_t)(ap * w + b) * K + k] += acc;
  else gb_last[ap * w + b] += acc;
}

__global__ void scatter_tgrad_teecnet_lin(fesr_model_dims d, const float* __restrict__ dT,
                                          const float* __restrict__ w_last, const float* __restrict__ b_last,
                                          float* __restrict__ glin_w, float* __restrict__ glin_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = d.w, K = d.k1 - 1;
  if (idx >= w * (w + 1)) return;
  const int ap = idx / (w + 1), a = idx % (w + 1);
  float acc = 0.f;
  for (int k = 0; k <= K; ++k)
    for (int b = 0; b < w; ++b) {
      const float t0 = (k < K) ? w_last[(size_t)(ap * w + b) * K + k] : b_last[ap * w + b];
      acc = fmaf(t0, dT[(size_t)(k * d.wp + a) * d.wp + b], acc);
    }
  if (a < w) glin_w[ap * w + a] += acc;
  else glin_b[ap] += acc;
}

// ---- edge MLP hidden layers (recomputed with stored activations for the backward)
__device__ __forceinline__ float act_f(float v, int leaky) { return leaky ? (v > 0.f ? v : 0.01f * v) : fmaxf(v, 0.f); }
__device__ __forceinline__ float act_grad(float a, int leaky) { return a > 0.f ? 1.f : (leaky ? 0.01f : 0.f); }

__global__ void mlp_layer0_kernel(const float* __restrict__ dattr, const float* __restrict__ w0,
                                  const float* __restrict__ b0, int64_t E, int D, int leaky, float* __restrict__ a0) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * D) return;
  const int o = (int)(idx % D);
  a0[idx] = act_f(fmaf(dattr[idx / D], w0[o], b0[o]), leaky);
}

__global__ void bias_act_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t E, int D, int leaky) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * D) return;
  a[idx] = act_f(a[idx] + b[idx % D], leaky);
}

// da_last[e, k] = dg[e, off(k)] * act'(a_last[e, k])   (un-permute the padded g layout).  The last hidden layer's
// activations ARE the forward's edge features g (same padded layout; rounding to tf32 keeps the sign the derivative
// depends on), so they are read from there instead of being recomputed with one more [E, K] x [K, K] product.
__global__ void dg_to_dpre_kernel(const float* __restrict__ dg, const float* __restrict__ g, int64_t E, int K,
                                  int kt, int ktp, int kp, int leaky, float* __restrict__ dpre) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * K) return;
  const int64_t e = idx / K;
  const int k = (int)(idx % K);
  const int off = (k / kt) * ktp + (k % kt);
  dpre[idx] = dg[e * kp + off] * act_grad(g[e * kp + off], leaky);
}

__global__ void act_grad_kernel(float* __restrict__ da, const float* __restrict__ a, int64_t count, int leaky) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < count) da[idx] *= act_grad(a[idx], leaky);
}

// ------------------------------------------------------------------------------- workspace
struct BackwardWs {
  float* dh[2];
  float* dpre;
  float* BZ;
  uint16_t* BZx;        // bf16 dZ of layers 2.. when the edge gradient runs once over all layers (layers 0, 1 share BZ)
  float* g_rev;
  float* dg;
  float* dT;
  float* tprime_r;      // tf32-rounded copy of T' [zk, wp] for the tcgen05 dZ product
  float* tprime_r_lo;   // ... and its lo part, dpre_hi / dpre_lo [n, wp]: the fp32 arm's three-term product
  float* dpre_hi;
  float* dpre_lo;
  float* dpre_s;        // dpre * S_l (scaled fp16 Z~ path), amax bit patterns and (S_l, 1 / S_l) per layer
  __half* g3_rev;       // fused reversed pass: planar fp16 g (reversed order), T~ in the fused K order, own / gathered rows
  __half* tfused_t;
  __half* own16;
  __half* q16;
  float* zero_bias;
  float* fused_P;       // fp32 partial sums between the launches of a multi-launch shape (TEECNet)
  unsigned* amax;
  float* scales;
  float* inv_deg;
  float* dbias;
  float* dattr;
  float* act[4];
  float* da[2];
  float* gemm_ws;
  size_t gemm_ws_bytes;
  float* colsum_ws;
  size_t bytes;
};

// A/B switch of the all-layer edge gradient; also decides whether the workspace holds the per-layer bf16 dZ buffers
static bool eg_layers_enabled() {
  static const bool on = !(getenv("FESR_EDGE_GRAD_LAYERS") && atoi(getenv("FESR_EDGE_GRAD_LAYERS")) == 0);
  return on;
}

static BackwardWs carve_backward(void* base, const fesr_model_dims& d, int64_t n, int64_t E) {
  Carver c(base);
  BackwardWs w;
  const size_t nn = (size_t)(n > 0 ? n : 1), ee = (size_t)(E > 0 ? E : 1);
  w.dh[0] = c.take<float>(nn * d.wp);
  w.dh[1] = c.take<float>(nn * d.wp);
  w.dpre = c.take<float>(nn * d.wp);
  w.BZ = c.take<float>(nn * d.zk);
  w.BZx = eg_layers_enabled() && d.layers > 2 && dz_tc_supported(d) && edge_grad_layers_supported(d, d.layers)
              ? c.take<uint16_t>((size_t)(d.layers - 2) * nn * d.zk) : nullptr;
  w.g_rev = c.take<float>(ee * d.kp);
  w.dg = c.take<float>(ee * d.kp);
  w.dT = c.take<float>((size_t)d.zk * d.wp);
  w.tprime_r = c.take<float>((size_t)d.zk * d.wp);
  w.tprime_r_lo = c.take<float>((size_t)d.zk * d.wp);
  w.dpre_hi = c.take<float>(nn * d.wp);
  w.dpre_lo = c.take<float>(nn * d.wp);
  w.dpre_s = c.take<float>(nn * d.wp);
  const bool fz = layer_fused_supported(d) && d.w <= 43 && ((d.kind == FESR_KERNELNN && d.kp == 48) || d.kind == FESR_TEECNET);
  w.g3_rev = fz ? c.take<__half>(ee * d.kp) : nullptr;
  w.tfused_t = fz ? c.take<__half>(layer_fused_tf_elems(d)) : nullptr;
  w.own16 = fz ? c.take<__half>(nn * d.wp) : nullptr;
  w.q16 = fz ? c.take<__half>(nn * d.wp) : nullptr;
  w.zero_bias = c.take<float>(64);
  w.fused_P = fz && d.kind == FESR_TEECNET ? c.take<float>(nn * d.wp) : nullptr;
  w.amax = c.take<unsigned>(64);
  w.scales = c.take<float>(128);
  w.inv_deg = c.take<float>(nn);
  w.dbias = c.take<float>(d.wp);
  w.dattr = c.take<float>(ee);
  int dmax = 1;
  for (int l = 0; l < d.n_hidden; ++l) {
    w.act[l] = c.take<float>(ee * d.hidden[l]);
    if (d.hidden[l] > dmax) dmax = d.hidden[l];
  }
  w.da[0] = c.take<float>(ee * dmax);
  w.da[1] = c.take<float>(ee * dmax);
  size_t g = gemm_ws_bytes(d.zk, d.wp, n);
  if (wgrad_mma_ws_bytes(d) > g) g = wgrad_mma_ws_bytes(d);
  for (int l = 0; l < d.n_hidden; ++l) {
    const size_t b = gemm_ws_bytes(d.hidden[l], l > 0 ? d.hidden[l - 1] : 1, E);
    if (b > g) g = b;
  }
  if (edge_mlp_bwd_ws_bytes(d, E) > g) g = edge_mlp_bwd_ws_bytes(d, E);
  size_t b2 = gemm_ws_bytes(d.out_ch, d.w, n);
  if (b2 > g) g = b2;
  b2 = gemm_ws_bytes(d.w, d.in_ch, n);
  if (b2 > g) g = b2;
  w.gemm_ws_bytes = g + 256;
  w.gemm_ws = c.take<float>(w.gemm_ws_bytes / sizeof(float));
  w.colsum_ws = c.take<float>(colsum_ws_bytes(256) / sizeof(float));
  w.bytes = c.used();
  return w;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

size_t fesr_backward_workspace_bytes(const fesr_model_dims* dims, int64_t n, int64_t E) {
  if (!dims || n < 0 || E < 0) return 0;
  return carve_backward(nullptr, *dims, n, E).bytes;
}

int fesr_nnconv_backward(const fesr_model_dims* dims, const fesr_params* params, const float* x,
                         const int32_t* rowptr, const int32_t* src_sorted, const int32_t* perm,
                         const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rev_to_fwd,
                         const float* edge_attr, int64_t n, int64_t E, int precision, const float* grad_y,
                         const void* forward_workspace, fesr_param_grads* grads, float* grad_x, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(dims && params && grads, "dims/params/grads NULL");
  FESR_CHECK_ARG(n >= 0 && E >= 0 && n < (1ll << 31) && E < (1ll << 31), "n/E out of range");
  FESR_CHECK_ARG(precision == FESR_PREC_FP32 || precision == FESR_PREC_TF32,
                 "backward supports fp32 | tf32 (f16 is a predict-only precision), got %d", precision);
  if (n == 0) return FESR_OK;
  FESR_CHECK_ARG(x && grad_y && rowptr && forward_workspace, "NULL pointer");
  FESR_CHECK_ARG(E == 0 || (src_sorted && rowptr_t && src_t && rev_to_fwd && edge_attr), "NULL edge arrays");
  const fesr_model_dims& d = *dims;
  const fesr_params& p = *params;
  ForwardWs fw = carve_forward(const_cast<void*>(forward_workspace), d, n, E, 1, z_stash_half(precision));
  BackwardWs w = carve_backward(workspace, d, n, E);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("backward workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  cudaStream_t s = as_stream(stream_);
  ProfScope prof(PROF_BACKWARD, s);
  const int T = 256;
  const int L = d.layers;
  const int relu = d.kind == FESR_KERNELNN;
  const int rnd = precision != FESR_PREC_FP32;
  int rc;
#define GEMM(...)                                                     \
  do {                                                                \
    GemmArgs ga__ = {__VA_ARGS__};                                    \
    ga__.tf32 = rnd;                                                  \
    if ((rc = launch_gemm(ga__, w.gemm_ws, w.gemm_ws_bytes, s))) return rc; \
  } while (0)

  inv_deg_kernel<<<(unsigned)ceil_div(n, T), T, 0, s>>>(rowptr, n, w.inv_deg);
  FESR_LAUNCH_CHECK();
  FESR_CUDA(cudaMemsetAsync(w.dT, 0, (size_t)d.zk * d.wp * sizeof(float), s));
  FESR_CUDA(cudaMemsetAsync(w.dbias, 0, (size_t)d.wp * sizeof(float), s));

  // ---- fc2 / fc_out:  y = h_L W2^T + b2
  const float* hL = fw.h[L];
  GEMM(grad_y, 1, d.out_ch, hL, d.wp, 1, grads->fc2_w, d.w, 1, d.out_ch, d.w, n, 1);
  if ((rc = launch_colsum(grad_y, n, d.out_ch, d.out_ch, 1, grads->fc2_b, w.colsum_ws, s))) return rc;
  FESR_CUDA(cudaMemsetAsync(w.dh[0], 0, (size_t)n * d.wp * sizeof(float), s));
  GEMM(grad_y, d.out_ch, 1, p.fc2_w, d.w, 1, w.dh[0], d.wp, 1, n, d.w, d.out_ch, 0);

  // tf32 arm: dZ = dpre T'^T on tcgen05 (FESR_DZ_TC=0: the mma.sync kernel, for A/B measurements)
  static const bool dz_tc_env = !(getenv("FESR_DZ_TC") && atoi(getenv("FESR_DZ_TC")) == 0);
  const bool dz_tc = rnd && dz_tc_env && dz_tc_supported(d) && E > 0;
  const bool dz_tc3 = !rnd && dz_tc_env && dz_tc_supported(d) && E > 0;      // fp32 arm: three-term product
  // fp32 arm: dT' += Z^T dpre as 3xTF32 on mma.sync (FESR_WGRAD_3X=0: the CUDA-core split-K GEMM)
  static const bool wgrad_tc_env = !(getenv("FESR_WGRAD_TC") && atoi(getenv("FESR_WGRAD_TC")) == 0);      // A/B switch
  static const bool wgrad3_env = !(getenv("FESR_WGRAD_3X") && atoi(getenv("FESR_WGRAD_3X")) == 0);
  const bool wgrad3 = !rnd && wgrad3_env && (d.wp == 16 || d.wp == 32 || d.wp == 48 || d.wp == 64);
  // ... written as bf16 (FESR_DZ_BF16=0: fp32), which the edge-gradient MMAs read as exact tf32 operands
  static const bool dz_bf16_env = !(getenv("FESR_DZ_BF16") && atoi(getenv("FESR_DZ_BF16")) == 0);
  const int dz_bf16 = dz_tc && dz_bf16_env && d.wp == 48 ? 1 : 0;
  if (dz_tc || dz_tc3) {
    const int64_t cnt = (int64_t)d.zk * d.wp;
    round_tf32_kernel<<<(unsigned)ceil_div(cnt, T), T, 0, s>>>(fw.prep.tprime, cnt, w.tprime_r, dz_tc3 ? w.tprime_r_lo : nullptr);
    FESR_LAUNCH_CHECK();
  }
  // tf32 arm with the fp16 Z stash: Z~ and dh through the fp16 kernels, dpre scaled per layer (FESR_ZT_HALF=0: fp32 Z~)
  static const bool zt_half_env = !(getenv("FESR_ZT_HALF") && atoi(getenv("FESR_ZT_HALF")) == 0);
  const bool zt_half = rnd && zt_half_env && z_stash_half(precision) && L <= 60 && d.wp % 16 == 0 && d.wp <= 64;
  if (zt_half) FESR_CUDA(cudaMemsetAsync(w.amax, 0, 64 * sizeof(unsigned), s));
  // ... and, for the single-launch KernelNN shape, through the fused layer kernel in sum mode on the reversed CSR: Z~ never
  // reaches HBM (FESR_ZT_FUSED=0: the two fp16 kernels)
  static const bool zt_fused_env = !(getenv("FESR_ZT_FUSED") && atoi(getenv("FESR_ZT_FUSED")) == 0);
  const bool zt_fused = zt_half && zt_fused_env && w.g3_rev != nullptr && E > 0;
  if (zt_fused) {
    FESR_CUDA(cudaMemsetAsync(w.zero_bias, 0, 64 * sizeof(float), s));
    // planar fp16 g in reversed order, gathered straight from the forward's g (the fp32 reversed copy is not needed here)
    g3_planar_kernel<<<(unsigned)ceil_div(E * (d.kp / 8), T), T, 0, s>>>(fw.g, rev_to_fwd, E, d.kp, d.kt, w.g3_rev);
    FESR_LAUNCH_CHECK();
  } else if (E > 0) {
    gather_rows_kernel<<<(unsigned)ceil_div(E * (d.kp / 4), T), T, 0, s>>>(fw.g, rev_to_fwd, E, d.kp / 4, w.g_rev);
    FESR_LAUNCH_CHECK();
  }
  // one edge-gradient pass over all layers at the end (dg written once) when every layer's bf16 dZ can be kept
  const bool eg_layers = eg_layers_enabled() && dz_bf16 && zt_fused && E > 0 && (L <= 2 || w.BZx != nullptr) &&
                         edge_grad_layers_supported(d, L);
  const void* dz_of[FESR_EG_MAX_LAYERS] = {nullptr};
  const float* h_of[FESR_EG_MAX_LAYERS] = {nullptr};
  if (eg_layers) {
    for (int l = 0; l < L; ++l) {
      dz_of[l] = l < 2 ? static_cast<const void*>(reinterpret_cast<uint16_t*>(w.BZ) + (size_t)l * n * d.zk)
                       : static_cast<const void*>(w.BZx + (size_t)(l - 2) * n * d.zk);
      h_of[l] = fw.h[l];
    }
  } else if (E > 0) {
    FESR_CUDA(cudaMemsetAsync(w.dg, 0, (size_t)E * d.kp * sizeof(float), s));
  }
  if (zt_fused) {
    // T~ in the fused K order; the constant-1 slot carries T~'s own constant rows (no centring of g in this arm)
    if ((rc = launch_prepare_tfused(d, fw.prep.ttilde, fw.prep.ttilde + (size_t)(d.k1 - 1) * d.wp * d.wp, w.tfused_t, s))) return rc;
  }
  const float* in_scale = nullptr;      // 1 / S of the layer processed before (dh arrives scaled by S)
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const bool mask_fused = rnd && d.wp <= 64 && d.wp % 4 == 0 && n > 0;      // tf32 arm: mask + dbias partial sums + absmax in one pass
    if (mask_fused) {
      int nb = (int)(n < COLSUM_BLOCKS ? n : COLSUM_BLOCKS);
      const int64_t rchunk = ceil_div(n, nb);
      nb = (int)ceil_div(n, rchunk);
      mask_colsum_kernel<<<nb, 256, 0, s>>>(w.dh[cur], fw.h[l + 1], n, d.wp, d.w, relu, in_scale, rchunk, w.dpre, w.colsum_ws,
                                            zt_half ? w.amax + l : nullptr);
      FESR_LAUNCH_CHECK();
      if ((rc = launch_colsum_final(w.colsum_ws, nb, d.wp, 1, w.dbias, s))) return rc;
    } else {
      mask_kernel<<<(unsigned)ceil_div(n * d.wp, T), T, 0, s>>>(w.dh[cur], fw.h[l + 1], n, d.wp, d.w, relu, rnd, w.dpre,
                                                                dz_tc3 ? w.dpre_hi : nullptr, dz_tc3 ? w.dpre_lo : nullptr, in_scale);
      FESR_LAUNCH_CHECK();
      if ((rc = launch_colsum(w.dpre, n, d.wp, d.wp, 1, w.dbias, w.colsum_ws, s))) return rc;
    }
    // dT' += Z_l^T dpre
    const bool wgrad_tc = wgrad_tc_env && zt_fused && mask_fused && z_stash_half(precision) && wgrad_tc_supported(d);
    if (wgrad_tc) {
      // the scaled fp16 rows of the fused reversed pass below are this product's operand too
      scale_rows_f16_kernel<<<(unsigned)ceil_div(n * d.wp / 4, T), T, 0, s>>>(w.dpre, n, d.wp, w.amax + l, 64.f, w.inv_deg, w.own16,
                                                                          w.q16, w.scales + 2 * l);
      FESR_LAUNCH_CHECK();
      if ((rc = launch_wgrad_tc(d, fw.Z[l], w.own16, w.scales + 2 * l + 1, n, w.dT, w.gemm_ws, s))) return rc;
    } else if (rnd) {
      if ((rc = launch_wgrad_mma(d, fw.Z[l], z_stash_half(precision), w.dpre, n, w.dT, w.gemm_ws, s))) return rc;
    } else if (wgrad3) {
      if ((rc = launch_wgrad_mma(d, fw.Z[l], 0, w.dpre, n, w.dT, w.gemm_ws, s, 3))) return rc;
    } else {
      GEMM(fw.Z[l], 1, d.zk, w.dpre, d.wp, 1, w.dT, d.wp, 1, d.zk, d.wp, n, 1);
    }
    if (E > 0) {
      // dZ = dpre T'^T ; dg += edge_grad(dZ, h_l)
      if (rnd && dz_tc) {
        if ((rc = launch_dz_tc(d, w.dpre, w.tprime_r, n, eg_layers ? const_cast<void*>(dz_of[l]) : w.BZ, dz_bf16, s))) return rc;
      } else if (rnd) {
        if ((rc = launch_dz_mma(d, w.dpre, fw.prep.tprime, n, w.BZ, s))) return rc;
      } else if (dz_tc3) {
        if ((rc = launch_dz_tc(d, w.dpre_hi, w.tprime_r, n, w.BZ, 0, s, w.dpre_lo, w.tprime_r_lo))) return rc;
      } else {
        GEMM(w.dpre, d.wp, 1, fw.prep.tprime, 1, d.wp, w.BZ, d.zk, 1, n, d.zk, d.wp, 0);
      }
      // (use_mma: 1 = tf32 arm, 3 = the fp32 arm's 3xTF32 form where the tensor-core kernel covers the shape)
      if (!eg_layers && (rc = launch_edge_grad(d, rowptr, src_sorted, w.BZ, fw.h[l], n, rnd ? 1 : 3, w.dg, s, dz_bf16))) return rc;
    }
    // dh_l = [sum over out-edges of g (x) dpre[dst]/deg[dst]  ++  dpre] T~
    if (zt_half) {
      const int64_t cnt = n * d.wp;
      if (!mask_fused) {
        absmax_kernel<<<(unsigned)(ceil_div(cnt, 256 * 8) < 4096 ? ceil_div(cnt, 256 * 8) : 4096), 256, 0, s>>>(w.dpre, cnt, w.amax + l);
        FESR_LAUNCH_CHECK();
      }
      if (zt_fused) {
        if (!wgrad_tc) {
          scale_rows_f16_kernel<<<(unsigned)ceil_div(cnt / 4, T), T, 0, s>>>(w.dpre, n, d.wp, w.amax + l, 64.f, w.inv_deg, w.own16,
                                                                         w.q16, w.scales + 2 * l);
          FESR_LAUNCH_CHECK();
        }
        rc = launch_layer_fused_f16(d, rowptr_t, src_t, w.g3_rev, E, w.q16, n, w.tfused_t, w.zero_bias, w.fused_P, w.dh[cur ^ 1],
                                    d.kind == FESR_KERNELNN ? 3 : 0, s,
                                    /*out_f32=*/1, /*sum_mode=*/1, w.own16);
        if (rc) return rc;
        in_scale = w.scales + 2 * l + 1;
        cur ^= 1;
        continue;
      }
      scale_copy_kernel<<<(unsigned)ceil_div(cnt, T), T, 0, s>>>(w.dpre, cnt, w.amax + l, 64.f, w.dpre_s, w.scales + 2 * l);
      FESR_LAUNCH_CHECK();
      rc = launch_zbuild_mma(d, rowptr_t, src_t, w.g_rev, w.dpre_s, n, w.BZ, 2, s, /*mean=*/0, w.inv_deg);
      if (rc) return rc;
      rc = launch_node_gemm_f16(d, fw.prep.ttilde_t_h, nullptr, EPI_NONE, w.BZ, n, w.dh[cur ^ 1], s);
      if (rc) return rc;
      in_scale = w.scales + 2 * l + 1;
      cur ^= 1;
      continue;
    }
    in_scale = nullptr;
    if (rnd)
      rc = launch_zbuild_mma(d, rowptr_t, src_t, w.g_rev, w.dpre, n, w.BZ, 1, s, /*mean=*/0, w.inv_deg);
    else
      rc = launch_zbuild(d, rowptr_t, src_t, w.g_rev, w.dpre, n, w.BZ, 0, s, /*mean=*/0, w.inv_deg);
    if (rc) return rc;
    // fp32 arm: 3xTF32 on tcgen05 where the forward uses it (gemm_tc.cu TERMS == 3), the CUDA-core GEMM otherwise
    static const bool fp32_simt = getenv("FESR_FP32_SIMT") && atoi(getenv("FESR_FP32_SIMT")) != 0;
    if (precision == FESR_PREC_FP32 && (fp32_simt || d.zk > 4096 || !(d.wp % 16 == 0 && d.wp <= 64)))
      rc = launch_node_gemm_fp32(d, fw.prep.ttilde, nullptr, EPI_NONE, w.BZ, n, w.dh[cur ^ 1], s);
    else if (precision == FESR_PREC_FP32)
      rc = launch_node_gemm_tf32(d, fw.prep.ttilde_t, nullptr, EPI_NONE, w.BZ, n, w.dh[cur ^ 1], s, 0, fw.prep.ttilde_t_lo, 3);
    else
      rc = launch_node_gemm_tf32(d, fw.prep.ttilde_t, nullptr, EPI_NONE, w.BZ, n, w.dh[cur ^ 1], s);
    if (rc) return rc;
    cur ^= 1;
  }
  if (eg_layers && (rc = launch_edge_grad_layers(d, rowptr, src_sorted, dz_of, h_of, L, n, w.dg, s))) return rc;

  // ---- conv bias / root / last edge-MLP layer / kernel.linear from dT'
  add_first_cols_kernel<<<1, 64, 0, s>>>(w.dbias, d.w, grads->bias);
  FESR_LAUNCH_CHECK();
  {
    const int last = d.n_hidden;
    const int64_t total = (int64_t)(d.k1 + 1) * d.w * d.w;
    if (d.kind == FESR_KERNELNN) {
      scatter_tgrad_kernelnn<<<(unsigned)ceil_div(total, T), T, 0, s>>>(d, w.dT, grads->mlp_w[last],
                                                                       grads->mlp_b[last], grads->root);
      FESR_LAUNCH_CHECK();
    } else {
      scatter_tgrad_teecnet_t0<<<(unsigned)ceil_div(total, T), T, 0, s>>>(d, w.dT, p.lin_w, p.lin_b,
                                                                         grads->mlp_w[last], grads->mlp_b[last],
                                                                         grads->root);
      FESR_LAUNCH_CHECK();
      scatter_tgrad_teecnet_lin<<<(unsigned)ceil_div(d.w * (d.w + 1), 64), 64, 0, s>>>(
          d, w.dT, p.mlp_w[last], p.mlp_b[last], grads->lin_w, grads->lin_b);
      FESR_LAUNCH_CHECK();
    }
  }

  // ---- fc1:  h_0 = x W1^T + b1
  if (in_scale != nullptr) {      // dh of layer 0 is still scaled
    scale_inplace_kernel<<<(unsigned)ceil_div(n * d.wp, T), T, 0, s>>>(w.dh[cur], n * d.wp, in_scale);
    FESR_LAUNCH_CHECK();
  }
  const float* dh0 = w.dh[cur];
  GEMM(dh0, 1, d.wp, x, d.in_ch, 1, grads->fc1_w, d.in_ch, 1, d.w, d.in_ch, n, 1);
  if ((rc = launch_colsum(dh0, n, d.w, d.wp, 1, grads->fc1_b, w.colsum_ws, s))) return rc;
  if (grad_x) GEMM(dh0, d.wp, 1, p.fc1_w, d.in_ch, 1, grad_x, d.in_ch, 1, n, d.in_ch, d.w, 0);

  // ---- edge MLP hidden layers
  static const bool no_fused_mlp = getenv("FESR_EDGE_MLP_BWD_GENERIC") != nullptr;       // A/B switch for profiling
  if (E > 0 && rnd && !no_fused_mlp && edge_mlp_bwd_supported(d)) {
    if ((rc = launch_edge_mlp_bwd(d, p, edge_attr, perm, w.dg, fw.g, E, grads, w.gemm_ws, s))) return rc;
  } else if (E > 0) {
    const int nh = d.n_hidden, leaky = d.leaky;
    gather_scalar_kernel<<<(unsigned)ceil_div(E, T), T, 0, s>>>(edge_attr, perm, E, w.dattr);
    FESR_LAUNCH_CHECK();
    mlp_layer0_kernel<<<(unsigned)ceil_div(E * d.hidden[0], T), T, 0, s>>>(w.dattr, p.mlp_w[0], p.mlp_b[0], E,
                                                                         d.hidden[0], leaky, w.act[0]);
    FESR_LAUNCH_CHECK();
    for (int l = 1; l < nh - 1; ++l) {      // (the last hidden layer's activations are g: not recomputed)
      const int din = d.hidden[l - 1], dout = d.hidden[l];
      GEMM(w.act[l - 1], din, 1, p.mlp_w[l], 1, din, w.act[l], dout, 1, E, dout, din, 0);
      bias_act_kernel<<<(unsigned)ceil_div(E * dout, T), T, 0, s>>>(w.act[l], p.mlp_b[l], E, dout, leaky);
      FESR_LAUNCH_CHECK();
    }
    const int K = d.hidden[nh - 1];
    dg_to_dpre_kernel<<<(unsigned)ceil_div(E * K, T), T, 0, s>>>(w.dg, fw.g, E, K, d.kt, d.ktp, d.kp, leaky, w.da[0]);
    FESR_LAUNCH_CHECK();
    int dc = 0;
    for (int l = nh - 1; l >= 0; --l) {
      const int dout = d.hidden[l], din = l > 0 ? d.hidden[l - 1] : 1;
      const float* prev = l > 0 ? w.act[l - 1] : w.dattr;
      GEMM(w.da[dc], 1, dout, prev, din, 1, grads->mlp_w[l], din, 1, dout, din, E, 1);
      if ((rc = launch_colsum(w.da[dc], E, dout, dout, 1, grads->mlp_b[l], w.colsum_ws, s))) return rc;
      if (l > 0) {
        GEMM(w.da[dc], dout, 1, p.mlp_w[l], din, 1, w.da[dc ^ 1], din, 1, E, din, dout, 0);
        act_grad_kernel<<<(unsigned)ceil_div(E * din, T), T, 0, s>>>(w.da[dc ^ 1], w.act[l - 1], E * (int64_t)din, leaky);
        FESR_LAUNCH_CHECK();
        dc ^= 1;
      }
    }
  }
#undef GEMM
  return FESR_OK;
}

}  // extern "C"
