// The two node-level GEMMs of the backward pass on warp-level tensor cores (tf32 arm):
//   W:  dT'[zk, wp] += Z^T[zk, n] . dpre[n, wp]     reduction over nodes; both operands have the
//       contraction index as their SLOW memory index ("MN-major"), so fragments are gathered
//       from padded shared-memory tiles; split over node ranges with a fixed-order reduce.
//   Z:  dZ[n, zk]    = dpre[n, wp] . T'^T[wp, zk]    thin K (= wp), bound by writing dZ.
// Both are HBM-bound (24 flop/B): mma.sync.m16n8k8 tf32 leaves the tensor pipe far from
// saturated, and unlike tcgen05 it reads MN-major operands without a transposing copy.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "backward.cuh"

namespace fesr {

__device__ __forceinline__ uint32_t bg_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void bg_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bg_cp16(void* dst_smem, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

// ---------------------------------------------------------------------------- W: dT' partials
constexpr int WG_BM = 128;     // zk columns per CTA (8 warps x 16)
constexpr int WG_BK = 32;      // nodes per stage
constexpr int WG_THREADS = 256;

// ZT = float, or __half for the fp16 Z stash of the tf32 arm (fp16 -> fp32 is exact and already tf32-representable)
// TERMS == 3: the fp32 arm -- fp32 Z and dpre are split into tf32 hi + lo in registers, lo.hi + hi.lo + hi.hi (3xTF32)
template <int WP, typename ZT, int TERMS>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_mma_kernel(const ZT* __restrict__ Z, const float* __restrict__ dpre, int64_t n, int zk, int64_t nchunk,
                 float* __restrict__ partial) {
  constexpr int NT = WP / 8;
  constexpr int SZ = WG_BM + 8, SD = WP + 8;
  constexpr int ZV = 16 / (int)sizeof(ZT);   // Z elements per 16-byte copy
  extern __shared__ __align__(16) float smem[];
  ZT* Zs = reinterpret_cast<ZT*>(smem);      // [2][BK][SZ]  (the float-sized region is kept for both element types)
  float* Ds = smem + 2 * WG_BK * SZ;         // [2][BK][SD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.x * WG_BM;
  const int64_t k_begin = (int64_t)blockIdx.y * nchunk, k_end = min(n, k_begin + nchunk);
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[nt][r] = 0.f;

  auto load_stage = [&](int st, int64_t k0) {
    ZT* zs = Zs + st * WG_BK * SZ;
    float* ds = Ds + st * WG_BK * SD;
    for (int t = tid; t < WG_BK * (WG_BM / ZV); t += WG_THREADS) {
      const int r = t / (WG_BM / ZV), c = (t % (WG_BM / ZV)) * ZV;
      ZT* dst = zs + r * SZ + c;
      if (k0 + r < k_end && m0 + c < zk) bg_cp16(dst, Z + (k0 + r) * (int64_t)zk + m0 + c);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int t = tid; t < WG_BK * (WP / 4); t += WG_THREADS) {
      const int r = t / (WP / 4), c = (t % (WP / 4)) * 4;
      float* dst = ds + r * SD + c;
      if (k0 + r < k_end) bg_cp16(dst, dpre + (k0 + r) * (int64_t)WP + c);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int st = 0;
  if (k_begin < k_end) load_stage(0, k_begin);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += WG_BK) {
    const bool more = k0 + WG_BK < k_end;
    if (more) load_stage(st ^ 1, k0 + WG_BK);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const ZT* zs = Zs + st * WG_BK * SZ + warp * 16;
    const float* ds = Ds + st * WG_BK * SD;
#pragma unroll
    for (int ks = 0; ks < WG_BK / 8; ++ks) {
      const int r0 = ks * 8 + tq, r1 = r0 + 4;
      const float av[4] = {(float)zs[r0 * SZ + gq], (float)zs[r0 * SZ + gq + 8], (float)zs[r1 * SZ + gq], (float)zs[r1 * SZ + gq + 8]};
      uint32_t a[4], alo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = bg_tf32(av[q]);
        if (TERMS == 3) alo[q] = bg_tf32(av[q] - __uint_as_float(a[q]));
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float b0f = ds[r0 * SD + nt * 8 + gq], b1f = ds[r1 * SD + nt * 8 + gq];
        const uint32_t b0 = bg_tf32(b0f), b1 = bg_tf32(b1f);
        if (TERMS == 3) {
          bg_mma(acc[nt], alo, b0, b1);
          bg_mma(acc[nt], a, bg_tf32(b0f - __uint_as_float(b0)), bg_tf32(b1f - __uint_as_float(b1)));
        }
        bg_mma(acc[nt], a, b0, b1);
      }
    }
    __syncthreads();
    st ^= 1;
  }
  // partial[blockIdx.y][zk][WP]
  float* out = partial + (int64_t)blockIdx.y * zk * WP;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int m = m0 + warp * 16 + gq + 8 * hh;
      if (m < zk)
        *reinterpret_cast<float2*>(out + (int64_t)m * WP + nt * 8 + 2 * tq) = make_float2(acc[nt][2 * hh], acc[nt][2 * hh + 1]);
    }
}

// The tf32 arm's hot shape (fp16 Z stash, tf32-rounded dpre): the kernel above is ISSUE-bound (ncu: issue slots 70 % active,
// 38 instructions per 6 MMAs -- scalar 16-bit fragment loads, a cvt per operand).  Here a warp owns 32 zk rows (two m-tiles
// that share every dpre fragment), the MMA's rows gq / gq + 8 are the ADJACENT zk columns 2 gq / 2 gq + 1 (one 32-bit load
// per fragment pair), and dpre needs no conversion: 24 instructions per 12 MMAs.
constexpr int WG2_BM = 256;    // zk columns per CTA (8 warps x 32)
template <int WP>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_mma2_kernel(const __half* __restrict__ Z, const float* __restrict__ dpre, int64_t n, int zk, int64_t nchunk,
                  float* __restrict__ partial) {
  constexpr int NT = WP / 8;
  constexpr int SZ = WG2_BM + 16, SD = WP + 8;       // row strides (halfs / floats): conflict-free fragment loads
  extern __shared__ __align__(16) float smem[];
  float* Ds = smem;                                              // [2][BK][SD]
  __half* Zs = reinterpret_cast<__half*>(smem + 2 * WG_BK * SD); // [2][BK][SZ]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.x * WG2_BM;
  const int64_t k_begin = (int64_t)blockIdx.y * nchunk, k_end = min(n, k_begin + nchunk);
  float acc[2][NT][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[j][nt][r] = 0.f;

  auto load_stage = [&](int st, int64_t k0) {
    __half* zs = Zs + st * WG_BK * SZ;
    float* ds = Ds + st * WG_BK * SD;
    for (int t = tid; t < WG_BK * (WG2_BM / 8); t += WG_THREADS) {
      const int r = t / (WG2_BM / 8), c = (t % (WG2_BM / 8)) * 8;
      __half* dst = zs + r * SZ + c;
      if (k0 + r < k_end && m0 + c < zk) bg_cp16(dst, Z + (k0 + r) * (int64_t)zk + m0 + c);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int t = tid; t < WG_BK * (WP / 4); t += WG_THREADS) {
      const int r = t / (WP / 4), c = (t % (WP / 4)) * 4;
      float* dst = ds + r * SD + c;
      if (k0 + r < k_end) bg_cp16(dst, dpre + (k0 + r) * (int64_t)WP + c);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int st = 0;
  if (k_begin < k_end) load_stage(0, k_begin);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += WG_BK) {
    const bool more = k0 + WG_BK < k_end;
    if (more) load_stage(st ^ 1, k0 + WG_BK);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const __half* zs = Zs + st * WG_BK * SZ + warp * 32 + 2 * gq;
    const float* ds = Ds + st * WG_BK * SD + gq;
#pragma unroll
    for (int ks = 0; ks < WG_BK / 8; ++ks) {
      const int r0 = ks * 8 + tq, r1 = r0 + 4;
      uint32_t a[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(zs + r0 * SZ + j * 16));   // rows gq, gq + 8 at node r0
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(zs + r1 * SZ + j * 16));   // ... at node r1
        a[j][0] = __float_as_uint(lo.x);
        a[j][1] = __float_as_uint(lo.y);
        a[j][2] = __float_as_uint(hi.x);
        a[j][3] = __float_as_uint(hi.y);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t b0 = __float_as_uint(ds[r0 * SD + nt * 8]), b1 = __float_as_uint(ds[r1 * SD + nt * 8]);
        bg_mma(acc[0][nt], a[0], b0, b1);
        bg_mma(acc[1][nt], a[1], b0, b1);
      }
    }
    __syncthreads();
    st ^= 1;
  }
  // partial[blockIdx.y][zk][WP]; MMA row gq <-> zk column 2 gq, row gq + 8 <-> 2 gq + 1 of the m-tile
  float* out = partial + (int64_t)blockIdx.y * zk * WP;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = m0 + warp * 32 + j * 16 + 2 * gq + hh;
        if (m < zk)
          *reinterpret_cast<float2*>(out + (int64_t)m * WP + nt * 8 + 2 * tq) = make_float2(acc[j][nt][2 * hh], acc[j][nt][2 * hh + 1]);
      }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int ks, int64_t count, float* __restrict__ dT) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < ks; ++z) s += partial[(int64_t)z * count + i];    // fixed order
  dT[i] += s;
}

int wgrad_mma_splits(int zk) {
  const int mtiles = (zk + WG_BM - 1) / WG_BM;      // (the 256-column kernel has half the m-tiles: twice the node ranges, same workspace bound)
  // 6 CTAs per SM's worth of node ranges (4 are resident at 49 KB of shared memory each): with 2 the two 16 KB stages
  // in flight per CTA do not cover the HBM latency (train step 18.7 -> 18.3 ms)
  int ks = (6 * num_sms() + mtiles - 1) / mtiles;
  return ks < 1 ? 1 : ks;
}

size_t wgrad_mma_ws_bytes(const fesr_model_dims& d) { return (size_t)wgrad_mma_splits(d.zk) * d.zk * d.wp * sizeof(float); }

int launch_wgrad_mma(const fesr_model_dims& d, const void* Z, int z_half, const float* dpre, int64_t n, float* dT, float* ws,
                     cudaStream_t s, int terms) {
  FESR_CHECK_ARG(terms == 1 || (terms == 3 && !z_half), "3xTF32 weight gradient takes the fp32 Z stash");
  if (n == 0) return FESR_OK;
  static const bool wide_env = !(getenv("FESR_WGRAD_WIDE") && atoi(getenv("FESR_WGRAD_WIDE")) == 0);      // A/B switch
  if (z_half && terms == 1 && d.wp == 48 && wide_env) {
    // dpre is tf32-rounded by the mask kernel in this arm (backward.cu): its bits are the operand.
    // One wave: 3 CTAs are resident per SM (69 registers x 256 threads), so at most 3 x SMs CTAs in all.
    int ks = (3 * num_sms()) / (int)ceil_div(d.zk, WG2_BM);
    if (ks > wgrad_mma_splits(d.zk)) ks = wgrad_mma_splits(d.zk);       // the workspace is sized for that many partials
    if (ks < 1) ks = 1;
    const int64_t nchunk = ceil_div(ceil_div(n, ks), WG_BK) * WG_BK;
    const int ks_eff = (int)ceil_div(n, nchunk);
    constexpr size_t smem = (size_t)2 * WG_BK * (48 + 8) * sizeof(float) + (size_t)2 * WG_BK * (WG2_BM + 16) * sizeof(__half);
    static bool attr = false;
    if (!attr) {
      FESR_CUDA(cudaFuncSetAttribute(wgrad_mma2_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    dim3 grid((unsigned)ceil_div(d.zk, WG2_BM), (unsigned)ks_eff);
    wgrad_mma2_kernel<48><<<grid, WG_THREADS, smem, s>>>(static_cast<const __half*>(Z), dpre, n, d.zk, nchunk, ws);
    FESR_LAUNCH_CHECK();
    const int64_t count = (int64_t)d.zk * d.wp;
    wgrad_reduce_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, s>>>(ws, ks_eff, count, dT);
    FESR_LAUNCH_CHECK();
    return FESR_OK;
  }
  const int ks = wgrad_mma_splits(d.zk);
  const int64_t nchunk = ceil_div(ceil_div(n, ks), WG_BK) * WG_BK;
  const int ks_eff = (int)ceil_div(n, nchunk);
  dim3 grid((unsigned)ceil_div(d.zk, WG_BM), (unsigned)ks_eff);
#define FESR_WG(WPV)                                                                                        \
  do {                                                                                                      \
    constexpr size_t smem = (size_t)2 * WG_BK * (WG_BM + 8 + WPV + 8) * sizeof(float);                      \
    static bool attr = false;                                                                               \
    if (!attr) {                                                                                            \
      FESR_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<WPV, float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      FESR_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<WPV, float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      FESR_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<WPV, __half, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = true;                                                                                          \
    }                                                                                                       \
    if (z_half)                                                                                             \
      wgrad_mma_kernel<WPV, __half, 1><<<grid, WG_THREADS, smem, s>>>(static_cast<const __half*>(Z), dpre, n, d.zk, nchunk, ws); \
    else if (terms == 3)                                                                                    \
      wgrad_mma_kernel<WPV, float, 3><<<grid, WG_THREADS, smem, s>>>(static_cast<const float*>(Z), dpre, n, d.zk, nchunk, ws); \
    else                                                                                                    \
      wgrad_mma_kernel<WPV, float, 1><<<grid, WG_THREADS, smem, s>>>(static_cast<const float*>(Z), dpre, n, d.zk, nchunk, ws);   \
  } while (0)
  switch (d.wp) {
    case 16: FESR_WG(16); break;
    case 32: FESR_WG(32); break;
    case 48: FESR_WG(48); break;
    case 64: FESR_WG(64); break;
    default: set_error("unsupported padded width %d", d.wp); return FESR_EINVAL;
  }
#undef FESR_WG
  FESR_LAUNCH_CHECK();
  const int64_t count = (int64_t)d.zk * d.wp;
  wgrad_reduce_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, s>>>(ws, ks_eff, count, dT);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// ---------------------------------------------------------------------------- Z: dZ = dpre T'^T
constexpr int ZG_BM = 128, ZG_BN = 128, ZG_THREADS = 256;

template <int WP>
__global__ void __launch_bounds__(ZG_THREADS)
dz_mma_kernel(const float* __restrict__ dpre, const float* __restrict__ tprime, int64_t n, int zk,
              float* __restrict__ dZ) {
  constexpr int KS = WP / 8;
  constexpr int S = WP + 4;                  // row stride: conflict-free fragment loads
  extern __shared__ __align__(16) float zg_smem[];
  float* As = zg_smem;                       // dpre tile  [node][k]
  float* Bs = zg_smem + ZG_BM * S;           // T' tile    [zk col][k]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t m0 = (int64_t)blockIdx.x * ZG_BM;
  const int n0 = blockIdx.y * ZG_BN;
  for (int t = tid; t < ZG_BM * (WP / 4); t += ZG_THREADS) {
    const int r = t / (WP / 4), c = (t % (WP / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < n) v = *reinterpret_cast<const float4*>(dpre + (m0 + r) * WP + c);
    As[r * S + c + 0] = __uint_as_float(bg_tf32(v.x));
    As[r * S + c + 1] = __uint_as_float(bg_tf32(v.y));
    As[r * S + c + 2] = __uint_as_float(bg_tf32(v.z));
    As[r * S + c + 3] = __uint_as_float(bg_tf32(v.w));
  }
  for (int t = tid; t < ZG_BN * (WP / 4); t += ZG_THREADS) {
    const int r = t / (WP / 4), c = (t % (WP / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + r < zk) v = *reinterpret_cast<const float4*>(tprime + (int64_t)(n0 + r) * WP + c);
    Bs[r * S + c + 0] = __uint_as_float(bg_tf32(v.x));
    Bs[r * S + c + 1] = __uint_as_float(bg_tf32(v.y));
    Bs[r * S + c + 2] = __uint_as_float(bg_tf32(v.z));
    Bs[r * S + c + 3] = __uint_as_float(bg_tf32(v.w));
  }
  __syncthreads();
  // warp -> 16 node rows x 128 columns (16 n-tiles)
  float acc[ZG_BN / 8][4];
#pragma unroll
  for (int nt = 0; nt < ZG_BN / 8; ++nt)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[nt][r] = 0.f;
  const float* as = As + (warp * 16) * S;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4];
    a[0] = __float_as_uint(as[gq * S + ks * 8 + tq]);
    a[1] = __float_as_uint(as[(gq + 8) * S + ks * 8 + tq]);
    a[2] = __float_as_uint(as[gq * S + ks * 8 + tq + 4]);
    a[3] = __float_as_uint(as[(gq + 8) * S + ks * 8 + tq + 4]);
#pragma unroll
    for (int nt = 0; nt < ZG_BN / 8; ++nt) {
      const uint32_t b0 = __float_as_uint(Bs[(nt * 8 + gq) * S + ks * 8 + tq]);
      const uint32_t b1 = __float_as_uint(Bs[(nt * 8 + gq) * S + ks * 8 + tq + 4]);
      bg_mma(acc[nt], a, b0, b1);
    }
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int64_t row = m0 + warp * 16 + gq + 8 * hh;
    if (row >= n) continue;
    float* out = dZ + row * (int64_t)zk + n0 + 2 * tq;
#pragma unroll
    for (int nt = 0; nt < ZG_BN / 8; ++nt)
      if (n0 + nt * 8 < zk) *reinterpret_cast<float2*>(out + nt * 8) = make_float2(acc[nt][2 * hh], acc[nt][2 * hh + 1]);
  }
}

int launch_dz_mma(const fesr_model_dims& d, const float* dpre, const float* tprime, int64_t n, float* dZ, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  dim3 grid((unsigned)ceil_div(n, ZG_BM), (unsigned)ceil_div(d.zk, ZG_BN));
#define FESR_ZG(WPV)                                                                                               \
  do {                                                                                                             \
    constexpr size_t smem = (size_t)(ZG_BM + ZG_BN) * (WPV + 4) * sizeof(float);                                   \
    static bool attr = false;                                                                                      \
    if (!attr) {                                                                                                   \
      FESR_CUDA(cudaFuncSetAttribute(dz_mma_kernel<WPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = true;                                                                                                 \
    }                                                                                                              \
    dz_mma_kernel<WPV><<<grid, ZG_THREADS, smem, s>>>(dpre, tprime, n, d.zk, dZ);                                  \
  } while (0)
  switch (d.wp) {
    case 16: FESR_ZG(16); break;
    case 32: FESR_ZG(32); break;
    case 48: FESR_ZG(48); break;
    case 64: FESR_ZG(64); break;
    default: set_error("unsupported padded width %d", d.wp); return FESR_EINVAL;
  }
#undef FESR_ZG
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // namespace fesr
