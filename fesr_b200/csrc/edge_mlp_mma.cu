// Hidden layers of KernelNN's edge MLP (Linear(1,w) act Linear(w,w) act, reference
// models/model.py:550 + :311-315) with the [32 edges, w] x [w, w] product on warp-level tensor
// cores.  fp32 arm: 3xTF32 (a = a_hi + a_lo, W = W_hi + W_lo, the three significant partial
// products accumulated in fp32; error ~2^-21, inside the 1e-5 gate).  Reduced-precision arms:
// the leading product only (g is rounded to 11 bits on store anyway).
// One warp owns 32 consecutive CSR edges: layer 0 is evaluated straight into MMA A-fragments
// (no shared-memory round trip).  The COLUMNS of W are permuted at load time into the padded /
// channel-grouped g row layout (slot = group*ktp + r; pad slots get zero weights, the constant-1
// channel gets zero weights and bias 1), so the accumulator fragments are g rows already: bias,
// activation, rounding, one shared-memory transpose for coalesced 128-bit stores.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace fesr {

__device__ __forceinline__ uint32_t em_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void em_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float em_act(float v, int leaky) { return leaky ? (v > 0.f ? v : 0.01f * v) : fmaxf(v, 0.f); }

// WPAD: padded input width (K of the GEMM); NTO: output n-tiles = kp / 8; TERMS: 1 or 3; OMODE: 0 fp32,
// 1 fp32 rounded to tf32, 2 fp16, 3 fp16 planar [KP/16][E][16] (the fused layer kernel's slot groups)
template <int WPAD, int NTO, int TERMS, int OMODE>
__global__ void __launch_bounds__(128)
edge_hidden2_mma_kernel(const float* __restrict__ w0g, const float* __restrict__ b0g, const float* __restrict__ w1g,
                        const float* __restrict__ b1g, int w, int leaky, int kt, int ktp, int k1,
                        const float* __restrict__ edge_attr, const int32_t* __restrict__ perm, int64_t E,
                        void* __restrict__ gv, int* ovf, const float* __restrict__ gcenter) {
  F16Guard guard;
  constexpr int KS = WPAD / 8, KP = NTO * 8, SB = KP + 8;
  constexpr int SST = KP + 4;                                          // stage row stride (floats)
  __shared__ __align__(16) float w0[WPAD], b0[WPAD], b1p[KP], gcs[KP];
  __shared__ __align__(16) uint32_t whi[WPAD][SB];                     // [in][slot], tf32 bit patterns
  __shared__ __align__(16) uint32_t wlo[TERMS == 3 ? WPAD : 1][SB];
  extern __shared__ __align__(16) float stage_dyn[];                   // [4 warps][32 edges][SST]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  for (int i = tid; i < WPAD; i += blockDim.x) {
    w0[i] = i < w ? w0g[i] : 0.f;
    b0[i] = i < w ? b0g[i] : 0.f;
  }
  for (int slot = tid; slot < KP; slot += blockDim.x) {
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    b1p[slot] = (r < kt && ch < k1 - 1 && ch < w) ? b1g[ch] : ((r < kt && ch == k1 - 1) ? 1.f : 0.f);
    gcs[slot] = (OMODE == 3 && gcenter != nullptr) ? gcenter[slot] : 0.f;
  }
  for (int i = tid; i < WPAD * KP; i += blockDim.x) {
    const int in = i / KP, slot = i % KP;
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    const float v = (in < w && r < kt && ch < k1 - 1 && ch < w) ? w1g[ch * w + in] : 0.f;
    const uint32_t hi = em_tf32(v);
    whi[in][slot] = hi;
    if (TERMS == 3) wlo[in][slot] = em_tf32(v - __uint_as_float(hi));
  }
  __syncthreads();

  float (*wst)[SST] = reinterpret_cast<float (*)[SST]>(stage_dyn + (size_t)warp * 32 * SST);
  const int64_t n_groups = (E + 31) / 32;
  for (int64_t grp = (int64_t)blockIdx.x * 4 + warp; grp < n_groups; grp += (int64_t)gridDim.x * 4) {
    const int64_t e_base = grp * 32;
    const int64_t e = e_base + lane;
    const float d_lane = (e < E) ? edge_attr[perm ? perm[e] : e] : 0.f;
    float dr[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      dr[mt][0] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq);
      dr[mt][1] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq + 8);
    }
    float acc[2][NTO][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int i0 = ks * 8 + tq, i1 = i0 + 4;
      const float wa = w0[i0], ba = b0[i0], wb = w0[i1], bb = b0[i1];
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        // A fragment of layer-0 activations: rows = edges gq / gq+8, cols = inputs i0 / i1
        const float v[4] = {em_act(fmaf(dr[mt][0], wa, ba), leaky), em_act(fmaf(dr[mt][1], wa, ba), leaky),
                            em_act(fmaf(dr[mt][0], wb, bb), leaky), em_act(fmaf(dr[mt][1], wb, bb), leaky)};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          ahi[mt][r] = em_tf32(v[r]);
          if (TERMS == 3) alo[mt][r] = em_tf32(v[r] - __uint_as_float(ahi[mt][r]));
        }
      }
      if (TERMS == 3) {      // the two correction terms first, each pass over distinct accumulators
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt) {
          const uint32_t h0 = whi[i0][nt * 8 + gq], h1 = whi[i1][nt * 8 + gq];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) em_mma(acc[mt][nt], alo[mt], h0, h1);
        }
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt) {
          const uint32_t l0 = wlo[i0][nt * 8 + gq], l1 = wlo[i1][nt * 8 + gq];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) em_mma(acc[mt][nt], ahi[mt], l0, l1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NTO; ++nt) {
        const uint32_t h0 = whi[i0][nt * 8 + gq], h1 = whi[i1][nt * 8 + gq];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) em_mma(acc[mt][nt], ahi[mt], h0, h1);
      }
    }
    // epilogue: accumulator columns ARE g-row slots
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) {
      const int c = nt * 8 + 2 * tq;
      const float bz0 = b1p[c], bz1 = b1p[c + 1];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v0 = em_act(acc[mt][nt][2 * hh] + bz0, leaky), v1 = em_act(acc[mt][nt][2 * hh + 1] + bz1, leaky);
          if (OMODE == 3) {
            v0 -= gcs[c];
            v1 -= gcs[c + 1];
          }
          if (OMODE == 1) {
            v0 = __uint_as_float(em_tf32(v0));
            v1 = __uint_as_float(em_tf32(v1));
          }
          *reinterpret_cast<float2*>(&wst[mt * 16 + gq + 8 * hh][c]) = make_float2(v0, v1);
        }
    }
    __syncwarp();
    if (OMODE >= 2) {                            // fp16 rows: 8 halfs = 16 bytes per store
      __half* gh = static_cast<__half*>(gv);
      constexpr int Q8 = KP / 8;
      for (int t = lane; t < 32 * Q8; t += 32) {
        const int r = t / Q8, c8 = t - r * Q8;
        if (e_base + r < E) {
          const float4 lo = *reinterpret_cast<const float4*>(&wst[r][8 * c8]);
          const float4 hi = *reinterpret_cast<const float4*>(&wst[r][8 * c8 + 4]);
          guard.note(lo.x, lo.y);
          guard.note(lo.z, lo.w);
          guard.note(hi.x, hi.y);
          guard.note(hi.z, hi.w);
          __half2 p0 = __floats2half2_rn(lo.x, lo.y), p1 = __floats2half2_rn(lo.z, lo.w);
          __half2 p2 = __floats2half2_rn(hi.x, hi.y), p3 = __floats2half2_rn(hi.z, hi.w);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0);
          pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2);
          pk.w = *reinterpret_cast<uint32_t*>(&p3);
          if (OMODE == 3) *reinterpret_cast<uint4*>(gh + ((int64_t)(c8 >> 1) * E + e_base + r) * 16 + (c8 & 1) * 8) = pk;
          else *reinterpret_cast<uint4*>(gh + (e_base + r) * KP + 8 * c8) = pk;
        }
      }
    } else {
      float* gf = static_cast<float*>(gv);
      constexpr int Q4 = KP / 4;
      for (int t = lane; t < 32 * Q4; t += 32) {
        const int r = t / Q4, c4 = t - r * Q4;
        if (e_base + r < E)
          *reinterpret_cast<float4*>(gf + (e_base + r) * KP + 4 * c4) = *reinterpret_cast<const float4*>(&wst[r][4 * c4]);
      }
    }
    __syncwarp();
  }
  if (OMODE >= 2) guard.flush(ovf);
}


// fp16 arms (OMODE 2: fp16 rows [E][KP]; 3: fp16 planar [KP/16][E][16], the fused layer kernel's slot groups):
// the same product on mma.sync.m16n8k16 (fp16 operands -- 11-bit mantissa like the tf32 leading term -- fp32
// accumulate): half the MMAs, weights as B fragments through ldmatrix from a [slot][in] fp16 copy, and the
// accumulator tiles leave through stmatrix into row-per-edge staging (no fp32 round trip).
__device__ __forceinline__ uint32_t em_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void em_mma16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t em_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t em_pack(float lo, float hi, F16Guard& guard) {     // range-checked (common.cuh)
  guard.note(lo, hi);
  return em_pack(lo, hi);
}

template <int WPAD, int NTO, int OMODE>     // ReLU only (KernelNN); LeakyReLU shapes use the tf32 kernel above
__global__ void __launch_bounds__(128, 6)
edge_hidden2_f16_kernel(const float* __restrict__ w0g, const float* __restrict__ b0g, const float* __restrict__ w1g,
                        const float* __restrict__ b1g, int w, int kt, int ktp, int k1,
                        const float* __restrict__ edge_attr, const int32_t* __restrict__ perm, int E,
                        __half* __restrict__ gh, int* ovf, const float* __restrict__ gcenter) {
  F16Guard guard;
  constexpr int KS = WPAD / 16, KP = NTO * 8;
  constexpr int WST = WPAD + 8, SST = KP + 8;                          // row strides in halfs (+16 B: conflict-free)
  __shared__ __align__(16) float w0[WPAD], b0[WPAD], b1p[KP], gcs[KP];  // gcs: per-slot centre (common.cuh), 0 without
  __shared__ __align__(16) __half wsm[KP][WST];                        // [slot][in]: K contiguous
  __shared__ __align__(16) __half stage[4][32][SST];                   // [warp][edge][slot]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3, lr = lane & 7, lm = lane >> 3;
  for (int i = tid; i < WPAD; i += blockDim.x) {
    w0[i] = i < w ? w0g[i] : 0.f;
    b0[i] = i < w ? b0g[i] : 0.f;
  }
  for (int slot = tid; slot < KP; slot += blockDim.x) {
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    b1p[slot] = (r < kt && ch < k1 - 1 && ch < w) ? b1g[ch] : ((r < kt && ch == k1 - 1) ? 1.f : 0.f);
    gcs[slot] = gcenter != nullptr ? gcenter[slot] : 0.f;
  }
  for (int i = tid; i < KP * WPAD; i += blockDim.x) {
    const int slot = i / WPAD, in = i % WPAD;
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    wsm[slot][in] = __float2half_rn((in < w && r < kt && ch < k1 - 1 && ch < w) ? w1g[ch * w + in] : 0.f);
  }
  __syncthreads();
  // ldmatrix rows of the B fragments: matrices (n-tile 2np, k lo), (2np, k hi), (2np + 1, k lo), (2np + 1, k hi)
  const uint32_t b_addr = em_smem(&wsm[(lm >> 1) * 8 + lr][(lm & 1) * 8]);
  // stmatrix rows of the C tiles of m-tile mt: matrices (hh 0, nt), (hh 1, nt), (hh 0, nt + 1), (hh 1, nt + 1)
  const uint32_t s_addr = em_smem(&stage[warp][(lm & 1) * 8 + lr][(lm >> 1) * 8]);

  // coalesced row stores: lanes 0..23 move 4 rows x 6 sixteen-byte chunks per round
  constexpr int Q8 = KP / 8, RPR = 32 / Q8;                            // chunks per row, rows per round
  const int sr = lane / Q8, sc8 = lane % Q8;
  const bool st_lane = lane < RPR * Q8;
  __half* const sdst = OMODE == 3 ? gh + (int64_t)(sc8 >> 1) * E * 16 + (sc8 & 1) * 8 : gh + 8 * sc8;
  constexpr int RSTR = OMODE == 3 ? 16 : KP;                           // row stride of the destination (halfs)
  const int n_groups = (E + 31) / 32;
  auto load_d = [&](int g) {      // edge length of this lane's edge of group g (clamped: the value of a missing edge is unused)
    const int e = min(g * 32 + lane, E - 1);
    return __ldg(edge_attr + (perm ? __ldg(perm + e) : e));
  };
  float d_next = load_d(min((int)(blockIdx.x * 4 + warp), n_groups - 1));
  for (int grp = blockIdx.x * 4 + warp; grp < n_groups; grp += gridDim.x * 4) {
    const int e_base = grp * 32;
    const float d_lane = d_next;
    d_next = load_d(min(grp + (int)gridDim.x * 4, n_groups - 1));     // next group's lengths fly under this group's math
    float dr[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      dr[mt][0] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq);
      dr[mt][1] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq + 8);
    }
    float acc[2][NTO][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      // layer 0 straight into A fragments: rows gq / gq + 8 (edges), columns 2tq, 2tq + 1 (+ 8) of this k-step
      uint32_t a[2][4];
      const int c = ks * 16 + 2 * tq;
      const float2 wl = *reinterpret_cast<const float2*>(&w0[c]), bl = *reinterpret_cast<const float2*>(&b0[c]);
      const float2 wh = *reinterpret_cast<const float2*>(&w0[c + 8]), bh = *reinterpret_cast<const float2*>(&b0[c + 8]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float d = dr[mt][hh];
          a[mt][hh] = em_pack(fmaxf(fmaf(d, wl.x, bl.x), 0.f), fmaxf(fmaf(d, wl.y, bl.y), 0.f), guard);
          a[mt][2 + hh] = em_pack(fmaxf(fmaf(d, wh.x, bh.x), 0.f), fmaxf(fmaf(d, wh.y, bh.y), 0.f), guard);
        }
#pragma unroll
      for (int np = 0; np < NTO / 2; ++np) {
        uint32_t b[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                     : "r"(b_addr + (uint32_t)((np * 16 * WST + ks * 16) * 2)));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          em_mma16(acc[mt][2 * np], a[mt], b[0], b[1]);
          em_mma16(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
        }
      }
    }
    // epilogue: + bias, activation, fp16, 8x8 tiles transposed into rows through stmatrix
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int np = 0; np < NTO / 2; ++np) {
        uint32_t h[2][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int nt = 2 * np + q;
          const float2 bz = *reinterpret_cast<const float2*>(&b1p[nt * 8 + 2 * tq]);
          const float2 gc = *reinterpret_cast<const float2*>(&gcs[nt * 8 + 2 * tq]);
          h[q][0] = em_pack(fmaxf(acc[mt][nt][0] + bz.x, 0.f) - gc.x, fmaxf(acc[mt][nt][1] + bz.y, 0.f) - gc.y, guard);
          h[q][1] = em_pack(fmaxf(acc[mt][nt][2] + bz.x, 0.f) - gc.x, fmaxf(acc[mt][nt][3] + bz.y, 0.f) - gc.y, guard);
        }
        asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(s_addr + (uint32_t)((mt * 16 * SST + np * 16) * 2)),
                     "r"(h[0][0]), "r"(h[0][1]), "r"(h[1][0]), "r"(h[1][1])
                     : "memory");
      }
    __syncwarp();
    if (st_lane) {
#pragma unroll
      for (int r = sr; r < 32; r += RPR)
        if (e_base + r < E)
          *reinterpret_cast<uint4*>(sdst + (int64_t)(e_base + r) * RSTR) = *reinterpret_cast<const uint4*>(&stage[warp][r][8 * sc8]);
    }
    __syncwarp();
  }
  guard.flush(ovf);
}

// TEECNet's edge MLP (DenseNet([1, 32, 64, 128, w*w], LeakyReLU), reference models/model.py:403 + :311-315): the
// three hidden layers 1 -> 32 -> 64 -> 128 for the reduced-precision arms.  One warp owns 32 consecutive CSR edges and
// the activations never leave its registers: layer 0 is evaluated straight into A fragments, the fp32 accumulator
// tiles of layer 1 (two adjacent n-tiles = one k-step) are re-packed into the A fragments of layer 2, and layer 2's
// 144 output slots (128 channels + the constant-1 channel in the padded / channel-grouped g row layout, the columns
// of W2 permuted at load time) are produced in three chunks of 48 so that the accumulators stay at 48 registers.
// Arithmetic: mma.sync.m16n8k16 on SPLIT fp16 operands, a = a_hi + a_lo and W = W_hi + W_lo with the three
// significant products accumulated in fp32 (the fp16 twin of 3xTF32, ~2^-21).  Plain fp16 operands are not enough
// here: the shipped TEECNet is sensitive to the hidden layers (single-term products move the predicted field by
// 1.3e-2 rel-L2, rounding W1 alone by 1.27e-2; the split form by 4e-5 -- measured on the CPU restatement).
// OMODE 1: fp32 rows rounded to tf32 [E][144]; OMODE 2: fp16 rows [E][144]; OMODE 3: fp16 planar [9][E][16] (the
// fused layer kernel's slot groups).
__device__ __forceinline__ void em_split(float v0, float v1, uint32_t& hi, uint32_t& lo, F16Guard& guard) {
  hi = em_pack(v0, v1, guard);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = em_pack(v0 - f.x, v1 - f.y);
}
__device__ __forceinline__ void em_ldsm4(uint32_t addr, uint32_t (&b)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
               : "r"(addr));
}

template <int OMODE>
__global__ void __launch_bounds__(128)
edge_hidden3_mma_kernel(const float* __restrict__ w0g, const float* __restrict__ b0g, const float* __restrict__ w1g,
                        const float* __restrict__ b1g, const float* __restrict__ w2g, const float* __restrict__ b2g,
                        int kt, int ktp, int k1, const float* __restrict__ edge_attr,
                        const int32_t* __restrict__ perm, int E, void* __restrict__ gv, int* ovf,
                        const float* __restrict__ gcenter) {
  F16Guard guard;
  constexpr int H0 = 32, H1 = 64, KP = 144;
  constexpr int CW = 48, NCH = KP / CW, NTC = CW / 8;                  // slots per chunk, chunks, n-tiles per chunk
  constexpr int KS0 = H0 / 16, KS1 = H1 / 16, NT1 = H1 / 8;
  constexpr int W1S = H0 + 8, W2S = H1 + 8;                            // row strides in halfs (+16 B: conflict-free)
  constexpr int SST = OMODE >= 2 ? (CW + 8) * 2 : (CW + 4) * 4;        // staged row stride in BYTES
  __shared__ __align__(16) float w0[H0], b0[H0], b1[H1], b2p[KP], gcs[KP];
  extern __shared__ __align__(16) uint8_t eh3_dyn[];
  __half (*w1h)[W1S] = reinterpret_cast<__half (*)[W1S]>(eh3_dyn);     // [out][in]: K contiguous
  __half (*w1l)[W1S] = w1h + H1;
  __half (*w2h)[W2S] = reinterpret_cast<__half (*)[W2S]>(w1l + H1);    // [slot][in]
  __half (*w2l)[W2S] = w2h + KP;
  uint8_t* stage = reinterpret_cast<uint8_t*>(w2l + KP);               // [4 warps][32 edges][SST]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3, lr = lane & 7, lm = lane >> 3;
  for (int i = tid; i < H0; i += blockDim.x) {
    w0[i] = w0g[i];
    b0[i] = b0g[i];
  }
  for (int i = tid; i < H1; i += blockDim.x) b1[i] = b1g[i];
  for (int i = tid; i < H1 * H0; i += blockDim.x) {
    const float v = w1g[i];
    const __half hi = __float2half_rn(v);
    w1h[i / H0][i % H0] = hi;
    w1l[i / H0][i % H0] = __float2half_rn(v - __half2float(hi));
  }
  for (int slot = tid; slot < KP; slot += blockDim.x) {
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    b2p[slot] = (r < kt && ch < k1 - 1) ? b2g[ch] : ((r < kt && ch == k1 - 1) ? 1.f : 0.f);
    gcs[slot] = (OMODE == 3 && gcenter != nullptr) ? gcenter[slot] : 0.f;
  }
  for (int i = tid; i < KP * H1; i += blockDim.x) {
    const int slot = i / H1, in = i % H1;
    const int q = slot / ktp, r = slot % ktp, ch = q * kt + r;
    const float v = (r < kt && ch < k1 - 1) ? w2g[ch * H1 + in] : 0.f;
    const __half hi = __float2half_rn(v);
    w2h[slot][in] = hi;
    w2l[slot][in] = __float2half_rn(v - __half2float(hi));
  }
  __syncthreads();
  auto lrelu = [](float v) { return v > 0.f ? v : 0.01f * v; };
  // ldmatrix rows of the B fragments: matrices (n-tile 2np, k lo), (2np, k hi), (2np + 1, k lo), (2np + 1, k hi)
  const uint32_t b1_addr = em_smem(&w1h[(lm >> 1) * 8 + lr][(lm & 1) * 8]);
  const uint32_t b2_addr = em_smem(&w2h[(lm >> 1) * 8 + lr][(lm & 1) * 8]);
  constexpr uint32_t LO1 = H1 * W1S * 2, LO2 = KP * W2S * 2;           // byte distance hi -> lo copy
  uint8_t* const wstage = stage + (size_t)warp * 32 * SST;
  // fp16 rows: stmatrix rows of the C tiles of m-tile mt: matrices (hh 0, nt), (hh 1, nt), (hh 0, nt + 1), (hh 1, nt + 1)
  const uint32_t s_addr = em_smem(wstage) + (uint32_t)(((lm & 1) * 8 + lr) * SST + (lm >> 1) * 16);
  constexpr int QB = OMODE >= 2 ? CW / 8 : CW / 4;                     // 16-byte chunks per staged row
  const int n_groups = (E + 31) / 32;
  auto load_d = [&](int g) {
    const int e = min(g * 32 + lane, E - 1);
    return __ldg(edge_attr + (perm ? __ldg(perm + e) : e));
  };
  float d_next = load_d(min((int)(blockIdx.x * 4 + warp), n_groups - 1));
  for (int grp = blockIdx.x * 4 + warp; grp < n_groups; grp += gridDim.x * 4) {
    const int e_base = grp * 32;
    const float d_lane = d_next;
    d_next = load_d(min(grp + (int)gridDim.x * 4, n_groups - 1));
    float dr[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      dr[mt][0] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq);
      dr[mt][1] = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq + 8);
    }
    // ---- layer 1: [32 edges, H0] x [H0, H1], layer 0 evaluated into the A fragments
    float acc1[2][NT1][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc1[mt][nt][r] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS0; ++ks) {
      uint32_t ah[2][4], al[2][4];
      const int c = ks * 16 + 2 * tq;
      const float2 wl = *reinterpret_cast<const float2*>(&w0[c]), bl = *reinterpret_cast<const float2*>(&b0[c]);
      const float2 wh = *reinterpret_cast<const float2*>(&w0[c + 8]), bh = *reinterpret_cast<const float2*>(&b0[c + 8]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float d = dr[mt][hh];
          em_split(lrelu(fmaf(d, wl.x, bl.x)), lrelu(fmaf(d, wl.y, bl.y)), ah[mt][hh], al[mt][hh], guard);
          em_split(lrelu(fmaf(d, wh.x, bh.x)), lrelu(fmaf(d, wh.y, bh.y)), ah[mt][2 + hh], al[mt][2 + hh], guard);
        }
#pragma unroll
      for (int np = 0; np < NT1 / 2; ++np) {
        uint32_t bh_[4], bl_[4];
        const uint32_t ad = b1_addr + (uint32_t)((np * 16 * W1S + ks * 16) * 2);
        em_ldsm4(ad, bh_);
        em_ldsm4(ad + LO1, bl_);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          em_mma16(acc1[mt][2 * np], al[mt], bh_[0], bh_[1]);
          em_mma16(acc1[mt][2 * np + 1], al[mt], bh_[2], bh_[3]);
          em_mma16(acc1[mt][2 * np], ah[mt], bl_[0], bl_[1]);
          em_mma16(acc1[mt][2 * np + 1], ah[mt], bl_[2], bl_[3]);
          em_mma16(acc1[mt][2 * np], ah[mt], bh_[0], bh_[1]);
          em_mma16(acc1[mt][2 * np + 1], ah[mt], bh_[2], bh_[3]);
        }
      }
    }
    // ---- layer-1 activations as the A fragments of layer 2: n-tiles (2ks, 2ks + 1) of the accumulator are the
    // (k lo, k hi) halves of k-step ks; rows gq / gq + 8 sit in accumulator elements {0,1} / {2,3}
    uint32_t a1h[2][KS1][4], a1l[2][KS1][4];
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks) {
      const float2 bl = *reinterpret_cast<const float2*>(&b1[ks * 16 + 2 * tq]);
      const float2 bh = *reinterpret_cast<const float2*>(&b1[ks * 16 + 8 + 2 * tq]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        em_split(lrelu(acc1[mt][2 * ks][0] + bl.x), lrelu(acc1[mt][2 * ks][1] + bl.y), a1h[mt][ks][0], a1l[mt][ks][0], guard);
        em_split(lrelu(acc1[mt][2 * ks][2] + bl.x), lrelu(acc1[mt][2 * ks][3] + bl.y), a1h[mt][ks][1], a1l[mt][ks][1], guard);
        em_split(lrelu(acc1[mt][2 * ks + 1][0] + bh.x), lrelu(acc1[mt][2 * ks + 1][1] + bh.y), a1h[mt][ks][2], a1l[mt][ks][2], guard);
        em_split(lrelu(acc1[mt][2 * ks + 1][2] + bh.x), lrelu(acc1[mt][2 * ks + 1][3] + bh.y), a1h[mt][ks][3], a1l[mt][ks][3], guard);
      }
    }
    // ---- layer 2 in chunks of CW output slots
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
      float acc[2][NTC][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
      const uint32_t bch = b2_addr + (uint32_t)(ch * CW * W2S * 2);
#pragma unroll
      for (int ks = 0; ks < KS1; ++ks)
#pragma unroll
        for (int np = 0; np < NTC / 2; ++np) {
          uint32_t bh_[4], bl_[4];
          const uint32_t ad = bch + (uint32_t)((np * 16 * W2S + ks * 16) * 2);
          em_ldsm4(ad, bh_);
          em_ldsm4(ad + LO2, bl_);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            em_mma16(acc[mt][2 * np], a1l[mt][ks], bh_[0], bh_[1]);
            em_mma16(acc[mt][2 * np + 1], a1l[mt][ks], bh_[2], bh_[3]);
            em_mma16(acc[mt][2 * np], a1h[mt][ks], bl_[0], bl_[1]);
            em_mma16(acc[mt][2 * np + 1], a1h[mt][ks], bl_[2], bl_[3]);
            em_mma16(acc[mt][2 * np], a1h[mt][ks], bh_[0], bh_[1]);
            em_mma16(acc[mt][2 * np + 1], a1h[mt][ks], bh_[2], bh_[3]);
          }
        }
      if (OMODE >= 2) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int np = 0; np < NTC / 2; ++np) {
            uint32_t h[2][2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int nt = 2 * np + q;
              const float2 bz = *reinterpret_cast<const float2*>(&b2p[ch * CW + nt * 8 + 2 * tq]);
              const float2 gc = *reinterpret_cast<const float2*>(&gcs[ch * CW + nt * 8 + 2 * tq]);
              h[q][0] = em_pack(lrelu(acc[mt][nt][0] + bz.x) - gc.x, lrelu(acc[mt][nt][1] + bz.y) - gc.y, guard);
              h[q][1] = em_pack(lrelu(acc[mt][nt][2] + bz.x) - gc.x, lrelu(acc[mt][nt][3] + bz.y) - gc.y, guard);
            }
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(s_addr + (uint32_t)(mt * 16 * SST + np * 32)),
                         "r"(h[0][0]), "r"(h[0][1]), "r"(h[1][0]), "r"(h[1][1])
                         : "memory");
          }
      } else {
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
          const float2 bz = *reinterpret_cast<const float2*>(&b2p[ch * CW + nt * 8 + 2 * tq]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float v0 = __uint_as_float(em_tf32(lrelu(acc[mt][nt][2 * hh] + bz.x)));
              const float v1 = __uint_as_float(em_tf32(lrelu(acc[mt][nt][2 * hh + 1] + bz.y)));
              *reinterpret_cast<float2*>(wstage + (size_t)(mt * 16 + gq + 8 * hh) * SST + (nt * 8 + 2 * tq) * 4) = make_float2(v0, v1);
            }
        }
      }
      __syncwarp();
      if (OMODE == 3) {          // planar: 16-byte chunk c of the 48-slot chunk ch = half (c & 1) of part 3 ch + c / 2
        __half* const gh = static_cast<__half*>(gv);
        for (int t = lane; t < 32 * QB; t += 32) {
          const int r = t / QB, c = t - r * QB;
          if (e_base + r < E)
            *reinterpret_cast<uint4*>(gh + ((size_t)(ch * 3 + (c >> 1)) * E + e_base + r) * 16 + (c & 1) * 8) =
                *reinterpret_cast<const uint4*>(wstage + (size_t)r * SST + c * 16);
        }
      } else {
        constexpr int ESZ = OMODE == 2 ? 2 : 4;
        uint8_t* const dst = static_cast<uint8_t*>(gv) + (size_t)ch * CW * ESZ;
        for (int t = lane; t < 32 * QB; t += 32) {
          const int r = t / QB, c = t - r * QB;
          if (e_base + r < E)
            *reinterpret_cast<uint4*>(dst + (size_t)(e_base + r) * KP * ESZ + c * 16) =
                *reinterpret_cast<const uint4*>(wstage + (size_t)r * SST + c * 16);
        }
      }
      __syncwarp();
    }
  }
  guard.flush(ovf);
}

// returns 1 when the shape is not covered here (the caller then uses the generic CUDA-core kernel)
int launch_edge_hidden3_mma(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                            int64_t E, float* g, cudaStream_t s, int omode) {
  if (E == 0) return FESR_OK;
  if (!(d.n_hidden == 3 && d.hidden[0] == 32 && d.hidden[1] == 64 && d.hidden[2] == 128 && d.kp == 144 && d.leaky)) return 1;
  if (omode < 1 || omode > 3) return 1;
  static const bool off = getenv("FESR_EDGE_FFMA") != nullptr;            // A/B switch for profiling
  if (off) return 1;
  for (int l = 0; l < 3; ++l)
    if (!p.mlp_w[l] || !p.mlp_b[l]) {
      set_error("NULL edge-MLP parameter %d", l);
      return FESR_EINVAL;
    }
  const int64_t blocks = ceil_div(ceil_div(E, 32), 4);
  const int grid = (int)(blocks < 3ll * num_sms() ? blocks : 3ll * num_sms());
  const size_t wbytes = (size_t)2 * 64 * 40 * 2 + (size_t)2 * 144 * 72 * 2;
  const size_t smem16 = wbytes + (size_t)4 * 32 * (48 + 8) * 2, smem32 = wbytes + (size_t)4 * 32 * (48 + 4) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden3_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden3_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden3_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    attr_set = true;
  }
  ProfScope prof(PROF_EDGE_HIDDEN, s);
  if (omode == 3)
    edge_hidden3_mma_kernel<3><<<grid, 128, smem16, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], p.mlp_w[2], p.mlp_b[2],
                                                        d.kt, d.ktp, d.k1, edge_attr, perm, (int)E, g, cur_ovf(), cur_gcenter());
  else if (omode == 2)
    edge_hidden3_mma_kernel<2><<<grid, 128, smem16, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], p.mlp_w[2], p.mlp_b[2],
                                                        d.kt, d.ktp, d.k1, edge_attr, perm, (int)E, g, cur_ovf(), cur_gcenter());
  else
    edge_hidden3_mma_kernel<1><<<grid, 128, smem32, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], p.mlp_w[2], p.mlp_b[2],
                                                        d.kt, d.ktp, d.k1, edge_attr, perm, (int)E, g, cur_ovf(), cur_gcenter());
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <int WPAD, int NTO>
static int launch_eh2m(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                       int64_t E, float* g, cudaStream_t s, int omode) {
  const int64_t blocks = ceil_div(ceil_div(E, 32), 4);
  const int grid = (int)(blocks < 8ll * num_sms() ? blocks : 8ll * num_sms());
  ProfScope prof(PROF_EDGE_HIDDEN, s);
  constexpr size_t stage_bytes = (size_t)4 * 32 * (NTO * 8 + 4) * sizeof(float);
  // persistent over edge groups: one wave of resident blocks (as for the fp16 kernels below; 8 blocks per SM with 5
  // resident at 96 registers ran as a full wave plus a 60 % one)
  auto wave = [&](const void* fn) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 128, stage_bytes) != cudaSuccess || occ < 1) return grid;
    const int64_t cap = (int64_t)occ * num_sms();
    return (int)(blocks < cap ? blocks : cap);
  };
#define FESR_EH(TERMS, OMODE)                                                                                   \
  edge_hidden2_mma_kernel<WPAD, NTO, TERMS, OMODE><<<wave(reinterpret_cast<const void*>(&edge_hidden2_mma_kernel<WPAD, NTO, TERMS, OMODE>)), 128, stage_bytes, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], \
                                                                        d.w, d.leaky, d.kt, d.ktp, d.k1, edge_attr, perm, E, g, cur_ovf(), cur_gcenter())
  static bool attr_set = false;
  if (!attr_set) {   // static + dynamic shared memory exceeds 48 KB for the widest rows
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden2_mma_kernel<WPAD, NTO, 3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden2_mma_kernel<WPAD, NTO, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden2_mma_kernel<WPAD, NTO, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden2_mma_kernel<WPAD, NTO, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    attr_set = true;
  }
  static const bool tf32_only = getenv("FESR_EDGE_TF32") != nullptr;     // A/B switch for profiling
  // the fp16 kernels are persistent over edge groups: exactly one wave of resident blocks (a grid of 8 blocks per SM
  // with 5 resident ran as a full wave plus a 60 % one)
  static int occ16 = 0;
  if (occ16 == 0) {
    FESR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ16, edge_hidden2_f16_kernel<WPAD, NTO, 3>, 128, 0));
    if (occ16 < 1) occ16 = 1;
  }
  const int grid16 = (int)(blocks < (int64_t)occ16 * num_sms() ? blocks : (int64_t)occ16 * num_sms());
  if (omode == 0) FESR_EH(3, 0);
  else if (omode == 1) FESR_EH(1, 1);
  else if (!tf32_only && !d.leaky && omode == 2)
    edge_hidden2_f16_kernel<WPAD, NTO, 2><<<grid16, 128, 0, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], d.w, d.kt, d.ktp, d.k1,
                                                              edge_attr, perm, (int)E, reinterpret_cast<__half*>(g), cur_ovf(), cur_gcenter());
  else if (!tf32_only && !d.leaky)
    edge_hidden2_f16_kernel<WPAD, NTO, 3><<<grid16, 128, 0, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], d.w, d.kt, d.ktp, d.k1,
                                                              edge_attr, perm, (int)E, reinterpret_cast<__half*>(g), cur_ovf(), cur_gcenter());
  else if (omode == 2) FESR_EH(1, 2);
  else FESR_EH(1, 3);
#undef FESR_EH
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// returns 1 when the shape is not covered here (the caller then uses the generic CUDA-core kernel)
int launch_edge_hidden2_mma(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                            int64_t E, float* g, cudaStream_t s, int round_tf32) {
  if (E == 0) return FESR_OK;
  const int wpad = d.w <= 16 ? 16 : (d.w <= 32 ? 32 : 48);
  if (d.w <= 48 && d.passes == 1) {
    if (wpad == 16 && d.kp == 16) return launch_eh2m<16, 2>(d, p, edge_attr, perm, E, g, s, round_tf32);
    if (wpad == 16 && d.kp == 32) return launch_eh2m<16, 4>(d, p, edge_attr, perm, E, g, s, round_tf32);
    if (wpad == 32 && d.kp == 32) return launch_eh2m<32, 4>(d, p, edge_attr, perm, E, g, s, round_tf32);
    if (wpad == 32 && d.kp == 48) return launch_eh2m<32, 6>(d, p, edge_attr, perm, E, g, s, round_tf32);
    if (wpad == 48 && d.kp == 48) return launch_eh2m<48, 6>(d, p, edge_attr, perm, E, g, s, round_tf32);
    if (wpad == 48 && d.kp == 64) return launch_eh2m<48, 8>(d, p, edge_attr, perm, E, g, s, round_tf32);
  }
  return 1;
}

}  // namespace fesr
