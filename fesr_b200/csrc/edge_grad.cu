// Backward of the outer-product aggregation with respect to the edge features g:
//   dg[e, k] += 1/deg(dst_e) * sum_a h[src_e, a] * dZ[dst_e, k*wp + a]
// One warp per destination node holds dZ_i in registers in exactly the lane tiling zbuild
// writes Z with (lane = 8*q + ag), loops over the node's incoming edges in CSR order, reduces the
// partial dot products over the 8 column groups with shuffles and the ag == 0 lanes accumulate
// into dg (one writer per element => deterministic).
#include <cuda_bf16.h>
#include <stdlib.h>

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

// Reduced-precision arm: the same contraction on mma.sync.m16n8k8 tf32 (fp32 accumulate), one warp per destination node:
//   dG_i [16 edges x channels] = H_i [16 edges x wp] . dZ_i^T [wp x channels]
// A fragments are the gathered h[src] rows, B fragments the node's dZ row block (channel-major, wp contiguous: the
// lane pattern 8 channels x 4 consecutive floats reads whole 32-byte sectors across the two halves of a k-step), the
// accumulators go straight into dg (one writer per element: deterministic).  ~9x fewer instructions than the
// shuffle-reduce kernel above, which stays the fp32 arm.
__device__ __forceinline__ uint32_t eg_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// a lane's 12 consecutive dZ values (a = 12 tq .. 12 tq + 11 of one channel) as tf32 operand bits: b[2 ks], b[2 ks + 1]
__device__ __forceinline__ void eg_ldz12(const float* p, uint32_t (&b)[12]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
    b[4 * q] = eg_tf32(v.x);
    b[4 * q + 1] = eg_tf32(v.y);
    b[4 * q + 2] = eg_tf32(v.z);
    b[4 * q + 3] = eg_tf32(v.w);
  }
}
// the same 12 values unrounded (3xTF32: the caller splits them into hi + lo)
__device__ __forceinline__ void eg_ldf12(const float* p, float (&b)[12]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
    b[4 * q] = v.x;
    b[4 * q + 1] = v.y;
    b[4 * q + 2] = v.z;
    b[4 * q + 3] = v.w;
  }
}
__device__ __forceinline__ void eg_ldz12(const __nv_bfloat16* p, uint32_t (&b)[12]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + q);      // bf16 -> tf32 bits: exact
    b[4 * q] = v.x << 16;
    b[4 * q + 1] = v.x & 0xffff0000u;
    b[4 * q + 2] = v.y << 16;
    b[4 * q + 3] = v.y & 0xffff0000u;
  }
}

// ZT = float, or __nv_bfloat16 for the bf16 dZ of the tcgen05 product (bf16 -> tf32 is exact: no second rounding)
// TERMS == 3 (fp32 arm, ZT = float): h rows and dZ are split into tf32 hi + lo in registers, lo.hi + hi.lo + hi.hi
template <int WP, typename ZT, int TERMS>
__global__ void __launch_bounds__(128)
edge_grad_mma_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                     const ZT* __restrict__ dZ, const float* __restrict__ h, int64_t n, int k1p, int kt, int ktp, int kp,
                     int zk, float* __restrict__ dg) {
  constexpr int KS = WP / 8;                // k-steps over the node-feature index a
  constexpr int NTC = 6;                    // channel n-tiles per pass (24 accumulator registers)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int64_t warp_global = (int64_t)blockIdx.x * 4 + warp, warp_stride = (int64_t)gridDim.x * 4;
  const int n_nt = (k1p + 7) / 8;
  // the 16 x kp tile of dg contributions of a chunk is staged per warp, so that dg is updated with coalesced
  // 16-byte read-modify-writes of whole rows instead of scattered 4-byte ones
  extern __shared__ __align__(16) float eg_tile[];
  float* tile = eg_tile + (size_t)warp * 16 * kp;
  const int q4 = kp >> 2;
  for (int64_t i = warp_global; i < n; i += warp_stride) {
    const int eb = rowptr[i], ee = rowptr[i + 1];
    {
      // the next node's dZ block (zk values: 34 lines as bf16) into L1 while this node computes: its fragment loads are
      // what the warp waits on (29 % of the kernel's stall samples)
      const int64_t inext = i + warp_stride;
      if (inext < n) {
        const char* pz = reinterpret_cast<const char*>(dZ + inext * (int64_t)zk);
        const int nline = (int)(((int64_t)zk * sizeof(ZT) + 127) / 128);
        for (int l = lane; l < nline; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(pz + (size_t)l * 128));
      }
    }
    if (ee == eb) continue;
    const float inv = 1.0f / (float)(ee - eb);
    const ZT* zrow = dZ + i * (int64_t)zk;
    for (int c0 = eb; c0 < ee; c0 += 16) {
      const int e0 = c0 + gq, e1 = c0 + gq + 8;
      // The contraction index of the MMAs is a permutation of the 48 node-feature indices, the same in A and B: k-step ks,
      // slot tq <-> a = 12 tq + 2 ks, slot tq + 4 <-> a = 12 tq + 2 ks + 1.  A lane then owns 12 CONSECUTIVE features of its
      // rows for all six k-steps: three 16-byte loads per h row, three 8- / 16-byte loads per dZ channel row (a quarter of
      // the load instructions of the natural order, each with four times the bytes in flight -- the kernel is bound by the
      // latency of these loads: 80 % of its stall samples are long-scoreboard).
      static_assert(WP == 48, "12 features per lane");
      const float* h0 = h + (int64_t)__ldg(src_sorted + min(e0, ee - 1)) * WP + 12 * tq;
      const float* h1 = h + (int64_t)__ldg(src_sorted + min(e1, ee - 1)) * WP + 12 * tq;
      uint32_t a[KS][4];
      uint32_t alo[TERMS == 3 ? KS : 1][4];
      if constexpr (TERMS == 3) {
        float v0[12], v1[12];
        eg_ldf12(h0, v0);
        eg_ldf12(h1, v1);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const float f[4] = {v0[2 * ks], v1[2 * ks], v0[2 * ks + 1], v1[2 * ks + 1]};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            a[ks][q] = eg_tf32(f[q]);
            alo[ks][q] = eg_tf32(f[q] - __uint_as_float(a[ks][q]));
          }
        }
      } else {
        uint32_t v0[12], v1[12];
        eg_ldz12(h0, v0);
        eg_ldz12(h1, v1);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          a[ks][0] = v0[2 * ks];
          a[ks][1] = v1[2 * ks];
          a[ks][2] = v0[2 * ks + 1];
          a[ks][3] = v1[2 * ks + 1];
        }
      }
      // the chunk's rows of dg are brought into the warp's tile by cp.async while the MMAs run (the read-modify-write of
      // dg used to be 36 % of the kernel's stall samples: a dependent global load in front of every row update)
      const int nrow = min(16, ee - c0);
      float4* drow = reinterpret_cast<float4*>(dg + (int64_t)c0 * kp);
      for (int t = lane; t < 16 * q4; t += 32) {
        if (t < nrow * q4) {
          const uint32_t dsts = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<float4*>(tile) + t);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts), "l"(drow + t) : "memory");
        } else {
          reinterpret_cast<float4*>(tile)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      bool landed = false;
      for (int nt0 = 0; nt0 < n_nt; nt0 += NTC) {
        float acc[NTC][4];
#pragma unroll
        for (int t = 0; t < NTC; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
        for (int t = 0; t < NTC; ++t) {
          if (nt0 + t < n_nt) {
            // B[a][channel]: dZ[chan = (nt0 + t) * 8 + gq][a = 12 tq + 2 ks (+ 1)]
            // (channels past k1p in the last tile: clamped row, masked on store)
            const ZT* zp = zrow + (int64_t)min((nt0 + t) * 8 + gq, k1p - 1) * WP + 12 * tq;
            if constexpr (TERMS == 3) {
              float bf[12];
              eg_ldf12(reinterpret_cast<const float*>(zp), bf);
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const uint32_t b0 = eg_tf32(bf[2 * ks]), b1 = eg_tf32(bf[2 * ks + 1]);
                const uint32_t l0 = eg_tf32(bf[2 * ks] - __uint_as_float(b0)), l1 = eg_tf32(bf[2 * ks + 1] - __uint_as_float(b1));
                asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                    : "r"(alo[ks][0]), "r"(alo[ks][1]), "r"(alo[ks][2]), "r"(alo[ks][3]), "r"(b0), "r"(b1));
                asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                    : "r"(a[ks][0]), "r"(a[ks][1]), "r"(a[ks][2]), "r"(a[ks][3]), "r"(l0), "r"(l1));
                asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                    : "r"(a[ks][0]), "r"(a[ks][1]), "r"(a[ks][2]), "r"(a[ks][3]), "r"(b0), "r"(b1));
              }
            } else {
              uint32_t b[12];
              eg_ldz12(zp, b);
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                    : "r"(a[ks][0]), "r"(a[ks][1]), "r"(a[ks][2]), "r"(a[ks][3]), "r"(b[2 * ks]), "r"(b[2 * ks + 1]));
              }
            }
          }
        }
        if (!landed) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          landed = true;
        }
        // accumulator (row gq / gq + 8, channels 2 tq, 2 tq + 1 of tile t) -> dg[edge][slot(channel)]
#pragma unroll
        for (int t = 0; t < NTC; ++t) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int chan = (nt0 + t) * 8 + 2 * tq + j;
            if (nt0 + t < n_nt && chan < k1p) {
              const int off = (chan / kt) * ktp + (chan % kt);
              tile[gq * kp + off] += acc[t][j] * inv;
              tile[(gq + 8) * kp + off] += acc[t][2 + j] * inv;
            }
          }
        }
      }
      __syncwarp();
      for (int t = lane; t < nrow * q4; t += 32) drow[t] = reinterpret_cast<const float4*>(tile)[t];
      __syncwarp();
    }
  }
}

// All layers in one pass (tf32 arm, bf16 dZ kept per layer by the caller): dg[e, :] = inv_deg[dst_e] * sum over the layers l
// of dZ_l[dst_e] . h_l[src_e].  dg is written once instead of read and re-written by every layer -- the [E, kp] fp32
// read-modify-write was half of the per-layer kernel's traffic (0.70 of 1.40 GB per launch at 527 k cells,
// profiles/r03f_train_kernels_summary.txt) -- and needs no zero fill: every edge row belongs to exactly one node.
// One warp per node and 16-edge chunk as above; the accumulators stay in registers across the layers.
struct EgLayers {
  const __nv_bfloat16* dZ[FESR_EG_MAX_LAYERS];
  const float* h[FESR_EG_MAX_LAYERS];
  int nl;
};

#ifndef EGL_MINB
#define EGL_MINB 5
#endif
__global__ void __launch_bounds__(128, EGL_MINB)
edge_grad_layers_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted, const EgLayers lay, int64_t n,
                        int k1p, int kt, int ktp, int kp, int zk, float* __restrict__ dg) {
  constexpr int WP = 48, KS = WP / 8, NTC = 6;      // k1p <= 48: six channel n-tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int64_t warp_global = (int64_t)blockIdx.x * 4 + warp, warp_stride = (int64_t)gridDim.x * 4;
  const int n_nt = (k1p + 7) / 8;
  extern __shared__ __align__(16) float eg_tile[];
  float* tile = eg_tile + (size_t)warp * 16 * kp;
  const int q4 = kp >> 2;
  const int nline = (int)(((int64_t)zk * sizeof(__nv_bfloat16) + 127) / 128);
  for (int64_t i = warp_global; i < n; i += warp_stride) {
    const int eb = rowptr[i], ee = rowptr[i + 1];
    if (ee == eb) continue;
    const float inv = 1.0f / (float)(ee - eb);
    for (int c0 = eb; c0 < ee; c0 += 16) {
      const int e0 = c0 + gq, e1 = c0 + gq + 8;
      const int64_t s0 = (int64_t)__ldg(src_sorted + min(e0, ee - 1)) * WP + 12 * tq;
      const int64_t s1 = (int64_t)__ldg(src_sorted + min(e1, ee - 1)) * WP + 12 * tq;
      for (int t = lane; t < 16 * q4; t += 32) reinterpret_cast<float4*>(tile)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      float acc[NTC][4];
#pragma unroll
      for (int t = 0; t < NTC; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
      for (int l = 0; l < lay.nl; ++l) {
        {
          // the block read next (this node's next layer, or the first layer of the warp's next node) into L1 meanwhile
          // (a second prefetch, two blocks ahead into L2, measured no gain)
          const __nv_bfloat16* nxt = nullptr;
          if (l + 1 < lay.nl) nxt = lay.dZ[l + 1] + i * (int64_t)zk;
          else if (c0 + 16 >= ee && i + warp_stride < n) nxt = lay.dZ[0] + (i + warp_stride) * (int64_t)zk;
          if (nxt != nullptr) {
            const char* pz = reinterpret_cast<const char*>(nxt);
            for (int q = lane; q < nline; q += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(pz + (size_t)q * 128));
          }
        }
        if (l + 1 < lay.nl) {      // ... and the next layer's two gathered h rows (L2 hits at best: random rows)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(lay.h[l + 1] + s0));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(lay.h[l + 1] + s1));
        }
        uint32_t v0[12], v1[12];
        eg_ldz12(lay.h[l] + s0, v0);
        eg_ldz12(lay.h[l] + s1, v1);
        const __nv_bfloat16* zrow = lay.dZ[l] + i * (int64_t)zk + 12 * tq;
        // the channel tiles' dZ fragments are loaded one tile ahead of their MMAs (under the 80-register cap the compiler
        // otherwise loads tile t + 1 after the MMAs of tile t: six exposed L1 / L2 latencies per layer, 43 % of the
        // kernel's stall samples)
        uint2 raw[2][3];
        auto ldraw = [&](int t, uint2 (&r)[3]) {
          const uint2* pz = reinterpret_cast<const uint2*>(zrow + (int64_t)min(t * 8 + gq, k1p - 1) * WP);
          r[0] = __ldg(pz);
          r[1] = __ldg(pz + 1);
          r[2] = __ldg(pz + 2);
        };
        ldraw(0, raw[0]);
#pragma unroll
        for (int t = 0; t < NTC; ++t) {
          if (t + 1 < NTC && t + 1 < n_nt) ldraw(t + 1, raw[(t + 1) & 1]);
          if (t < n_nt) {
            uint32_t b[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {      // bf16 -> tf32 bits: exact
              b[4 * q] = raw[t & 1][q].x << 16;
              b[4 * q + 1] = raw[t & 1][q].x & 0xffff0000u;
              b[4 * q + 2] = raw[t & 1][q].y << 16;
              b[4 * q + 3] = raw[t & 1][q].y & 0xffff0000u;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                  : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                  : "r"(v0[2 * ks]), "r"(v1[2 * ks]), "r"(v0[2 * ks + 1]), "r"(v1[2 * ks + 1]), "r"(b[2 * ks]), "r"(b[2 * ks + 1]));
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < NTC; ++t) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int chan = t * 8 + 2 * tq + j;
          if (t < n_nt && chan < k1p) {
            const int off = (chan / kt) * ktp + (chan % kt);
            tile[gq * kp + off] = acc[t][j] * inv;
            tile[(gq + 8) * kp + off] = acc[t][2 + j] * inv;
          }
        }
      }
      __syncwarp();
      const int nrow = min(16, ee - c0);
      float4* drow = reinterpret_cast<float4*>(dg + (int64_t)c0 * kp);
      for (int t = lane; t < nrow * q4; t += 32) drow[t] = reinterpret_cast<const float4*>(tile)[t];
      __syncwarp();
    }
  }
}

bool edge_grad_layers_supported(const fesr_model_dims& d, int n_layers) {
  return d.wp == 48 && d.k1p <= 48 && n_layers >= 1 && n_layers <= FESR_EG_MAX_LAYERS;
}

int launch_edge_grad_layers(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* const* dZ_bf16,
                            const float* const* h, int n_layers, int64_t n, float* dg, cudaStream_t s) {
  FESR_CHECK_ARG(edge_grad_layers_supported(d, n_layers), "all-layer edge gradient: wp == 48, k1p <= 48, <= %d layers",
                 FESR_EG_MAX_LAYERS);
  if (n == 0) return FESR_OK;
  ProfScope prof(PROF_BACKWARD, s);
  EgLayers lay;
  for (int l = 0; l < FESR_EG_MAX_LAYERS; ++l) {
    lay.dZ[l] = static_cast<const __nv_bfloat16*>(dZ_bf16[l < n_layers ? l : 0]);
    lay.h[l] = h[l < n_layers ? l : 0];
  }
  lay.nl = n_layers;
  const int64_t blocks = ceil_div(n, 4);
  // one persistent wave: EGL_MINB blocks are resident per SM (16 per SM = 3.2 waves left a tail: 6.08 vs 5.94 ms per train
  // step at 527 k cells; 10 per SM 5.95, 32 per SM 5.99)
  static const int per_sm = getenv("FESR_EGL_BLOCKS_PER_SM") && atoi(getenv("FESR_EGL_BLOCKS_PER_SM")) > 0
                                ? atoi(getenv("FESR_EGL_BLOCKS_PER_SM")) : EGL_MINB;
  const int grid = (int)(blocks < (int64_t)per_sm * num_sms() ? blocks : (int64_t)per_sm * num_sms());
  const size_t smem = (size_t)4 * 16 * d.kp * sizeof(float);
  edge_grad_layers_kernel<<<grid, 128, smem, s>>>(rowptr, src_sorted, lay, n, d.k1p, d.kt, d.ktp, d.kp, d.zk, dg);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
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

int launch_edge_grad(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* dZv,
                     const float* h, int64_t n, int use_mma, float* dg, cudaStream_t s, int dz_bf16) {
  const float* dZ = static_cast<const float*>(dZv);
  // dg is zero-initialised by the caller and always accumulated into
  if (n == 0) return FESR_OK;
  ProfScope prof(PROF_BACKWARD, s);
  static const bool no_mma = getenv("FESR_EDGE_GRAD_FFMA") != nullptr;      // A/B switch for profiling
  static const bool no_3x = getenv("FESR_EDGE_GRAD_3X") && atoi(getenv("FESR_EDGE_GRAD_3X")) == 0;      // A/B switch
  if (use_mma == 3 && (no_mma || no_3x || d.wp != 48)) use_mma = 0;
  if (use_mma && !no_mma && d.wp == 48) {
    // one persistent wave: as many blocks as are resident (16 per SM = 2.7-4 waves left a tail, see launch_edge_grad_layers)
    const int64_t blocks = ceil_div(n, 4);
    const size_t smem = (size_t)4 * 16 * d.kp * sizeof(float);      // <= 36.9 KB (kp = 144)
    auto grid_of = [&](const void* fn) {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 128, smem) != cudaSuccess || occ < 1) occ = 4;
      const int64_t cap = (int64_t)occ * num_sms();
      return (int)(blocks < cap ? blocks : cap);
    };
    if (dz_bf16) {
      auto* fn = edge_grad_mma_kernel<48, __nv_bfloat16, 1>;
      fn<<<grid_of(reinterpret_cast<const void*>(fn)), 128, smem, s>>>(rowptr, src_sorted, static_cast<const __nv_bfloat16*>(dZv), h, n,
                                                                   d.k1p, d.kt, d.ktp, d.kp, d.zk, dg);
    } else if (use_mma == 3) {
      auto* fn = edge_grad_mma_kernel<48, float, 3>;
      fn<<<grid_of(reinterpret_cast<const void*>(fn)), 128, smem, s>>>(rowptr, src_sorted, dZ, h, n, d.k1p, d.kt, d.ktp, d.kp, d.zk, dg);
    } else {
      auto* fn = edge_grad_mma_kernel<48, float, 1>;
      fn<<<grid_of(reinterpret_cast<const void*>(fn)), 128, smem, s>>>(rowptr, src_sorted, dZ, h, n, d.k1p, d.kt, d.ktp, d.kp, d.zk, dg);
    }
    FESR_LAUNCH_CHECK();
    return FESR_OK;
  }
  FESR_CHECK_ARG(!dz_bf16, "bf16 dZ is read by the tensor-core edge-gradient kernel only");
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
