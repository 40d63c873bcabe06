// Gather + segmented mean with the per-node outer products on warp-level tensor-core MMAs.
//
// Same contract as zbuild.cu (Z_i = 1/deg_i sum_{e->i} g_e (x) h[src_e]  ++  h_i), used by the
// reduced-precision arms (FESR_PREC_TF32 / FESR_PREC_F16) where Z is consumed as an 11-bit-mantissa
// operand anyway.  Per destination node the sum of outer products is a tiny GEMM
//     Z_i [GROW x WP] = G_i^T [GROW x deg] . H_i [deg x WP]          (deg ~ 12, GROW = WP = 48)
// whose contraction length is one node's in-degree: far below tcgen05's M >= 64, CTA-wide issue
// granularity, so it is issued as mma.sync.m16n8k8 (tf32 in, fp32 accumulate) by the warp that
// owns the node.  Against the FFMA kernel this cuts the instruction count per edge ~2.5x and
// leaves the kernel bound by HBM (g read + Z write).  The fp32 arm keeps the FFMA kernel.
//
// Staging: every edge row is ONE bulk asynchronous copy (cp.async.bulk global->shared, the
// non-tensor TMA path) issued by one lane -- lane j copies the g row and the gathered h[src_j]
// row of edge j -- completing on a per-warp mbarrier; slabs are double buffered and the source
// ids are fetched two chunks ahead.  Slab rows are padded by 8 floats so that the MMA fragment
// loads are bank-conflict free.  In the forward, g and h are already tf32-rounded by their
// producers, so the fragments are loaded without any conversion.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fesr {

constexpr int ZM_WARPS = 8;
constexpr int ZM_DEGC = 16;   // edges per chunk = 2 MMA k-steps
constexpr int ZM_TASK = 8;

__device__ __forceinline__ uint32_t zm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zm_bulk(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   zm_smem(dst_smem)),
               "l"(src), "r"(bytes), "r"(zm_smem(bar))
               : "memory");
}
__device__ __forceinline__ void zm_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zm_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zm_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "ZM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra ZM_DONE;\n\t"
      "bra ZM_WAIT;\n\t"
      "ZM_DONE:\n\t"
      "}" ::"r"(zm_smem(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t zm_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint32_t zm_h2_sat(float lo, float hi) {   // {lo, hi} -> packed f16x2, saturating
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void zm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// same MMA with a zero C operand: starts an accumulator without a separate zeroing pass
__device__ __forceinline__ void zm_mma0(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

struct ZmItem {
  int k, p, c0, eb, ee;
};

// MT = (floats of g per edge per pass) / 16; ZMODE 1 = fp32 Z rounded to tf32, 2 = fp16 Z;
// BWD: backward use (sum instead of mean, per-gathered-row scale, inputs not pre-rounded)
template <int MT, int WP, int ZMODE, bool BWD>
__global__ void __launch_bounds__(ZM_WARPS * 32, 2)
zbuild_mma_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                  const float* __restrict__ g, const float* __restrict__ h, int64_t n, int passes, int kp, int kt,
                  int ktp, int zk_main, int zk, const float* __restrict__ gather_scale, void* __restrict__ Zv, int* ovf) {
  F16Guard guard;
  constexpr int GROW = 16 * MT;
  constexpr int NT = WP / 8;
  // slab row strides (floats): h rows padded by 8 (conflict-free B fragments); g rows unpadded so
  // that a chunk is one contiguous bulk copy (A fragment loads are then 2-way conflicted)
  constexpr int SG = GROW, SH = WP + 8;
  constexpr int BUF = ZM_DEGC * (SG + SH) + ZM_DEGC;        // + per-edge gather scale (backward)
  constexpr int LPR = WP / 4, RPI = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slab = smem + warp * (2 * BUF);
  const uint32_t slab_u32 = zm_smem(slab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ZM_WARPS * 2 * BUF) + warp * 2;
  const int gq = lane >> 2, tq = lane & 3;                  // MMA fragment coordinates
  const int hj = lane / LPR, hc = lane % LPR;
  const bool h_lane = lane < RPI * LPR;
  const unsigned FULL = 0xffffffffu;

  for (int t = lane; t < 2 * BUF; t += 32) slab[t] = 0.f;   // stale slab contents must stay finite
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(zm_smem(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(zm_smem(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t phase0 = 0, phase1 = 0;

  // output rows owned by this lane: slot = mt*16 + gq (+8) -> channel (skipping the pad slots)
  int chan[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int slot = mt * 16 + gq + 8 * hh;
      const int q = slot / ktp, r = slot % ktp;
      chan[mt][hh] = (r < kt) ? q * kt + r : -1;
    }

  const int64_t n_tasks = (n + ZM_TASK - 1) / ZM_TASK;
  const int64_t warp_global = (int64_t)blockIdx.x * ZM_WARPS + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * ZM_WARPS;

  for (int64_t task = warp_global; task < n_tasks; task += warp_stride) {
    const int64_t i0 = task * ZM_TASK;
    const int nn = (int)min((int64_t)ZM_TASK, n - i0);
    const int rp = (lane <= nn) ? __ldg(rowptr + i0 + lane) : 0;

    auto node_item = [&](int k) {
      ZmItem it;
      it.k = k;
      it.p = 0;
      it.eb = __shfl_sync(FULL, rp, min(k, ZM_TASK));
      it.ee = __shfl_sync(FULL, rp, min(k + 1, ZM_TASK));
      it.c0 = it.eb;
      return it;
    };
    auto advance = [&](ZmItem it) {
      if (it.k >= nn) return it;
      it.c0 += ZM_DEGC;
      if (it.c0 >= it.ee) {
        it.c0 = it.eb;
        if (++it.p == passes) return node_item(it.k + 1);
      }
      return it;
    };
    auto load_src = [&](const ZmItem& it) {
      const int e = it.c0 + lane;
      return (it.k < nn && lane < ZM_DEGC && e < it.ee) ? __ldg(src_sorted + e) : 0;
    };
    auto load_scale = [&](const ZmItem& it, int src_reg) {
      return (BWD && it.k < nn && lane < ZM_DEGC && it.c0 + lane < it.ee) ? __ldg(gather_scale + src_reg) : 1.f;
    };
    // lane j (< m) copies edge row j: the g row (GROW floats of this pass) and the gathered h row
    auto issue = [&](const ZmItem& it, int buf, int src_reg, float sc_reg) {
      float* sg = slab + buf * BUF;
      float* sh = sg + ZM_DEGC * SG;
      const int m = min(ZM_DEGC, it.ee - it.c0);
      // g: the chunk's rows are one contiguous block when a row is exactly one pass wide -> ONE bulk
      // copy (unpadded rows, SG == GROW); otherwise one bulk copy per row.  h[src]: 16-byte cp.async,
      // LPR lanes per gathered row.
      if (lane == 0) zm_expect(&bars[buf], (uint32_t)(m * GROW * 4));
      __syncwarp();
      if (kp == GROW) {
        if (lane == 0 && m > 0) zm_bulk(sg, g + (int64_t)it.c0 * GROW, (uint32_t)(m * GROW * 4), &bars[buf]);
      } else if (lane < m) {
        zm_bulk(sg + lane * SG, g + (int64_t)(it.c0 + lane) * kp + it.p * GROW, GROW * 4, &bars[buf]);
      }
      if (lane >= m && lane < ((m + 7) & ~7)) {
        // unused edge slots of an issued k-step: g row = 0 (the stale h row is finite)
#pragma unroll
        for (int c = 0; c < GROW; c += 4) *reinterpret_cast<float4*>(sg + lane * SG + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint32_t sh_lane = slab_u32 + (uint32_t)(buf * BUF + ZM_DEGC * SG + hj * SH + 4 * hc) * 4u;
      const float* h_lane_ptr = h + 4 * hc;
#pragma unroll 2
      for (int j0 = 0; j0 < m; j0 += RPI) {
        const int j = j0 + hj;
        const int s = __shfl_sync(FULL, src_reg, j & 31);
        if (h_lane && j < m)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sh_lane + (uint32_t)(j0 * SH * 4)),
                       "l"(h_lane_ptr + (int64_t)s * WP)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (BWD && lane < ZM_DEGC) sh[ZM_DEGC * SH + lane] = sc_reg;
    };

    ZmItem cur = node_item(0);
    int buf = 0;
    {
      const int s0 = load_src(cur);
      issue(cur, 0, s0, load_scale(cur, s0));
    }
    ZmItem nxt = advance(cur);
    int src_nxt = load_src(nxt);
    float sc_nxt = load_scale(nxt, src_nxt);
    float acc[MT][NT][4];
    while (cur.k < nn) {
      const bool has_next = nxt.k < nn;
      if (has_next) issue(nxt, buf ^ 1, src_nxt, sc_nxt);
      const ZmItem nn2 = advance(nxt);
      const int src_nn2 = load_src(nn2);
      if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (buf == 0) {
        zm_wait(&bars[0], phase0);
        phase0 ^= 1;
      } else {
        zm_wait(&bars[1], phase1);
        phase1 ^= 1;
      }
      __syncwarp();
      const bool fresh = cur.c0 == cur.eb;                 // first chunk of (node, pass)
      if (fresh && cur.ee == cur.eb) {                      // zero in-degree: the row is all zeros
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
      }
      const float* sg = slab + buf * BUF;
      const float* sh = sg + ZM_DEGC * SG;
      const int m = min(ZM_DEGC, cur.ee - cur.c0);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        if (ks * 8 < m) {
          const int e0 = ks * 8 + tq, e1 = e0 + 4;            // this lane's two edge rows of the k-step
          uint32_t a[MT][4];
          if constexpr (BWD) {
            const float s0 = sh[ZM_DEGC * SH + e0], s1 = sh[ZM_DEGC * SH + e1];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              a[mt][0] = zm_tf32(sg[e0 * SG + mt * 16 + gq] * s0);
              a[mt][1] = zm_tf32(sg[e0 * SG + mt * 16 + gq + 8] * s0);
              a[mt][2] = zm_tf32(sg[e1 * SG + mt * 16 + gq] * s1);
              a[mt][3] = zm_tf32(sg[e1 * SG + mt * 16 + gq + 8] * s1);
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              a[mt][0] = __float_as_uint(sg[e0 * SG + mt * 16 + gq]);
              a[mt][1] = __float_as_uint(sg[e0 * SG + mt * 16 + gq + 8]);
              a[mt][2] = __float_as_uint(sg[e1 * SG + mt * 16 + gq]);
              a[mt][3] = __float_as_uint(sg[e1 * SG + mt * 16 + gq + 8]);
            }
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            uint32_t b0, b1;
            if constexpr (BWD) {
              b0 = zm_tf32(sh[e0 * SH + nt * 8 + gq]);
              b1 = zm_tf32(sh[e1 * SH + nt * 8 + gq]);
            } else {
              b0 = __float_as_uint(sh[e0 * SH + nt * 8 + gq]);
              b1 = __float_as_uint(sh[e1 * SH + nt * 8 + gq]);
            }
            if (ks == 0 && fresh) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) zm_mma0(acc[mt][nt], a[mt], b0, b1);
            } else {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) zm_mma(acc[mt][nt], a[mt], b0, b1);
            }
          }
        }
      }
      if (cur.c0 + ZM_DEGC >= cur.ee) {
        const int deg = cur.ee - cur.eb;
        const float inv = BWD ? 1.0f : 1.0f / (float)(deg > 0 ? deg : 1);
        const int64_t i = i0 + cur.k;
        const int kbase = cur.p * 4 * kt;                  // first channel of this pass
        if constexpr (ZMODE == 2) {
          __half* zh = reinterpret_cast<__half*>(Zv) + i * (int64_t)zk;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (chan[mt][hh] < 0) continue;
              uint32_t* row = reinterpret_cast<uint32_t*>(zh + (kbase + chan[mt][hh]) * WP + 2 * tq);
#pragma unroll
              for (int nt = 0; nt < NT; ++nt)
              {
                const float z0 = acc[mt][nt][2 * hh] * inv, z1 = acc[mt][nt][2 * hh + 1] * inv;
                guard.note(z0, z1);
                row[nt * 4] = zm_h2_sat(z0, z1);
              }
            }
          if (cur.p == passes - 1) {
            for (int c = lane * 2; c < zk - zk_main; c += 64) {
              const float h0 = (c < WP) ? h[i * WP + c] : 0.f, h1 = (c + 1 < WP) ? h[i * WP + c + 1] : 0.f;
              guard.note(h0, h1);
              *reinterpret_cast<uint32_t*>(zh + zk_main + c) = zm_h2_sat(h0, h1);
            }
          }
        } else {
          float* zf = reinterpret_cast<float*>(Zv) + i * (int64_t)zk;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (chan[mt][hh] < 0) continue;
              float* row = zf + (kbase + chan[mt][hh]) * WP + 2 * tq;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                const uint32_t u0 = zm_tf32(acc[mt][nt][2 * hh] * inv), u1 = zm_tf32(acc[mt][nt][2 * hh + 1] * inv);
                *reinterpret_cast<float2*>(row + nt * 8) = make_float2(__uint_as_float(u0), __uint_as_float(u1));
              }
            }
          if (cur.p == passes - 1) {
            for (int c = lane; c < zk - zk_main; c += 32)
              zf[zk_main + c] = (c < WP) ? __uint_as_float(zm_tf32(h[i * WP + c])) : 0.f;
          }
        }
      }
      // the slab just consumed is overwritten by the async proxy two items from now
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      cur = nxt;
      nxt = nn2;
      src_nxt = src_nn2;
      sc_nxt = load_scale(nxt, src_nxt);
      buf ^= 1;
    }
  }
  if constexpr (ZMODE == 2) guard.flush(ovf);
}

template <int MT, int WP, int ZMODE, bool BWD>
static int launch_zm(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                     const float* h, int64_t n, void* Z, cudaStream_t s, const float* gsc) {
  constexpr size_t smem =
      (size_t)ZM_WARPS * 2 * (ZM_DEGC * (16 * MT + WP + 8) + ZM_DEGC) * sizeof(float) + ZM_WARPS * 2 * sizeof(uint64_t);
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(zbuild_mma_kernel<MT, WP, ZMODE, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_set = true;
  }
  const int64_t blocks_needed = ceil_div(ceil_div(n, ZM_TASK), ZM_WARPS);
  const int64_t cap = (int64_t)num_sms() * 2;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  ProfScope prof(PROF_ZBUILD, s);
  zbuild_mma_kernel<MT, WP, ZMODE, BWD><<<grid, ZM_WARPS * 32, smem, s>>>(rowptr, src_sorted, g, h, n, d.passes, d.kp, d.kt,
                                                                         d.ktp, d.zk_main, d.zk, gsc, Z, cur_ovf());
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <int MT, int WP>
static int zm_dispatch_mode(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                            const float* h, int64_t n, void* Z, int zmode, cudaStream_t s, int mean, const float* gsc) {
  const bool bwd = (mean == 0);
  if (bwd) {
    if (!gsc) {
      set_error("backward zbuild needs a gather scale");
      return FESR_EINVAL;
    }
    // zmode 2: the scaled fp16 Z~ of the tf32 arm's backward (backward.cu)
    if (zmode == 2) return launch_zm<MT, WP, 2, true>(d, rowptr, src_sorted, g, h, n, Z, s, gsc);
    return launch_zm<MT, WP, 1, true>(d, rowptr, src_sorted, g, h, n, Z, s, gsc);
  }
  if (zmode == 2) return launch_zm<MT, WP, 2, false>(d, rowptr, src_sorted, g, h, n, Z, s, nullptr);
  return launch_zm<MT, WP, 1, false>(d, rowptr, src_sorted, g, h, n, Z, s, nullptr);
}

template <int MT>
static int zm_dispatch_wp(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                          const float* h, int64_t n, void* Z, int zmode, cudaStream_t s, int mean, const float* gsc) {
  switch (d.wp) {
    case 16: return zm_dispatch_mode<MT, 16>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 32: return zm_dispatch_mode<MT, 32>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 48: return zm_dispatch_mode<MT, 48>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 64: return zm_dispatch_mode<MT, 64>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
  }
  set_error("unsupported padded width %d", d.wp);
  return FESR_EINVAL;
}

int launch_zbuild_mma(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                      const float* h, int64_t n, void* Z, int zmode, cudaStream_t s, int mean, const float* gsc) {
  if (n == 0) return FESR_OK;
  switch (4 * d.ktp) {
    case 16: return zm_dispatch_wp<1>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 32: return zm_dispatch_wp<2>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 48: return zm_dispatch_wp<3>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 64: return zm_dispatch_wp<4>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
  }
  set_error("unsupported g row width %d", 4 * d.ktp);
  return FESR_EINVAL;
}

}  // namespace fesr
