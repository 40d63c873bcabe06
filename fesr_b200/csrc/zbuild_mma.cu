// Gather + segmented mean with the per-node outer products on warp-level tensor-core MMAs.
//
// Same contract as zbuild.cu (Z_i = 1/deg_i sum_{e->i} g_e (x) h[src_e]  ++  h_i), used by the
// reduced-precision arms (FESR_PREC_TF32 / FESR_PREC_F16) where Z is consumed as an 11-bit-mantissa
// operand anyway.  Per destination node the sum of outer products is a tiny GEMM
//     Z_i [GROW x WP] = G_i^T [GROW x deg] . H_i [deg x WP]          (deg ~ 12, GROW = WP = 48)
// with a contraction length of one node's in-degree: far below tcgen05's M >= 64 / one-CTA-wide
// issue granularity, so it is issued as mma.sync.m16n8k8 (tf32 in, fp32 accumulate) by the warp
// that owns the node.  This cuts the instruction count of the kernel from ~135 to ~30 per edge
// and leaves it bound by HBM (g read + Z write).  The fp32 arm keeps the FFMA kernel (zbuild.cu).
//
// Staging is as in zbuild.cu: per-warp double-buffered slabs, g rows streamed with cp.async.cg,
// h[src] rows gathered with cp.async.ca, source ids fetched two chunks ahead; slab rows are padded
// by 8 floats so that the MMA fragment loads are bank-conflict free.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fesr {

constexpr int ZM_WARPS = 8;
constexpr int ZM_DEGC = 16;   // edges per chunk = 2 MMA k-steps
constexpr int ZM_TASK = 8;

__device__ __forceinline__ void zm_cp_async16(float* dst_smem, const float* src, bool l1) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  if (l1)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t zm_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void zm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct ZmItem {
  int k, p, c0, eb, ee;
};

// MT = (floats of g per edge per pass) / 16, NT = WP / 8
template <int MT, int WP>
__global__ void __launch_bounds__(ZM_WARPS * 32, 2)
zbuild_mma_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                  const float* __restrict__ g, const float* __restrict__ h, int64_t n, int passes, int kp, int kt,
                  int ktp, int zk_main, int zk, int zmode, int mean, const float* __restrict__ gather_scale,
                  void* __restrict__ Zv) {
  constexpr int GROW = 16 * MT;
  constexpr int NT = WP / 8;
  constexpr int SG = GROW + 8, SH = WP + 8;                 // padded slab row strides (floats)
  constexpr int BUF = ZM_DEGC * (SG + SH) + ZM_DEGC;        // + per-edge gather scale
  constexpr int LPR = WP / 4, RPI = 32 / LPR;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slab = smem + warp * (2 * BUF);
  const int gq = lane >> 2, tq = lane & 3;                  // MMA fragment coordinates
  const int hj = lane / LPR, hc = lane % LPR;
  const bool h_lane = lane < RPI * LPR;
  const unsigned FULL = 0xffffffffu;

  for (int t = lane; t < 2 * BUF; t += 32) slab[t] = 0.f;   // stale slab contents must stay finite
  __syncwarp();

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
      return (gather_scale && it.k < nn && lane < ZM_DEGC && it.c0 + lane < it.ee) ? __ldg(gather_scale + src_reg) : 1.f;
    };
    auto issue = [&](const ZmItem& it, int buf, int src_reg, float sc_reg) {
      float* sg = slab + buf * BUF;
      float* sh = sg + ZM_DEGC * SG;
      const int m = min(ZM_DEGC, it.ee - it.c0);
      if (gather_scale && lane < ZM_DEGC) sh[ZM_DEGC * SH + lane] = sc_reg;
      for (int t = lane; t < m * (GROW / 4); t += 32) {
        const int j = t / (GROW / 4), c = t % (GROW / 4);
        zm_cp_async16(sg + j * SG + 4 * c, g + (int64_t)(it.c0 + j) * kp + it.p * GROW + 4 * c, false);
      }
      // zero the g rows of the unused edge slots of the k-steps that will be issued
      const int mpad = (m + 7) & ~7;
      for (int t = lane + m * (GROW / 4); t < mpad * (GROW / 4); t += 32) {
        const int j = t / (GROW / 4), c = t % (GROW / 4);
        *reinterpret_cast<float4*>(sg + j * SG + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int j0 = 0; j0 < m; j0 += RPI) {
        const int j = j0 + hj;
        const int s = __shfl_sync(FULL, src_reg, j & 31);
        if (h_lane && j < m) zm_cp_async16(sh + j * SH + 4 * hc, h + (int64_t)s * WP + 4 * hc, true);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
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
      __syncwarp();
      if (cur.c0 == cur.eb) {
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
      for (int ks = 0; ks * 8 < m; ++ks) {
        const int e0 = ks * 8 + tq, e1 = e0 + 4;            // this lane's two edge rows of the k-step
        float s0 = 1.f, s1 = 1.f;
        if (gather_scale) {
          s0 = sh[ZM_DEGC * SH + e0];
          s1 = sh[ZM_DEGC * SH + e1];
        }
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          a[mt][0] = zm_tf32(sg[e0 * SG + mt * 16 + gq] * s0);
          a[mt][1] = zm_tf32(sg[e0 * SG + mt * 16 + gq + 8] * s0);
          a[mt][2] = zm_tf32(sg[e1 * SG + mt * 16 + gq] * s1);
          a[mt][3] = zm_tf32(sg[e1 * SG + mt * 16 + gq + 8] * s1);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint32_t b0 = zm_tf32(sh[e0 * SH + nt * 8 + gq]);
          const uint32_t b1 = zm_tf32(sh[e1 * SH + nt * 8 + gq]);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) zm_mma(acc[mt][nt], a[mt], b0, b1);
        }
      }
      if (cur.c0 + ZM_DEGC >= cur.ee) {
        const int deg = cur.ee - cur.eb;
        const float inv = mean ? 1.0f / (float)(deg > 0 ? deg : 1) : 1.0f;
        const int64_t i = i0 + cur.k;
        const int kbase = cur.p * 4 * kt;                  // first channel of this pass
        if (zmode == 2) {
          __half* zh = reinterpret_cast<__half*>(Zv) + i * (int64_t)zk;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (chan[mt][hh] < 0) continue;
              __half* row = zh + (kbase + chan[mt][hh]) * WP + 2 * tq;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                const float v0 = fminf(fmaxf(acc[mt][nt][2 * hh] * inv, -65504.f), 65504.f);
                const float v1 = fminf(fmaxf(acc[mt][nt][2 * hh + 1] * inv, -65504.f), 65504.f);
                *reinterpret_cast<__half2*>(row + nt * 8) = __floats2half2_rn(v0, v1);
              }
            }
          if (cur.p == passes - 1) {
            for (int c = lane * 2; c < zk - zk_main; c += 64) {
              const float h0 = (c < WP) ? h[i * WP + c] : 0.f, h1 = (c + 1 < WP) ? h[i * WP + c + 1] : 0.f;
              *reinterpret_cast<__half2*>(zh + zk_main + c) =
                  __floats2half2_rn(fminf(fmaxf(h0, -65504.f), 65504.f), fminf(fmaxf(h1, -65504.f), 65504.f));
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
      __syncwarp();
      cur = nxt;
      nxt = nn2;
      src_nxt = src_nn2;
      sc_nxt = load_scale(nxt, src_nxt);
      buf ^= 1;
    }
  }
}

template <int MT, int WP>
static int launch_zm(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                     const float* h, int64_t n, void* Z, int zmode, cudaStream_t s, int mean, const float* gsc) {
  constexpr size_t smem = (size_t)ZM_WARPS * 2 * (ZM_DEGC * (16 * MT + 8 + WP + 8) + ZM_DEGC) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(zbuild_mma_kernel<MT, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t blocks_needed = ceil_div(ceil_div(n, ZM_TASK), ZM_WARPS);
  const int64_t cap = (int64_t)num_sms() * 2;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  ProfScope prof(PROF_ZBUILD, s);
  zbuild_mma_kernel<MT, WP><<<grid, ZM_WARPS * 32, smem, s>>>(rowptr, src_sorted, g, h, n, d.passes, d.kp, d.kt, d.ktp,
                                                             d.zk_main, d.zk, zmode, mean, gsc, Z);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <int MT>
static int zm_dispatch_wp(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const float* g,
                          const float* h, int64_t n, void* Z, int zmode, cudaStream_t s, int mean, const float* gsc) {
  switch (d.wp) {
    case 16: return launch_zm<MT, 16>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 32: return launch_zm<MT, 32>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 48: return launch_zm<MT, 48>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
    case 64: return launch_zm<MT, 64>(d, rowptr, src_sorted, g, h, n, Z, zmode, s, mean, gsc);
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
