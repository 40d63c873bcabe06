// FESR_PREC_F16 arm of the gather + segmented mean: g, h and Z all live in fp16 (11-bit mantissa,
// the same operand precision as the tf32 arm, half the bytes everywhere).
//
//   Z_i[k, a] = 1/deg_i sum_{e->i} g_e[k] h[src_e, a]   ++   h_i            (contract of zbuild.cu)
//
// Per destination node the sum of outer products is G_i^T [GROW x deg] . H_i [deg x WP]; with
// 16-bit operands it is one k-step of mma.sync.m16n8k16 (f16 in, fp32 accumulate) per 16 edges:
// 18 MMAs per node for GROW = WP = 48, fragments fetched with ldmatrix.trans straight from the
// row-per-edge staging slabs (6 ldmatrix.x4 per node).  Because a staged edge is only 2 x 96 B,
// every warp keeps THREE chunks in flight (two prefetched) inside the same shared-memory budget,
// which is what the HBM latency-bandwidth product asks for with 16 resident warps per SM.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fesr {

constexpr int ZH_WARPS = 8;
constexpr int ZH_DEGC = 16;    // edges per chunk = one m16n8k16 k-step
constexpr int ZH_TASK = 8;
constexpr int ZH_NBUF = 3;

__device__ __forceinline__ uint32_t zh_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zh_ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void zh_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void zh_mma0(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t zh_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct ZhItem {
  int k, p, c0, eb, ee;
};

// MT = (g elements per edge per pass) / 16, NT = WP / 8 (even)
template <int MT, int WP>
__global__ void __launch_bounds__(ZH_WARPS * 32, 2)
zbuild_f16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                  const __half* __restrict__ g, const __half* __restrict__ h, int64_t n, int passes, int kp, int kt,
                  int ktp, int zk_main, int zk, __half* __restrict__ Z, int* ovf) {
  F16Guard guard;
  constexpr int GROW = 16 * MT;
  constexpr int NT = WP / 8;
  constexpr int SG = GROW + 8, SH = WP + 8;                 // slab row strides in halfs (+16 B: conflict-free ldmatrix)
  constexpr int BUF = ZH_DEGC * (SG + SH);                  // halfs per buffer
  constexpr int GCH = GROW / 8, HCH = WP / 8;               // 16-byte chunks per row
  extern __shared__ __align__(16) __half smem_h[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* slab = smem_h + warp * (ZH_NBUF * BUF);
  const uint32_t slab_u32 = zh_smem(slab);
  const int gq = lane >> 2, tq = lane & 3;
  const unsigned FULL = 0xffffffffu;

  for (int t = lane; t < ZH_NBUF * BUF / 2; t += 32) reinterpret_cast<uint32_t*>(slab)[t] = 0u;   // keep stale data finite
  __syncwarp();

  int chan[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int slot = mt * 16 + gq + 8 * hh;
      const int q = slot / ktp, r = slot % ktp;
      chan[mt][hh] = (r < kt) ? q * kt + r : -1;
    }
  // ldmatrix row addresses (bytes, relative to a buffer): this lane supplies row (lane & 7) of matrix (lane >> 3)
  //   A (G^T) tile mt : matrices {k 0-7, m 0-7}, {k 0-7, m 8-15}, {k 8-15, m 0-7}, {k 8-15, m 8-15}
  //   B (H) tiles nt, nt+1 : matrices {k 0-7, nt}, {k 8-15, nt}, {k 0-7, nt+1}, {k 8-15, nt+1}
  const int lr = lane & 7, lm = lane >> 3;
  const uint32_t a_off = (uint32_t)(((lm >> 1) * 8 + lr) * SG + (lm & 1) * 8) * 2u;
  const uint32_t b_off = (uint32_t)(ZH_DEGC * SG + ((lm & 1) * 8 + lr) * SH + (lm >> 1) * 8) * 2u;

  const int64_t n_tasks = (n + ZH_TASK - 1) / ZH_TASK;
  const int64_t warp_global = (int64_t)blockIdx.x * ZH_WARPS + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * ZH_WARPS;

  for (int64_t task = warp_global; task < n_tasks; task += warp_stride) {
    const int64_t i0 = task * ZH_TASK;
    const int nn = (int)min((int64_t)ZH_TASK, n - i0);
    const int rp = (lane <= nn) ? __ldg(rowptr + i0 + lane) : 0;

    auto node_item = [&](int k) {
      ZhItem it;
      it.k = k;
      it.p = 0;
      it.eb = __shfl_sync(FULL, rp, min(k, ZH_TASK));
      it.ee = __shfl_sync(FULL, rp, min(k + 1, ZH_TASK));
      it.c0 = it.eb;
      return it;
    };
    auto advance = [&](ZhItem it) {
      if (it.k >= nn) return it;
      it.c0 += ZH_DEGC;
      if (it.c0 >= it.ee) {
        it.c0 = it.eb;
        if (++it.p == passes) return node_item(it.k + 1);
      }
      return it;
    };
    auto load_src = [&](const ZhItem& it) {
      const int e = it.c0 + lane;
      return (it.k < nn && lane < ZH_DEGC && e < it.ee) ? __ldg(src_sorted + e) : 0;
    };
    // 16-byte cp.async: GCH chunks per g row (streamed), HCH per gathered h row; always one commit
    auto issue = [&](const ZhItem& it, int buf, int src_reg) {
      if (it.k < nn) {
        const uint32_t base = slab_u32 + (uint32_t)(buf * BUF) * 2u;
        const int m = min(ZH_DEGC, it.ee - it.c0);
        const __half* grow = g + (int64_t)it.c0 * kp + it.p * GROW;
        for (int t = lane; t < ZH_DEGC * GCH; t += 32) {
          const int j = t / GCH, c = t % GCH;
          const uint32_t d = base + (uint32_t)(j * SG + c * 8) * 2u;
          if (j < m) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(grow + (int64_t)j * kp + c * 8) : "memory");
          } else {   // unused edge slots of the k-step: g row = 0 (the stale h row is finite)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0) : "memory");
          }
        }
        for (int t = lane; t < ZH_DEGC * HCH; t += 32) {
          const int j = t / HCH, c = t % HCH;
          const int s = __shfl_sync(FULL, src_reg, j);
          if (j < m) {
            const uint32_t d = base + (uint32_t)(ZH_DEGC * SG + j * SH + c * 8) * 2u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(h + (int64_t)s * WP + c * 8) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // software pipeline: item i is consumed while items i+1 and i+2 are in flight
    ZhItem cur = node_item(0);
    ZhItem nx1 = advance(cur);
    ZhItem nx2 = advance(nx1);
    issue(cur, 0, load_src(cur));
    issue(nx1, 1, load_src(nx1));
    int src_nx2 = load_src(nx2);
    int buf = 0;
    float acc[MT][NT][4];
    while (cur.k < nn) {
      const int buf2 = (buf + 2) % ZH_NBUF;
      issue(nx2, buf2, src_nx2);                    // (empty commit group when there is no such item)
      const ZhItem nx3 = advance(nx2);
      const int src_nx3 = load_src(nx3);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncwarp();
      const bool fresh = cur.c0 == cur.eb;
      const uint32_t base = slab_u32 + (uint32_t)(buf * BUF) * 2u;
      if (cur.ee > cur.c0) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) zh_ldsm4t(base + a_off + (uint32_t)(mt * 16) * 2u, a[mt]);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b[4];
          zh_ldsm4t(base + b_off + (uint32_t)(np * 16) * 2u, b);
          if (fresh) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              zh_mma0(acc[mt][2 * np], a[mt], b[0], b[1]);
              zh_mma0(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              zh_mma(acc[mt][2 * np], a[mt], b[0], b[1]);
              zh_mma(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
            }
          }
        }
      } else if (fresh) {                           // zero in-degree
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
      }
      if (cur.c0 + ZH_DEGC >= cur.ee) {
        const int deg = cur.ee - cur.eb;
        const float inv = 1.0f / (float)(deg > 0 ? deg : 1);
        const int64_t i = i0 + cur.k;
        const int kbase = cur.p * 4 * kt;
        __half* zh = Z + i * (int64_t)zk;
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
              row[nt * 4] = zh_h2_sat(z0, z1);
            }
          }
        if (cur.p == passes - 1) {                  // root block (h_i is fp16 already) + zero tail
          const uint32_t* hi = reinterpret_cast<const uint32_t*>(h + i * WP);
          uint32_t* zr = reinterpret_cast<uint32_t*>(zh + zk_main);
          for (int c = lane; c < (zk - zk_main) / 2; c += 32) zr[c] = (c < WP / 2) ? hi[c] : 0u;
        }
      }
      __syncwarp();
      cur = nx1;
      nx1 = nx2;
      nx2 = nx3;
      src_nx2 = src_nx3;
      buf = (buf + 1) % ZH_NBUF;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  guard.flush(ovf);
}

template <int MT, int WP>
static int launch_zh(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g,
                     const void* h, int64_t n, void* Z, cudaStream_t s) {
  constexpr size_t smem = (size_t)ZH_WARPS * ZH_NBUF * ZH_DEGC * (16 * MT + 8 + WP + 8) * sizeof(__half);
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(zbuild_f16_kernel<MT, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t blocks_needed = ceil_div(ceil_div(n, ZH_TASK), ZH_WARPS);
  const int64_t cap = (int64_t)num_sms() * 2;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  ProfScope prof(PROF_ZBUILD, s);
  zbuild_f16_kernel<MT, WP><<<grid, ZH_WARPS * 32, smem, s>>>(rowptr, src_sorted, static_cast<const __half*>(g),
                                                             static_cast<const __half*>(h), n, d.passes, d.kp, d.kt,
                                                             d.ktp, d.zk_main, d.zk, static_cast<__half*>(Z), cur_ovf());
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <int MT>
static int zh_dispatch_wp(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g,
                          const void* h, int64_t n, void* Z, cudaStream_t s) {
  switch (d.wp) {
    case 16: return launch_zh<MT, 16>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 32: return launch_zh<MT, 32>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 48: return launch_zh<MT, 48>(d, rowptr, src_sorted, g, h, n, Z, s);
    case 64: return launch_zh<MT, 64>(d, rowptr, src_sorted, g, h, n, Z, s);
  }
  set_error("unsupported padded width %d", d.wp);
  return FESR_EINVAL;
}

int launch_zbuild_f16(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g_half,
                      const void* h_half, int64_t n, void* Z_half, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  switch (4 * d.ktp) {
    case 16: return zh_dispatch_wp<1>(d, rowptr, src_sorted, g_half, h_half, n, Z_half, s);
    case 32: return zh_dispatch_wp<2>(d, rowptr, src_sorted, g_half, h_half, n, Z_half, s);
    case 48: return zh_dispatch_wp<3>(d, rowptr, src_sorted, g_half, h_half, n, Z_half, s);
    case 64: return zh_dispatch_wp<4>(d, rowptr, src_sorted, g_half, h_half, n, Z_half, s);
  }
  set_error("unsupported g row width %d", 4 * d.ktp);
  return FESR_EINVAL;
}

}  // namespace fesr
