// Hidden layers of KernelNN's edge MLP (Linear(1,w) act Linear(w,w) act, reference
// models/model.py:550 + :311-315) with the [32 edges, w] x [w, w] product on warp-level tensor
// cores at fp32-class accuracy (3xTF32: a = a_hi + a_lo, W = W_hi + W_lo, the three significant
// partial products accumulated in fp32; error ~2^-21, inside the 1e-5 gate of the fp32 arm).
// One warp owns 32 consecutive CSR edges: layer 0 is evaluated straight into MMA A-fragments
// (no shared-memory round trip), W_hi/W_lo live in shared memory with a conflict-free row pad,
// the g rows are staged in shared memory and written with coalesced 128-bit stores.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fesr {

__device__ __forceinline__ uint32_t em_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void em_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float em_act(float v, int leaky) { return leaky ? (v > 0.f ? v : 0.01f * v) : fmaxf(v, 0.f); }

template <int WPAD>
__global__ void __launch_bounds__(128)
edge_hidden2_mma_kernel(const float* __restrict__ w0g, const float* __restrict__ b0g, const float* __restrict__ w1g,
                        const float* __restrict__ b1g, int w, int leaky, int kt, int ktp, int kp, int k1,
                        const float* __restrict__ edge_attr, const int32_t* __restrict__ perm, int64_t E,
                        int round_tf32, float* __restrict__ g) {
  constexpr int KS = WPAD / 8, NT = WPAD / 8, SB = WPAD + 8;
  __shared__ __align__(16) float w0[WPAD], b0[WPAD], b1[WPAD];
  __shared__ __align__(16) uint32_t whi[WPAD][SB], wlo[WPAD][SB];     // [in][out], tf32 bit patterns
  __shared__ int off_of[WPAD + 1];
  extern __shared__ __align__(16) float stage[];                        // [4 warps][32 edges][kp + 4]
  const int sstride = kp + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  for (int i = tid; i < WPAD; i += blockDim.x) {
    w0[i] = i < w ? w0g[i] : 0.f;
    b0[i] = i < w ? b0g[i] : 0.f;
    b1[i] = i < w ? b1g[i] : 0.f;
  }
  for (int i = tid; i < WPAD * WPAD; i += blockDim.x) {
    const int in = i / WPAD, out = i % WPAD;
    const float v = (in < w && out < w) ? w1g[out * w + in] : 0.f;
    const uint32_t hi = em_tf32(v);
    whi[in][out] = hi;
    wlo[in][out] = em_tf32(v - __uint_as_float(hi));
  }
  for (int k = tid; k <= WPAD; k += blockDim.x) off_of[k] = (k / kt) * ktp + (k % kt);
  __syncthreads();

  float* wst = stage + (size_t)warp * 32 * sstride;
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
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
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
          alo[mt][r] = em_tf32(v[r] - __uint_as_float(ahi[mt][r]));
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t bh0 = whi[i0][nt * 8 + gq], bh1 = whi[i1][nt * 8 + gq];
        const uint32_t bl0 = wlo[i0][nt * 8 + gq], bl1 = wlo[i1][nt * 8 + gq];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          em_mma(acc[mt][nt], alo[mt], bh0, bh1);
          em_mma(acc[mt][nt], ahi[mt], bl0, bl1);
          em_mma(acc[mt][nt], ahi[mt], bh0, bh1);
        }
      }
    }
    // epilogue: + bias, activation, (tf32 round), into the padded / permuted g row layout
    for (int t = lane; t < 32 * (sstride / 4); t += 32) reinterpret_cast<float4*>(wst)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int c = nt * 8 + 2 * tq + (r & 1);
          const int row = mt * 16 + gq + 8 * (r >> 1);
          if (c < w) {
            float v = em_act(acc[mt][nt][r] + b1[c], leaky);
            if (round_tf32 == 1) v = __uint_as_float(em_tf32(v));
            wst[row * sstride + off_of[c]] = v;
          }
        }
    wst[lane * sstride + off_of[k1 - 1]] = 1.f;
    __syncwarp();
    if (round_tf32 == 2) {                       // fp16 rows (FESR_PREC_F16): 8 halfs = 16 bytes per store
      __half* gh = reinterpret_cast<__half*>(g);
      const int q8 = kp >> 3;
      for (int t = lane; t < 32 * q8; t += 32) {
        const int r = t / q8, c8 = t - r * q8;
        if (e_base + r < E) {
          const float4 lo = *reinterpret_cast<const float4*>(wst + r * sstride + 8 * c8);
          const float4 hi = *reinterpret_cast<const float4*>(wst + r * sstride + 8 * c8 + 4);
          __half2 p0 = __floats2half2_rn(lo.x, lo.y), p1 = __floats2half2_rn(lo.z, lo.w);
          __half2 p2 = __floats2half2_rn(hi.x, hi.y), p3 = __floats2half2_rn(hi.z, hi.w);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0);
          pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2);
          pk.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(gh + (e_base + r) * kp + 8 * c8) = pk;
        }
      }
    } else {
      const int q4 = kp >> 2;
      for (int t = lane; t < 32 * q4; t += 32) {
        const int r = t / q4, c4 = t - r * q4;
        if (e_base + r < E)
          *reinterpret_cast<float4*>(g + (e_base + r) * kp + 4 * c4) = *reinterpret_cast<const float4*>(wst + r * sstride + 4 * c4);
      }
    }
    __syncwarp();
  }
}

template <int WPAD>
static int launch_eh2m(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                       int64_t E, float* g, cudaStream_t s, int round_tf32) {
  const size_t stage_bytes = (size_t)4 * 32 * (d.kp + 4) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(edge_hidden2_mma_kernel<WPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_set = true;
  }
  const int64_t blocks = ceil_div(ceil_div(E, 32), 4);
  const int grid = (int)(blocks < 8ll * num_sms() ? blocks : 8ll * num_sms());
  ProfScope prof(PROF_EDGE_HIDDEN, s);
  edge_hidden2_mma_kernel<WPAD><<<grid, 128, stage_bytes, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], p.mlp_b[1], d.w, d.leaky,
                                                              d.kt, d.ktp, d.kp, d.k1, edge_attr, perm, E, round_tf32, g);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

int launch_edge_hidden2_mma(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                            int64_t E, float* g, cudaStream_t s, int round_tf32) {
  if (E == 0) return FESR_OK;
  if (d.w <= 16) return launch_eh2m<16>(d, p, edge_attr, perm, E, g, s, round_tf32);
  if (d.w <= 32) return launch_eh2m<32>(d, p, edge_attr, perm, E, g, s, round_tf32);
  if (d.w <= 48) return launch_eh2m<48>(d, p, edge_attr, perm, E, g, s, round_tf32);
  return launch_eh2m<64>(d, p, edge_attr, perm, E, g, s, round_tf32);
}

}  // namespace fesr
