// Backward of KernelNN's edge-MLP hidden layers (Linear(1,w) ReLU Linear(w,w) ReLU, reference models/model.py:550 +
// :311-315) in ONE kernel for the reduced-precision arm.  Replaces, per train step, the chain
//   mlp_layer0 -> dg_to_dpre -> [E,w]x[w,w] GEMM (da0) -> act_grad -> two split-K GEMMs over E (dW1, dW0) -> three
//   column sums (db1, db0, ...)
// that streams five [E, w] fp32 arrays through HBM several times, by one pass over dg and g:
//   da1 = dg (.) relu'(g)                      (g = the forward's edge features = the layer-1 activations)
//   da0 = da1 W1,  dpre0 = da0 (.) relu'(a0)   (a0 = relu(d w0 + b0), recomputed from the edge length)
//   dW1 += da1^T a0,  db1 += 1^T da1,  dw0 += dpre0^T d,  db0 += 1^T dpre0
// One warp owns groups of 32 consecutive CSR edges; both products run on mma.sync.m16n8k8 tf32 (fp32 accumulate):
// da0 with the edges on M, the weight gradient with the edges on K and its accumulators (48 x 48 (+ a ones column
// for db1)) living in registers for the whole kernel.  Every warp writes its partial sums once; a second kernel adds
// them in warp order: deterministic.
#include "backward.cuh"

namespace fesr {

__device__ __forceinline__ uint32_t eb_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void eb_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int EB_W = 48;                       // padded width
constexpr int EB_PART = EB_W * EB_W + 3 * EB_W;   // floats per warp partial: dW1 [48][48] | db1 | dw0 | db0
constexpr int EB_S1 = EB_W + 4;                // da1 tile row stride
constexpr int EB_SW = EB_W + 8;                // W1 row stride

__global__ void __launch_bounds__(128)
edge_mlp_bwd_kernel(const float* __restrict__ w0g, const float* __restrict__ b0g, const float* __restrict__ w1g, int w,
                    int kt, int ktp, int kp, const float* __restrict__ edge_attr, const int32_t* __restrict__ perm,
                    const float* __restrict__ dg, const float* __restrict__ g, int E, float* __restrict__ partial) {
  __shared__ __align__(16) float w0s[EB_W], b0s[EB_W];
  __shared__ __align__(16) float w1s[EB_W][EB_SW];          // [o][i], tf32 bit patterns
  __shared__ __align__(16) float da1s[4][32][EB_S1];        // [warp][edge][channel o] (un-permuted), tf32 bit patterns
  // per-warp staging of the NEXT group's dg and g rows (cp.async, issued once this group's rows have been turned into
  // the da1 tile): the loads used to sit in front of every group -- 12 dependent 16-byte load pairs per lane with 8 warps
  // per SM (231 registers): 1.15 TB/s, 80 % of the stall samples long-scoreboard -- and now fly under the MMAs
  extern __shared__ __align__(16) float eb_stage[];          // [4 warps][2 arrays][32 edges][kp]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  for (int i = tid; i < EB_W; i += blockDim.x) {
    w0s[i] = i < w ? w0g[i] : 0.f;
    b0s[i] = i < w ? b0g[i] : 0.f;
  }
  for (int t = tid; t < EB_W * EB_W; t += blockDim.x) {
    const int o = t / EB_W, i = t % EB_W;
    w1s[o][i] = __uint_as_float(eb_tf32((o < w && i < w) ? w1g[o * w + i] : 0.f));
  }
  __syncthreads();
  // this lane's layer-0 constants for the B fragments of the weight-gradient product (input channel i = nt*8 + gq)
  float w0r[6], b0r[6];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    w0r[nt] = w0s[nt * 8 + gq];
    b0r[nt] = b0s[nt * 8 + gq];
  }
  float accW[3][7][4];                          // dW1[o = mt*16 + gq (+8)][i = nt*8 + 2tq (+1)]; nt = 6: column 0 = db1[o]
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) accW[mt][nt][0] = accW[mt][nt][1] = accW[mt][nt][2] = accW[mt][nt][3] = 0.f;
  float dw0[6][2], db0[6][2];                   // columns i = nt*8 + 2tq (+1), summed over this lane's rows
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) dw0[nt][0] = dw0[nt][1] = db0[nt][0] = db0[nt][1] = 0.f;
  const uint32_t one_b = (gq == 0) ? eb_tf32(1.0f) : 0u;      // B fragment of the ones column (n = 0 of tile 6)

  float (*tile)[EB_S1] = da1s[warp];
  const int n_groups = (E + 31) / 32;
  const int q4 = kp >> 2;
  float4* sdg = reinterpret_cast<float4*>(eb_stage) + (size_t)warp * 2 * 32 * q4;
  float4* sg = sdg + 32 * q4;
  auto stage_group = [&](int grp_) {        // the group's rows are contiguous: [32][kp] floats of dg and of g
    if (grp_ < n_groups) {
      const int eb_ = grp_ * 32;
      const int nrow = min(32, E - eb_);
      const float4* pd = reinterpret_cast<const float4*>(dg + (int64_t)eb_ * kp);
      const float4* pg = reinterpret_cast<const float4*>(g + (int64_t)eb_ * kp);
      for (int t = lane; t < nrow * q4; t += 32) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdg + t)), "l"(pd + t) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sg + t)), "l"(pg + t) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_group(blockIdx.x * 4 + warp);
  for (int grp = blockIdx.x * 4 + warp; grp < n_groups; grp += gridDim.x * 4) {
    const int e_base = grp * 32;
    const int e_l = e_base + lane;
    const float d_lane = e_l < E ? __ldg(edge_attr + (perm ? __ldg(perm + e_l) : e_l)) : 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- da1 = dg (.) relu'(g), un-permuted into the tile (slot layout -> channel layout)
    for (int t = lane; t < 32 * q4; t += 32) {
      const int r = t / q4, c4 = t - r * q4;
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), gv = dv;
      if (e_base + r < E) {
        dv = sdg[t];
        gv = sg[t];
      }
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int slot = 4 * c4 + j;
        const int q = slot / ktp, rr = slot - q * ktp, ch = q * kt + rr;
        if (rr < kt && ch < w) tile[r][ch] = __uint_as_float(eb_tf32(gg[j] > 0.f ? dd[j] : 0.f));
      }
    }
    if (w < EB_W)
      for (int t = lane; t < 32 * (EB_W - w); t += 32) tile[t / (EB_W - w)][w + t % (EB_W - w)] = 0.f;
    __syncwarp();
    stage_group(grp + (int)gridDim.x * 4);        // the staging rows have been consumed: bring the next group's
    // ---- da0 = da1 W1 (edges on M), dpre0 = da0 (.) relu'(a0); its column sums feed dw0 / db0
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[6][4];
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        const uint32_t a0 = __float_as_uint(tile[mt * 16 + gq][ks * 8 + tq]), a1 = __float_as_uint(tile[mt * 16 + gq + 8][ks * 8 + tq]);
        const uint32_t a2 = __float_as_uint(tile[mt * 16 + gq][ks * 8 + tq + 4]), a3 = __float_as_uint(tile[mt * 16 + gq + 8][ks * 8 + tq + 4]);
#pragma unroll
        for (int nt = 0; nt < 6; ++nt)
          eb_mma(acc[nt], a0, a1, a2, a3, __float_as_uint(w1s[ks * 8 + tq][nt * 8 + gq]), __float_as_uint(w1s[ks * 8 + tq + 4][nt * 8 + gq]));
      }
      const float d0 = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq), d1 = __shfl_sync(0xffffffffu, d_lane, mt * 16 + gq + 8);
#pragma unroll
      for (int nt = 0; nt < 6; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int i = nt * 8 + 2 * tq + j;
          const float wi = w0s[i], bi = b0s[i];
          const float p0 = fmaf(d0, wi, bi) > 0.f ? acc[nt][j] : 0.f;          // row gq
          const float p1 = fmaf(d1, wi, bi) > 0.f ? acc[nt][2 + j] : 0.f;      // row gq + 8
          dw0[nt][j] += p0 * d0 + p1 * d1;
          db0[nt][j] += p0 + p1;
        }
    }
    // ---- dW1 += da1^T a0 (edges on K; a0 recomputed into the B fragments), db1 through the ones column
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float dk0 = __shfl_sync(0xffffffffu, d_lane, ks * 8 + tq), dk1 = __shfl_sync(0xffffffffu, d_lane, ks * 8 + tq + 4);
      uint32_t b0f[6], b1f[6];
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        b0f[nt] = eb_tf32(fmaxf(fmaf(dk0, w0r[nt], b0r[nt]), 0.f));
        b1f[nt] = eb_tf32(fmaxf(fmaf(dk1, w0r[nt], b0r[nt]), 0.f));
      }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        const uint32_t a0 = __float_as_uint(tile[ks * 8 + tq][mt * 16 + gq]), a1 = __float_as_uint(tile[ks * 8 + tq][mt * 16 + gq + 8]);
        const uint32_t a2 = __float_as_uint(tile[ks * 8 + tq + 4][mt * 16 + gq]), a3 = __float_as_uint(tile[ks * 8 + tq + 4][mt * 16 + gq + 8]);
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) eb_mma(accW[mt][nt], a0, a1, a2, a3, b0f[nt], b1f[nt]);
        eb_mma(accW[mt][6], a0, a1, a2, a3, one_b, one_b);
      }
    }
    __syncwarp();
  }
  // ---- this warp's partial sums
  float* out = partial + (size_t)(blockIdx.x * 4 + warp) * EB_PART;
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 6; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
        *reinterpret_cast<float2*>(out + (mt * 16 + gq + 8 * hh) * EB_W + nt * 8 + 2 * tq) =
            make_float2(accW[mt][nt][2 * hh], accW[mt][nt][2 * hh + 1]);
  if (tq == 0) {
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      out[EB_W * EB_W + mt * 16 + gq] = accW[mt][6][0];
      out[EB_W * EB_W + mt * 16 + gq + 8] = accW[mt][6][2];
    }
  }
  // dw0 / db0: sum over the 8 row lanes (gq) in a fixed tree
#pragma unroll
  for (int nt = 0; nt < 6; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float a = dw0[nt][j], b = db0[nt][j];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (gq == 0) {
        out[EB_W * EB_W + EB_W + nt * 8 + 2 * tq + j] = a;
        out[EB_W * EB_W + 2 * EB_W + nt * 8 + 2 * tq + j] = b;
      }
    }
}

// grads += sum over the warp partials, in warp order
__global__ void edge_mlp_bwd_reduce_kernel(const float* __restrict__ partial, int n_part, int w, float* __restrict__ gw1,
                                           float* __restrict__ gb1, float* __restrict__ gw0, float* __restrict__ gb0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= EB_PART) return;
  float* dst = nullptr;
  if (idx < EB_W * EB_W) {
    const int o = idx / EB_W, i = idx % EB_W;
    if (o < w && i < w) dst = gw1 + o * w + i;
  } else {
    const int r = idx - EB_W * EB_W, which = r / EB_W, c = r % EB_W;
    if (c < w) dst = (which == 0 ? gb1 : which == 1 ? gw0 : gb0) + c;
  }
  if (!dst) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * EB_PART + idx];
  *dst += s;
}

static int eb_grid(int64_t E) {
  const int64_t blocks = ceil_div(ceil_div(E, 32), 4);
  return (int)(blocks < 2ll * num_sms() ? blocks : 2ll * num_sms());
}

bool edge_mlp_bwd_supported(const fesr_model_dims& d) {
  return d.kind == FESR_KERNELNN && d.n_hidden == 2 && d.hidden[0] == d.w && d.hidden[1] == d.w && d.w <= EB_W && !d.leaky &&
         d.kp % 4 == 0 && d.passes == 1;
}
size_t edge_mlp_bwd_ws_bytes(const fesr_model_dims& d, int64_t E) {
  if (!edge_mlp_bwd_supported(d)) return 0;
  return (size_t)eb_grid(E > 0 ? E : 1) * 4 * EB_PART * sizeof(float);
}

// dg, g: [E, kp] (slot layout); adds into grads->mlp_w[0..1], mlp_b[0..1]
int launch_edge_mlp_bwd(const fesr_model_dims& d, const fesr_params& p, const float* edge_attr, const int32_t* perm,
                        const float* dg, const float* g, int64_t E, fesr_param_grads* grads, float* ws, cudaStream_t s) {
  if (E == 0) return FESR_OK;
  const int grid = eb_grid(E);
  ProfScope prof(PROF_BACKWARD, s);
  const size_t smem = (size_t)4 * 2 * 32 * d.kp * sizeof(float);       // 48 KB at kp = 48 (+ 38 KB static): 2 CTAs per SM
  static bool attr = false;
  if (!attr) {
    FESR_CUDA(cudaFuncSetAttribute(edge_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  FESR_CHECK_ARG(smem <= 64 * 1024, "edge-MLP backward: staging exceeds 64 KB");
  edge_mlp_bwd_kernel<<<grid, 128, smem, s>>>(p.mlp_w[0], p.mlp_b[0], p.mlp_w[1], d.w, d.kt, d.ktp, d.kp, edge_attr, perm, dg, g, (int)E, ws);
  FESR_LAUNCH_CHECK();
  edge_mlp_bwd_reduce_kernel<<<(unsigned)ceil_div(EB_PART, 256), 256, 0, s>>>(ws, grid * 4, d.w, grads->mlp_w[1], grads->mlp_b[1],
                                                                            grads->mlp_w[0], grads->mlp_b[0]);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // namespace fesr
