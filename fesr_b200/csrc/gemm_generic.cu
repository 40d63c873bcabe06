// Generic strided fp32 GEMM with deterministic split-K, used by the backward pass for the
// reductions over nodes / edges (weight gradients) and the thin input-gradient products.
//   C[M,N] (+)= A[M,K] * B[K,N]       element (i,j) of X at X[i*sXrow + j*sXcol]
#include "backward.cuh"

namespace fesr {

constexpr int GG_BM = 64, GG_BN = 64, GG_BK = 16, GG_THREADS = 256;

__device__ __forceinline__ uint32_t gg_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// TF32 = true: the 64 x 64 x 16 slab product on mma.sync.m16n8k8 (warp = 16 rows x 32 columns) instead of the 4 x 4
// register tiles; same tiles, same split-K, same fixed-order reduce
template <bool TF32>
__global__ void __launch_bounds__(GG_THREADS)
gemm_generic_kernel(GemmArgs a, int64_t kchunk, float* __restrict__ partial) {
  __shared__ __align__(16) float As[GG_BK][GG_BM + 4];
  __shared__ __align__(16) float Bs[GG_BK][GG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GG_BM, n0 = (int64_t)blockIdx.y * GG_BN;
  const int64_t kb = (int64_t)blockIdx.z * kchunk;
  const int64_t ke = min(a.K, kb + kchunk);
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;      // TF32: this warp's corner of the tile
  float acc[4][4];                                            // TF32: acc[n-tile][c0..c3]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_m_fast = a.sAm == 1, b_n_fast = a.sBn == 1;
  for (int64_t k0 = kb; k0 < ke; k0 += GG_BK) {
#pragma unroll
    for (int it = 0; it < (GG_BM * GG_BK) / GG_THREADS; ++it) {
      const int idx = tid + it * GG_THREADS;
      const int mm = a_m_fast ? (idx % GG_BM) : (idx / GG_BK);
      const int kk = a_m_fast ? (idx / GG_BM) : (idx % GG_BK);
      const int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < a.M && gk < ke) ? a.A[gm * a.sAm + gk * a.sAk] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < (GG_BN * GG_BK) / GG_THREADS; ++it) {
      const int idx = tid + it * GG_THREADS;
      const int nn = b_n_fast ? (idx % GG_BN) : (idx / GG_BK);
      const int kk = b_n_fast ? (idx / GG_BN) : (idx % GG_BK);
      const int64_t gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < a.N && gk < ke) ? a.B[gk * a.sBk + gn * a.sBn] : 0.f;
    }
    __syncthreads();
    if (TF32) {
#pragma unroll
      for (int ks = 0; ks < GG_BK / 8; ++ks) {
        const uint32_t a0 = gg_tf32(As[ks * 8 + tq][wm + gq]), a1 = gg_tf32(As[ks * 8 + tq][wm + gq + 8]);
        const uint32_t a2 = gg_tf32(As[ks * 8 + tq + 4][wm + gq]), a3 = gg_tf32(As[ks * 8 + tq + 4][wm + gq + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t b0 = gg_tf32(Bs[ks * 8 + tq][wn + nt * 8 + gq]), b1 = gg_tf32(Bs[ks * 8 + tq + 4][wn + nt * 8 + gq]);
          asm volatile(
              "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
      }
    } else {
#pragma unroll
    for (int kk = 0; kk < GG_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // TF32: i = n-tile, j = accumulator element (rows gq / gq + 8, columns 2 tq / 2 tq + 1)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gm = TF32 ? m0 + wm + gq + 8 * (j >> 1) : m0 + tm + i;
      const int64_t gn = TF32 ? n0 + wn + i * 8 + 2 * tq + (j & 1) : n0 + tn + j;
      if (gm >= a.M || gn >= a.N) continue;
      if (partial) {
        partial[((int64_t)blockIdx.z * a.M + gm) * a.N + gn] = acc[i][j];
      } else {
        float* c = a.C + gm * a.sCm + gn * a.sCn;
        *c = a.accumulate ? *c + acc[i][j] : acc[i][j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ks, GemmArgs a) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= a.M * a.N) return;
  float s = 0.f;
  for (int z = 0; z < ks; ++z) s += partial[(int64_t)z * a.M * a.N + idx];   // fixed order
  const int64_t m = idx / a.N, n = idx % a.N;
  float* c = a.C + m * a.sCm + n * a.sCn;
  *c = a.accumulate ? *c + s : s;
}

int gemm_pick_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(M, GG_BM) * ceil_div(N, GG_BN);
  int64_t want = (4 * (int64_t)num_sms() + tiles - 1) / tiles;
  const int64_t maxk = ceil_div(K, 8 * GG_BK);
  if (want > maxk) want = maxk;
  if (want > 256) want = 256;
  return (int)(want < 1 ? 1 : want);
}

size_t gemm_ws_bytes(int64_t M, int64_t N, int64_t K) {
  const int ks = gemm_pick_splits(M, N, K);
  return ks > 1 ? (size_t)ks * M * N * sizeof(float) : 0;
}

int launch_gemm(const GemmArgs& a, float* ws, size_t ws_bytes, cudaStream_t s) {
  if (a.M == 0 || a.N == 0) return FESR_OK;
  int ks = gemm_pick_splits(a.M, a.N, a.K);
  if (ks > 1 && (!ws || ws_bytes < (size_t)ks * a.M * a.N * sizeof(float))) ks = 1;
  int64_t kchunk = ceil_div(ceil_div(a.K, ks), GG_BK) * GG_BK;
  if (kchunk == 0) kchunk = GG_BK;
  ks = (int)ceil_div(a.K > 0 ? a.K : 1, kchunk);
  dim3 grid((unsigned)ceil_div(a.M, GG_BM), (unsigned)ceil_div(a.N, GG_BN), (unsigned)ks);
  if (a.tf32) gemm_generic_kernel<true><<<grid, GG_THREADS, 0, s>>>(a, kchunk, ks > 1 ? ws : nullptr);
  else gemm_generic_kernel<false><<<grid, GG_THREADS, 0, s>>>(a, kchunk, ks > 1 ? ws : nullptr);
  FESR_LAUNCH_CHECK();
  if (ks > 1) {
    splitk_reduce_kernel<<<(unsigned)ceil_div(a.M * a.N, 256), 256, 0, s>>>(ws, ks, a);
    FESR_LAUNCH_CHECK();
  }
  return FESR_OK;
}

// out[j] (+)= sum_i X[i*ld + j], j < cols; deterministic two-stage (rows split over blocks)
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ X, int64_t rows, int cols, int64_t ld, int64_t rchunk,
                      float* __restrict__ partial) {
  // 64 columns x 4 row groups per block; a thread walks its rows 16 apart with four independent accumulators (four
  // loads in flight), the groups are combined in a fixed order: deterministic
  __shared__ float red[4][64];
  const int jl = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int j = blockIdx.y * 64 + jl;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < cols) {
    const int64_t r0 = (int64_t)blockIdx.x * rchunk, r1 = min(rows, r0 + rchunk);
    int64_t i = r0 + rg;
    for (; i + 12 < r1; i += 16) {
      s0 += X[i * ld + j];
      s1 += X[(i + 4) * ld + j];
      s2 += X[(i + 8) * ld + j];
      s3 += X[(i + 12) * ld + j];
    }
    for (; i < r1; i += 4) s0 += X[i * ld + j];
  }
  red[rg][jl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rg == 0 && j < cols) partial[(int64_t)blockIdx.x * cols + j] = (red[0][jl] + red[1][jl]) + (red[2][jl] + red[3][jl]);
}
// 64 columns x 16 groups per block: group q adds the partials b = q, q + 16, ... with four independent accumulators (a
// single chain of nb dependent adds took 32 us at nb = 512), the groups are combined in a fixed order: deterministic
constexpr int CF_GROUPS = 16;
__global__ void __launch_bounds__(64 * CF_GROUPS)
colsum_final_kernel(const float* __restrict__ partial, int nb, int cols, int accumulate, float* __restrict__ out) {
  __shared__ float red[CF_GROUPS][64];
  const int jl = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + jl;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < cols) {
    int b = q;
    for (; b + 3 * CF_GROUPS < nb; b += 4 * CF_GROUPS) {
      s0 += partial[(int64_t)b * cols + j];
      s1 += partial[(int64_t)(b + CF_GROUPS) * cols + j];
      s2 += partial[(int64_t)(b + 2 * CF_GROUPS) * cols + j];
      s3 += partial[(int64_t)(b + 3 * CF_GROUPS) * cols + j];
    }
    for (; b < nb; b += CF_GROUPS) s0 += partial[(int64_t)b * cols + j];
  }
  red[q][jl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (q == 0 && j < cols) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < CF_GROUPS; g += 4) s += (red[g][jl] + red[g + 1][jl]) + (red[g + 2][jl] + red[g + 3][jl]);
    out[j] = accumulate ? out[j] + s : s;
  }
}

// second stage alone (the caller produced `nb` rows of partial column sums itself)
int launch_colsum_final(const float* partial, int nb, int cols, int accumulate, float* out, cudaStream_t s) {
  if (cols == 0) return FESR_OK;
  colsum_final_kernel<<<(unsigned)ceil_div(cols, 64), 64 * CF_GROUPS, 0, s>>>(partial, nb, cols, accumulate, out);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

size_t colsum_ws_bytes(int cols) { return (size_t)COLSUM_BLOCKS * cols * sizeof(float); }

int launch_colsum(const float* X, int64_t rows, int cols, int64_t ld, int accumulate, float* out, float* ws,
                  cudaStream_t s) {
  if (cols == 0) return FESR_OK;
  int nb = (int)(rows < COLSUM_BLOCKS ? (rows > 0 ? rows : 1) : COLSUM_BLOCKS);
  const int64_t rchunk = ceil_div(rows > 0 ? rows : 1, nb);
  nb = (int)ceil_div(rows > 0 ? rows : 1, rchunk);
  dim3 grid(nb, (unsigned)ceil_div(cols, 64));
  colsum_partial_kernel<<<grid, 256, 0, s>>>(X, rows, cols, ld, rchunk, ws);
  FESR_LAUNCH_CHECK();
  colsum_final_kernel<<<(unsigned)ceil_div(cols, 64), 64 * CF_GROUPS, 0, s>>>(ws, nb, cols, accumulate, out);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // namespace fesr
