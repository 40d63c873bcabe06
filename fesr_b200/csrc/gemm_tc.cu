// Tensor-core node contraction  h' = act(Z x T' + bias)  on tcgen05 (kind::tf32), sm_100a.
//
// This is the last Linear layer of the reference's edge MLP ([E,K] x [K,w*w], models/model.py:
// 428/528) plus the root weight (:445/:533), moved from "per edge per layer" to "per node per
// layer" by the reordering described in zbuild.cu -- a dense [n, zk] x [zk, wp] GEMM:
//   A = Z       [n,  zk]  fp32 (tf32-rounded by the producer), K-major, TMA-loaded 128x32 boxes
//   B = T'^T    [wp, zk]  tf32-rounded, K-major, TMA-loaded wp x 32 boxes (L2 resident)
//   D           [128, wp] fp32 accumulators in TMEM (two stages, so the epilogue of tile t
//                         overlaps the main loop of tile t+1)
// Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected lane), warps 2..5 = epilogue (tcgen05.ld -> +bias -> activation -> global).
// The kernel is HBM-bound on streaming Z (AI = 2*wp/4 = 24 flop/B), so the 8-deep TMA ring
// (176 KB in flight per SM) is what matters, not MMA issue rate.
//
// TERMS (multi-term operands; the tensor pipe has the slack, the kernel streams Z once either way):
//   1  Z x T'                                   training forward (the backward differentiates exactly this product)
//   2  Z x T'hi + Z x T'lo                      predict, tf32 / f16 arms: the weights enter with ~21 bits -- rounding T'
//                                               to one 11-bit term is a COHERENT error (same weights on every node, every
//                                               layer) and was the largest term of these arms' error
//   3  Zhi x T'hi + Zlo x T'hi + Zhi x T'lo     the fp32 arm (3xTF32): Z arrives as fp32; four converter warps split every
//                                               landed A tile into tf32 hi (low mantissa bits cleared, in place) and lo
//                                               (a second tile) before the MMA warp may touch the stage; fp32 accumulate
//                                               in TMEM.  The tensor core's fp32 accumulation rounds toward zero, so a
//                                               K = 2176 chain drifts by ~K 2^-24 (measured: 6e-6 on the KernelNN
//                                               field against 2e-7 for RN FFMA): K is therefore cut into TC_CHUNKS
//                                               accumulators of <= 17 k-blocks that the epilogue adds in RN fp32.
//                                               Shapes with zk > 4096 (TEECNet, K = 6400) stay on the CUDA-core GEMM
//                                               (gemm_simt.cu), which is also the test cross-check (FESR_FP32_SIMT=1).
#include <cuda.h>
#include <cuda_fp16.h>

#include "backward.cuh"

namespace fesr {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;           // fp32 elements = 128 bytes = one SWIZZLE_128B row
constexpr int TC_THREADS = 192;     // 6 warps (+ 4 converter warps when TERMS == 3)
// as many pipeline stages (<= 8) as fit the 227 KB of shared memory an SM offers one CTA
__host__ __device__ constexpr int tc_stage_bytes(int terms, int wp) {
  return (terms == 3 ? 2 : 1) * TC_BM * TC_BK * 4 + (terms >= 2 ? 2 : 1) * wp * TC_BK * 4;
}
__host__ __device__ constexpr int tc_stages(int terms, int wp) {
  return (227 * 1024 - 1280) / tc_stage_bytes(terms, wp) < 8 ? (227 * 1024 - 1280) / tc_stage_bytes(terms, wp) : 8;
}
constexpr int TC_ACC_COLS = 64;     // TMEM columns per accumulator (>= wp)
constexpr int TC_CHUNKS = 4;        // TERMS == 3: accumulators per tile, each over 1 / TC_CHUNKS of K

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version = 1):
// rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                 // LBO (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}

template <int WP, bool HALF, int TERMS>
__global__ void __launch_bounds__(TC_THREADS + (TERMS == 3 ? 128 : 0), 1)
node_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmBlo, const float* __restrict__ bias_p, int64_t n, int zk,
                      int w, int epi, int round_out, float* __restrict__ h_out, int* ovf) {
  static_assert(!(HALF && TERMS == 3), "the three-term form is the fp32 arm");
  F16Guard guard;
  constexpr int TC_STAGES = tc_stages(TERMS, WP);
  constexpr int NCH = TERMS == 3 ? TC_CHUNKS : 1;                 // accumulators per tile
  constexpr uint32_t STAGE_COLS = NCH * TC_ACC_COLS;              // TMEM columns of one accumulator stage
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
  constexpr uint32_t B_BYTES = WP * TC_BK * 4;      // wp * 128 B
  constexpr uint32_t TX_BYTES = A_BYTES + (TERMS >= 2 ? 2 : 1) * B_BYTES;                   // what TMA lands per stage
  constexpr uint32_t STAGE_BYTES = (TERMS == 3 ? 2 : 1) * A_BYTES + (TERMS >= 2 ? 2 : 1) * B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_alo = smem_a + TC_STAGES * A_BYTES;                                          // TERMS == 3 only
  uint8_t* smem_b = smem_alo + (TERMS == 3 ? TC_STAGES * A_BYTES : 0);
  uint8_t* smem_blo = smem_b + TC_STAGES * B_BYTES;                                          // TERMS >= 2 only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* split_bar = empty_bar + TC_STAGES;      // TERMS == 3: the converter warps are done with the stage
  uint64_t* tmem_full = split_bar + TC_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + TC_BM - 1) / TC_BM;
  constexpr int ELEMS_PER_KB = HALF ? 2 * TC_BK : TC_BK;   // one k-block is always 128 bytes per row
  const int n_kb = zk / ELEMS_PER_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (TERMS >= 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM allocation: 2 accumulator stages
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * STAGE_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = (int)(tile * TC_BM);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], TX_BYTES);
          tma_load_2d(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], kb * ELEMS_PER_KB, row0);
          tma_load_2d(smem_b + stage * B_BYTES, &tmB, &full_bar[stage], kb * ELEMS_PER_KB, 0);
          if (TERMS >= 2) tma_load_2d(smem_blo + stage * B_BYTES, &tmBlo, &full_bar[stage], kb * ELEMS_PER_KB, 0);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32, A = B = TF32 (2) or F16 (0), both K-major, N = WP, M = 128
    constexpr uint32_t fmt = HALF ? 0u : 2u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(WP >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int kb_per_chunk = (n_kb + NCH - 1) / NCH;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(TERMS == 3 ? &split_bar[stage] : &full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int chunk = kb / kb_per_chunk;
        const uint32_t tmem_d = tmem_base + as * STAGE_COLS + chunk * TC_ACC_COLS;
        const int kb0 = kb - chunk * kb_per_chunk;          // first k-block of an accumulator starts it from zero
        if (elect_one()) {
          const uint64_t adesc = make_sw128_desc(smem_u32(smem_a + stage * A_BYTES));
          const uint64_t bdesc = make_sw128_desc(smem_u32(smem_b + stage * B_BYTES));
          const uint64_t alodesc = make_sw128_desc(smem_u32(smem_alo + stage * A_BYTES));
          const uint64_t blodesc = make_sw128_desc(smem_u32(smem_blo + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {  // 32 bytes (8 tf32 / 16 f16) per MMA: advance the start address by 2 (x16 B)
            if constexpr (HALF) {
              umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb0 | k) != 0);
              if (TERMS >= 2) umma_f16(tmem_d, adesc + 2 * k, blodesc + 2 * k, idesc, 1);
            } else {
              // small terms first, so that they are not absorbed one by one into a large running sum
              if (TERMS == 3) umma_tf32(tmem_d, alodesc + 2 * k, bdesc + 2 * k, idesc, (kb0 | k) != 0);
              if (TERMS >= 2) umma_tf32(tmem_d, adesc + 2 * k, blodesc + 2 * k, idesc, TERMS == 3 || (kb0 | k) != 0);
              umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, TERMS >= 2 || (kb0 | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);              // frees the smem slot when these MMAs retire
          if (kb == n_kb - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == TC_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 6) {
    // ===== converter warps (TERMS == 3): split every landed fp32 A tile into tf32 hi (in place) + lo (second tile).
    // The swizzle only permutes 16-byte chunks inside a tile, and hi / lo are elementwise, so the tile is walked as
    // flat float4's: the lo tile gets the same bytes-in-place layout as the A tile.
    if constexpr (TERMS == 3) {
      const int ct = threadIdx.x - TC_THREADS;          // 0..127
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          float4* a4 = reinterpret_cast<float4*>(smem_a + stage * A_BYTES);
          float4* l4 = reinterpret_cast<float4*>(smem_alo + stage * A_BYTES);
#pragma unroll
          for (int q = 0; q < (int)(A_BYTES / 16 / 128); ++q) {
            const int i = q * 128 + ct;
            float4 v = a4[i];
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            lo.x = v.x - hi.x;          // exact: at most 13 significant bits; the tensor core keeps its top 11
            lo.y = v.y - hi.y;
            lo.z = v.z - hi.z;
            lo.w = v.w - hi.w;
            a4[i] = hi;
            l4[i] = lo;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> tensor-core reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[stage]);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row = tile * TC_BM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * STAGE_COLS;
#pragma unroll
      for (int c0 = 0; c0 < WP; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if constexpr (NCH > 1) {          // add the K-chunk accumulators in round-to-nearest fp32
          const int kbc = (n_kb + NCH - 1) / NCH, used = (n_kb + kbc - 1) / kbc;      // short K: fewer accumulators in use
#pragma unroll
          for (int ch = 1; ch < NCH; ++ch) {
            if (ch >= used) break;
            uint32_t q[16];
            tmem_ld16(taddr + ch * TC_ACC_COLS + c0, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(q[j])));
          }
        }
        if (row < n && round_out == 2) {          // fp16 h (FESR_PREC_F16): 16 columns = 2 x 16-byte stores
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float x0 = __uint_as_float(r[j]), x1 = __uint_as_float(r[j + 1]);
            const int c = c0 + j;
            if (epi != EPI_NONE) {
              x0 += bias_p[c];
              x1 += bias_p[c + 1];
            }
            if (epi == EPI_BIAS_CONST1) {
              if (c == w) x0 = 1.f;
              if (c + 1 == w) x1 = 1.f;
            } else if (epi == EPI_BIAS_RELU) {
              x0 = fmaxf(x0, 0.f);
              x1 = fmaxf(x1, 0.f);
            }
            guard.note(x0, x1);
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[j / 2]) : "f"(x1), "f"(x0));
          }
          __half* hh = reinterpret_cast<__half*>(h_out) + row * WP + c0;
          *reinterpret_cast<uint4*>(hh) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(hh + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else if (row < n) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int c = c0 + j + t;
              float x = __uint_as_float(r[j + t]);
              if (epi != EPI_NONE) x += bias_p[c];
              if (epi == EPI_BIAS_CONST1) {
                if (c == w) x = 1.f;
              } else if (epi == EPI_BIAS_RELU) {
                x = fmaxf(x, 0.f);
              }
              if (round_out == 1) {
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
                x = __uint_as_float(u);
              }
              v[t] = x;
            }
            *reinterpret_cast<float4*>(h_out + row * WP + c0 + j) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    guard.flush(ovf);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * STAGE_COLS));
  }
}

static int encode_map(CUtensorMap* map, const void* base, bool half, uint64_t inner, uint64_t outer,
                      uint32_t box_inner, uint32_t box_outer, bool bf16 = false, bool swizzle32 = false) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * ((half || bf16) ? 2 : 4)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  // the driver entry point is resolved at run time so that libfesr.so has no link-time
  // dependency on libcuda.so (it must load on a machine without a driver for the ABI checks)
  typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn_t encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FESR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return FESR_ECUDA;
    }
    encode = reinterpret_cast<encode_fn_t>(fn);
  }
  CUresult r = encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                      const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FESR_ECUDA;
  }
  return FESR_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward, tf32 arm:  dZ[n, zk] = dpre[n, wp] . T'^T  (reference: autograd of the last edge-MLP Linear + NNConv message,
// models/model.py:311-315, 527-529, in the node-level form of DESIGN.md section 5).  K = wp is thin (6 MMAs per 128 x 256
// output tile); the kernel exists to WRITE dZ (1.35 GB per layer at 527 k cells) at HBM rate:
//   B = a 256-row chunk of T' [zk, wp] (K-major, tf32-rounded copy) stays in shared memory for the whole CTA;
//   A = dpre tiles [128, wp] stream through a 3-stage TMA ring (two 128 x 32 boxes per tile: columns 48..63 of the second
//       box are out of bounds and arrive as zeros, which pads K to the 64-float swizzle atom pair);
//   D = [128, 256] fp32 in TMEM, two stages (all 512 columns): the epilogue of tile t overlaps the MMAs of tile t + 1;
//   epilogue: tcgen05.ld (thread = node row, 32 columns) -> a swizzled [128 x 32] staging tile in shared memory -> one
//       TMA store per tile and 32-column box (two staging buffers; cp.async.bulk.wait_group.read frees them).  Storing
//       the rows straight from the registers (32 lanes = 32 rows, 16 bytes each) ran at the mma.sync kernel's 2.1 TB/s:
//       half-filled sectors at twice the request rate.
// Grid: ceil(zk / 256) column chunks x as many node ranges as fill the SMs.
// TERMS == 3 (the fp32 arm): both operands arrive as tf32 hi + lo pairs (dpre_hi / dpre_lo written by the mask kernel,
// T'hi / T'lo by the rounding kernel) and the chain is lo.hi + hi.lo + hi.hi -- K is only 48, so the truncating fp32
// accumulation of the tensor core has nothing to drift over; 128-column chunks and a 2-stage ring fit the four tiles.
__host__ __device__ constexpr int dz_bn(int terms) { return terms == 3 ? 128 : 256; }
__host__ __device__ constexpr int dz_stages(int terms) { return terms == 3 ? 2 : 3; }
constexpr int DZ_THREADS = 192;

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}

template <int WP, bool BF16OUT, int TERMS>
__global__ void __launch_bounds__(DZ_THREADS, 1)
dz_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
             const __grid_constant__ CUtensorMap tmD, int64_t n, int zk, int n_chunks) {
  constexpr int DZ_BN = dz_bn(TERMS), DZ_STAGES = dz_stages(TERMS);
  constexpr int NT = TERMS == 3 ? 2 : 1;               // operand copies (hi, lo)
  static_assert(WP > 32 && WP <= 64 && WP % 8 == 0, "two 32-float k-blocks");
  constexpr uint32_t A_SLAB = TC_BM * TC_BK * 4;       // 16 KB: 128 rows x 128 B
  constexpr uint32_t B_SLAB = DZ_BN * TC_BK * 4;       // 256 (128) rows x 128 B
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                              // [NT][2 k-blocks][BN][128 B]
  uint8_t* smem_a = smem_b + NT * 2 * B_SLAB;          // [DZ_STAGES][NT][2 k-blocks][128][128 B]
  uint8_t* smem_d = smem_a + DZ_STAGES * NT * 2 * A_SLAB;   // [2][128][128 B] staging tiles of the epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_d + 2 * A_SLAB);
  uint64_t* empty_bar = full_bar + DZ_STAGES;
  uint64_t* b_bar = empty_bar + DZ_STAGES;
  uint64_t* tmem_full = b_bar + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % n_chunks;             // column chunk of this CTA
  const int range = blockIdx.x / n_chunks, n_ranges = gridDim.x / n_chunks;
  const int64_t n_tiles = (n + TC_BM - 1) / TC_BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (TERMS == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DZ_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(b_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the T' chunk once, then the dpre tiles of this CTA's node range =====
    if (elect_one()) {
      mbar_expect_tx(b_bar, NT * 2 * B_SLAB);
      tma_load_2d(smem_b, &tmB, b_bar, 0, chunk * DZ_BN);
      tma_load_2d(smem_b + B_SLAB, &tmB, b_bar, TC_BK, chunk * DZ_BN);
      if (TERMS == 3) {
        tma_load_2d(smem_b + 2 * B_SLAB, &tmBlo, b_bar, 0, chunk * DZ_BN);
        tma_load_2d(smem_b + 3 * B_SLAB, &tmBlo, b_bar, TC_BK, chunk * DZ_BN);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = range; tile < n_tiles; tile += n_ranges) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], NT * 2 * A_SLAB);
        tma_load_2d(smem_a + (stage * NT * 2) * A_SLAB, &tmA, &full_bar[stage], 0, (int)(tile * TC_BM));
        tma_load_2d(smem_a + (stage * NT * 2 + 1) * A_SLAB, &tmA, &full_bar[stage], TC_BK, (int)(tile * TC_BM));
        if (TERMS == 3) {
          tma_load_2d(smem_a + (stage * NT * 2 + 2) * A_SLAB, &tmAlo, &full_bar[stage], 0, (int)(tile * TC_BM));
          tma_load_2d(smem_a + (stage * NT * 2 + 3) * A_SLAB, &tmAlo, &full_bar[stage], TC_BK, (int)(tile * TC_BM));
        }
        if (++stage == DZ_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D = F32, A = B = TF32 K-major, N = 256, M = 128; WP / 8 k-steps of 8 per tile =====
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(DZ_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    mbar_wait(b_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int64_t tile = range; tile < n_tiles; tile += n_ranges, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      mbar_wait(&full_bar[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t tmem_d = tmem_base + as * DZ_BN;
#pragma unroll
        for (int k = 0; k < WP / 8; ++k) {
          const uint64_t adesc = make_sw128_desc(smem_u32(smem_a + (stage * NT * 2 + (k >> 2)) * A_SLAB)) + 2 * (k & 3);
          const uint64_t bdesc = make_sw128_desc(smem_u32(smem_b + (k >> 2) * B_SLAB)) + 2 * (k & 3);
          if (TERMS == 3) {      // small terms first
            const uint64_t alodesc = make_sw128_desc(smem_u32(smem_a + (stage * NT * 2 + 2 + (k >> 2)) * A_SLAB)) + 2 * (k & 3);
            const uint64_t blodesc = make_sw128_desc(smem_u32(smem_b + (2 + (k >> 2)) * B_SLAB)) + 2 * (k & 3);
            umma_tf32(tmem_d, alodesc, bdesc, idesc, k != 0);
            umma_tf32(tmem_d, adesc, blodesc, idesc, 1);
            umma_tf32(tmem_d, adesc, bdesc, idesc, 1);
          } else {
            umma_tf32(tmem_d, adesc, bdesc, idesc, k != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[as]);
      }
      __syncwarp();
      if (++stage == DZ_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4; thread = node row =====
    const int quad = warp & 3;
    const int rloc = quad * 32 + lane;                   // row of the tile
    const bool issuer = warp == 2 && lane == 0;          // the thread that issues (and waits for) the TMA stores
    const int col0 = chunk * DZ_BN;
    const int ncol = min(DZ_BN, zk - col0);              // the last chunk may be short (a multiple of 32: zk % 32 == 0)
    const uint32_t stg = smem_u32(smem_d) + (uint32_t)(rloc * 128);
    const uint32_t sw = (uint32_t)(rloc & 7);
    int it = 0, box = 0;
    for (int64_t tile = range; tile < n_tiles; tile += n_ranges, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * DZ_BN;
      constexpr int BOXC = BF16OUT ? 64 : 32;            // columns of one 128-byte staging row
      for (int c0 = 0; c0 < ncol; c0 += BOXC, ++box) {
        const uint32_t buf = (uint32_t)(box & 1) * A_SLAB;
        // the store that last read this staging buffer (two boxes ago) has finished reading it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t r[32];
        if constexpr (BF16OUT) {
          // bf16 dZ (a subset of tf32: the edge-gradient MMAs consume it without a second rounding)
          uint32_t q0[32], q1[32];
          tmem_ld32(taddr + c0, q0);
          tmem_ld32(taddr + c0 + 32, q1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[j]) : "f"(__uint_as_float(q0[2 * j + 1])), "f"(__uint_as_float(q0[2 * j])));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[16 + j]) : "f"(__uint_as_float(q1[2 * j + 1])), "f"(__uint_as_float(q1[2 * j])));
          }
        } else {
          tmem_ld32(taddr + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + buf + ((((uint32_t)j) ^ sw) << 4)), "r"(r[4 * j]),
                       "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmD),
                       "r"(smem_u32(smem_d) + buf), "r"(col0 + c0), "r"((int)(tile * TC_BM))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int WP, bool HALF, int TERMS>
static int launch_tc(const fesr_model_dims& d, const void* B_kmajor, const void* B_lo, const float* bias_p, int epi,
                     const void* Z, int64_t n, float* h_out, cudaStream_t s, int round_out) {
  constexpr int STAGES = tc_stages(TERMS, WP);
  constexpr size_t smem = (size_t)STAGES * tc_stage_bytes(TERMS, WP) + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(smem <= 227 * 1024, "stage ring exceeds the shared memory of an SM");
  constexpr uint32_t box_inner = HALF ? 2 * TC_BK : TC_BK;
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(node_gemm_tf32_kernel<WP, HALF, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_set = true;
  }
  FESR_CHECK_ARG(d.zk % (int)box_inner == 0, "zk must be a multiple of %u", box_inner);
  FESR_CHECK_ARG(TERMS == 1 || B_lo != nullptr, "multi-term node GEMM without a low-order weight copy");
  CUtensorMap tmA, tmB, tmBlo;
  int rc;
  if ((rc = encode_map(&tmA, Z, HALF, (uint64_t)d.zk, (uint64_t)n, box_inner, TC_BM))) return rc;
  if ((rc = encode_map(&tmB, B_kmajor, HALF, (uint64_t)d.zk, (uint64_t)d.wp, box_inner, WP))) return rc;
  if ((rc = encode_map(&tmBlo, TERMS >= 2 ? B_lo : B_kmajor, HALF, (uint64_t)d.zk, (uint64_t)d.wp, box_inner, WP))) return rc;
  const int64_t n_tiles = ceil_div(n, TC_BM);
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  ProfScope prof(PROF_NODE_GEMM, s);
  node_gemm_tf32_kernel<WP, HALF, TERMS><<<grid, TC_THREADS + (TERMS == 3 ? 128 : 0), smem, s>>>(
      tmA, tmB, tmBlo, bias_p, n, d.zk, d.w, epi, round_out, h_out, cur_ovf());
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

template <bool HALF, int TERMS>
static int dispatch_tc(const fesr_model_dims& d, const void* B, const void* B_lo, const float* bias_p, int epi, const void* Z,
                       int64_t n, float* h_out, cudaStream_t s, int round_out) {
  if (n == 0) return FESR_OK;
  switch (d.wp) {
    case 16: return launch_tc<16, HALF, TERMS>(d, B, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
    case 32: return launch_tc<32, HALF, TERMS>(d, B, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
    case 48: return launch_tc<48, HALF, TERMS>(d, B, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
    case 64: return launch_tc<64, HALF, TERMS>(d, B, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
  }
  set_error("unsupported padded width %d", d.wp);
  return FESR_EINVAL;
}

int launch_node_gemm_tf32(const fesr_model_dims& d, const float* B_kmajor, const float* bias_p, int epi, const float* Z,
                          int64_t n, float* h_out, cudaStream_t s, int round_out, const float* B_lo, int terms) {
  if (terms == 3) return dispatch_tc<false, 3>(d, B_kmajor, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
  if (terms == 2) return dispatch_tc<false, 2>(d, B_kmajor, B_lo, bias_p, epi, Z, n, h_out, s, round_out);
  return dispatch_tc<false, 1>(d, B_kmajor, nullptr, bias_p, epi, Z, n, h_out, s, round_out);
}

int launch_node_gemm_f16(const fesr_model_dims& d, const void* B_kmajor_h, const float* bias_p, int epi, const void* Z_h,
                         int64_t n, float* h_out, cudaStream_t s, int round_out, const void* B_lo_h) {
  if (B_lo_h != nullptr) return dispatch_tc<true, 2>(d, B_kmajor_h, B_lo_h, bias_p, epi, Z_h, n, h_out, s, round_out);
  return dispatch_tc<true, 1>(d, B_kmajor_h, nullptr, bias_p, epi, Z_h, n, h_out, s, round_out);
}

// dZ = dpre . T'^T on tcgen05 (tf32 arm of the backward); tprime_r: tf32-rounded copy of T' [zk, wp] row-major;
// dpre must be tf32-rounded too (the tensor core truncates instead of rounding).  false: shape not covered (caller
// falls back to the mma.sync kernel).
bool dz_tc_supported(const fesr_model_dims& d) { return d.wp == 48 && d.zk % 64 == 0; }

template <bool BF16OUT, int TERMS>
static int launch_dz_tc_t(const fesr_model_dims& d, const float* dpre, const float* dpre_lo, const float* tprime_r,
                          const float* tprime_r_lo, int64_t n, void* dZ, cudaStream_t s) {
  constexpr int BN = dz_bn(TERMS), STAGES = dz_stages(TERMS), NT = TERMS == 3 ? 2 : 1;
  constexpr size_t smem = 1024 + (size_t)NT * 2 * BN * TC_BK * 4 + (size_t)(STAGES * NT + 1) * 2 * TC_BM * TC_BK * 4 + 256;
  static_assert(smem <= 227 * 1024, "dZ tiles exceed the shared memory of an SM");
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(dz_tc_kernel<48, BF16OUT, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmAlo, tmBlo, tmD;
  int rc;
  if ((rc = encode_map(&tmA, dpre, false, (uint64_t)d.wp, (uint64_t)n, TC_BK, TC_BM))) return rc;
  if ((rc = encode_map(&tmB, tprime_r, false, (uint64_t)d.wp, (uint64_t)d.zk, TC_BK, BN))) return rc;
  if ((rc = encode_map(&tmAlo, TERMS == 3 ? dpre_lo : dpre, false, (uint64_t)d.wp, (uint64_t)n, TC_BK, TC_BM))) return rc;
  if ((rc = encode_map(&tmBlo, TERMS == 3 ? tprime_r_lo : tprime_r, false, (uint64_t)d.wp, (uint64_t)d.zk, TC_BK, BN))) return rc;
  if ((rc = encode_map(&tmD, dZ, false, (uint64_t)d.zk, (uint64_t)n, BF16OUT ? 2 * TC_BK : TC_BK, TC_BM, BF16OUT))) return rc;
  const int n_chunks = (int)ceil_div(d.zk, BN);
  const int64_t n_tiles = ceil_div(n, TC_BM);
  int ranges = num_sms() / n_chunks;
  if (ranges < 1) ranges = 1;
  if (ranges > n_tiles) ranges = (int)n_tiles;
  dz_tc_kernel<48, BF16OUT, TERMS><<<n_chunks * ranges, DZ_THREADS, smem, s>>>(tmA, tmB, tmAlo, tmBlo, tmD, n, d.zk, n_chunks);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

int launch_dz_tc(const fesr_model_dims& d, const float* dpre, const float* tprime_r, int64_t n, void* dZ, int out_bf16,
                 cudaStream_t s, const float* dpre_lo, const float* tprime_r_lo) {
  if (n == 0) return FESR_OK;
  if (dpre_lo != nullptr) {
    FESR_CHECK_ARG(tprime_r_lo != nullptr && !out_bf16, "the three-term dZ product takes both lo operands and writes fp32");
    return launch_dz_tc_t<false, 3>(d, dpre, dpre_lo, tprime_r, tprime_r_lo, n, dZ, s);
  }
  if (out_bf16) return launch_dz_tc_t<true, 1>(d, dpre, nullptr, tprime_r, nullptr, n, dZ, s);
  return launch_dz_tc_t<false, 1>(d, dpre, nullptr, tprime_r, nullptr, n, dZ, s);
}

// ------------------------------------------------------------------------------------------------
// Backward, tf32 arm:  dT'[zk, wp] += Z^T[zk, n] . dpre[n, wp]  (reference: autograd of the edge-MLP's last Linear through the
// NNConv message, models/model.py:311-315, 527-529, in the node-level form of DESIGN.md section 5).  K = n is the long
// dimension, so both operands are MN-major for the tensor core: the rows TMA brings in (one node = one K index) are
// contiguous along M (Z: channels) and N (dpre: columns).
//   A = Z^T: the fp16 Z stash [n, zk] row-major; a stage holds 64 nodes x (up to) 768 channels as twelve SWIZZLE_128B boxes of
//       64 channels (128 B) x 64 nodes -- the canonical MN-major layout: 8 nodes per 1024-byte atom (SBO), the next 64
//       channels one box further (LBO);
//   B = the scaled fp16 dpre rows [n, 48] (scale_rows_f16_kernel's `own` rows: dpre S_l with the per-layer power of two S_l)
//       as three SWIZZLE_32B boxes of 16 columns x 64 nodes (atoms of 8 nodes x 32 B);
//   D = up to six [128 channels, 48] fp32 accumulators in TMEM, live for the CTA's whole node range (split-K over CTAs).
// The 2176 channels of the shipped KernelNN take 17 m-tiles = three channel splits; grid = splits x node ranges, one CTA per
// SM; every CTA writes its [channels, 48] partial and a second kernel adds them up in a fixed order, times 1 / S_l.
// The operands carry the same 11 significant bits as the tf32 mma.sync kernel this replaces (bwd_gemm_mma.cu); that kernel
// was bound by the mma.sync issue rate (tensor pipe 59 % busy at 3.2 TB/s), this one by the HBM stream of Z.
constexpr int WT_KB = 64;                          // nodes per stage
constexpr int WT_TILES = 6;                        // m-tiles per CTA: 6 x 64 TMEM columns
constexpr int WT_STAGES = 2;
constexpr int WT_THREADS = 192;
constexpr uint32_t WT_ZBOX = WT_KB * 128;          // 8 KB
constexpr uint32_t WT_DBOX = WT_KB * 32;           // 2 KB
constexpr uint32_t WT_STAGE_BYTES = 2 * WT_TILES * WT_ZBOX + 3 * WT_DBOX;      // 102 KB (a multiple of 1024)

// MN-major shared-memory matrix descriptor: lbo = bytes between swizzle-wide groups along M / N, sbo = bytes between
// 8-row groups along K; layout 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmD, int64_t n, int zk, int msplit,
                int64_t nchunk, float* __restrict__ partial) {
  constexpr int WP = 48;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WT_STAGES * WT_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WT_STAGES;
  uint64_t* done_bar = empty_bar + WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x % msplit, range = blockIdx.x / msplit;
  const int box0 = split * 2 * WT_TILES;                       // first 64-channel box of this CTA
  const int nbox = min(2 * WT_TILES, zk / 64 - box0);
  const int tiles = (nbox + 1) / 2;                            // (an odd box count leaves the last tile's upper half unread)
  const int64_t k_begin = (int64_t)range * nchunk, k_end = min(n, k_begin + nchunk);
  const int iters = k_begin < k_end ? (int)((k_end - k_begin + WT_KB - 1) / WT_KB) : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int stage = it % WT_STAGES;
        mbar_wait(&empty_bar[stage], ((it / WT_STAGES) & 1) ^ 1);
        uint8_t* zs = smem + stage * WT_STAGE_BYTES;
        uint8_t* ds = zs + 2 * WT_TILES * WT_ZBOX;
        const int node0 = (int)(k_begin + (int64_t)it * WT_KB);        // rows past n arrive as zeros
        mbar_expect_tx(&full_bar[stage], nbox * WT_ZBOX + 3 * WT_DBOX);
        for (int j = 0; j < 3; ++j) tma_load_2d(ds + j * WT_DBOX, &tmD, &full_bar[stage], j * 16, node0);
        for (int b = 0; b < nbox; ++b) tma_load_2d(zs + b * WT_ZBOX, &tmZ, &full_bar[stage], (box0 + b) * 64, node0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D = F32, A = B = F16, both MN-major, M = 128, N = 48, K = 16: four k-steps per stage and m-tile =====
    constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(WP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const int stage = it % WT_STAGES;
      mbar_wait(&full_bar[stage], (it / WT_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t zs = smem_u32(smem + stage * WT_STAGE_BYTES);
        const uint32_t ds = zs + 2 * WT_TILES * WT_ZBOX;
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
          for (int j = 0; j < WT_KB / 16; ++j) {
            const uint64_t adesc = make_mn_desc(zs + t * 2 * WT_ZBOX + j * 16 * 128, WT_ZBOX, 1024, 2);
            const uint64_t bdesc = make_mn_desc(ds + j * 16 * 32, WT_DBOX, 256, 6);
            umma_f16(tmem_base + t * 64, adesc, bdesc, idesc, (it | j) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == iters - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4; thread = one channel row of the partial =====
    const int quad = warp & 3;
    if (iters > 0) {
      mbar_wait(done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    float* out = partial + (int64_t)range * zk * WP;
    for (int t = 0; t < tiles; ++t) {
      const int m0 = box0 * 64 + t * 128 + quad * 32;            // warp-uniform; zk is a multiple of 64
      if (m0 >= zk) continue;
      uint32_t r[3][16];
      if (iters > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * 64;
        tmem_ld16(taddr, r[0]);
        tmem_ld16(taddr + 16, r[1]);
        tmem_ld16(taddr + 32, r[2]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 48; ++j) r[j / 16][j % 16] = 0u;
      }
      float4* row = reinterpret_cast<float4*>(out + (int64_t)(m0 + lane) * WP);
#pragma unroll
      for (int j = 0; j < 12; ++j)
        row[j] = make_float4(__uint_as_float(r[j / 4][(j % 4) * 4]), __uint_as_float(r[j / 4][(j % 4) * 4 + 1]),
                             __uint_as_float(r[j / 4][(j % 4) * 4 + 2]), __uint_as_float(r[j / 4][(j % 4) * 4 + 3]));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// dT[i] += inv_scale * sum over the node ranges, in range order
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, int ranges, int64_t count,
                                       const float* __restrict__ inv_scale, float* __restrict__ dT) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < ranges; ++z) s += partial[(int64_t)z * count + i];
  dT[i] += s * __ldg(inv_scale);
}

bool wgrad_tc_supported(const fesr_model_dims& d) { return d.wp == 48 && d.zk % 64 == 0; }

int launch_wgrad_tc(const fesr_model_dims& d, const void* Z_half, const void* dpre_half, const float* inv_scale, int64_t n,
                    float* dT, float* ws, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  FESR_CHECK_ARG(wgrad_tc_supported(d), "weight gradient on tcgen05: wp == 48 and zk %% 64 == 0");
  constexpr size_t smem = 1024 + (size_t)WT_STAGES * WT_STAGE_BYTES + 256;
  static_assert(smem <= 227 * 1024, "weight-gradient stages exceed the shared memory of an SM");
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  CUtensorMap tmZ, tmD;
  int rc;
  if ((rc = encode_map(&tmZ, Z_half, true, (uint64_t)d.zk, (uint64_t)n, 64, WT_KB))) return rc;
  if ((rc = encode_map(&tmD, dpre_half, true, (uint64_t)d.wp, (uint64_t)n, 16, WT_KB, false, /*swizzle32=*/true))) return rc;
  const int msplit = (int)ceil_div(d.zk / 64, 2 * WT_TILES);
  int ranges = num_sms() / msplit;
  const int cap = (int)(wgrad_mma_ws_bytes(d) / ((size_t)d.zk * d.wp * sizeof(float)));      // partials the workspace holds
  if (ranges > cap) ranges = cap;
  if (ranges < 1) ranges = 1;
  const int64_t nchunk = ceil_div(ceil_div(n, ranges), WT_KB) * WT_KB;
  ranges = (int)ceil_div(n, nchunk);
  wgrad_tc_kernel<<<msplit * ranges, WT_THREADS, smem, s>>>(tmZ, tmD, n, d.zk, msplit, nchunk, ws);
  FESR_LAUNCH_CHECK();
  const int64_t count = (int64_t)d.zk * d.wp;
  wgrad_tc_reduce_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, s>>>(ws, ranges, count, inv_scale, dT);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // namespace fesr
