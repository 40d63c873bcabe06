// Device helpers shared by the fused layer kernels (layer_fused.cu: 8-node tiles; layer_fused16.cu: 16-node tiles).
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace fesr {

constexpr int FL_NODES = 8;                   // nodes per tile
constexpr int FL_CAP = 128;                   // staged edges per ring slot (a tile with more edges takes several)
constexpr int FL_DEGC = 16;                   // edges per staged chunk = one m16n8k16 k-step
constexpr int FL_NKB = 13;                    // k-blocks per part: 12 outer-product blocks + root block
constexpr int FL_KP = FL_NKB * 64;            // 832 fp16 of K per part
constexpr int FL_ACOLS = FL_KP / 2;           // TMEM columns of the A operand
constexpr int FL_WP = 48;

__device__ __forceinline__ uint32_t fl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef FL_WAIT_MODE
#define FL_WAIT_MODE 0
#endif
// tools/dev only (-DFL_TRACE): per-role clock64 accounting of the mbarrier waits, written to a global buffer, and
// what-if switches in bits 8..11 of `relu` (0x100 no h gathers, 0x200 no fix-up row, 0x400 no outer-product MMAs,
// 0x800 one k-block of the contraction) -- results are wrong by construction, only the time matters.  Production
// builds compile both out (FL_WHATIF(x) == false).
#ifdef FL_TRACE
#define FL_WHATIF(bit) ((relu & (bit)) != 0)
static __device__ long long fl_trace_buf[148 * 24 * 12];
#define FL_TWAIT(slot, call) { const long long t0__ = clock64(); call; tw[slot] += clock64() - t0__; }
#define FL_TMARK(slot) { const long long t1__ = clock64(); tw[slot] += t1__ - tmark; tmark = t1__; }
#else
#define FL_TMARK(slot)
#define FL_WHATIF(bit) false
#define FL_TWAIT(slot, call) call;
#endif
__device__ __forceinline__ void fl_mbar_wait(uint32_t bar, uint32_t parity) {
#if FL_WAIT_MODE == 0
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "FL_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra FL_DONE;\n\t"
      "bra FL_WAIT;\n\t"
      "FL_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680)     // suspend-time hint: a waiting warp sleeps instead of spinning
      : "memory");
#elif FL_WAIT_MODE == 1
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "FL_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FL_DONE;\n\t"
      "bra FL_WAIT;\n\t"
      "FL_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "FL_WAIT:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FL_DONE;\n\t"
      "bra FL_WAIT;\n\t"
      "FL_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
#endif
}
__device__ __forceinline__ bool fl_elect_one() {
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
__device__ __forceinline__ void fl_ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// fp16-accumulate MMA: D/C are two f16x2 registers (rows gq and gq + 8, columns 2tq, 2tq + 1)
__device__ __forceinline__ void fl_mma16(uint32_t (&c)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
      : "+r"(c[0]), "+r"(c[1])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// same with a zero C operand (the first k-step of an accumulation: no zeroing of the accumulators)
__device__ __forceinline__ void fl_mma16z(uint32_t (&c)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%8};"
      : "=r"(c[0]), "=r"(c[1])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0u));
}
__device__ __forceinline__ __half2 fl_as_h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ uint32_t fl_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t fl_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t fl_hmul2(uint32_t a, __half2 b) {
  const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&a), b);
  return *reinterpret_cast<const uint32_t*>(&r);
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same format as gemm_tc.cu)
__device__ __forceinline__ uint64_t fl_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void fl_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// same with the accumulate flag known at compile time (no predicate set-up on the issuing thread's critical path)
template <bool ACC>
__device__ __forceinline__ void fl_umma_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void fl_umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fl_tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void fl_tmem_st16(uint32_t addr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// layer_fused16.cu: the 16-node-tile kernel (same contract as launch_fl<PPL, NBUF> of layer_fused.cu)
template <int PPL>
int launch_fl16(const int32_t* rowptr, const int32_t* src_sorted, const __half* g3, int64_t E, const __half* h_in,
                int64_t n, int part0, int has_root, const __half* tf, const float* bias_p, const float* p_in,
                float* p_out, __half* h_out, int rs, int fix_b, int relu, cudaStream_t s, const __half* h_own = nullptr);

}  // namespace fesr
