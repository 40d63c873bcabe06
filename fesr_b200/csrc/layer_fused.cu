// One message-passing layer as ONE kernel (FESR_PREC_F16 arm, KernelNN shape):
//
//   h'_i = act( sum_{(s,a)} Z_i[s,a] T'[(s,a), :] + h_i root + bias ),   Z_i = 1/deg_i sum_{e->i} g_e (x) h[src_e]
//
// (reference: NNConv_old.forward/message/update + PyG mean aggregation, models/model.py:521-536, and
// the last Linear of the edge MLP, :311-315 -- reordered as in DESIGN.md section 2.)  The unfused
// path writes Z (4.4 KB per node) to HBM and reads it back in the node GEMM; that round trip is
// two thirds of a layer's time.  Here Z never leaves the SM:
//
//   builder warps   stage 16 edges at a time (cp.async: the g slots of this launch's parts + the
//                   gathered h[src] row), form the per-node outer-product sum on mma.sync.m16n8k16
//                   (D = H_i^T [a x edges] . G_i [edges x slots], fp32 accumulate), scale by 1/deg
//                   and write the fp16 result straight into the B-operand tile of the node
//                   contraction in shared memory (K-major, SWIZZLE_128B, K ordered so that every
//                   warp-wide store is one conflict-free 128-byte row);
//   MMA warp        tcgen05.mma kind::f16, M = 128, N = 16 nodes x PPL parts, A = T' FROM TENSOR
//                   MEMORY (loaded once per CTA with tcgen05.st), B = the Z tile, D in TMEM;
//   epilogue warps  tcgen05.ld, combine the parts, (+ partial sums of earlier launches), bias,
//                   ReLU, fp16 h' rows.
//
// Nodes sit on the MMA N dimension because a full-K Z tile of >= 64 nodes (what M would need)
// does not fit in shared memory, and T' sits in TMEM because re-reading it from shared memory for
// every 16-node tile would cost more shared-memory bandwidth than Z itself.  TMEM holds 512
// columns = 1024 fp16 of K per lane, so K (2304 + 48) is cut into PARTS of 16 g-slots
// (768 = 12 k-blocks of 64) plus one root k-block.  Several parts share the 128 TMEM lanes
// (rows) at once: part p lives in lanes [p*RS, p*RS + w) and multiplies only the B columns that
// hold part p of the tile's nodes.  With w <= 43 all three parts of the shipped model fit
// (3*43 = 129: the single row that does not fit, (part 2, channel 42), is evaluated on CUDA
// cores from the builder's registers), so a layer is ONE launch and every edge is staged once.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fesr {

constexpr int FL_NODES = 8;                   // nodes per tile
constexpr int FL_DEGC = 16;                   // edges per staged chunk = one m16n8k16 k-step
constexpr int FL_NKB = 13;                    // k-blocks per part: 12 outer-product blocks + root block
constexpr int FL_KP = FL_NKB * 64;            // 832 fp16 of K per part
constexpr int FL_ACOLS = FL_KP / 2;           // TMEM columns of the A operand
constexpr int FL_SHB = 112;                   // staged h row stride in bytes (96 + 16: conflict-free ldmatrix)
constexpr int FL_WP = 48;

__device__ __forceinline__ uint32_t fl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fl_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void fl_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void fl_mma0(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t fl_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 fl_h2_to_f2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
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

// ------------------------------------------------------------------------------------------------
// T' in the fused kernel's K order, fp16:  TF[part][b][K],  K = kb*64 + a_in*8 + s_in with
//   kb < 12 : a = (kb/2)*8 + a_in,  slot = part*16 + (kb%2)*8 + s_in  -> T'[(chan(slot), a), b]
//   kb = 12 : a = a_in*8 + s_in (root block, last part only; the tail 16 entries are zero)
// so that the 64 values one builder store instruction produces (8 a's x 8 slots) are one k-block row.
__global__ void prepare_tfused_kernel(fesr_model_dims d, int n_parts, const float* __restrict__ tprime,
                                      __half* __restrict__ tf) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_parts * FL_WP * FL_KP;
  if (idx >= total) return;
  const int K = (int)(idx % FL_KP);
  const int b = (int)((idx / FL_KP) % FL_WP);
  const int part = (int)(idx / ((int64_t)FL_KP * FL_WP));
  const int kb = K >> 6, e = K & 63;
  float v = 0.f;
  if (kb < 12) {
    const int a = (kb >> 1) * 8 + (e >> 3);
    const int slot = part * 16 + (kb & 1) * 8 + (e & 7);
    const int q = slot / d.ktp, r = slot % d.ktp;
    const int chan = q * d.kt + r;
    if (r < d.kt && chan < d.k1 && a < d.wp && b < d.wp) v = tprime[((size_t)chan * d.wp + a) * d.wp + b];
  } else if (part == n_parts - 1 && e < d.wp) {
    v = tprime[((size_t)d.zk_main + e) * d.wp + b];
  }
  tf[idx] = __float2half_rn(v);
}

// PPL: parts per launch (1..3); NBUF: staged nodes in flight per node slot.  A tile is FL_NODES = 8 nodes;
// node slot j is served by a GROUP of PPL builder warps (warp j*PPL + pp builds part pp: 6 MMAs, 12 row
// stores), so 8*PPL builder warps hide each other's load / ldmatrix / MMA latencies.  MMA N = 8 * PPL
// rounded up to 16 (row pp*8 + j of the Z tile = part pp of node j).
template <int PPL, int NBUF>
__global__ void __launch_bounds__((8 * PPL + 5) * 32, 1)
layer_fused_f16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                       const __half* __restrict__ g3, int64_t E, const __half* __restrict__ h_in, int64_t n,
                       int part0, int has_root, const __half* __restrict__ tf, const float* __restrict__ bias_p,
                       const float* p_in, float* p_out, __half* __restrict__ h_out,
                       int rs, int fix_b, int relu) {
  constexpr int FL_BW = FL_NODES * PPL;                // builder warps
  constexpr int FL_THREADS = (FL_BW + 5) * 32;         // + 4 epilogue warps + 1 MMA warp
  constexpr int N = (PPL * FL_NODES + 15) / 16 * 16;   // MMA N
  constexpr int SLAB = N * 128;                        // bytes per k-block of the Z tile
  constexpr int ZBYTES = FL_NKB * SLAB;
  constexpr int SGB = 32 * PPL + 16;                   // staged g row stride (bytes), odd multiple of 16
  constexpr int GSZ = FL_DEGC * SGB;
  constexpr int STG = GSZ + (FL_DEGC + 1) * FL_SHB;    // + 16 gathered h rows + the node's own h row
  constexpr int HI = (FL_DEGC * 6 + 32 * PPL - 1) / (32 * PPL);   // h cp.async instructions per warp per chunk

  extern __shared__ __align__(1024) uint8_t fl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* zbuf = smem;                                            // [2][ZBYTES]
  uint8_t* stage = zbuf + 2 * ZBYTES;                              // [FL_NODES][NBUF][STG]
  float* comb = reinterpret_cast<float*>(stage + FL_NODES * NBUF * STG);   // [PPL][8][48]
  float* fixs = comb + PPL * FL_NODES * FL_WP;                     // [4][8]  fix-up row values
  float* invs = fixs + 4 * FL_NODES;                               // [4][8]  1 / max(deg, 1)
  uint32_t* wfs = reinterpret_cast<uint32_t*>(invs + 4 * FL_NODES);       // [12][32] + [8][4]  fix-up row weights
  uint64_t* bars = reinterpret_cast<uint64_t*>(wfs + (PPL == 3 ? 12 * 32 + 32 : 0));
  const uint32_t zfull = fl_smem(&bars[0]), zempty = fl_smem(&bars[2]), dfull = fl_smem(&bars[4]),
                 dempty = fl_smem(&bars[6]), aready = fl_smem(&bars[8]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[9]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t n_tiles = (n + FL_NODES - 1) / FL_NODES;
  const int n_it = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // grid <= n_tiles

  // the Z tiles start as zeros: unused rows, root-block rows of the parts without a root and the zero
  // tail of the root block are never written afterwards
  for (int t = threadIdx.x; t < 2 * ZBYTES / 16; t += FL_THREADS)
    reinterpret_cast<uint4*>(zbuf)[t] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = threadIdx.x; t < FL_NODES * NBUF * STG / 16; t += FL_THREADS)
    reinterpret_cast<uint4*>(stage)[t] = make_uint4(0u, 0u, 0u, 0u);       // stale slab contents stay finite
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      fl_mbar_init(zfull + 8 * s, FL_BW);
      fl_mbar_init(zempty + 8 * s, 1);
      fl_mbar_init(dfull + 8 * s, 1);
      fl_mbar_init(dempty + 8 * s, 4);
    }
    fl_mbar_init(aready, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PPL == 3 && fix_b >= 0) {
    // fix-up row (see header): weights of output channel fix_b against a builder lane's last-part values,
    // in the order the lane holds them: wfs[(mt*2+hh)*2+nl][lane], then the root block as 6 x 16 bytes
    const __half* wr = tf + ((size_t)(part0 + PPL - 1) * FL_WP + fix_b) * FL_KP;
    for (int t = threadIdx.x; t < 12 * 32; t += FL_THREADS) {
      const int q = t >> 5, l = t & 31;
      wfs[t] = *reinterpret_cast<const uint32_t*>(wr + (q * 64 + (l >> 2) * 8 + 2 * (l & 3)));
    }
    for (int t = threadIdx.x; t < 32; t += FL_THREADS)
      wfs[12 * 32 + t] = t < 24 ? *reinterpret_cast<const uint32_t*>(wr + 12 * 64 + 2 * t) : 0u;
  }
  if (warp == FL_BW) {   // TMEM: all 512 columns (A operand 416, two D stages of N)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fl_smem(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // zero fill is read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < FL_BW) {
    // =========================================================================== builders
    // Group j = warp / PPL owns node j of every tile of this CTA; this warp builds part pp = warp % PPL.
    // Item = one node (its first 16 edges are prefetched NBUF - 1 items ahead; the rare longer rows
    // finish synchronously).  Each warp of a group stages its own share of an item: the g slot group of
    // its part and every PPL-th block of 32 h-row chunks.
    const int j = warp / PPL, pp = warp % PPL;
    const int gq = lane >> 2, tq = lane & 3;
    const uint32_t st_u32 = fl_smem(stage + (size_t)j * NBUF * STG);
    const int lr = lane & 7, lm = lane >> 3;
    // ldmatrix row addresses (bytes, relative to a staged chunk):
    //   A = H^T tile mt: matrices {a 0-7, e 0-7}, {a 8-15, e 0-7}, {a 0-7, e 8-15}, {a 8-15, e 8-15}
    //   B = G of this part: matrices {e 0-7, nt 0}, {e 8-15, nt 0}, {e 0-7, nt 1}, {e 8-15, nt 1}
    const uint32_t a_off = (uint32_t)(GSZ + ((lm >> 1) * 8 + lr) * FL_SHB + (lm & 1) * 16);
    const uint32_t b_off = (uint32_t)(((lm & 1) * 8 + lr) * SGB + (lm >> 1) * 16 + pp * 32);
    // this lane's row of the Z tile: row pp*8 + j, 16-byte chunk gq (swizzled), bytes 4*tq
    const uint32_t zlane = fl_smem(zbuf) + (uint32_t)((pp * FL_NODES + j) * 128 + ((gq ^ j) << 4) + (tq << 2));

    // staging constants of this lane
    const int gj = lane >> 1;                                        // g: chunk (lane & 1) of edge row gj
    const uint32_t gdst = (uint32_t)(gj * SGB + pp * 32 + (lane & 1) * 16);
    const __half* gsrc = g3 + ((size_t)(part0 + pp) * E + gj) * 16 + (lane & 1) * 8;
    int hj[HI];
    uint32_t hdst[HI], hsrc[HI];
#pragma unroll
    for (int i = 0; i < HI; ++i) {
      const int t = (i * PPL + pp) * 32 + lane, r = t / 6, c = t % 6;
      hj[i] = (t < FL_DEGC * 6) ? r : FL_DEGC;                       // FL_DEGC: no such row
      hdst[i] = (uint32_t)(GSZ + r * FL_SHB + c * 16);
      hsrc[i] = (uint32_t)(c * 8);
    }
    const bool last_part = pp == PPL - 1;


    // rowptr of this group's nodes: 16 tile iterations per register (lanes 2i, 2i+1 = begin, end)
    auto load_block = [&](int blk) {
      const int itx = blk * 16 + (lane >> 1);
      int64_t node = ((int64_t)blockIdx.x + (int64_t)itx * gridDim.x) * FL_NODES + j + (lane & 1);
      node = node < n ? node : n;
      return __ldg(rowptr + node);
    };
    int rpb0 = load_block(0), rpb1 = load_block(1);
    auto bounds = [&](int itx, int& eb, int& ee) {
      const int rpv = ((itx >> 4) & 1) ? rpb1 : rpb0;
      eb = __shfl_sync(FULL, rpv, (itx & 15) * 2);
      ee = __shfl_sync(FULL, rpv, (itx & 15) * 2 + 1);
    };
    // source ids of item itx (lanes 0..15; clamped address: the value is not looked at before it is used)
    auto load_src = [&](int itx) {
      int eb, ee;
      bounds(itx, eb, ee);
      int e = eb + (lane & 15);
      e = e < ee ? e : (ee > 0 ? ee - 1 : 0);
      return __ldg(src_sorted + e);
    };
    // this warp's share of edges [c0, c0 + m) of a node: g slot group of its part (planar [part][E][16]; a
    // miss pulls 256 B into L2, i.e. the rows of the next edges of the same stream) and gathered h chunks
    auto stage_chunk = [&](uint32_t base, int c0, int m, int src_reg) {
      if (gj < m) {
        asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(base + gdst), "l"(gsrc + (size_t)c0 * 16) : "memory");
      } else {   // unused edge slots of the k-step: g = 0 (the stale h row is finite)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + gdst), "r"(0) : "memory");
      }
#pragma unroll
      for (int i = 0; i < HI; ++i) {
        const int s = __shfl_sync(FULL, src_reg, hj[i] & 15);
        if (hj[i] < m)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + hdst[i]), "l"(h_in + (size_t)s * FL_WP + hsrc[i]) : "memory");
      }
    };
    auto issue = [&](int itx, int buf, int src_reg) {
      if (itx < n_it) {
        int eb, ee;
        bounds(itx, eb, ee);
        const uint32_t base = st_u32 + (uint32_t)(buf * STG);
        stage_chunk(base, eb, min(FL_DEGC, ee - eb), src_reg);
        const int64_t node = ((int64_t)blockIdx.x + (int64_t)itx * gridDim.x) * FL_NODES + j;
        if (last_part && lane < 6 && node < n)   // the node's own row: root block (+ fix-up)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(GSZ + FL_DEGC * FL_SHB + lane * 16)),
                       "l"(h_in + (size_t)node * FL_WP + lane * 8)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto group_sync = [&]() {   // the PPL warps of node slot j
      if (PPL > 1) asm volatile("bar.sync %0, %1;" ::"r"(2 + j), "n"(32 * PPL) : "memory");
      else __syncwarp();
    };

#pragma unroll
    for (int i = 0; i < NBUF - 1; ++i) issue(i, i, load_src(i));
    // source ids are fetched three iterations before their gathers are issued: the dependent load never
    // sits on the critical path
    int src_a = load_src(NBUF - 1), src_b = load_src(NBUF), src_c = load_src(NBUF + 1);
    int buf = 0;
    float acc[3][2][4];
    for (int it = 0; it < n_it; ++it) {
      const int src_n = load_src(it + NBUF + 2);
      asm volatile("cp.async.wait_group %0;" ::"n"(NBUF - 2) : "memory");   // this warp's share of item `it` has landed
      group_sync();          // ... and everyone's; every warp of the group is also done reading item it - 1,
      {                      // whose buffer the next prefetch overwrites
        const int bl = buf == 0 ? NBUF - 1 : buf - 1;
        issue(it + NBUF - 1, bl, src_a);
      }
      if ((it & 15) == 0 && it > 0) {          // rowptr of the block after this one
        const int nb = load_block((it >> 4) + 1);
        if (((it >> 4) + 1) & 1) rpb1 = nb;
        else rpb0 = nb;
      }
      int eb, ee;
      bounds(it, eb, ee);
      const uint32_t base = st_u32 + (uint32_t)(buf * STG);
      if (ee > eb) {
        uint32_t b[4];
        fl_ldsm4t(base + b_off, b);
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          uint32_t a[4];
          fl_ldsm4t(base + a_off + (uint32_t)(mt * 32), a);
          fl_mma0(acc[mt][0], a, b[0], b[1]);
          fl_mma0(acc[mt][1], a, b[2], b[3]);
        }
        for (int c0 = eb + FL_DEGC; c0 < ee; c0 += FL_DEGC) {   // rows longer than one chunk (not prefetched)
          group_sync();
          const int m = min(FL_DEGC, ee - c0);
          const int sr = __ldg(src_sorted + c0 + min(lane & 15, m - 1));
          stage_chunk(base, c0, m, sr);
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          group_sync();
          fl_ldsm4t(base + b_off, b);
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            uint32_t a[4];
            fl_ldsm4t(base + a_off + (uint32_t)(mt * 32), a);
            fl_mma(acc[mt][0], a, b[0], b[1]);
            fl_mma(acc[mt][1], a, b[2], b[3]);
          }
        }
      } else {   // zero in-degree
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int nl = 0; nl < 2; ++nl)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nl][r] = 0.f;
      }
      // ---- this part's row of the Z tile: raw sums (the epilogue applies 1/deg), fp16, one conflict-free
      // 128-byte row per store instruction
      const int zb = it & 1;
      const int deg = ee - eb;
      fl_mbar_wait(zempty + 8 * zb, ((it >> 1) & 1) ^ 1);          // the tensor core is done with this buffer
      const uint32_t zrow = zlane + (uint32_t)(zb * ZBYTES);
      float fsum = 0.f;
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int nl = 0; nl < 2; ++nl) {
            const uint32_t v = fl_h2_sat(acc[mt][nl][2 * hh], acc[mt][nl][2 * hh + 1]);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(zrow + (uint32_t)(((mt * 2 + hh) * 2 + nl) * SLAB)), "r"(v) : "memory");
          }
      if (last_part) {
        if (PPL == 3 && fix_b >= 0) {
#pragma unroll
          for (int mt = 0; mt < 3; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
              for (int nl = 0; nl < 2; ++nl) {
                const float2 wf = fl_h2_to_f2(wfs[((mt * 2 + hh) * 2 + nl) * 32 + lane]);
                fsum = fmaf(acc[mt][nl][2 * hh], wf.x, fmaf(acc[mt][nl][2 * hh + 1], wf.y, fsum));
              }
        }
        if (has_root && lane < 8) {
          // root block: h_i * max(deg, 1) so that the epilogue's 1/deg leaves h_i
          uint4 hv = make_uint4(0u, 0u, 0u, 0u);
          if (lane < 6) {
            const uint32_t sa = base + (uint32_t)(GSZ + FL_DEGC * FL_SHB + lane * 16);
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(sa));
          }
          const __half2 dg = __float2half2_rn((float)(deg > 0 ? deg : 1));
          uint4 hs;
          hs.x = fl_hmul2(hv.x, dg);
          hs.y = fl_hmul2(hv.y, dg);
          hs.z = fl_hmul2(hv.z, dg);
          hs.w = fl_hmul2(hv.w, dg);
          const uint32_t addr = fl_smem(zbuf) + (uint32_t)(zb * ZBYTES + 12 * SLAB + ((PPL - 1) * FL_NODES + j) * 128 + ((lane ^ j) << 4));
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(hs.x), "r"(hs.y), "r"(hs.z), "r"(hs.w) : "memory");
          if (PPL == 3) {
            const float2 h0 = fl_h2_to_f2(hs.x), h1 = fl_h2_to_f2(hs.y), h2 = fl_h2_to_f2(hs.z), h3 = fl_h2_to_f2(hs.w);
            const uint4 wroot = *reinterpret_cast<const uint4*>(wfs + 12 * 32 + 4 * lane);     // lanes 6, 7: zeros
            const float2 w0 = fl_h2_to_f2(wroot.x), w1 = fl_h2_to_f2(wroot.y), w2 = fl_h2_to_f2(wroot.z), w3 = fl_h2_to_f2(wroot.w);
            fsum += h0.x * w0.x + h0.y * w0.y + h1.x * w1.x + h1.y * w1.y + h2.x * w2.x + h2.y * w2.y + h3.x * w3.x + h3.y * w3.y;
          }
        }
        if (PPL == 3 && fix_b >= 0) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(FULL, fsum, o);
        }
        if (lane == 0) {
          fixs[(it & 3) * FL_NODES + j] = fsum;
          invs[(it & 3) * FL_NODES + j] = 1.0f / (float)(deg > 0 ? deg : 1);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(zfull + 8 * zb);
      src_a = src_b;
      src_b = src_c;
      src_c = src_n;
      buf = buf == NBUF - 1 ? 0 : buf + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp < FL_BW + 4) {
    // =========================================================================== epilogue
    const int qd = warp - FL_BW;                 // TMEM lane quadrant (== warp % 4)
    const int L = qd * 32 + lane;                // TMEM lane = A/D row
    const int p = L / rs, b = L % rs;
    const bool row_ok = p < PPL && b < FL_WP;
    // ---- T' rows of this launch's parts into TMEM (A operand): lane L <- TF[part0 + p][b][:]
    {
      const __half* src = tf + ((size_t)(part0 + (row_ok ? p : 0)) * FL_WP + (row_ok ? b : 0)) * FL_KP;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < FL_ACOLS; c0 += 16) {
        uint32_t r[16];
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok) t = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0) + v4);
          r[4 * v4] = t.x;
          r[4 * v4 + 1] = t.y;
          r[4 * v4 + 2] = t.z;
          r[4 * v4 + 3] = t.w;
        }
        fl_tmem_st16(taddr + c0, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(aready);
    }
    const int te = threadIdx.x - FL_BW * 32;     // 0..127
    constexpr int OUTI = (FL_NODES * 24 + 127) / 128;     // (node, channel pair) outputs per thread
    for (int it = 0; it < n_it; ++it) {
      const int as = it & 1;
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      fl_mbar_wait(dfull + 8 * as, (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + FL_ACOLS + as * N;
      uint32_t r[N / 16][16];
#pragma unroll
      for (int c = 0; c < N / 16; ++c) fl_tmem_ld16(taddr + c * 16, r[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // per-node scalars of this thread's outputs, read BEFORE the D stage is handed back (the builders
      // reuse a slot four tiles later, which needs that arrival first)
      float fxv[OUTI] = {}, inv[OUTI] = {};
#pragma unroll
      for (int i = 0; i < OUTI; ++i)
        if (te + 128 * i < FL_NODES * 24) {
          fxv[i] = fixs[(it & 3) * FL_NODES + (te + 128 * i) / 24];
          inv[i] = invs[(it & 3) * FL_NODES + (te + 128 * i) / 24];
        }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(dempty + 8 * as);     // the D stage is free again
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < FL_NODES; ++j) {
          uint32_t v = r[j / 16][j % 16];
          if (PPL > 1 && p == 1) v = r[(FL_NODES + j) / 16][(FL_NODES + j) % 16];
          if (PPL > 2 && p == 2) v = r[(2 * FL_NODES + j) / 16][(2 * FL_NODES + j) % 16];
          comb[(p * FL_NODES + j) * FL_WP + b] = __uint_as_float(v);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // (node j, channel pair) outputs: 8 x 24 pairs over 128 threads
#pragma unroll
      for (int i = 0; i < OUTI; ++i) {
        const int o = te + 128 * i;
        if (o >= FL_NODES * 24) break;
        const int j = o / 24, bb = (o % 24) * 2;
        const int64_t row = tile * FL_NODES + j;
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = bb + u;
          float s = 0.f;
          if (c < rs) {
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
              if (PPL == 3 && pp == PPL - 1 && c == fix_b) s += fxv[i];
              else s += comb[(pp * FL_NODES + j) * FL_WP + c];
            }
          }
          v[u] = s * inv[i];
        }
        if (row < n) {
          if (p_in) {
            const float2 pv = *reinterpret_cast<const float2*>(p_in + row * FL_WP + bb);
            v[0] += pv.x;
            v[1] += pv.y;
          }
          if (p_out) {
            *reinterpret_cast<float2*>(p_out + row * FL_WP + bb) = make_float2(v[0], v[1]);
          } else {
            v[0] += bias_p[bb];
            v[1] += bias_p[bb + 1];
            if (relu) {
              v[0] = fmaxf(v[0], 0.f);
              v[1] = fmaxf(v[1], 0.f);
            }
            *reinterpret_cast<uint32_t*>(h_out + row * FL_WP + bb) = fl_h2_sat(v[0], v[1]);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  } else {
    // =========================================================================== MMA issuer
    // instruction descriptor: D = F32, A = B = F16, both K-major, N, M = 128
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int nkb = has_root ? FL_NKB : FL_NKB - 1;
    fl_mbar_wait(aready, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int it = 0; it < n_it; ++it) {
      const int zb = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      fl_mbar_wait(dempty + 8 * zb, ph ^ 1);
      fl_mbar_wait(zfull + 8 * zb, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (fl_elect_one()) {
        const uint32_t tmem_d = tmem_base + FL_ACOLS + zb * N;
        const uint32_t zaddr = fl_smem(zbuf + (size_t)zb * ZBYTES);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t bdesc = fl_sw128_desc(zaddr + (uint32_t)(kb * SLAB));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            fl_umma_ts(tmem_d, tmem_base + (uint32_t)(kb * 32 + k * 8), bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        fl_umma_commit(zempty + 8 * zb);
        fl_umma_commit(dfull + 8 * zb);
      }
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == FL_BW) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int PPL, int NBUF>
static size_t fl_smem_bytes() {
  constexpr int NODES = FL_NODES;
  constexpr int N = (PPL * NODES + 15) / 16 * 16;
  constexpr size_t z = (size_t)2 * FL_NKB * N * 128;
  constexpr size_t st = (size_t)FL_NODES * NBUF * (FL_DEGC * (32 * PPL + 16) + (FL_DEGC + 1) * FL_SHB);
  constexpr size_t misc = (size_t)PPL * NODES * FL_WP * 4 + 8 * NODES * 4 + (PPL == 3 ? (12 * 32 + 32) * 4 : 0) + 10 * 8 + 16;
  return 1024 + z + st + misc;
}

template <int PPL, int NBUF>
static int launch_fl(const int32_t* rowptr, const int32_t* src_sorted, const __half* g3, int64_t E, const __half* h_in,
                     int64_t n, int part0, int has_root, const __half* tf, const float* bias_p, const float* p_in,
                     float* p_out, __half* h_out, int rs, int fix_b, int relu, cudaStream_t s) {
  size_t smem = fl_smem_bytes<PPL, NBUF>();
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: the kernel owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(layer_fused_f16_kernel<PPL, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t n_tiles = ceil_div(n, FL_NODES);
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  layer_fused_f16_kernel<PPL, NBUF><<<grid, (FL_NODES * PPL + 5) * 32, smem, s>>>(rowptr, src_sorted, g3, E, h_in, n, part0, has_root, tf,
                                                                  bias_p, p_in, p_out, h_out, rs, fix_b, relu);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// Shapes the fused layer covers: KernelNN-like rows (one pass of 48 g slots, padded width 48).
bool layer_fused_supported(const fesr_model_dims& d) {
  return d.kind == FESR_KERNELNN && d.passes == 1 && d.kp == 48 && d.wp == FL_WP;
}
int layer_fused_parts(const fesr_model_dims& d) { return d.kp / 16; }
size_t layer_fused_tf_elems(const fesr_model_dims& d) { return (size_t)layer_fused_parts(d) * FL_WP * FL_KP; }

int launch_prepare_tfused(const fesr_model_dims& d, const float* tprime, void* tf, cudaStream_t s) {
  const int64_t total = (int64_t)layer_fused_tf_elems(d);
  ProfScope prof(PROF_PREPARE, s);
  prepare_tfused_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(d, layer_fused_parts(d), tprime, static_cast<__half*>(tf));
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// One layer: h_out = relu(Z(h_in) T' + bias); g3 planar fp16 [parts][E][16]; P: fp32 [n, 48] scratch
// (only touched when the parts need more than one launch).  mode: parts per launch (0 = best).
int launch_layer_fused_f16(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g3,
                           int64_t E, const void* h_in, int64_t n, const void* tf, const float* bias_p, float* P,
                           void* h_out, int mode, cudaStream_t s) {
  if (n == 0) return FESR_OK;
  if (!layer_fused_supported(d)) {
    set_error("fused layer: unsupported model shape");
    return FESR_EINVAL;
  }
  const __half* gh = static_cast<const __half*>(g3);
  const __half* hh = static_cast<const __half*>(h_in);
  const __half* tfh = static_cast<const __half*>(tf);
  __half* ho = static_cast<__half*>(h_out);
  const int relu = 1;
  ProfScope prof(PROF_ZBUILD, s);
  int rc;
  if (mode == 0) mode = d.w <= 43 ? 3 : 2;
  if (mode == 3 && d.w <= 43) {
    // all three parts in one launch: part p in TMEM lanes [43p, 43p + 43); (part 2, channel 42) on CUDA cores
    return launch_fl<3, 4>(rowptr, src_sorted, gh, E, hh, n, 0, 1, tfh, bias_p, nullptr, nullptr, ho, 43,
                           d.w == 43 ? 42 : -1, relu, s);
  }
  if (mode == 2) {
    if ((rc = launch_fl<2, 4>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 64, -1, relu, s))) return rc;
    return launch_fl<1, 4>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 128, -1, relu, s);
  }
  if ((rc = launch_fl<1, 4>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 128, -1, relu, s))) return rc;
  if ((rc = launch_fl<1, 4>(rowptr, src_sorted, gh, E, hh, n, 1, 0, tfh, bias_p, P, P, nullptr, 128, -1, relu, s))) return rc;
  return launch_fl<1, 4>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 128, -1, relu, s);
}

}  // namespace fesr
