// One message-passing layer as ONE kernel, 16-node tiles (FESR_PREC_F16 arm; round 2 successor of layer_fused.cu).
//
//   h'_i = act( sum_{(s,a)} Z_i[s,a] T'[(s,a), :] + h_i root + bias ),   Z_i = 1/deg_i sum_{e->i} g_e (x) h[src_e]
//
// (reference: NNConv_old.forward/message/update + PyG mean aggregation, models/model.py:521-536, and the last Linear of
// the edge MLP, :311-315 -- reordered as in DESIGN.md section 2.)  Same operand placement as layer_fused.cu -- T' is the
// A operand of tcgen05.mma and lives in TMEM, part p of K in lanes [p*rs, p*rs + w), the Z tile is the B operand in
// shared memory (K-major, SWIZZLE_128B), nodes on N -- but what the per-role trace of that kernel showed
// (profiles/r02_layer_fused_role_trace.txt: three producer warps and the MMA-issuing thread saturated at once, on
// per-TILE instruction streams) is answered by the tile, not by the math:
//
//   * a tile is 16 consecutive nodes: N = 16 x PPL = 48 for the shipped model, so the 52 tcgen05.mma of a tile serve
//     twice the nodes (and no column of the instruction is padding), and every per-tile hand-off, header, barrier
//     round trip and epilogue barrier is paid once per 16 nodes;
//   * ONE Z tile (80 KB) instead of two 8-node tiles: the 16 consumer warps (warp = node) accumulate tile t + 1 in
//     registers while the tensor core contracts tile t, and only the 9 stmatrix of a warp sit between two MMA chains;
//     the fix-up row (reduced to 8 partial sums per node; the epilogue adds them) and max(deg, 1) are written AFTER
//     the hand-off, behind their own barrier;
//   * each producer warp stages WHOLE segments (<= 112 edges: an interior tile of the Kuhn mesh has 16 x 14 = 224 = two
//     segments) into ring slots it alone refills (segment s -> producer s % 3, slot s % 6): no per-tile work is
//     triplicated, no two warps ever touch one slot's barrier phase;
//   * the nodes' own h rows (root block) are read by the consumer warps straight from global memory at the top of the
//     tile -- they are needed a tile later -- instead of being staged;
//   * nothing the consumer loop needs is spilled: a local-memory reload misses the few KB of L1 that the gathers stream
//     through and costs ~300 cycles (ncu source view: 10 % of the consumers' samples sat on two LDL) -- the tile split
//     comes from the host as kernel arguments, and the accumulators are not zeroed (first k-step with a zero C).
//
// Measured on the B200, same box, 526 848 cells (tools/dev/fl_time.py): 8-node kernel 0.1770 ms; this kernel 0.1616 ms.
// Segment size matters through CONTENTION, not through overhead: 224-edge segments (one per tile: all 16 warps start
// their ldmatrix / mma.sync burst together) 0.186 ms, 56-edge segments 0.187 ms, 112-edge segments (two cohorts of 8
// warps half a phase apart) 0.1616 ms.  What bounds it now (ncu source-level stall sampling, profiles/r02i_*): the
// consumers -- ~4 500 warp-cycles per node of dependent, latency-bound instructions (ldmatrix -> mma.sync at 8 cycles per
// instruction and sub-partition on the legacy tensor path, shared with the tcgen05 chain: 13 % math-pipe throttle;
// stmatrix drain + proxy fence; header / address arithmetic), 16 nodes in flight per SM.  More nodes in flight would need
// more consumer warps (the register file and 1024 threads per CTA forbid it) or a second Z tile (shared memory).
//
// TMA tile::gather4 for the h[src] rows was measured first (tools/dev/gather4_probe.cu): it works with a {48, 1} box
// (SWIZZLE_NONE: rows at 96 B pitch; SWIZZLE_128B: rows at 128 B pitch, chunk ^ (row & 7)), but sustains only one
// 112-row segment per 683 cycles per SM against 305 for 16-byte cp.async from three warps, so the gathers stay LDGSTS.
#include "layer_fused.cuh"

namespace fesr {

constexpr int F2_NODES = 16;                  // nodes per tile
#ifndef F2_CAP_EDGES
#define F2_CAP_EDGES 112
#endif
constexpr int F2_CAP = F2_CAP_EDGES;          // staged edges per ring slot (A/B: 56 | 112 | 224)
#ifndef F2_RING_SLOTS
#define F2_RING_SLOTS (3 * 224 / F2_CAP_EDGES)
#endif
constexpr int F2_NBUF = F2_RING_SLOTS;        // ring slots
constexpr int F2_NPROD = 3;                   // producer warps (slot s belongs to producer s)
constexpr int F2_BW = F2_NODES;               // consumer warps
constexpr int F2_THREADS = (F2_BW + 4 + 1 + F2_NPROD) * 32;
constexpr int F2_HALF = F2_CAP < 112 ? F2_CAP : 112;      // a producer stages a segment in pieces of <= 112 rows (register budget of the source ids)
constexpr int F2_GOPS = (F2_HALF * 6 + 31) / 32;          // 16-byte gather chunk-ops per producer lane and piece
constexpr int F2_SPP = F2_NBUF / F2_NPROD;    // ring slots per producer
static_assert(F2_CAP % F2_HALF == 0 && F2_CAP % FL_DEGC == 0 && F2_NBUF % F2_NPROD == 0, "whole pieces per segment, whole slots per producer");

template <int PPL>
__global__ void __launch_bounds__(F2_THREADS, 1)
layer_fused16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                     const __half* __restrict__ g3, int64_t E, const __half* __restrict__ h_in, int64_t n,
                     int part0, int has_root, const __half* __restrict__ tf, const float* __restrict__ bias_p,
                     const float* p_in, float* p_out, __half* __restrict__ h_out,
                     int rs, int fix_b, int relu, int* ovf, int t_per, int t_extra, const __half* __restrict__ h_own) {
  // relu: bit 0 ReLU, bit 1 fp32 output rows, bit 2 SUM mode (the backward's reversed-graph pass: no 1/deg, the root
  // row enters unscaled); h_own: the nodes' own rows when they are not the gathered array's (backward: dpre against
  // dpre / deg[dst]), NULL = h_in
  // warp ids of the roles: consumers 0..15, epilogue 16..19 (TMEM lane quadrant = warp % 4), MMA issuer 20, producers 21..23
  constexpr int W_EPI0 = F2_BW;
  constexpr int W_MMA = F2_BW + 4;
  constexpr int W_PROD0 = F2_BW + 5;
  constexpr int N = PPL * F2_NODES;                    // MMA N (16 | 32 | 48)
  constexpr int SLAB = N * 128;                        // bytes per k-block of the Z tile
  constexpr int ZBYTES = FL_NKB * SLAB;
  constexpr int GPL = F2_CAP * 32;                     // bytes of one staged g plane
  constexpr int HOFF = PPL * GPL;                      // gathered h rows (96 B each)
  constexpr int HDR = HOFF + F2_CAP * 96;              // header: 17 row bounds, segment begin / end, last flag
  constexpr int STG = HDR + 96;
  constexpr int CMB = PPL * F2_NODES * FL_WP;          // floats of the combine buffer

  extern __shared__ __align__(1024) uint8_t f2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(f2_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* zbuf = smem;                                            // [13][N][128 B]
  uint8_t* stage = zbuf + ZBYTES;                                  // [NBUF][STG]
  float* comb = reinterpret_cast<float*>(stage + F2_NBUF * STG);   // [PPL][16][48]
  float* fxp = comb + CMB;                                         // [4][16][8]  partial sums of the fix-up row, tiles in flight
  float* degs = fxp + 4 * F2_NODES * 8;                            // [4][16]  max(deg, 1)
  uint32_t* zrow = reinterpret_cast<uint32_t*>(degs + 4 * F2_NODES);    // 32 B of zeros: masked g rows
  uint32_t* wfs = zrow + 8;                                        // [12][32] + [8][4]  fix-up row weights (PPL == 3)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wfs + (PPL == 3 ? 12 * 32 + 32 : 0));
  const uint32_t mdone = fl_smem(&bars[0]);                        // [2]  tensor core done with the tile of accumulator stage s
  const uint32_t dfree = fl_smem(&bars[2]);                        // [2]  accumulator stage s has been read (4 warps)
  const uint32_t zready = fl_smem(&bars[4]);                       // the Z tile is written (16 warps)
  const uint32_t fready = fl_smem(&bars[5]);                       // [4]  fix-up values / 1/deg of tile t & 3 are written (16 warps)
  const uint32_t sfull = fl_smem(&bars[9]);                        // [NBUF]
  const uint32_t sempty = sfull + 8 * F2_NBUF;                     // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[9 + 2 * F2_NBUF]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
#ifdef FL_TRACE
  long long tw[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long t_begin = 0, tmark = 0;
#endif
  // programmatic dependent launch: see layer_fused.cu
  asm volatile("griddepcontrol.launch_dependents;");
  // CTA b owns a contiguous chunk of tiles [t_first, t_first + n_it)
  // (t_per = n_tiles / grid and t_extra = n_tiles % grid come from the host: values the compiler can re-derive from the
  // constant bank instead of spilling -- a local-memory reload misses the L1 the gathers stream through)
  const int t_first = (int)blockIdx.x * t_per + min((int)blockIdx.x, t_extra);
  const int n_it = t_per + ((int)blockIdx.x < t_extra ? 1 : 0);       // grid <= n_tiles

  // the Z tile starts as zeros: root-block rows of the parts without a root and the zero tail of the root block are
  // never written afterwards
  for (int t = threadIdx.x; t < ZBYTES / 16; t += F2_THREADS)
    reinterpret_cast<uint4*>(zbuf)[t] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = threadIdx.x; t < F2_NBUF * STG / 16; t += F2_THREADS)
    reinterpret_cast<uint4*>(stage)[t] = make_uint4(0u, 0u, 0u, 0u);       // stale slab contents stay finite
  if (threadIdx.x < 8) zrow[threadIdx.x] = 0u;
  if (PPL == 3 && fix_b >= 0) {
    // fix-up row (layer_fused.cu): weights of output channel fix_b against a consumer lane's last-part values
    const __half* wr = tf + ((size_t)(part0 + PPL - 1) * FL_WP + fix_b) * FL_KP;
    for (int t = threadIdx.x; t < 12 * 32; t += F2_THREADS) {
      const int q = t >> 5, l = t & 31;
      wfs[t] = *reinterpret_cast<const uint32_t*>(wr + (q * 64 + (l >> 2) * 8 + 2 * (l & 3)));
    }
    for (int t = threadIdx.x; t < 32; t += F2_THREADS)
      wfs[12 * 32 + t] = t < 24 ? *reinterpret_cast<const uint32_t*>(wr + 12 * 64 + 2 * t) : 0u;
  }
  if (threadIdx.x == 0) {
    fl_mbar_init(mdone, 1);
    fl_mbar_init(mdone + 8, 1);
    fl_mbar_init(dfree, 4);
    fl_mbar_init(dfree + 8, 4);
    fl_mbar_init(zready, F2_BW);
    for (int s = 0; s < 4; ++s) fl_mbar_init(fready + 8 * s, F2_BW);
    for (int s = 0; s < F2_NBUF; ++s) {
      fl_mbar_init(sfull + 8 * s, 32 + 1);      // cp.async completions of the owner's 32 lanes + its explicit arrive
      fl_mbar_init(sempty + 8 * s, F2_BW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (A operand 416, two accumulator stages of N)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fl_smem(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // zero fill is read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- T' rows of this launch's parts into TMEM (A operand): lane L <- TF[part0 + L / rs][L % rs][:]; warps 0-3 load
  // the first half of the columns, warps 8-11 the second half (a warp reaches the TMEM lanes 32*(warp%4)...)
  if (warp < F2_BW && (warp & 7) < 4) {
    const int L = (warp & 3) * 32 + lane;
    const int p = L / rs, b = L % rs;
    const bool row_ok = p < PPL && b < FL_WP;
    const __half* src = tf + ((size_t)(part0 + (row_ok ? p : 0)) * FL_WP + (row_ok ? b : 0)) * FL_KP;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int cbeg = (warp >> 3) * (FL_ACOLS / 2);
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + FL_ACOLS / 2; c0 += 16) {
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // everything above touched only this launch's constants; h_in / P / h_out belong to the previous launch until it
  // has completed (no-op when this kernel was not launched as a programmatic dependent)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef FL_TRACE
  t_begin = clock64();
  tmark = t_begin;
#endif

  if (warp < F2_BW) {
    // =========================================================================== consumers (warp j = node j of the tile)
    const int j = warp;
    const int lr = lane & 7, lm = lane >> 3;
    const uint32_t zrow_u32 = fl_smem(zrow) + (uint32_t)((lm >> 1) * 16);
    // ldmatrix rows supplied by this lane (edge index within the 16-edge k-step):
    //   A = H^T tile mt: matrices {a 0-7, e 0-7}, {a 8-15, e 0-7}, {a 0-7, e 8-15}, {a 8-15, e 8-15}
    //   B = G of part pp: matrices {e 0-7, nt 0}, {e 8-15, nt 0}, {e 0-7, nt 1}, {e 8-15, nt 1}
    const int ka = (lm >> 1) * 8 + lr, kb = (lm & 1) * 8 + lr;
    // stmatrix row address: matrix lm = (nl, hh) of a-tile mt is k-block mt*4 + hh*2 + nl; this lane supplies row lr
    // (= a within the octet) -> 16-byte chunk lr (swizzled by the row index & 7) of Z-tile row pp*16 + j
    const uint32_t zb = fl_smem(zbuf);
    const uint32_t zst = zb + (uint32_t)(((lm & 1) * 2 + (lm >> 1)) * SLAB + j * 128 + ((lr ^ (j & 7)) << 4));
    const uint32_t zroot = zb + (uint32_t)(12 * SLAB + ((PPL - 1) * F2_NODES + j) * 128 + (((lane & 7) ^ (j & 7)) << 4));
    const bool do_fix = PPL == 3 && fix_b >= 0 && !FL_WHATIF(0x200);
    const uint32_t wfl = fl_smem(wfs) + (uint32_t)(lane * 4);
    const uint32_t st_u32 = fl_smem(stage);
    const uint32_t fxp_u32 = fl_smem(fxp), deg_u32 = fl_smem(degs);
    const uint4* h16 = reinterpret_cast<const uint4*>(h_own != nullptr ? h_own : h_in);      // own rows only (root block)
    const bool sum_mode = (relu & 4) != 0;
    int buf = 0;
    uint32_t par = 0;             // parity of the ring's current pass
    const int n32 = (int)n;
    int node = t_first * F2_NODES + j;
    for (int it = 0; it < n_it; ++it) {
      // D[pp][mt][nl][hh]: fp16x2 accumulators (rows gq + 8*hh of a-tile mt, slots 2tq, 2tq+1 of n-tile nl)
      // (not zeroed: the node's first 16-edge step writes them with a zero C operand)
      uint32_t d[PPL][3][2][2];
      bool fresh = true;
      // own h row (root block), chunk `lane` of 6: requested now, used after the tile's segments
      uint4 hv = make_uint4(0u, 0u, 0u, 0u);
      if (has_root && lane < 6 && node < n32) hv = __ldg(h16 + (uint32_t)node * 6u + (uint32_t)lane);
      node += F2_NODES;
      int deg = 0;
      int lastw;
      do {
        const uint32_t base = st_u32 + (uint32_t)(buf * STG);
        FL_TWAIT(0, fl_mbar_wait(sfull + 8 * buf, par))
        int seg_lo, seg_hi, pad;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(seg_lo), "=r"(seg_hi), "=r"(lastw), "=r"(pad) : "r"(base + (uint32_t)(HDR + 80)));
        {
          int eb, ee;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(eb) : "r"(base + (uint32_t)(HDR + 4 * j)));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ee) : "r"(base + (uint32_t)(HDR + 4 * j + 4)));
          deg = ee - eb;
          const int lo = max(eb, seg_lo), hi = min(ee, seg_hi);
          for (int c = lo; c < hi && !FL_WHATIF(0x400); c += FL_DEGC) {
            const int rem = hi - c;
            const int r0 = c - seg_lo;
            const int ra = min(r0 + ka, seg_hi - seg_lo - 1);            // rows past the node's range: any finite row
            const uint32_t a_addr = base + (uint32_t)(HOFF + ra * 96 + (((lm & 1) ^ ((ra >> 2) & 1)) << 4));
            const uint32_t b_addr = kb < rem ? base + (uint32_t)((r0 + kb) * 32 + (lm >> 1) * 16) : zrow_u32;
            const uint32_t b_step = kb < rem ? (uint32_t)GPL : 0u;
            uint32_t a[3][4];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) fl_ldsm4t(a_addr + (uint32_t)(mt * 32), a[mt]);
            if (fresh) {
#pragma unroll
              for (int pp = 0; pp < PPL; ++pp) {
                uint32_t b[4];
                fl_ldsm4t(b_addr + pp * b_step, b);
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) {
                  fl_mma16z(d[pp][mt][0], a[mt], b[0], b[1]);
                  fl_mma16z(d[pp][mt][1], a[mt], b[2], b[3]);
                }
              }
              fresh = false;
            } else {
#pragma unroll
              for (int pp = 0; pp < PPL; ++pp) {
                uint32_t b[4];
                fl_ldsm4t(b_addr + pp * b_step, b);
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) {
                  fl_mma16(d[pp][mt][0], a[mt], b[0], b[1]);
                  fl_mma16(d[pp][mt][1], a[mt], b[2], b[3]);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) fl_mbar_arrive(sempty + 8 * buf);     // this warp is done reading the segment
        if (++buf == F2_NBUF) {
          buf = 0;
          par ^= 1;
        }
      } while (lastw == 0);
      if (fresh) {        // a node without edges
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp)
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) d[pp][mt][0][0] = d[pp][mt][0][1] = d[pp][mt][1][0] = d[pp][mt][1][1] = 0u;
      }
      // ---- the Z tile is free once the tensor core has finished the previous tile: this tile's Z rows, raw sums
      // (the epilogue applies 1/deg)
      if (it > 0) FL_TWAIT(1, fl_mbar_wait(mdone + 8 * ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1)))
      FL_TMARK(9)
#pragma unroll
      for (int pp = 0; pp < PPL; ++pp)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
          asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(zst + (uint32_t)(pp * F2_NODES * 128 + mt * 4 * SLAB)),
                       "r"(d[pp][mt][0][0]), "r"(d[pp][mt][0][1]), "r"(d[pp][mt][1][0]), "r"(d[pp][mt][1][1])
                       : "memory");
      uint4 hs = make_uint4(0u, 0u, 0u, 0u);
      if (has_root && lane < 8) {
        // root block: h_i * max(deg, 1) so that the epilogue's 1/deg leaves h_i (lanes 6, 7: the zero tail)
        const float degf = sum_mode ? 1.f : (float)(deg > 0 ? deg : 1);
        const __half2 dg = __float2half2_rn(degf);
        hs.x = fl_hmul2(hv.x, dg);
        hs.y = fl_hmul2(hv.y, dg);
        hs.z = fl_hmul2(hv.z, dg);
        hs.w = fl_hmul2(hv.w, dg);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(zroot), "r"(hs.x), "r"(hs.y), "r"(hs.z), "r"(hs.w) : "memory");
        if (PPL >= 2) {
          // the same row scaled by 2^-8 for the part before the last: its root block holds the low-order term of
          // `root` (prepare_tfused_kernel), so h_i root is applied with two-term fp16 weights
          const __half2 dl = __float2half2_rn(degf * FESR_LO_SCALE);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(zroot - (uint32_t)(F2_NODES * 128)), "r"(fl_hmul2(hv.x, dl)),
                       "r"(fl_hmul2(hv.y, dl)), "r"(fl_hmul2(hv.z, dl)), "r"(fl_hmul2(hv.w, dl))
                       : "memory");
        }
      }
      // Z rows visible to the tensor core: tell the MMA issuer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(zready);
      FL_TMARK(2)
      // ---- off the critical path: the fix-up row (output channel fix_b of the last part, on CUDA cores from the
      // accumulator registers), reduced to 8 partial sums per node -- the epilogue adds them -- and max(deg, 1)
      if (do_fix) {
        __half2 f0 = __float2half2_rn(0.f), f1 = f0;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          f0 = __hfma2(fl_as_h2(d[PPL - 1][mt][0][0]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 0) * 128)), f0);     // (hh 0, nl 0)
          f1 = __hfma2(fl_as_h2(d[PPL - 1][mt][1][0]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 1) * 128)), f1);     // (hh 0, nl 1)
          f0 = __hfma2(fl_as_h2(d[PPL - 1][mt][0][1]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 2) * 128)), f0);     // (hh 1, nl 0)
          f1 = __hfma2(fl_as_h2(d[PPL - 1][mt][1][1]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 3) * 128)), f1);     // (hh 1, nl 1)
        }
        if (has_root && lane < 8) {
          uint4 wr;     // root weights of this lane's chunk (lanes 6, 7: zeros)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wr.x), "=r"(wr.y), "=r"(wr.z), "=r"(wr.w) : "r"(wfl + (uint32_t)(12 * 128 + lane * 12)));
          f0 = __hfma2(fl_as_h2(hs.x), fl_as_h2(wr.x), f0);
          f1 = __hfma2(fl_as_h2(hs.y), fl_as_h2(wr.y), f1);
          f0 = __hfma2(fl_as_h2(hs.z), fl_as_h2(wr.z), f0);
          f1 = __hfma2(fl_as_h2(hs.w), fl_as_h2(wr.w), f1);
        }
        __syncwarp();          // reconverge after the lane < 8 branch: the shuffles below must not take the divergent path
        const float2 a0 = __half22float2(f0), a1 = __half22float2(f1);
        float fsum = (a0.x + a0.y) + (a1.x + a1.y);
        fsum += __shfl_xor_sync(FULL, fsum, 16);
        fsum += __shfl_xor_sync(FULL, fsum, 8);
        if (lane < 8) asm volatile("st.shared.f32 [%0], %1;" ::"r"(fxp_u32 + (uint32_t)((((it & 3) * F2_NODES + j) * 8 + lane) * 4)), "f"(fsum) : "memory");
      }
      if (lane == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(deg_u32 + (uint32_t)(((it & 3) * F2_NODES + j) * 4)), "f"(sum_mode ? 1.f : (float)(deg > 0 ? deg : 1)) : "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(fready + 8 * (it & 3));
      FL_TMARK(3)
    }
  } else if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
    // =========================================================================== epilogue (4 warps, every tile)
    const int qd = warp - W_EPI0;                     // TMEM lane quadrant (== warp % 4)
    const int L = qd * 32 + lane;                     // TMEM lane = row (p, b) = (L / rs, L % rs)
    const int ep = L / rs, eb_ = L % rs;
    const bool row_ok = ep < PPL && eb_ < FL_WP;
    // the (at most two) parts this warp's 32 lanes belong to: only their 16 columns of the accumulator are loaded
    const int p_lo = min((qd * 32) / rs, PPL - 1), p_hi = min((qd * 32 + 31) / rs, PPL - 1);
    const int te = qd * 32 + lane;
    const int oq = te >> 3;                           // output phase: node oq of the tile, channels 6 cg .. 6 cg + 5
    const int oc = (te & 7) * 6;
    const uint32_t cmb = fl_smem(comb);
    const uint32_t fxp_u32 = fl_smem(fxp), deg_u32 = fl_smem(degs);
    const uint32_t cst = cmb + (uint32_t)((ep * F2_NODES * FL_WP + eb_) * 4);
    const uint32_t cld = cmb + (uint32_t)((oq * FL_WP + oc) * 4);
    float ob[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) ob[u] = bias_p[oc + u];
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      FL_TWAIT(0, fl_mbar_wait(mdone + 8 * s, (uint32_t)((it >> 1) & 1)))
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r0[16], r1[16];
      const uint32_t dcol = tmem_base + ((uint32_t)(qd * 32) << 16) + FL_ACOLS + s * N;
      fl_tmem_ld16(dcol + p_lo * F2_NODES, r0);
      if (p_hi != p_lo) fl_tmem_ld16(dcol + p_hi * F2_NODES, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(dfree + 8 * s);
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < F2_NODES; ++q) {
          const uint32_t v = (p_hi != p_lo && ep == p_hi) ? r1[q] : r0[q];
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(cst + (uint32_t)(q * FL_WP * 4)), "r"(v) : "memory");
        }
      }
      // per-node scalars of this thread's outputs (written by the consumers after their hand-off)
      FL_TWAIT(1, fl_mbar_wait(fready + 8 * (it & 3), (uint32_t)((it >> 2) & 1)))
      float fxv = 0.f;
      if (PPL == 3 && fix_b >= oc && fix_b < oc + 6) {      // the thread that owns the fix-up channel adds its node's 8 partial sums
        float4 q0, q1;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(fxp_u32 + (uint32_t)((((it & 3) * F2_NODES + oq) * 8) * 4)));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(fxp_u32 + (uint32_t)((((it & 3) * F2_NODES + oq) * 8 + 4) * 4)));
        fxv = ((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w));
      }
      float inv;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(inv) : "r"(deg_u32 + (uint32_t)(((it & 3) * F2_NODES + oq) * 4)));
      inv = __frcp_rn(inv);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int64_t row = (int64_t)(t_first + it) * F2_NODES + oq;
      float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int pp = 0; pp < PPL; ++pp)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          float2 cv;
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(cv.x), "=f"(cv.y) : "r"(cld + (uint32_t)((pp * F2_NODES * FL_WP + 2 * u) * 4)));
          const int c0 = oc + 2 * u;
          if (c0 < rs) v[2 * u] += (PPL == 3 && pp == PPL - 1 && c0 == fix_b) ? fxv : cv.x;
          if (c0 + 1 < rs) v[2 * u + 1] += (PPL == 3 && pp == PPL - 1 && c0 + 1 == fix_b) ? fxv : cv.y;
        }
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the combine buffer may be overwritten by the next tile
      if (row < n) {
#pragma unroll
        for (int u = 0; u < 6; ++u) v[u] *= inv;
        if (p_in) {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const float2 pv = *reinterpret_cast<const float2*>(p_in + row * FL_WP + oc + 2 * u);
            v[2 * u] += pv.x;
            v[2 * u + 1] += pv.y;
          }
        }
        if (p_out) {
#pragma unroll
          for (int u = 0; u < 3; ++u) *reinterpret_cast<float2*>(p_out + row * FL_WP + oc + 2 * u) = make_float2(v[2 * u], v[2 * u + 1]);
        } else {
          F16Guard guard;      // fp16 range guard (common.cuh): an inf / NaN out of the fp16-accumulated Z tile reaches these sums too
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            v[u] += ob[u];
            if (relu & 1) v[u] = fmaxf(v[u], 0.f);
          }
#pragma unroll
          for (int u = 0; u < 3; ++u) guard.note(v[2 * u], v[2 * u + 1]);
          guard.flush(ovf);
          if (relu & 2) {     // the model's last layer: fp32 rows for fc2 (one rounding less on the way out)
#pragma unroll
            for (int u = 0; u < 3; ++u)
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(h_out) + row * FL_WP + oc + 2 * u) = make_float2(v[2 * u], v[2 * u + 1]);
          } else {
#pragma unroll
            for (int u = 0; u < 3; ++u) *reinterpret_cast<uint32_t*>(h_out + row * FL_WP + oc + 2 * u) = fl_h2_sat(v[2 * u], v[2 * u + 1]);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // =========================================================================== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D = F32, A = B = F16 K-major, N, M = 128
    const int nkb = has_root ? FL_NKB : FL_NKB - 1;
    const uint64_t bdesc0 = fl_sw128_desc(fl_smem(zbuf));
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      FL_TWAIT(0, fl_mbar_wait(dfree + 8 * s, (uint32_t)(((it >> 1) & 1) ^ 1)))      // the epilogue has read this stage's previous result
      FL_TWAIT(1, fl_mbar_wait(zready, (uint32_t)(it & 1)))
      FL_TMARK(9)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (fl_elect_one()) {
        const uint32_t tmem_d = tmem_base + FL_ACOLS + s * N;
        // 52 MMAs from ONE thread, fully unrolled: the k-block / k-step offsets are immediates added to one base
        // descriptor (the start-address field holds addr >> 4 and cannot carry: shared memory ends below 256 KB)
#pragma unroll
        for (int kk = 0; kk < FL_NKB; ++kk) {
          if ((kk == FL_NKB - 1 && nkb < FL_NKB) || (FL_WHATIF(0x800) && kk >= 1)) break;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t ta = tmem_base + (uint32_t)(kk * 32 + q * 8);
            const uint64_t bd = bdesc0 + (uint64_t)((kk * SLAB + q * 32) >> 4);
            if (kk == 0 && q == 0) fl_umma_ts_c<false>(tmem_d, ta, bd, idesc);
            else fl_umma_ts_c<true>(tmem_d, ta, bd, idesc);
          }
        }
        fl_umma_commit(mdone + 8 * s);
      }
      __syncwarp();
      FL_TMARK(2)
    }
  } else {
    // =========================================================================== producers
    // Every producer walks the CTA's tiles and counts their segments (<= F2_CAP edges each; an empty tile has one empty
    // segment); segment number sg of the CTA goes to producer sg % 3, ring slot sg % F2_NBUF.  A producer stages its segments
    // alone: header (the tile's 17 row bounds, segment begin / end, last flag), one cp.async.bulk per g slot group, and
    // the 16-byte chunks of the gathered h[src] rows -- chunk-op t = i*32 + lane is chunk t % 6 of row t / 6, stored at
    // chunk ^ bit 2 of the row (conflict-free ldmatrix without padding).
    const int pw = warp - W_PROD0;
    const uint32_t st_u32 = fl_smem(stage);
    const uint4* h16 = reinterpret_cast<const uint4*>(h_in);
    const __half* gplane = g3 + (size_t)part0 * E * 16;
    const int n32 = (int)n, lane16 = lane < 17 ? lane : 16;
    const int tile0 = t_first;
    auto load_rp = [&](int itx) { return __ldg(rowptr + min((tile0 + itx) * F2_NODES + lane16, n32)); };
    int rp[4];                    // row bounds of tiles it .. it + 3 (lanes 0..16), indexed statically inside the x4 unroll
    rp[0] = load_rp(0);
    rp[1] = load_rp(1);
    rp[2] = load_rp(2);
    int cnt = 0;                  // (segment number) % 3
    int kown = 0;                 // own segments staged so far
#ifndef F2_PROD_UNROLL
#define F2_PROD_UNROLL 4
#endif
    for (int it0 = 0; it0 < n_it; it0 += F2_PROD_UNROLL) {
#pragma unroll
      for (int u = 0; u < F2_PROD_UNROLL; ++u) {
        const int it = it0 + u;
        if (it >= n_it) break;
#if F2_PROD_UNROLL == 4
        rp[(u + 3) & 3] = load_rp(it + 3);
#endif
        if (pw == 0 && lane == 0)      // rowptr lines of the tiles well ahead: into L2 now
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rowptr + min((tile0 + it + 32) * F2_NODES, n32)));
#if F2_PROD_UNROLL == 4
        const int rp0 = rp[u];
#else
        const int rp0 = load_rp(it);
#endif
        const int e_lo = __shfl_sync(FULL, rp0, 0), e_hi = __shfl_sync(FULL, rp0, 16);
        int seg_lo = e_lo;
        bool last;
        do {
          // A segment ends on a NODE boundary: the largest row bound within F2_CAP edges of its start (a node with more
          // edges than that is cut every F2_CAP edges, a multiple of the 16-edge k-step).  A node's edges are therefore
          // accumulated in the same 16-edge steps wherever its tile or shard begins: the result for a node does not depend
          // on the batch around it (block-diagonal independence, sharded == single-rank bit for bit).
          const unsigned fits = __ballot_sync(FULL, lane <= F2_NODES && rp0 <= seg_lo + F2_CAP);     // rp is monotone: lanes 0..b
          const int bnd = __shfl_sync(FULL, rp0, 31 - __clz(fits));
          const int seg_hi = bnd > seg_lo ? bnd : min(seg_lo + F2_CAP, e_hi);
          last = seg_hi == e_hi;
          if (cnt == pw) {
            const int nseg = seg_hi - seg_lo;
            const int sl = pw + F2_NPROD * (kown % F2_SPP);
            const uint32_t base = st_u32 + (uint32_t)(sl * STG);
            const uint32_t fb = sfull + 8 * sl;
            // source ids of this lane's chunk-ops of the first half: requested before the wait for the slot
            int sidx[F2_GOPS];
#pragma unroll
            for (int i = 0; i < F2_GOPS; ++i) sidx[i] = nseg > 0 ? __ldg(src_sorted + min(seg_lo + (i * 32 + lane) / 6, seg_hi - 1)) : 0;
            FL_TWAIT(0, fl_mbar_wait(sempty + 8 * sl, (uint32_t)(((kown / F2_SPP) & 1) ^ 1)))     // a fresh barrier passes a parity-1 wait
            FL_TMARK(9)
            if (lane < 20) {
              int hw = rp0;
              if (lane >= 17) hw = 0;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)(HDR + 4 * lane)), "r"(hw) : "memory");
            }
            if (lane < 4) {
              int hw = seg_lo;
              if (lane == 1) hw = seg_hi;
              if (lane == 2) hw = last ? 1 : 0;
              if (lane == 3) hw = 0;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)(HDR + 80 + 4 * lane)), "r"(hw) : "memory");
            }
            __syncwarp();
            // g slot groups: one bulk copy per part (the explicit arrive also releases the header stores)
            if (lane == 0) {
              if (nseg > 0) {
                const uint32_t bytes = (uint32_t)nseg * 32u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(bytes * PPL) : "memory");
#pragma unroll
                for (int pp = 0; pp < PPL; ++pp)
                  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + (uint32_t)(pp * GPL)),
                               "l"(gplane + ((size_t)pp * E + seg_lo) * 16), "r"(bytes), "r"(fb)
                               : "memory");
              } else {
                fl_mbar_arrive(fb);
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < F2_GOPS; ++i) {
              const int t = i * 32 + lane;
              const int row = t / 6, c = t - row * 6;
              if (row < nseg && !FL_WHATIF(0x100))
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(HOFF + row * 96 + ((c ^ ((row >> 2) & 1)) << 4))),
                             "l"(h16 + ((uint32_t)sidx[i] * 6u + (uint32_t)c))
                             : "memory");
            }
            if (F2_CAP > F2_HALF && nseg > F2_HALF) {       // second half of the segment
#pragma unroll
              for (int i = 0; i < F2_GOPS; ++i) sidx[i] = __ldg(src_sorted + min(seg_lo + F2_HALF + (i * 32 + lane) / 6, seg_hi - 1));
#pragma unroll
              for (int i = 0; i < F2_GOPS; ++i) {
                const int t = i * 32 + lane;
                const int row = F2_HALF + t / 6, c = t - (t / 6) * 6;
                if (row < nseg && !FL_WHATIF(0x100))
                  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(HOFF + row * 96 + ((c ^ ((row >> 2) & 1)) << 4))),
                               "l"(h16 + ((uint32_t)sidx[i] * 6u + (uint32_t)c))
                               : "memory");
              }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(fb) : "memory");
            FL_TMARK(2)
            ++kown;
          }
          cnt = cnt == F2_NPROD - 1 ? 0 : cnt + 1;
          seg_lo = seg_hi;
        } while (!last);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }

#ifdef FL_TRACE
  if (lane == 0) {
    long long* o = fl_trace_buf + ((size_t)blockIdx.x * 24 + warp) * 12;
    o[0] = clock64() - t_begin;
    o[1] = n_it;
    for (int q = 0; q < 10; ++q) o[2 + q] = tw[q];
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

#ifdef FL_TRACE
extern "C" int fesr_dev_fl16_trace(long long* host_out /* [148*24*12] */) {
  return cudaMemcpyFromSymbol(host_out, fl_trace_buf, sizeof(fl_trace_buf)) == cudaSuccess ? 0 : -1;
}
#endif

template <int PPL>
static size_t f2_smem_bytes() {
  constexpr size_t z = (size_t)FL_NKB * PPL * F2_NODES * 128;
  constexpr size_t st = (size_t)F2_NBUF * (PPL * F2_CAP * 32 + F2_CAP * 96 + 96);
  constexpr size_t misc = (size_t)PPL * F2_NODES * FL_WP * 4 + 4 * F2_NODES * 8 * 4 + 4 * F2_NODES * 4 + 32 + (PPL == 3 ? (12 * 32 + 32) * 4 : 0) +
                          (9 + 2 * F2_NBUF) * 8 + 16;
  return 1024 + z + st + misc;
}

template <int PPL>
int launch_fl16(const int32_t* rowptr, const int32_t* src_sorted, const __half* g3, int64_t E, const __half* h_in,
                int64_t n, int part0, int has_root, const __half* tf, const float* bias_p, const float* p_in,
                float* p_out, __half* h_out, int rs, int fix_b, int relu, cudaStream_t s, const __half* h_own) {
  int* ovf = cur_ovf();
  size_t smem = f2_smem_bytes<PPL>();
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: the kernel owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(layer_fused16_kernel<PPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t n_tiles = ceil_div(n, F2_NODES);
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  static const bool pdl = !(getenv("FESR_PDL") && atoi(getenv("FESR_PDL")) == 0);     // A/B switch for profiling
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(F2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const int t_per = (int)(n_tiles / grid), t_extra = (int)(n_tiles % grid);
  FESR_CUDA(cudaLaunchKernelEx(&cfg, layer_fused16_kernel<PPL>, rowptr, src_sorted, g3, E, h_in, n, part0, has_root, tf, bias_p,
                               p_in, p_out, h_out, rs, fix_b, relu, ovf, t_per, t_extra, h_own));
  count_launch();
  return FESR_OK;
}

template int launch_fl16<1>(const int32_t*, const int32_t*, const __half*, int64_t, const __half*, int64_t, int, int, const __half*,
                            const float*, const float*, float*, __half*, int, int, int, cudaStream_t, const __half*);
template int launch_fl16<2>(const int32_t*, const int32_t*, const __half*, int64_t, const __half*, int64_t, int, int, const __half*,
                            const float*, const float*, float*, __half*, int, int, int, cudaStream_t, const __half*);
template int launch_fl16<3>(const int32_t*, const int32_t*, const __half*, int64_t, const __half*, int64_t, int, int, const __half*,
                            const float*, const float*, float*, __half*, int, int, int, cudaStream_t, const __half*);

}  // namespace fesr
