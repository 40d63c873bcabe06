// One message-passing layer as ONE kernel (FESR_PREC_F16 arm, KernelNN shape):
//
//   h'_i = act( sum_{(s,a)} Z_i[s,a] T'[(s,a), :] + h_i root + bias ),   Z_i = 1/deg_i sum_{e->i} g_e (x) h[src_e]
//
// (reference: NNConv_old.forward/message/update + PyG mean aggregation, models/model.py:521-536, and
// the last Linear of the edge MLP, :311-315 -- reordered as in DESIGN.md section 2.)  The two-kernel
// path (zbuild_f16.cu + gemm_tc.cu) writes Z (4.4 KB per node) to HBM and reads it back; that round
// trip is two thirds of a layer's time.  Here Z never leaves the SM.  Warp-specialised CTA, one per SM,
// tiles of 8 consecutive nodes:
//
//   producers (3 warps)   stage the tile's CSR edge range into a shared-memory ring: one cp.async.bulk per
//                         g slot group, cp.async gathers of the h[src] rows, the 8 own h rows, a header;
//   consumers (2 x 8)     two groups play ping-pong over the tiles, warp j = node j: per 16 edges 6
//                         ldmatrix + 18 mma.sync.m16n8k16 (D = H_i^T [a x edges] . G_i [edges x slots],
//                         fp16 accumulate), then 9 stmatrix.x4 straight into the B-operand tile of the node
//                         contraction (K-major, SWIZZLE_128B, K ordered so that one 8x8 accumulator
//                         fragment is one 16-byte chunk of a k-block row: conflict-free);
//   MMA issuer (1 warp)   tcgen05.mma kind::f16, M = 128, N = 8 nodes x PPL parts, A = T' FROM TENSOR
//                         MEMORY (loaded once per CTA with tcgen05.st), B = the Z tile, D in TMEM;
//   epilogue (4 warps)    tcgen05.ld, combine the parts, 1/deg, (+ partial sums of earlier launches),
//                         bias, ReLU, fp16 h' rows.
//
// Nodes sit on the MMA N dimension because a full-K Z tile of >= 64 nodes (what M would need)
// does not fit in shared memory, and T' sits in TMEM because re-reading it from shared memory for
// every tile would cost more shared-memory bandwidth than Z itself.  TMEM holds 512 columns = 1024
// fp16 of K per lane, so K (2304 + 48) is cut into PARTS of 16 g-slots (768 = 12 k-blocks of 64)
// plus one root k-block.  Several parts share the 128 TMEM lanes (rows) at once: part p lives in
// lanes [p*RS, p*RS + w) and multiplies only the B columns that hold part p of the tile's nodes.
// With w <= 43 all three parts of the shipped model fit (3*43 = 129: the single row that does not
// fit, (part 2, channel 42), is evaluated on CUDA cores from the consumer's accumulator registers),
// so a layer is ONE launch and every edge is staged once.
//
// What bounds it (B200, 526 848-cell duct, clock64 accounting per role, profiles/r01_layer_fused_*):
// every role is a latency-bound instruction stream (about one instruction per 12-17 cycles per warp
// with 6 warps per scheduler), so the structure minimises instructions on the critical path and hands
// nothing between roles except through mbarriers that the next tile's work has already covered.
#include "layer_fused.cuh"

namespace fesr {

// ------------------------------------------------------------------------------------------------
// T' in the fused kernel's K order, fp16:  TF[part][b][K],  K = kb*64 + a_in*8 + s_in with
//   kb < 12 : a = (kb/2)*8 + a_in,  slot = part*16 + (kb%2)*8 + s_in  -> T'[(chan(slot), a), b]
//   kb = 12 : a = a_in*8 + s_in (root block, last part only; the tail 16 entries are zero)
// so that the 64 values one builder store instruction produces (8 a's x 8 slots) are one k-block row.
//
// Two-term weights where the rounding of T' to fp16 would otherwise dominate the arm's error.  The
// edge features are centred, g~ = g - g(0): sum_k g_k T'_k = sum_k g~_k T'_k + M with M = T'_K + sum_k g(0)_k T'_k
// (prepare_center_kernel, fp32).  For these models g varies by < 1 % over the mesh's edge lengths, so M -- applied
// to the neighbour mean of h through the constant-1 slot -- carries almost the whole layer, and so does `root`
// (applied to h_i): both get a hi + lo pair of fp16 weights at no run-time cost,
//   constant-1 slot : fp16(M)                  lo slot (a padding slot, g = 2^-8)      : fp16((M - fp16(M)) 2^8)
//   root block      : fp16(root)               root block of the part before the last : fp16((root - fp16(root)) 2^8)
//                                              (its B rows get h_i deg 2^-8 from the consumer)
// The other slots hold fp16(T'_k) against g~_k: their rounding errors now multiply a < 1 % signal.
__global__ void prepare_tfused_kernel(fesr_model_dims d, int n_parts, const float* __restrict__ tprime,
                                      const float* __restrict__ mfull, __half* __restrict__ tf) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_parts * FL_WP * FL_KP;
  if (idx >= total) return;
  const int K = (int)(idx % FL_KP);
  const int b = (int)((idx / FL_KP) % FL_WP);
  const int part = (int)(idx / ((int64_t)FL_KP * FL_WP));
  const int kb = K >> 6, e = K & 63;
  const float inv_lo = 1.0f / FESR_LO_SCALE;
  float v = 0.f;
  if (kb < 12) {
    const int a = (kb >> 1) * 8 + (e >> 3);
    const int slot = part * 16 + (kb & 1) * 8 + (e & 7);
    const int q = slot / d.ktp, r = slot % d.ktp;
    const int chan = q * d.kt + r;
    if (a < d.wp && b < d.wp) {
      if (slot == d.kt) {                                        // lo slot: low-order term of M
        const float m = mfull[a * d.wp + b];
        v = (m - __half2float(__float2half_rn(m))) * inv_lo;
      } else if (r < d.kt && chan == d.k1 - 1) {                 // constant-1 slot: M
        v = mfull[a * d.wp + b];
      } else if (r < d.kt && chan < d.k1) {
        v = tprime[((size_t)chan * d.wp + a) * d.wp + b];
      }
    }
  } else if (e < d.wp) {
    const float rt = tprime[((size_t)d.zk_main + e) * d.wp + b];
    if (part == n_parts - 1) v = rt;
    else if (part == n_parts - 2) v = (rt - __half2float(__float2half_rn(rt))) * inv_lo;      // low-order term of root
  }
  tf[idx] = __float2half_rn(v);
}

// PPL: parts per launch (1..3); NBUF: staged tile segments in flight.
//
// A tile is FL_NODES = 8 consecutive nodes; their CSR edges are one contiguous range, staged as one
// SEGMENT (<= FL_CAP edges; a tile with more edges takes several) into a ring of NBUF buffers:
//   * FL_NPROD PRODUCER warps work on a segment together: producer p bulk-copies the g slot group of
//     part p (planar [part][E][16] fp16 -> one cp.async.bulk of 32 B x edges) and issues every
//     FL_NPROD-th block of 32 sixteen-byte h[src] gather chunks (cp.async, rows of 96 B with chunk c
//     stored at c ^ bit 2 of the row index: conflict-free ldmatrix without padding) plus, on the tile's
//     last segment, the 8 nodes' own h rows.  Completion lands on the ring's `sfull` mbarrier
//     (complete_tx / cp.async.mbarrier.arrive.noinc); a header carries the 9 row bounds of the tile.
//   * two GROUPS of 8 CONSUMER warps play ping-pong over the tiles (group = tile parity, warp j = node j).
//     Per tile a group: per 16 edges of its node 6 ldmatrix (row addresses inside the segment, rows past the
//     node's range read a zero row) and 18 mma.sync.m16n8k16 (fp16 accumulate: one rounding, exactly what
//     the fp16 Z tile gets); then -- once the tensor core has finished the group's PREVIOUS tile, a wait that
//     the gather work above has already covered -- the epilogue of that previous tile (tcgen05.ld, combine
//     the parts, 1/deg, bias, ReLU, fp16 h' rows), 9 stmatrix.x4 of the new Z rows into the group's B-operand
//     tile, a group barrier, and one elected lane issues the tile's tcgen05.mma chain (A = T' in TMEM,
//     D = the group's TMEM accumulator) + commit.  No role hands anything to another role except the ring.
// MMA N = 8 * PPL rounded up to 16 (row pp*8 + j of the Z tile = part pp of node j).
template <int PPL, int NBUF>
__global__ void __launch_bounds__((16 + 3 + 1 + 4) * 32, 1)
layer_fused_f16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src_sorted,
                       const __half* __restrict__ g3, int64_t E, const __half* __restrict__ h_in, int64_t n,
                       int part0, int has_root, const __half* __restrict__ tf, const float* __restrict__ bias_p,
                       const float* p_in, float* p_out, __half* __restrict__ h_out,
                       int rs, int fix_b, int relu, int* ovf) {
  constexpr int FL_BW = 2 * FL_NODES;                  // consumer warps: two groups of 8
  constexpr int FL_NPROD = 3;                          // producer warps
  constexpr int FL_THREADS = (FL_BW + FL_NPROD + 1 + 4) * 32;   // + the MMA issuer warp + 4 epilogue warps
  // warp ids of the roles: consumers 0..15, epilogue 16..19 (TMEM lane quadrant = warp % 4), MMA issuer 20,
  // producers 21..23 -- FL_ROLE_ORDER 0 is the older order (producers 16..18, issuer 19, epilogue 20..23)
#ifndef FL_ROLE_ORDER
#define FL_ROLE_ORDER 1
#endif
  constexpr int W_EPI0 = FL_ROLE_ORDER ? FL_BW : FL_BW + FL_NPROD + 1;
  constexpr int W_MMA = FL_ROLE_ORDER ? FL_BW + 4 : FL_BW + FL_NPROD;
  constexpr int W_PROD0 = FL_ROLE_ORDER ? FL_BW + 5 : FL_BW;
  constexpr int N = (PPL * FL_NODES + 15) / 16 * 16;   // MMA N
  constexpr int SLAB = N * 128;                        // bytes per k-block of the Z tile
  constexpr int ZBYTES = FL_NKB * SLAB;
  constexpr int GPL = FL_CAP * 32;                     // bytes of one staged g plane
  constexpr int HOFF = PPL * GPL;                      // gathered h rows (96 B each)
  constexpr int OWN = HOFF + FL_CAP * 96;              // the 8 nodes' own h rows
  constexpr int HDR = OWN + FL_NODES * 96;             // header: 9 row bounds, segment begin / end, last flag
  constexpr int STG = HDR + 64;
  constexpr int CMB = PPL * FL_NODES * FL_WP;          // floats of one group's combine buffer

  extern __shared__ __align__(1024) uint8_t fl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* zbuf = smem;                                            // [2 groups][ZBYTES]
  uint8_t* stage = zbuf + 2 * ZBYTES;                              // [NBUF][STG]
  float* comb = reinterpret_cast<float*>(stage + NBUF * STG);      // [PPL][8][48]
  float* fixs = comb + CMB;                                        // [2 groups][2][8]  fix-up row values
  float* invs = fixs + 4 * FL_NODES;                               // [2 groups][2][8]  1 / max(deg, 1)
  uint32_t* zrow = reinterpret_cast<uint32_t*>(invs + 4 * FL_NODES);    // 32 B of zeros: masked g rows
  uint32_t* wfs = zrow + 8;                                        // [12][32] + [8][4]  fix-up row weights (PPL == 3)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wfs + (PPL == 3 ? 12 * 32 + 32 : 0));
  const uint32_t mdone = fl_smem(&bars[0]);                        // [2]  tensor core done with a group's tile
  const uint32_t zready = fl_smem(&bars[2]);                       // [2]  a group's Z tile is written (8 warps)
  const uint32_t dfree = fl_smem(&bars[4]);                        // [2]  a group's TMEM accumulator has been read (4 warps)
  const uint32_t sfull = fl_smem(&bars[6]);                        // [NBUF]
  const uint32_t sempty = sfull + 8 * NBUF;                        // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[6 + 2 * NBUF]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
#ifdef FL_TRACE
  long long tw[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long t_begin = 0, tmark = 0;
#endif
  // Programmatic dependent launch: the next launch in the stream (the next layer) may be scheduled onto an SM as
  // soon as this CTA has left it, and runs its prologue (zero fill, barriers, TMEM allocation, T' into TMEM --
  // nothing that depends on this layer) under the tail of this grid; it waits for this grid's completion below.
  asm volatile("griddepcontrol.launch_dependents;");
  // CTA b owns a CONTIGUOUS chunk of tiles [t_first, t_first + n_it): consecutive tiles share rowptr / source-id cache
  // lines and the CTA's gathers stay inside a compact range of h rows.  (Tiles used to be dealt out round-robin with a
  // stride of gridDim.x: every rowptr read was then a lone DRAM miss whose latency the producers -- who need rowptr(t)
  // one tile period after requesting it, to address the source ids -- could not cover: the per-role trace of
  // profiles/r02_layer_fused_role_trace.txt showed them busy 2 350 of 2 430 cycles per tile without ever waiting.)
  const int64_t n_tiles = (n + FL_NODES - 1) / FL_NODES;
  const int64_t t_per = n_tiles / gridDim.x, t_extra = n_tiles % gridDim.x;
  const int64_t t_first = blockIdx.x * t_per + ((int64_t)blockIdx.x < t_extra ? (int64_t)blockIdx.x : t_extra);
  const int n_it = (int)(t_per + ((int64_t)blockIdx.x < t_extra ? 1 : 0));       // grid <= n_tiles

  // the Z tiles start as zeros: unused rows, root-block rows of the parts without a root and the zero
  // tail of the root block are never written afterwards
  for (int t = threadIdx.x; t < 2 * ZBYTES / 16; t += FL_THREADS)
    reinterpret_cast<uint4*>(zbuf)[t] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = threadIdx.x; t < NBUF * STG / 16; t += FL_THREADS)
    reinterpret_cast<uint4*>(stage)[t] = make_uint4(0u, 0u, 0u, 0u);       // stale slab contents stay finite
  if (threadIdx.x < 8) zrow[threadIdx.x] = 0u;
  if (PPL == 3 && fix_b >= 0) {
    // fix-up row (see header): weights of output channel fix_b against a consumer lane's last-part values,
    // in the order the lane holds them: wfs[(mt*2+hh)*2+nl][lane], then the root block as 6 x 16 bytes (+ zeros)
    const __half* wr = tf + ((size_t)(part0 + PPL - 1) * FL_WP + fix_b) * FL_KP;
    for (int t = threadIdx.x; t < 12 * 32; t += FL_THREADS) {
      const int q = t >> 5, l = t & 31;
      wfs[t] = *reinterpret_cast<const uint32_t*>(wr + (q * 64 + (l >> 2) * 8 + 2 * (l & 3)));
    }
    for (int t = threadIdx.x; t < 32; t += FL_THREADS)
      wfs[12 * 32 + t] = t < 24 ? *reinterpret_cast<const uint32_t*>(wr + 12 * 64 + 2 * t) : 0u;
  }
  if (threadIdx.x == 0) {
    fl_mbar_init(mdone, 1);
    fl_mbar_init(mdone + 8, 1);
    fl_mbar_init(zready, FL_NODES);
    fl_mbar_init(zready + 8, FL_NODES);
    fl_mbar_init(dfree, 4);
    fl_mbar_init(dfree + 8, 4);
    for (int s = 0; s < NBUF; ++s) {
      fl_mbar_init(sfull + 8 * s, FL_NPROD * 32 + FL_NPROD);   // cp.async completions of every producer lane + one arrive per producer
      fl_mbar_init(sempty + 8 * s, FL_NODES);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (A operand 416, one D stage of N per group)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fl_smem(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // zero fill is read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // TMEM lane L holds row (p, b) = (L / rs, L % rs): part p, output channel b
  // ---- T' rows of this launch's parts into TMEM (A operand): lane L <- TF[part0 + p][b][:]; warps 0-3 load
  // the first half of the columns, warps 8-11 the second half (a warp reaches the TMEM lanes 32*(warp%4)...)
  if (warp < FL_BW && (warp & 7) < 4) {
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

  if (warp < FL_BW) {
    // =========================================================================== consumers
    const int grp = warp >> 3, j = warp & 7;
    const int lr = lane & 7, lm = lane >> 3;
    const uint32_t zrow_u32 = fl_smem(zrow) + (uint32_t)((lm >> 1) * 16);
    // ldmatrix rows supplied by this lane (edge index within the 16-edge k-step):
    //   A = H^T tile mt: matrices {a 0-7, e 0-7}, {a 8-15, e 0-7}, {a 0-7, e 8-15}, {a 8-15, e 8-15}
    //   B = G of part pp: matrices {e 0-7, nt 0}, {e 8-15, nt 0}, {e 0-7, nt 1}, {e 8-15, nt 1}
    const int ka = (lm >> 1) * 8 + lr, kb = (lm & 1) * 8 + lr;
    // stmatrix row address: the 4 matrices of a-tile mt are stored in accumulator register order
    // (nl 0, hh 0), (nl 0, hh 1), (nl 1, hh 0), (nl 1, hh 1); matrix lm = (nl, hh) is k-block mt*4 + hh*2 + nl;
    // this lane supplies row lr (= a within the octet) -> 16-byte chunk lr (swizzled) of Z-tile row pp*8 + j
    const uint32_t zgrp = fl_smem(zbuf) + (uint32_t)(grp * ZBYTES);
    const uint32_t zst = zgrp + (uint32_t)(((lm & 1) * 2 + (lm >> 1)) * SLAB + j * 128 + ((lr ^ j) << 4));
    const uint32_t zroot = zgrp + (uint32_t)(12 * SLAB + ((PPL - 1) * FL_NODES + j) * 128 + ((lane ^ j) << 4));
    const uint32_t md = mdone + 8 * grp;
    const bool do_fix = PPL == 3 && fix_b >= 0 && !FL_WHATIF(0x200);
    const uint32_t wfl = fl_smem(wfs) + (uint32_t)(lane * 4);

    // the group's own ring: buffers grp*NBG .. grp*NBG + NBG - 1 hold the segments of its tiles
    constexpr int NBG = NBUF / 2;
    const uint32_t rg_stage = fl_smem(stage) + (uint32_t)(grp * NBG * STG);
    const uint32_t rg_full = sfull + 8 * grp * NBG, rg_empty = sempty + 8 * grp * NBG;
    int buf = 0;
    uint32_t par = 0;             // parity of the ring's current pass
    int n_mine = 0;               // own tiles issued so far
    for (int it = grp; it < n_it; it += 2) {
      // D[pp][mt][nl][hh]: fp16x2 accumulators (rows gq + 8*hh of a-tile mt, slots 2tq, 2tq+1 of n-tile nl)
      uint32_t d[PPL][3][2][2];
#pragma unroll
      for (int pp = 0; pp < PPL; ++pp)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) d[pp][mt][0][0] = d[pp][mt][0][1] = d[pp][mt][1][0] = d[pp][mt][1][1] = 0u;
      int deg = 0;
      int lastw;
      uint4 hv = make_uint4(0u, 0u, 0u, 0u);       // own h row chunk (root block)
      do {
        const uint32_t base = rg_stage + (uint32_t)(buf * STG);
        FL_TWAIT(0, fl_mbar_wait(rg_full + 8 * buf, par))
        int seg_lo, seg_hi, pad;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(seg_lo), "=r"(seg_hi), "=r"(lastw), "=r"(pad) : "r"(base + (uint32_t)(HDR + 48)));
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
          if (lastw != 0 && has_root && lane < 6)
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "r"(base + (uint32_t)(OWN + j * 96 + lane * 16)));
        }
        __syncwarp();
        if (lane == 0) fl_mbar_arrive(rg_empty + 8 * buf);     // this warp is done reading the segment
        if (++buf == NBG) {
          buf = 0;
          par ^= 1;
        }
      } while (lastw == 0);
      // ---- once the tensor core has finished this group's previous tile (a wait the gather work above has
      // covered) the Z buffer is free: this tile's Z rows, raw sums (the epilogue applies 1/deg)
      if (n_mine > 0) FL_TWAIT(1, fl_mbar_wait(md, (uint32_t)((n_mine - 1) & 1)))
#pragma unroll
      for (int pp = 0; pp < PPL; ++pp)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
          asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(zst + (uint32_t)(pp * FL_NODES * 128 + mt * 4 * SLAB)),
                       "r"(d[pp][mt][0][0]), "r"(d[pp][mt][0][1]), "r"(d[pp][mt][1][0]), "r"(d[pp][mt][1][1])
                       : "memory");
      __half2 f0 = __float2half2_rn(0.f), f1 = f0;
      if (do_fix) {
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          f0 = __hfma2(fl_as_h2(d[PPL - 1][mt][0][0]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 0) * 128)), f0);     // (hh 0, nl 0)
          f1 = __hfma2(fl_as_h2(d[PPL - 1][mt][1][0]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 1) * 128)), f1);     // (hh 0, nl 1)
          f0 = __hfma2(fl_as_h2(d[PPL - 1][mt][0][1]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 2) * 128)), f0);     // (hh 1, nl 0)
          f1 = __hfma2(fl_as_h2(d[PPL - 1][mt][1][1]), fl_as_h2(fl_lds32(wfl + (mt * 4 + 3) * 128)), f1);     // (hh 1, nl 1)
        }
      }
      if (has_root && lane < 8) {
        // root block: h_i * max(deg, 1) so that the epilogue's 1/deg leaves h_i (lanes 6, 7: the zero tail)
        const __half2 dg = __float2half2_rn((float)(deg > 0 ? deg : 1));
        uint4 hs;
        hs.x = fl_hmul2(hv.x, dg);
        hs.y = fl_hmul2(hv.y, dg);
        hs.z = fl_hmul2(hv.z, dg);
        hs.w = fl_hmul2(hv.w, dg);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(zroot), "r"(hs.x), "r"(hs.y), "r"(hs.z), "r"(hs.w) : "memory");
        if (PPL >= 2) {
          // the same row scaled by 2^-8 for the part before the last: its root block holds the low-order term of
          // `root` (prepare_tfused_kernel), so h_i root is applied with two-term fp16 weights
          const __half2 dl = __float2half2_rn((float)(deg > 0 ? deg : 1) * FESR_LO_SCALE);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(zroot - (uint32_t)(FL_NODES * 128)), "r"(fl_hmul2(hv.x, dl)),
                       "r"(fl_hmul2(hv.y, dl)), "r"(fl_hmul2(hv.z, dl)), "r"(fl_hmul2(hv.w, dl))
                       : "memory");
        }
        if (do_fix) {
          uint4 wr;     // root weights of this lane's chunk (lanes 6, 7: zeros)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wr.x), "=r"(wr.y), "=r"(wr.z), "=r"(wr.w) : "r"(wfl + (uint32_t)(12 * 128 + lane * 12)));
          f0 = __hfma2(fl_as_h2(hs.x), fl_as_h2(wr.x), f0);
          f1 = __hfma2(fl_as_h2(hs.y), fl_as_h2(wr.y), f1);
          f0 = __hfma2(fl_as_h2(hs.z), fl_as_h2(wr.z), f0);
          f1 = __hfma2(fl_as_h2(hs.w), fl_as_h2(wr.w), f1);
        }
      }
      float fsum = 0.f;
      __syncwarp();          // reconverge after the lane < 8 branch: the shuffles below must not take the divergent path
      if (do_fix) {
        const float2 a0 = __half22float2(f0), a1 = __half22float2(f1);
        fsum = (a0.x + a0.y) + (a1.x + a1.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(FULL, fsum, o);
      }
      if (lane == 0) {
        fixs[(grp * 2 + (n_mine & 1)) * FL_NODES + j] = fsum;
        invs[(grp * 2 + (n_mine & 1)) * FL_NODES + j] = __frcp_rn((float)(deg > 0 ? deg : 1));
      }
      // Z rows visible to the tensor core (and this warp's TMEM loads of the previous epilogue retired):
      // tell the group's MMA issuer warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(zready + 8 * grp);
      ++n_mine;
    }
  } else if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
    // =========================================================================== epilogue (4 warps, every tile)
    const int qd = warp - W_EPI0;                     // TMEM lane quadrant (== warp % 4)
    const int L = qd * 32 + lane;                     // TMEM lane = row (p, b) = (L / rs, L % rs)
    const int ep = L / rs, eb_ = L % rs;
    const bool row_ok = ep < PPL && eb_ < FL_WP;
    const int te = qd * 32 + lane;
    constexpr int OUTI = (FL_NODES * 24 + 127) / 128;      // (node, channel pair) outputs per thread
    const uint32_t cmb = fl_smem(comb);
    int oj[OUTI], oc[OUTI];
    float ob0[OUTI], ob1[OUTI];
#pragma unroll
    for (int i = 0; i < OUTI; ++i) {
      const int o = te + 128 * i;
      oj[i] = o < FL_NODES * 24 ? o / 24 : -1;
      oc[i] = (o % 24) * 2;
      ob0[i] = bias_p[oc[i]];
      ob1[i] = bias_p[oc[i] + 1];
    }
    for (int it = 0; it < n_it; ++it) {
      const int grp = it & 1, k = it >> 1;
      FL_TWAIT(0, fl_mbar_wait(mdone + 8 * grp, (uint32_t)(k & 1)))
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[N / 16][16];
#pragma unroll
      for (int c = 0; c < N / 16; ++c) fl_tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + FL_ACOLS + grp * N + c * 16, r[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // per-node scalars of this thread's outputs, read BEFORE the accumulator is handed back (the consumers
      // reuse the slot two own tiles later, which needs that arrival first)
      float fxv[OUTI], inv[OUTI];
#pragma unroll
      for (int i = 0; i < OUTI; ++i) {
        const int q = oj[i] < 0 ? 0 : oj[i];
        fxv[i] = fixs[(grp * 2 + (k & 1)) * FL_NODES + q];
        inv[i] = invs[(grp * 2 + (k & 1)) * FL_NODES + q];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) fl_mbar_arrive(dfree + 8 * grp);
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < FL_NODES; ++q) {
          uint32_t v = r[q / 16][q % 16];
          if (PPL > 1 && ep == 1) v = r[(FL_NODES + q) / 16][(FL_NODES + q) % 16];
          if (PPL > 2 && ep == 2) v = r[(2 * FL_NODES + q) / 16][(2 * FL_NODES + q) % 16];
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(cmb + (uint32_t)(((ep * FL_NODES + q) * FL_WP + eb_) * 4)), "r"(v) : "memory");
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int64_t row0 = (t_first + it) * FL_NODES;
#pragma unroll
      for (int i = 0; i < OUTI; ++i) {
        const int64_t row = row0 + oj[i];
        if (oj[i] >= 0 && row < n) {
          const int c0 = oc[i];
          float v0 = 0.f, v1 = 0.f;
#pragma unroll
          for (int pp = 0; pp < PPL; ++pp) {
            float2 cv;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(cv.x), "=f"(cv.y) : "r"(cmb + (uint32_t)(((pp * FL_NODES + oj[i]) * FL_WP + c0) * 4)));
            if (c0 < rs) v0 += (PPL == 3 && pp == PPL - 1 && c0 == fix_b) ? fxv[i] : cv.x;
            if (c0 + 1 < rs) v1 += (PPL == 3 && pp == PPL - 1 && c0 + 1 == fix_b) ? fxv[i] : cv.y;
          }
          v0 *= inv[i];
          v1 *= inv[i];
          if (p_in) {
            const float2 pv = *reinterpret_cast<const float2*>(p_in + row * FL_WP + c0);
            v0 += pv.x;
            v1 += pv.y;
          }
          if (p_out) {
            *reinterpret_cast<float2*>(p_out + row * FL_WP + c0) = make_float2(v0, v1);
          } else {
            v0 += ob0[i];
            v1 += ob1[i];
            if (relu & 1) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            {   // fp16 range guard (common.cuh): an inf / NaN out of the fp16-accumulated Z tile reaches these sums too;
                // raised on the spot -- no loop-carried register in this role's 80-register budget
              F16Guard guard;
              guard.note(v0, v1);
              guard.flush(ovf);
            }
            if (relu & 2)      // the model's last layer: fp32 rows for fc2 (one rounding less on the way out)
              *reinterpret_cast<float2*>(reinterpret_cast<float*>(h_out) + row * FL_WP + c0) = make_float2(v0, v1);
            else
              *reinterpret_cast<uint32_t*>(h_out + row * FL_WP + c0) = fl_h2_sat(v0, v1);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  } else if (warp == W_MMA) {
    // =========================================================================== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D = F32, A = B = F16 K-major, N, M = 128
    const int nkb = has_root ? FL_NKB : FL_NKB - 1;
    for (int it = 0; it < n_it; ++it) {
      const int grp = it & 1;
      FL_TWAIT(0, fl_mbar_wait(dfree + 8 * grp, (uint32_t)(((it >> 1) & 1) ^ 1)))     // the epilogue has read the previous result
      FL_TWAIT(1, fl_mbar_wait(zready + 8 * grp, (uint32_t)((it >> 1) & 1)))
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (fl_elect_one()) {
        const uint32_t zgrp = fl_smem(zbuf) + (uint32_t)(grp * ZBYTES);
        const uint32_t tmem_d = tmem_base + FL_ACOLS + grp * N;
        // The 52 MMAs of a tile are issued by ONE thread: every instruction it spends per MMA on descriptor arithmetic
        // or predicates stretches the chain (it was busy 2 180 cycles per tile, 28 per MMA: issue-bound, not tensor
        // bound).  Fully unrolled, the k-block / k-step offsets are immediates added to one base descriptor (the
        // start-address field holds addr >> 4 and cannot carry: shared memory ends below 256 KB).
        const uint64_t bdesc0 = fl_sw128_desc(zgrp);
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
        fl_umma_commit(mdone + 8 * grp);
      }
      __syncwarp();
    }
  } else {
    // =========================================================================== producers
    const int pw = warp - W_PROD0;
    const uint32_t st_u32 = fl_smem(stage);
    // This lane's gather chunk-ops of a segment: op i is 16-byte chunk tc of staged row tr[i]; the flat index
    // t = (i * FL_NPROD + pw) * 32 + lane runs over rows x 6 chunks.  Destination offsets are lane constants.
    // (One 96-byte cp.async.bulk per row was tried: ~30 cycles per small bulk copy, slower than these.)
    // With 32 * FL_NPROD = 96 = 16 rows x 6 chunks per round, op i of a lane is chunk c0 of row r0 + 16 i, and its
    // swizzled destination (chunk ^ bit 2 of the row) is g0 + 1536 i: three lane constants describe all ops.
    constexpr int GOPS = (FL_CAP * 6 + 32 * FL_NPROD - 1) / (32 * FL_NPROD);
    static_assert(FL_NPROD == 3, "the lane constants below assume 96 chunk-ops per round");
    const int t0 = pw * 32 + lane;
    const int r0 = t0 / 6;
    const uint32_t c0 = (uint32_t)(t0 % 6);
    const uint32_t g0 = (uint32_t)(HOFF + ((t0 ^ ((r0 >> 2) & 1)) << 4));
    // the 8 nodes' own rows: 48 chunk-ops, lanes of producers 0 and 1
    const int ot = pw * 32 + lane;
    const bool own_lane = ot < FL_NODES * 6;
    const int orow = ot / 6;
    const uint32_t odst = (uint32_t)(OWN + orow * 96 + (ot % 6) * 16);
    const uint4* h16 = reinterpret_cast<const uint4*>(h_in);         // h rows as 6 sixteen-byte chunks
    const __half* gplane = g3 + (size_t)(part0 + (pw < PPL ? pw : 0)) * E * 16;   // this producer's g slot group
    const int n32 = (int)n, lane9 = lane < 9 ? lane : 8;
    const int tile0 = (int)t_first, tstep = 1;
    // rowptr of the tile's 9 node boundaries in lanes 0..8 (node ids fit 31 bits)
    auto load_rp = [&](int itx) { return __ldg(rowptr + min((tile0 + itx * tstep) * FL_NODES + lane9, n32)); };
    // source ids of this lane's gather ops for the first segment of a tile (clamped addresses)
    auto load_srcs = [&](int rpv, int (&sv)[GOPS]) {
      const int e_lo = __shfl_sync(FULL, rpv, 0), e_max = max(__shfl_sync(FULL, rpv, 8) - 1, 0);
#pragma unroll
      for (int i = 0; i < GOPS; ++i) sv[i] = __ldg(src_sorted + min(e_lo + r0 + 16 * i, e_max));
    };
    // Prefetch registers, indexed by (tile & 3) inside a loop unrolled by 4 so that no in-flight load is ever
    // moved between registers (a move would wait for it): rowptr of tile t is requested 3 tiles ahead, the
    // source ids of tile t 2 tiles ahead (their addresses need rowptr(t), one tile old by then).
    int rp[4];
    int sc[4][GOPS];
    rp[0] = load_rp(0);
    rp[1] = load_rp(1);
    rp[2] = load_rp(2);
    load_srcs(rp[0], sc[0]);
    load_srcs(rp[1], sc[1]);
    constexpr int NBG = NBUF / 2; // ring of the tile-parity group: buffers g*NBG ..
    int bufs[2] = {0, 0};
    uint32_t pars[2] = {1u, 1u};  // a fresh `sempty` barrier passes a parity-1 wait
    for (int it0 = 0; it0 < n_it; it0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u;
        if (it >= n_it) break;
        rp[(u + 3) & 3] = load_rp(it + 3);
        if (pw == 0 && lane == 0)      // rowptr lines of the tiles well ahead: into L2 now, so that load_rp finds them there
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rowptr + min((tile0 + it + 32) * FL_NODES, n32)));
        load_srcs(rp[(u + 2) & 3], sc[(u + 2) & 3]);
        FL_TMARK(2)
        const int rp0 = rp[u];
        const int node0 = (tile0 + it * tstep) * FL_NODES;
        const int e_lo = __shfl_sync(FULL, rp0, 0), e_hi = __shfl_sync(FULL, rp0, 8);
        FL_TMARK(3)
        int seg_lo = e_lo;
        bool last;
        do {
          const int seg_hi = min(seg_lo + FL_CAP, e_hi);
          last = seg_hi == e_hi;
          const int nseg = seg_hi - seg_lo;
          const int sl = (u & 1) * NBG + bufs[u & 1];
          const uint32_t base = st_u32 + (uint32_t)(sl * STG);
          const uint32_t fb = sfull + 8 * sl;
          FL_TWAIT(0, fl_mbar_wait(sempty + 8 * sl, pars[u & 1]))
          FL_TMARK(4)
          if (pw == 0 && lane < 16) {   // header: rp[0..8] | .. | seg_lo, seg_hi, last
            int hw = rp0;
            if (lane == 12) hw = seg_lo;
            if (lane == 13) hw = seg_hi;
            if (lane == 14) hw = last ? 1 : 0;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)(HDR + 4 * lane)), "r"(hw) : "memory");
          }
          __syncwarp();
          // g slot group of part pw: one bulk copy (the explicit arrive also releases the header stores)
          if (lane == 0) {
            if (pw < PPL && nseg > 0) {
              const uint32_t bytes = (uint32_t)nseg * 32u;
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(bytes) : "memory");
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + (uint32_t)(pw * GPL)),
                           "l"(gplane + (size_t)seg_lo * 16), "r"(bytes), "r"(fb)
                           : "memory");
            } else {
              fl_mbar_arrive(fb);
            }
          }
          __syncwarp();
          FL_TMARK(5)
          // gathered h rows (later segments of a big tile: source ids not prefetched)
          int sidx[GOPS];
#pragma unroll
          for (int i = 0; i < GOPS; ++i) sidx[i] = sc[u][i];
          if (seg_lo != e_lo) {
#pragma unroll
            for (int i = 0; i < GOPS; ++i) sidx[i] = __ldg(src_sorted + min(seg_lo + r0 + 16 * i, seg_hi - 1));
          }
          const uint32_t gb = base + g0;
#ifdef FL_TRACE
          { int acc__ = 0;
#pragma unroll
            for (int i = 0; i < GOPS; ++i) acc__ += sidx[i];
            if (acc__ == 0x7fffffff) tw[9] += 1; }      // forces the source ids to have arrived before the next mark
#endif
          FL_TMARK(6)
#pragma unroll
          for (int i = 0; i < GOPS; ++i)
            if (r0 + 16 * i < nseg && !FL_WHATIF(0x100))
              asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(gb + (uint32_t)(1536 * i)), "l"(h16 + ((uint32_t)sidx[i] * 6u + c0)) : "memory");
          // the nodes' own rows ride with the tile's last segment
          if (last && own_lane && node0 + orow < n32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + odst), "l"(h16 + (uint32_t)((node0 + orow) * 6 + ot % 6)) : "memory");
          FL_TMARK(7)
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(fb) : "memory");
          FL_TMARK(8)
          if (++bufs[u & 1] == NBG) {
            bufs[u & 1] = 0;
            pars[u & 1] ^= 1;
          }
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
extern "C" int fesr_dev_fl_trace(long long* host_out /* [148*24*12] */) {
  return cudaMemcpyFromSymbol(host_out, fl_trace_buf, sizeof(fl_trace_buf)) == cudaSuccess ? 0 : -1;
}
#endif

template <int PPL, int NBUF>
static size_t fl_smem_bytes() {
  constexpr int N = (PPL * FL_NODES + 15) / 16 * 16;
  constexpr size_t z = (size_t)2 * FL_NKB * N * 128;
  constexpr size_t st = (size_t)NBUF * (PPL * FL_CAP * 32 + FL_CAP * 96 + FL_NODES * 96 + 64);
  constexpr size_t misc = (size_t)PPL * FL_NODES * FL_WP * 4 + 8 * FL_NODES * 4 + 32 + (PPL == 3 ? (12 * 32 + 32) * 4 : 0) + (6 + 2 * NBUF) * 8 + 16;
  return 1024 + z + st + misc;
}

template <int PPL, int NBUF>
static int launch_fl(const int32_t* rowptr, const int32_t* src_sorted, const __half* g3, int64_t E, const __half* h_in,
                     int64_t n, int part0, int has_root, const __half* tf, const float* bias_p, const float* p_in,
                     float* p_out, __half* h_out, int rs, int fix_b, int relu, cudaStream_t s) {
  int* ovf = cur_ovf();
  size_t smem = fl_smem_bytes<PPL, NBUF>();
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: the kernel owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    FESR_CUDA(cudaFuncSetAttribute(layer_fused_f16_kernel<PPL, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t n_tiles = ceil_div(n, FL_NODES);
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  static const bool pdl = !(getenv("FESR_PDL") && atoi(getenv("FESR_PDL")) == 0);     // A/B switch for profiling
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((2 * FL_NODES + 3 + 1 + 4) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  FESR_CUDA(cudaLaunchKernelEx(&cfg, layer_fused_f16_kernel<PPL, NBUF>, rowptr, src_sorted, g3, E, h_in, n, part0, has_root, tf, bias_p,
                               p_in, p_out, h_out, rs, fix_b, relu, ovf));
  count_launch();
  return FESR_OK;
}

// Shapes the fused layer covers: KernelNN-like rows (one pass of 48 g slots, padded width 48) in one launch (w <= 43)
// or two / three; TEECNet rows (129 channels -> 144 slots = 9 parts, w + 1 <= 44 columns incl. the constant-1 column)
// in three launches of three parts that accumulate through the fp32 scratch P.
bool layer_fused_supported(const fesr_model_dims& d) {
  if (d.kind == FESR_KERNELNN) return d.passes == 1 && (d.kp == 48 || d.kp == 64) && d.wp == FL_WP;
  return d.kind == FESR_TEECNET && d.kp == 144 && d.wp == FL_WP && d.w <= 43;
}
int layer_fused_parts(const fesr_model_dims& d) { return d.kp / 16; }
size_t layer_fused_tf_elems(const fesr_model_dims& d) { return (size_t)layer_fused_parts(d) * FL_WP * FL_KP; }

int launch_prepare_tfused(const fesr_model_dims& d, const float* tprime, const float* mfull, void* tf, cudaStream_t s) {
  const int64_t total = (int64_t)layer_fused_tf_elems(d);
  ProfScope prof(PROF_PREPARE, s);
  prepare_tfused_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(d, layer_fused_parts(d), tprime, mfull,
                                                                      static_cast<__half*>(tf));
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// One layer: h_out = relu(Z(h_in) T' + bias); g3 planar fp16 [parts][E][16]; P: fp32 [n, 48] scratch
// (only touched when the parts need more than one launch).  mode: parts per launch (0 = best).
int launch_layer_fused_f16(const fesr_model_dims& d, const int32_t* rowptr, const int32_t* src_sorted, const void* g3,
                           int64_t E, const void* h_in, int64_t n, const void* tf, const float* bias_p, float* P,
                           void* h_out, int mode, cudaStream_t s, int out_f32, int sum_mode, const void* h_own) {
  if (n == 0) return FESR_OK;
  if (!layer_fused_supported(d)) {
    set_error("fused layer: unsupported model shape");
    return FESR_EINVAL;
  }
  const __half* gh = static_cast<const __half*>(g3);
  const __half* hh = static_cast<const __half*>(h_in);
  const __half* tfh = static_cast<const __half*>(tf);
  __half* ho = static_cast<__half*>(h_out);
  // bit 0: ReLU (TEECNet has no activation between the layers, models/model.py:280-282); bit 1: fp32 output rows
  static const int what_if = getenv("FESR_FL_EXP") ? (atoi(getenv("FESR_FL_EXP")) & 0xf) << 8 : 0;      // tools/dev only
  const int relu = (d.kind == FESR_KERNELNN && !sum_mode ? 1 : 0) | (out_f32 ? 2 : 0) | (sum_mode ? 4 : 0) | what_if;
  ProfScope prof(PROF_LAYER_FUSED, s);
  int rc;
  // 16-node tiles (layer_fused16.cu) unless FESR_FL_TILE=8 asks for the 8-node kernel of this file (A/B measurements)
  static const bool tile16 = !(getenv("FESR_FL_TILE") && atoi(getenv("FESR_FL_TILE")) == 8);
  const bool sum_ok = tile16 && d.w <= 43 && ((d.kind == FESR_KERNELNN && d.kp == 48 && (mode == 0 || mode == 3)) || d.kind == FESR_TEECNET);
  if (sum_mode && !sum_ok) {
    set_error("fused layer: the sum mode covers the w <= 43 shapes of the 16-node kernel only");
    return FESR_EINVAL;
  }
  if (tile16) {
    if (d.kind == FESR_TEECNET) {
      const int fix_b = d.w == 43 ? 42 : -1;
      if ((rc = launch_fl16<3>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 43, fix_b, relu, s))) return rc;
      if ((rc = launch_fl16<3>(rowptr, src_sorted, gh, E, hh, n, 3, 0, tfh, bias_p, P, P, nullptr, 43, fix_b, relu, s))) return rc;
      return launch_fl16<3>(rowptr, src_sorted, gh, E, hh, n, 6, 1, tfh, bias_p, P, nullptr, ho, 43, fix_b, relu, s,
                            static_cast<const __half*>(h_own));
    }
    if (d.kp == 64) {
      if ((rc = launch_fl16<2>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 64, -1, relu, s))) return rc;
      return launch_fl16<2>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 64, -1, relu, s);
    }
    if (mode == 0) mode = d.w <= 43 ? 3 : 2;
    if (mode == 3 && d.w <= 43)
      return launch_fl16<3>(rowptr, src_sorted, gh, E, hh, n, 0, 1, tfh, bias_p, nullptr, nullptr, ho, 43, d.w == 43 ? 42 : -1, relu, s,
                            static_cast<const __half*>(h_own));
    if (mode == 2) {
      if ((rc = launch_fl16<2>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 64, -1, relu, s))) return rc;
      return launch_fl16<1>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 128, -1, relu, s);
    }
    if ((rc = launch_fl16<1>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 128, -1, relu, s))) return rc;
    if ((rc = launch_fl16<1>(rowptr, src_sorted, gh, E, hh, n, 1, 0, tfh, bias_p, P, P, nullptr, 128, -1, relu, s))) return rc;
    return launch_fl16<1>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 128, -1, relu, s);
  }
  if (d.kind == FESR_TEECNET) {
    // 9 parts, three per launch, partial sums through P; the constant-1 column h[:, w] comes out of the epilogue as
    // 0 (no T'' row feeds it) + bias_p[w], which prepare_small_kernel sets to 1 for TEECNet
    const int fix_b = d.w == 43 ? 42 : -1;
    if ((rc = launch_fl<3, 4>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 43, fix_b, relu, s))) return rc;
    if ((rc = launch_fl<3, 4>(rowptr, src_sorted, gh, E, hh, n, 3, 0, tfh, bias_p, P, P, nullptr, 43, fix_b, relu, s))) return rc;
    return launch_fl<3, 4>(rowptr, src_sorted, gh, E, hh, n, 6, 1, tfh, bias_p, P, nullptr, ho, 43, fix_b, relu, s);
  }
  if (d.kp == 64) {
    // w = 48 (the reference's default config width): 49 edge channels -> 64 slots = 4 parts, two launches of two
    // parts (part p in TMEM lanes [64p, 64p + 48)), partial sums through P
    if ((rc = launch_fl<2, 4>(rowptr, src_sorted, gh, E, hh, n, 0, 0, tfh, bias_p, nullptr, P, nullptr, 64, -1, relu, s))) return rc;
    return launch_fl<2, 4>(rowptr, src_sorted, gh, E, hh, n, 2, 1, tfh, bias_p, P, nullptr, ho, 64, -1, relu, s);
  }
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
