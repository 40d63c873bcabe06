// Per-subdomain node compaction, edge de-duplication, edge lengths and the destination CSR,
// for all subdomains at once (bit-exact with oracle/graph.py build_subdomains).
//
// Replaces the per-partition Python loops of the reference (dataset/GraphDataset.py:1245-1284
// calling vtk_to_pyg :838-869 once per partition): every (subdomain, vertex) and every
// (dst, src) candidate becomes a 64-bit key; two radix sorts + flag/scan/compact give the
// sorted unique node list (ascending global id per subdomain) and the edge list already in
// (subdomain, dst, src) order, i.e. the CSR the message-passing kernels consume.
#include "common.cuh"
#include "sortutil.cuh"

namespace fesr {

struct SubWs {
  int32_t* pair_leaf;   // [P]
  uint64_t* vkeys_a;    // [4P]
  uint64_t* vkeys_b;    // [4P] sorted
  int32_t* vflags;      // [4P]
  int32_t* vscan;       // [4P+1]
  uint64_t* ukeys;      // [<=4P] unique (leaf<<32 | gid), sorted
  int32_t* vptr;        // [S+1]
  uint64_t* ekeys_a;    // [12P]
  uint64_t* ekeys_b;    // [12P] sorted
  int32_t* eflags;      // [12P]
  int32_t* escan;       // [12P+1]
  int32_t* totals;      // [2] device copy of n_tot, e_tot
  void* sort_temp;
  size_t sort_bytes;
  void* scan_temp;
  size_t scan_bytes;
  size_t bytes;
};

static SubWs carve_sub(void* base, int64_t P, int32_t S) {
  Carver c(base);
  SubWs w;
  const size_t p = (size_t)(P > 0 ? P : 1);
  w.pair_leaf = c.take<int32_t>(p);
  w.vkeys_a = c.take<uint64_t>(4 * p);
  w.vkeys_b = c.take<uint64_t>(4 * p);
  w.vflags = c.take<int32_t>(4 * p);
  w.vscan = c.take<int32_t>(4 * p + 1);
  w.ukeys = c.take<uint64_t>(4 * p);
  w.vptr = c.take<int32_t>((size_t)S + 1);
  w.ekeys_a = c.take<uint64_t>(12 * p);
  w.ekeys_b = c.take<uint64_t>(12 * p);
  w.eflags = c.take<int32_t>(12 * p);
  w.escan = c.take<int32_t>(12 * p + 1);
  w.totals = c.take<int32_t>(4);
  w.sort_bytes = sort_keys_temp_bytes(12 * (int64_t)p);
  w.sort_temp = c.take<char>(w.sort_bytes);
  w.scan_bytes = scan_temp_bytes(12 * (int64_t)p + 1);
  w.scan_temp = c.take<char>(w.scan_bytes);
  w.bytes = c.used();
  return w;
}

__global__ void pair_leaf_kernel(const int32_t* __restrict__ leaf_ptr, int S, int64_t P, int32_t* __restrict__ pair_leaf) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  int lo = 0, hi = S;                 // last s with leaf_ptr[s] <= p
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (leaf_ptr[mid] <= p) lo = mid; else hi = mid;
  }
  pair_leaf[p] = lo;
}

__global__ void vertex_keys_kernel(const int32_t* __restrict__ cells, const int32_t* __restrict__ leaf_cells,
                                   const int32_t* __restrict__ pair_leaf, int64_t P, uint64_t* __restrict__ keys) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int4 v = reinterpret_cast<const int4*>(cells)[leaf_cells[p]];
  const uint64_t hi = (uint64_t)pair_leaf[p] << 32;
  keys[4 * p + 0] = hi | (uint32_t)v.x;
  keys[4 * p + 1] = hi | (uint32_t)v.y;
  keys[4 * p + 2] = hi | (uint32_t)v.z;
  keys[4 * p + 3] = hi | (uint32_t)v.w;
}

__global__ void compact_keys_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ flags,
                                    const int32_t* __restrict__ scan, int64_t m, uint64_t* __restrict__ out) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  if (flags[j]) out[scan[j]] = keys[j];
}

// node_ptr[s] = number of unique (leaf, vertex) keys with leaf < s
__global__ void node_ptr_kernel(const int32_t* __restrict__ vptr, const int32_t* __restrict__ vscan, int64_t m, int S,
                                int32_t* __restrict__ node_ptr, int32_t* __restrict__ totals) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > S) return;
  node_ptr[s] = vscan[vptr[s]];          // vscan has m+1 entries; vptr[s] == m -> total
  if (s == S) totals[0] = vscan[m];
}

__device__ __forceinline__ int find_node(const uint64_t* __restrict__ ukeys, int lo, int hi, uint64_t key) {
  while (lo < hi) {                       // lower_bound; key is guaranteed present
    const int mid = (lo + hi) >> 1;
    if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// 12 directed candidates per (subdomain, cell) pair; a repeated vertex yields a self loop that
// the flag pass drops (the oracle drops them as well)
__global__ void edge_keys_kernel(const int32_t* __restrict__ cells, const int32_t* __restrict__ leaf_cells,
                                 const int32_t* __restrict__ pair_leaf, const int32_t* __restrict__ node_ptr,
                                 const uint64_t* __restrict__ ukeys, int64_t P, uint64_t* __restrict__ ekeys) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int s = pair_leaf[p];
  const int4 v = reinterpret_cast<const int4*>(cells)[leaf_cells[p]];
  const int vid[4] = {v.x, v.y, v.z, v.w};
  const int lo = node_ptr[s], hi = node_ptr[s + 1];
  uint32_t idx[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) idx[a] = (uint32_t)find_node(ukeys, lo, hi, ((uint64_t)s << 32) | (uint32_t)vid[a]);
  int o = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (a != b) ekeys[12 * p + (o++)] = ((uint64_t)idx[b] << 32) | idx[a];   // dst = b, src = a
}

__global__ void edge_flags_kernel(const uint64_t* __restrict__ keys, int64_t m, int32_t* __restrict__ flags) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint64_t k = keys[j];
  const bool self = (uint32_t)(k >> 32) == (uint32_t)(k & 0xffffffffull);
  flags[j] = (!self && (j == 0 || k != keys[j - 1])) ? 1 : 0;
}

// edge_ptr[s] = number of kept edges whose dst < node_ptr[s]
__global__ void edge_ptr_kernel(const uint64_t* __restrict__ ekeys, const int32_t* __restrict__ escan, int64_t m,
                                const int32_t* __restrict__ node_ptr, int S, int32_t* __restrict__ edge_ptr,
                                int32_t* __restrict__ totals) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > S) return;
  const uint64_t bound = (uint64_t)(uint32_t)node_ptr[s] << 32;
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ekeys[mid] < bound) lo = mid + 1; else hi = mid;
  }
  edge_ptr[s] = escan[lo];
  if (s == S) totals[1] = escan[m];
}

__global__ void gid_out_kernel(const uint64_t* __restrict__ ukeys, int64_t n_tot, int64_t* __restrict__ gids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_tot) gids[i] = (int64_t)(ukeys[i] & 0xffffffffull);
}

// compact the edges and compute fp32 lengths exactly as numpy's norm does on float32:
// d = p[src]-p[dst]; sqrt((dx*dx + dy*dy) + dz*dz), each op rounded to nearest
__global__ void edge_out_kernel(const uint64_t* __restrict__ ekeys, const int32_t* __restrict__ eflags,
                                const int32_t* __restrict__ escan, int64_t m, const uint64_t* __restrict__ ukeys,
                                const float* __restrict__ pos, int32_t* __restrict__ edge_src,
                                int32_t* __restrict__ edge_dst, float* __restrict__ edge_attr) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m || !eflags[j]) return;
  const uint64_t k = ekeys[j];
  const int32_t d = (int32_t)(k >> 32), s = (int32_t)(k & 0xffffffffull);
  const int e = escan[j];
  edge_src[e] = s;
  edge_dst[e] = d;
  const int64_t gs = (int64_t)(ukeys[s] & 0xffffffffull), gd = (int64_t)(ukeys[d] & 0xffffffffull);
  const float dx = __fsub_rn(pos[gs * 3 + 0], pos[gd * 3 + 0]);
  const float dy = __fsub_rn(pos[gs * 3 + 1], pos[gd * 3 + 1]);
  const float dz = __fsub_rn(pos[gs * 3 + 2], pos[gd * 3 + 2]);
  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  edge_attr[e] = __fsqrt_rn(ss);
}

__global__ void rowptr_from_dst_kernel(const int32_t* __restrict__ dst, int64_t E, int64_t n, int32_t* __restrict__ rowptr) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > E) return;
  const int64_t d = (e < E) ? dst[e] : n;
  const int64_t dprev = (e > 0) ? dst[e - 1] : -1;
  for (int64_t i = dprev + 1; i <= d && i <= n; ++i) rowptr[i] = (int32_t)e;
}

static int bits_for(int64_t v) {
  int b = 1;
  while (b < 63 && (1ll << b) < v) ++b;
  return b;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

size_t fesr_subdomain_workspace_bytes(int64_t total_pairs, int64_t N, int32_t n_sub) {
  (void)N;
  if (total_pairs < 0 || n_sub < 0) return 0;
  return carve_sub(nullptr, total_pairs, n_sub).bytes;
}

int fesr_subdomain_count(const int32_t* cells, const int32_t* leaf_ptr, const int32_t* leaf_cells, int32_t n_sub,
                         int64_t total_pairs, int64_t N, int32_t* node_ptr, int32_t* edge_ptr, int64_t* host_totals,
                         void* workspace, size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(n_sub >= 1 && total_pairs >= 0 && 12 * total_pairs < (1ll << 31), "bad sizes (12*pairs must fit int32)");
  FESR_CHECK_ARG(N >= 0 && N < (1ll << 31), "N out of range");
  FESR_CHECK_ARG(node_ptr && edge_ptr && host_totals && leaf_ptr, "NULL pointer");
  cudaStream_t s = as_stream(stream_);
  const int S = n_sub;
  const int64_t P = total_pairs;
  if (P == 0) {
    FESR_CUDA(cudaMemsetAsync(node_ptr, 0, (size_t)(S + 1) * sizeof(int32_t), s));
    FESR_CUDA(cudaMemsetAsync(edge_ptr, 0, (size_t)(S + 1) * sizeof(int32_t), s));
    host_totals[0] = host_totals[1] = 0;
    return FESR_OK;
  }
  FESR_CHECK_ARG(cells && leaf_cells, "NULL pointer");
  SubWs w = carve_sub(workspace, P, S);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("subdomain workspace too small: need %zu, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  const int T = 256;
  int rc;
  pair_leaf_kernel<<<(unsigned)ceil_div(P, T), T, 0, s>>>(leaf_ptr, S, P, w.pair_leaf);
  FESR_LAUNCH_CHECK();
  vertex_keys_kernel<<<(unsigned)ceil_div(P, T), T, 0, s>>>(cells, leaf_cells, w.pair_leaf, P, w.vkeys_a);
  FESR_LAUNCH_CHECK();
  const int64_t mv = 4 * P;
  if ((rc = sort_keys_u64(w.vkeys_a, w.vkeys_b, mv, 0, 32 + bits_for(S), w.sort_temp, w.sort_bytes, s))) return rc;
  if ((rc = launch_head_flags(w.vkeys_b, mv, w.vflags, s))) return rc;
  // vscan has mv+1 slots: exclusive scan of the mv flags, then vscan[mv] = total
  if ((rc = exclusive_scan_i32(w.vflags, w.vscan, mv, w.scan_temp, w.scan_bytes, s))) return rc;
  if ((rc = launch_scan_total(w.vflags, w.vscan, mv, s))) return rc;
  compact_keys_kernel<<<(unsigned)ceil_div(mv, T), T, 0, s>>>(w.vkeys_b, w.vflags, w.vscan, mv, w.ukeys);
  FESR_LAUNCH_CHECK();
  if ((rc = launch_ptr_from_sorted(w.vkeys_b, mv, 32, S, w.vptr, s))) return rc;
  node_ptr_kernel<<<(unsigned)ceil_div(S + 1, T), T, 0, s>>>(w.vptr, w.vscan, mv, S, node_ptr, w.totals);
  FESR_LAUNCH_CHECK();
  // edges
  edge_keys_kernel<<<(unsigned)ceil_div(P, T), T, 0, s>>>(cells, leaf_cells, w.pair_leaf, node_ptr, w.ukeys, P, w.ekeys_a);
  FESR_LAUNCH_CHECK();
  const int64_t me = 12 * P;
  if ((rc = sort_keys_u64(w.ekeys_a, w.ekeys_b, me, 0, 32 + bits_for(mv), w.sort_temp, w.sort_bytes, s))) return rc;
  edge_flags_kernel<<<(unsigned)ceil_div(me, T), T, 0, s>>>(w.ekeys_b, me, w.eflags);
  FESR_LAUNCH_CHECK();
  if ((rc = exclusive_scan_i32(w.eflags, w.escan, me, w.scan_temp, w.scan_bytes, s))) return rc;
  if ((rc = launch_scan_total(w.eflags, w.escan, me, s))) return rc;
  edge_ptr_kernel<<<(unsigned)ceil_div(S + 1, T), T, 0, s>>>(w.ekeys_b, w.escan, me, node_ptr, S, edge_ptr, w.totals);
  FESR_LAUNCH_CHECK();
  int32_t tot[2] = {0, 0};
  FESR_CUDA(cudaMemcpyAsync(tot, w.totals, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FESR_CUDA(cudaStreamSynchronize(s));
  host_totals[0] = tot[0];
  host_totals[1] = tot[1];
  return FESR_OK;
}

int fesr_subdomain_fill(const float* pos, const int32_t* node_ptr, const int32_t* edge_ptr, int32_t n_sub,
                        int64_t total_pairs, int64_t n_tot, int64_t e_tot, int64_t* global_ids, int32_t* edge_src,
                        int32_t* edge_dst, float* edge_attr, int32_t* rowptr, void* workspace,
                        size_t workspace_bytes, void* stream_) {
  (void)node_ptr;
  (void)edge_ptr;
  FESR_CHECK_ARG(n_sub >= 1 && total_pairs >= 0 && n_tot >= 0 && e_tot >= 0, "bad sizes");
  FESR_CHECK_ARG(rowptr, "NULL rowptr");
  cudaStream_t s = as_stream(stream_);
  const int T = 256;
  if (total_pairs == 0 || n_tot == 0) {
    FESR_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_tot + 1) * sizeof(int32_t), s));
    return FESR_OK;
  }
  FESR_CHECK_ARG(pos && global_ids && (e_tot == 0 || (edge_src && edge_dst && edge_attr)), "NULL pointer");
  SubWs w = carve_sub(workspace, total_pairs, n_sub);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("subdomain workspace too small: need %zu, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  gid_out_kernel<<<(unsigned)ceil_div(n_tot, T), T, 0, s>>>(w.ukeys, n_tot, global_ids);
  FESR_LAUNCH_CHECK();
  const int64_t me = 12 * total_pairs;
  if (e_tot > 0) {
    edge_out_kernel<<<(unsigned)ceil_div(me, T), T, 0, s>>>(w.ekeys_b, w.eflags, w.escan, me, w.ukeys, pos, edge_src,
                                                           edge_dst, edge_attr);
    FESR_LAUNCH_CHECK();
  }
  rowptr_from_dst_kernel<<<(unsigned)ceil_div(e_tot + 1, T), T, 0, s>>>(edge_dst, e_tot, n_tot, rowptr);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // extern "C"
