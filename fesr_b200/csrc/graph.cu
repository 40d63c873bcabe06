// CSR-by-destination build and the global-node occurrence CSR used by the stitch.
// Sorting is CUB's stable LSD radix sort (toolkit header library); everything that
// touches the sorted data is hand-written.
#include <cub/cub.cuh>

#include "common.cuh"
#include "sortutil.cuh"

namespace fesr {

// key = dst << 32 | src, val = original edge id
__global__ void edge_keys_kernel(const int64_t* __restrict__ edge_index, int64_t E,
                                 uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  uint64_t s = (uint64_t)edge_index[e];
  uint64_t d = (uint64_t)edge_index[E + e];
  keys[e] = (d << 32) | (s & 0xffffffffull);
  vals[e] = (int32_t)e;
}

__global__ void split_keys_kernel(const uint64_t* __restrict__ keys, int64_t E, int64_t n,
                                  int32_t* __restrict__ src_sorted, int32_t* __restrict__ rowptr) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > E) return;
  // thread e fills rowptr(prev_dst, dst_e] = e ; thread E closes the tail
  int64_t d = (e < E) ? (int64_t)(keys[e] >> 32) : n;
  int64_t dprev = (e > 0) ? (int64_t)(keys[e - 1] >> 32) : -1;
  if (e < E) src_sorted[e] = (int32_t)(keys[e] & 0xffffffffull);
  for (int64_t i = dprev + 1; i <= d && i <= n; ++i) rowptr[i] = (int32_t)e;
}

__global__ void gid_keys_kernel(const int64_t* __restrict__ gids, int64_t n, uint64_t* __restrict__ keys,
                                int32_t* __restrict__ vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint64_t)gids[i];
  vals[i] = (int32_t)i;
}

__global__ void ptr_from_sorted_kernel(const uint64_t* __restrict__ keys, int64_t m, int64_t N,
                                       int32_t* __restrict__ ptr) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > m) return;
  int64_t d = (e < m) ? (int64_t)keys[e] : N;
  int64_t dprev = (e > 0) ? (int64_t)keys[e - 1] : -1;
  for (int64_t i = dprev + 1; i <= d && i <= N; ++i) ptr[i] = (int32_t)e;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

size_t fesr_csr_workspace_bytes(int64_t n, int64_t E) {
  (void)n;
  if (E < 0) return 0;
  return sort_pairs_u64_bytes(E);
}

int fesr_csr_build(const int64_t* edge_index, int64_t E, int64_t n, int32_t* rowptr, int32_t* src_sorted,
                   int32_t* perm, void* workspace, size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(E >= 0 && n >= 0 && n < (1ll << 31) && E < (1ll << 31), "n/E out of int32 range");
  FESR_CHECK_ARG(rowptr && (E == 0 || (edge_index && src_sorted && perm)), "NULL pointer");
  cudaStream_t stream = as_stream(stream_);
  if (E == 0) {
    FESR_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n + 1) * sizeof(int32_t), stream));
    return FESR_OK;
  }
  FESR_CHECK_ARG(workspace && workspace_bytes >= fesr_csr_workspace_bytes(n, E), "workspace too small");
  SortBuffers sb = carve_sort_buffers(workspace, E);
  const int T = 256;
  edge_keys_kernel<<<(unsigned)ceil_div(E, T), T, 0, stream>>>(edge_index, E, sb.keys_in, sb.vals_in);
  FESR_LAUNCH_CHECK();
  int bits = 32;
  while (bits < 63 && (1ll << (bits - 32)) < n) ++bits;   // dst occupies bits [32, bits)
  int rc = sort_pairs_u64(sb, E, 0, bits, stream);
  if (rc) return rc;
  FESR_CUDA(cudaMemcpyAsync(perm, sb.vals_out, (size_t)E * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  split_keys_kernel<<<(unsigned)ceil_div(E + 1, T), T, 0, stream>>>(sb.keys_out, E, n, src_sorted, rowptr);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

size_t fesr_occurrence_workspace_bytes(int64_t n_tot, int64_t N) {
  (void)N;
  if (n_tot < 0) return 0;
  return sort_pairs_u64_bytes(n_tot);
}

int fesr_occurrence_build(const int64_t* global_ids, int64_t n_tot, int64_t N, int32_t* occ_ptr,
                          int32_t* occ_idx, void* workspace, size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(n_tot >= 0 && N >= 0 && N < (1ll << 31) && n_tot < (1ll << 31), "sizes out of int32 range");
  FESR_CHECK_ARG(occ_ptr && (n_tot == 0 || (global_ids && occ_idx)), "NULL pointer");
  cudaStream_t stream = as_stream(stream_);
  if (n_tot == 0) {
    FESR_CUDA(cudaMemsetAsync(occ_ptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
    return FESR_OK;
  }
  FESR_CHECK_ARG(workspace && workspace_bytes >= fesr_occurrence_workspace_bytes(n_tot, N), "workspace too small");
  SortBuffers sb = carve_sort_buffers(workspace, n_tot);
  const int T = 256;
  gid_keys_kernel<<<(unsigned)ceil_div(n_tot, T), T, 0, stream>>>(global_ids, n_tot, sb.keys_in, sb.vals_in);
  FESR_LAUNCH_CHECK();
  int bits = 1;
  while (bits < 63 && (1ll << bits) < N) ++bits;
  int rc = sort_pairs_u64(sb, n_tot, 0, bits, stream);
  if (rc) return rc;
  FESR_CUDA(cudaMemcpyAsync(occ_idx, sb.vals_out, (size_t)n_tot * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  ptr_from_sorted_kernel<<<(unsigned)ceil_div(n_tot + 1, T), T, 0, stream>>>(sb.keys_out, n_tot, N, occ_ptr);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // extern "C"
