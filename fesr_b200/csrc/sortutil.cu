#include <cub/cub.cuh>

#include "sortutil.cuh"

namespace fesr {

size_t sort_temp_bytes(int64_t m) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, m > 0 ? m : 1);
  return bytes;
}

size_t sort_pairs_u64_bytes(int64_t m) {
  if (m <= 0) m = 1;
  size_t b = 0;
  b += align_up((size_t)m * sizeof(uint64_t), 256) * 2;
  b += align_up((size_t)m * sizeof(int32_t), 256) * 2;
  b += align_up(sort_temp_bytes(m), 256);
  return b + 256;
}

SortBuffers carve_sort_buffers(void* workspace, int64_t m) {
  if (m <= 0) m = 1;
  Carver c(workspace);
  SortBuffers sb;
  sb.keys_in = c.take<uint64_t>(m);
  sb.keys_out = c.take<uint64_t>(m);
  sb.vals_in = c.take<int32_t>(m);
  sb.vals_out = c.take<int32_t>(m);
  sb.temp_bytes = sort_temp_bytes(m);
  sb.temp = c.take<char>(sb.temp_bytes);
  return sb;
}

int sort_pairs_u64(const SortBuffers& sb, int64_t m, int begin_bit, int end_bit, cudaStream_t stream) {
  size_t bytes = sb.temp_bytes;
  FESR_CUDA(cub::DeviceRadixSort::SortPairs(sb.temp, bytes, sb.keys_in, sb.keys_out, sb.vals_in, sb.vals_out, m,
                                            begin_bit, end_bit, stream));
  return FESR_OK;
}

size_t sort_keys_temp_bytes(int64_t m) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, m > 0 ? m : 1);
  return bytes;
}

int sort_keys_u64(const uint64_t* in, uint64_t* out, int64_t m, int begin_bit, int end_bit, void* temp,
                  size_t temp_bytes, cudaStream_t stream) {
  FESR_CUDA(cub::DeviceRadixSort::SortKeys(temp, temp_bytes, in, out, m, begin_bit, end_bit, stream));
  return FESR_OK;
}

__global__ void head_flags_kernel(const uint64_t* __restrict__ keys, int64_t m, int32_t* __restrict__ flags) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

int launch_head_flags(const uint64_t* keys, int64_t m, int32_t* flags, cudaStream_t stream) {
  if (m == 0) return FESR_OK;
  head_flags_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, stream>>>(keys, m, flags);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

__global__ void ptr_from_sorted_shift_kernel(const uint64_t* __restrict__ keys, int64_t m, int shift, int64_t nseg,
                                             int32_t* __restrict__ ptr) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > m) return;
  const int64_t d = (e < m) ? (int64_t)(keys[e] >> shift) : nseg;
  const int64_t dprev = (e > 0) ? (int64_t)(keys[e - 1] >> shift) : -1;
  for (int64_t i = dprev + 1; i <= d && i <= nseg; ++i) ptr[i] = (int32_t)e;
}

int launch_ptr_from_sorted(const uint64_t* keys, int64_t m, int shift, int64_t nseg, int32_t* ptr, cudaStream_t stream) {
  ptr_from_sorted_shift_kernel<<<(unsigned)ceil_div(m + 1, 256), 256, 0, stream>>>(keys, m, shift, nseg, ptr);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

__global__ void scan_total_kernel(const int32_t* __restrict__ flags, int32_t* __restrict__ scan, int64_t m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) scan[m] = (m > 0) ? scan[m - 1] + flags[m - 1] : 0;
}

int launch_scan_total(const int32_t* flags, int32_t* scan, int64_t m, cudaStream_t stream) {
  scan_total_kernel<<<1, 32, 0, stream>>>(flags, scan, m);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

size_t scan_temp_bytes(int64_t m) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, m > 0 ? m : 1);
  return bytes;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t m, void* temp, size_t temp_bytes, cudaStream_t stream) {
  FESR_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, m, stream));
  return FESR_OK;
}

}  // namespace fesr
