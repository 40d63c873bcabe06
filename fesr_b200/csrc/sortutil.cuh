// Stable 64-bit key / 32-bit value radix sort on caller-provided workspace (CUB underneath).
#pragma once
#include "common.cuh"

namespace fesr {

struct SortBuffers {
  uint64_t* keys_in;
  uint64_t* keys_out;
  int32_t* vals_in;
  int32_t* vals_out;
  void* temp;
  size_t temp_bytes;
};

size_t sort_temp_bytes(int64_t m);           // CUB temp storage for m pairs
size_t sort_pairs_u64_bytes(int64_t m);      // everything carve_sort_buffers needs
SortBuffers carve_sort_buffers(void* workspace, int64_t m);
// sorts keys_in/vals_in -> keys_out/vals_out on bits [begin_bit, end_bit); stable
int sort_pairs_u64(const SortBuffers& sb, int64_t m, int begin_bit, int end_bit, cudaStream_t stream);

// keys-only variant (u64), stable
size_t sort_keys_temp_bytes(int64_t m);
int sort_keys_u64(const uint64_t* in, uint64_t* out, int64_t m, int begin_bit, int end_bit, void* temp,
                  size_t temp_bytes, cudaStream_t stream);

// exclusive prefix sum of int32 (in != out allowed to alias); temp from workspace
size_t scan_temp_bytes(int64_t m);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t m, void* temp, size_t temp_bytes, cudaStream_t stream);

// scan[m] = scan[m-1] + flags[m-1]  (closes an exclusive scan of m flags)
int launch_scan_total(const int32_t* flags, int32_t* scan, int64_t m, cudaStream_t stream);
// out[j] = 1 if keys[j] != keys[j-1] (j = 0 -> 1)
int launch_head_flags(const uint64_t* keys, int64_t m, int32_t* flags, cudaStream_t stream);
// ptr[i] = first position j with (keys[j] >> shift) >= i, for i in [0, nseg]; ptr[nseg] = m
int launch_ptr_from_sorted(const uint64_t* keys, int64_t m, int shift, int64_t nseg, int32_t* ptr, cudaStream_t stream);

}  // namespace fesr
