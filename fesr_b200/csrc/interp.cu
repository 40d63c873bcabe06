// Low-resolution -> high-resolution field transfer: Gaussian-kernel point interpolation.
//
// Replaces AnsysDataset._lagrangian_interpolation (reference dataset/GraphDataset.py:1041-1105), which drives
// vtkPointInterpolator with a vtkGaussianKernel(radius = 3 * mesh_spacing, sharpness = 2): for every target point
//   out[t, :] = sum_j w_j v[j, :] / sum_j w_j     over the source points with |p_j - q_t|^2 <= radius^2,
//   w_j = exp(-(sharpness / radius)^2 |p_j - q_t|^2),
// and targets without a source point in range get `null_value` (VTK's NULL_VALUE strategy, 0.0).
//
// The source points are binned into a hash grid of cell size 1.001 x `radius`: a 63-bit key of the three 21-bit biased cell
// coordinates, one stable radix sort, positions / values gathered into the sorted order.  A target thread visits
// the 3 x 3 (ix, iy) columns around its cell; the three iz cells of a column are one contiguous key range, found by
// one binary search and scanned linearly.  Accumulation order = (column, sorted position): fixed, so the result is
// deterministic.  The in-range test is done in fp64 on the exactly representable fp32 inputs (a point on the rim
// carries exp(-sharpness^2) of the weight, so the decision must not depend on fp32 rounding); weights and sums fp32.
#include <math.h>

#include "common.cuh"
#include "sortutil.cuh"

namespace fesr {

constexpr int IP_BIAS = 1 << 20;

__device__ __forceinline__ int ip_cell(float x, float inv_r) {
  const int c = (int)floorf(x * inv_r);
  return min(max(c, -IP_BIAS + 2), IP_BIAS - 3) + IP_BIAS;            // 21 bits, room for the +-1 neighbours
}
__device__ __forceinline__ uint64_t ip_key(int ix, int iy, int iz) {
  return ((uint64_t)ix << 42) | ((uint64_t)iy << 21) | (uint64_t)iz;
}

__global__ void interp_keys_kernel(const float* __restrict__ pos, int64_t n, float inv_r, uint64_t* __restrict__ keys,
                                   int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ip_key(ip_cell(pos[3 * i], inv_r), ip_cell(pos[3 * i + 1], inv_r), ip_cell(pos[3 * i + 2], inv_r));
  idx[i] = (int32_t)i;
}

template <int C>
__global__ void interp_gather_kernel(const float* __restrict__ pos, const float* __restrict__ val,
                                     const int32_t* __restrict__ order, int64_t n, float4* __restrict__ spos,
                                     float* __restrict__ sval) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t j = order[i];
  spos[i] = make_float4(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], 0.f);
#pragma unroll
  for (int c = 0; c < C; ++c) sval[i * C + c] = val[j * C + c];
}

template <int C>
__global__ void interp_query_kernel(const uint64_t* __restrict__ keys, const float4* __restrict__ spos,
                                    const float* __restrict__ sval, int64_t n_src, const float* __restrict__ dst,
                                    int64_t n_dst, float inv_r, double r2, float f2, float null_value,
                                    float* __restrict__ out, int32_t* __restrict__ count) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_dst) return;
  const float qx = dst[3 * t], qy = dst[3 * t + 1], qz = dst[3 * t + 2];
  const int cx = ip_cell(qx, inv_r), cy = ip_cell(qy, inv_r), cz = ip_cell(qz, inv_r);
  float num[C];
#pragma unroll
  for (int c = 0; c < C; ++c) num[c] = 0.f;
  float den = 0.f;
  int cnt = 0;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const uint64_t klo = ip_key(cx + dx, cy + dy, cz - 1), khi = ip_key(cx + dx, cy + dy, cz + 1);
      int64_t lo = 0, hi = n_src;                                      // lower bound of klo
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < klo) lo = mid + 1;
        else hi = mid;
      }
      for (int64_t j = lo; j < n_src && keys[j] <= khi; ++j) {
        const float4 p = spos[j];
        const double ex = (double)p.x - (double)qx, ey = (double)p.y - (double)qy, ez = (double)p.z - (double)qz;
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        if (d2 <= r2) {
          const float w = expf(-f2 * (float)d2);
#pragma unroll
          for (int c = 0; c < C; ++c) num[c] = fmaf(w, sval[j * C + c], num[c]);
          den += w;
          ++cnt;
        }
      }
    }
  const bool ok = cnt > 0 && den != 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) out[t * C + c] = ok ? num[c] / den : null_value;
  if (count) count[t] = cnt;
}

struct InterpWs {
  SortBuffers sb;
  float4* spos;
  float* sval;
  size_t bytes;
};

static InterpWs carve_interp(void* base, int64_t n_src, int channels) {
  InterpWs w;
  const int64_t m = n_src > 0 ? n_src : 1;
  const size_t sort_bytes = (sort_pairs_u64_bytes(m) + 255) & ~(size_t)255;
  uint8_t* p = static_cast<uint8_t*>(base);
  if (base) w.sb = carve_sort_buffers(base, m);
  w.spos = reinterpret_cast<float4*>(p + sort_bytes);
  const size_t pos_bytes = ((size_t)m * sizeof(float4) + 255) & ~(size_t)255;
  w.sval = reinterpret_cast<float*>(p + sort_bytes + pos_bytes);
  w.bytes = sort_bytes + pos_bytes + (((size_t)m * channels * sizeof(float) + 255) & ~(size_t)255);
  return w;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

size_t fesr_interp_workspace_bytes(int64_t n_src, int32_t channels) {
  if (n_src < 0 || channels < 1) return 0;
  return carve_interp(nullptr, n_src, channels).bytes;
}

int fesr_interp_gaussian(const float* src_pos, const float* src_val, int32_t channels, int64_t n_src,
                         const float* dst_pos, int64_t n_dst, float radius, float sharpness, float null_value,
                         float* out, int32_t* count, void* workspace, size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4, got %d", channels);
  FESR_CHECK_ARG(n_src >= 0 && n_dst >= 0 && n_src < (1ll << 31) && n_dst < (1ll << 31), "sizes out of range");
  FESR_CHECK_ARG(radius > 0.f && isfinite(radius) && isfinite(sharpness), "radius must be positive and finite");
  if (n_dst == 0) return FESR_OK;
  FESR_CHECK_ARG(dst_pos && out && (n_src == 0 || (src_pos && src_val)), "NULL pointer");
  InterpWs w = carve_interp(workspace, n_src, channels);
  FESR_CHECK_ARG(workspace && workspace_bytes >= w.bytes, "interpolation workspace too small: need %zu bytes", w.bytes);
  cudaStream_t s = as_stream(stream_);
  const int T = 256;
  // cells 0.1 % larger than the radius: two points within `radius` of each other are then in adjacent cells even
  // after the fp32 rounding of x * inv_r (cell coordinates up to ~1e3)
  const float inv_r = 1.0f / (radius * 1.001f);
  if (n_src > 0) {
    interp_keys_kernel<<<(unsigned)ceil_div(n_src, T), T, 0, s>>>(src_pos, n_src, inv_r, w.sb.keys_in, w.sb.vals_in);
    FESR_LAUNCH_CHECK();
    int rc = sort_pairs_u64(w.sb, n_src, 0, 63, s);
    if (rc) return rc;
    const unsigned g = (unsigned)ceil_div(n_src, T);
    if (channels == 4) interp_gather_kernel<4><<<g, T, 0, s>>>(src_pos, src_val, w.sb.vals_out, n_src, w.spos, w.sval);
    else if (channels == 3) interp_gather_kernel<3><<<g, T, 0, s>>>(src_pos, src_val, w.sb.vals_out, n_src, w.spos, w.sval);
    else interp_gather_kernel<1><<<g, T, 0, s>>>(src_pos, src_val, w.sb.vals_out, n_src, w.spos, w.sval);
    FESR_LAUNCH_CHECK();
  }
  const double r2 = (double)radius * (double)radius;
  const float f2 = (sharpness / radius) * (sharpness / radius);
  const unsigned g = (unsigned)ceil_div(n_dst, T);
  if (channels == 4)
    interp_query_kernel<4><<<g, T, 0, s>>>(w.sb.keys_out, w.spos, w.sval, n_src, dst_pos, n_dst, inv_r, r2, f2, null_value, out, count);
  else if (channels == 3)
    interp_query_kernel<3><<<g, T, 0, s>>>(w.sb.keys_out, w.spos, w.sval, n_src, dst_pos, n_dst, inv_r, r2, f2, null_value, out, count);
  else
    interp_query_kernel<1><<<g, T, 0, s>>>(w.sb.keys_out, w.spos, w.sval, n_src, dst_pos, n_dst, inv_r, r2, f2, null_value, out, count);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // extern "C"
