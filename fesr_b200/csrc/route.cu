// ALDS routing on the device: PCA transform of the first `rows` nodes of every subdomain,
// standardisation, nearest k-means centroid (reference models/encoder.py:143-157,
// models/classifier.py:26-27,48-50; arithmetic of scikit-learn's PCA.transform /
// StandardScaler.transform / KMeans.predict restated in fp64).  One block per subdomain,
// fixed-shape tree reductions => deterministic labels.
#include "common.cuh"

namespace fesr {

constexpr int RT_THREADS = 256;
constexpr int RT_MAXCOMP = 8;

__device__ __forceinline__ int nearest_centroid(const double* z, int n_comp, const double* __restrict__ sc_mean,
                                                const double* __restrict__ sc_scale,
                                                const double* __restrict__ centroids, int n_clusters) {
  int best = 0;
  double bestd = 0.0;
  for (int k = 0; k < n_clusters; ++k) {
    double d2 = 0.0;
    for (int c = 0; c < n_comp; ++c) {
      const double zs = (z[c] - sc_mean[c]) / sc_scale[c];
      const double df = zs - centroids[(size_t)k * n_comp + c];
      d2 = fma(df, df, d2);
    }
    if (k == 0 || d2 < bestd) {      // first minimum on ties, like numpy argmin
      bestd = d2;
      best = k;
    }
  }
  return best;
}

__global__ void route_kernel(const float* __restrict__ x, int channels, const int32_t* __restrict__ node_ptr, int rows,
                             const double* __restrict__ pca_mean, const double* __restrict__ comps, int n_comp,
                             const double* __restrict__ sc_mean, const double* __restrict__ sc_scale,
                             const double* __restrict__ centroids, int n_clusters, int32_t* __restrict__ labels,
                             double* __restrict__ latent) {
  __shared__ double red[RT_MAXCOMP][RT_THREADS];
  const int s = blockIdx.x;
  const int64_t n0 = node_ptr[s];
  const int avail = node_ptr[s + 1] - node_ptr[s];
  const int F = rows * channels;
  double acc[RT_MAXCOMP];
#pragma unroll
  for (int c = 0; c < RT_MAXCOMP; ++c) acc[c] = 0.0;
  for (int f = threadIdx.x; f < F; f += RT_THREADS) {
    const int r = f / channels;
    const double v = (r < avail ? (double)x[n0 * channels + f] : 0.0) - pca_mean[f];
#pragma unroll
    for (int c = 0; c < RT_MAXCOMP; ++c)
      if (c < n_comp) acc[c] = fma(v, comps[(size_t)c * F + f], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < RT_MAXCOMP; ++c) red[c][threadIdx.x] = acc[c];
  __syncthreads();
  for (int off = RT_THREADS / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off)
      for (int c = 0; c < n_comp; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double z[RT_MAXCOMP];
    for (int c = 0; c < n_comp; ++c) {
      z[c] = red[c][0];
      if (latent) latent[(size_t)s * n_comp + c] = z[c];
    }
    if (labels && n_clusters > 0) labels[s] = nearest_centroid(z, n_comp, sc_mean, sc_scale, centroids, n_clusters);
  }
}

// classifier alone (KMeansClassifier.cluster on a latent array), one thread per subdomain
__global__ void cluster_kernel(const double* __restrict__ latent, int n_sub, int n_comp,
                               const double* __restrict__ sc_mean, const double* __restrict__ sc_scale,
                               const double* __restrict__ centroids, int n_clusters, int32_t* __restrict__ labels) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sub) return;
  double z[RT_MAXCOMP];
  for (int c = 0; c < n_comp; ++c) z[c] = latent[(size_t)s * n_comp + c];
  labels[s] = nearest_centroid(z, n_comp, sc_mean, sc_scale, centroids, n_clusters);
}

}  // namespace fesr

using namespace fesr;

extern "C" int fesr_route(const float* x, int32_t channels, const int32_t* node_ptr, int32_t n_sub, int32_t rows,
                          const double* pca_mean, const double* pca_components, int32_t n_comp,
                          const double* scaler_mean, const double* scaler_scale, const double* centroids,
                          int32_t n_clusters, int32_t* labels, double* latent, void* stream_) {
  FESR_CHECK_ARG(n_sub >= 0 && rows >= 1 && channels >= 1, "bad sizes");
  FESR_CHECK_ARG(n_comp >= 1 && n_comp <= RT_MAXCOMP, "n_components must be in [1, %d]", RT_MAXCOMP);
  if (n_sub == 0) return FESR_OK;
  FESR_CHECK_ARG(x && node_ptr && pca_mean && pca_components, "NULL pointer");
  FESR_CHECK_ARG(labels || latent, "nothing to compute");
  FESR_CHECK_ARG(!labels || n_clusters == 0 || (scaler_mean && scaler_scale && centroids), "NULL classifier arrays");
  ProfScope prof(PROF_GRAPH, as_stream(stream_));
  route_kernel<<<n_sub, RT_THREADS, 0, as_stream(stream_)>>>(x, channels, node_ptr, rows, pca_mean, pca_components,
                                                            n_comp, scaler_mean, scaler_scale, centroids, n_clusters,
                                                            labels, latent);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

// dst[r] = src[index[r]] (SCATTER: dst[index[r]] = src[r]); rows of q4 float4's, one thread per float4
template <bool SCATTER>
__global__ void move_rows_kernel(const float4* __restrict__ src, const int64_t* __restrict__ index, int64_t rows, int q4,
                                 float4* __restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= rows * q4) return;
  const int64_t r = t / q4;
  const int c = (int)(t - r * q4);
  const int64_t j = __ldg(index + r);
  if (SCATTER) dst[j * q4 + c] = src[t];
  else dst[t] = __ldg(src + j * q4 + c);
}

template <bool SCATTER>
static int move_rows(const float* src, const int64_t* index, int64_t rows, int32_t row_floats, float* dst, void* stream_) {
  FESR_CHECK_ARG(rows >= 0 && row_floats > 0 && row_floats % 4 == 0, "rows of a multiple of 4 floats");
  if (rows == 0) return FESR_OK;
  FESR_CHECK_ARG(src && index && dst, "NULL pointer");
  FESR_CHECK_ARG(((uintptr_t)src | (uintptr_t)dst) % 16 == 0, "16-byte aligned rows");
  const int q4 = row_floats / 4;
  move_rows_kernel<SCATTER><<<(unsigned)ceil_div(rows * q4, 256), 256, 0, as_stream(stream_)>>>(
      reinterpret_cast<const float4*>(src), index, rows, q4, reinterpret_cast<float4*>(dst));
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

extern "C" int fesr_gather_rows(const float* src, const int64_t* index, int64_t rows, int32_t row_floats, float* dst,
                                void* stream_) {
  return move_rows<false>(src, index, rows, row_floats, dst, stream_);
}

extern "C" int fesr_scatter_rows(const float* src, const int64_t* index, int64_t rows, int32_t row_floats, float* dst,
                                 void* stream_) {
  return move_rows<true>(src, index, rows, row_floats, dst, stream_);
}

extern "C" int fesr_cluster(const double* latent, int32_t n_sub, int32_t n_comp, const double* scaler_mean,
                            const double* scaler_scale, const double* centroids, int32_t n_clusters, int32_t* labels,
                            void* stream_) {
  FESR_CHECK_ARG(n_sub >= 0 && n_comp >= 1 && n_comp <= RT_MAXCOMP && n_clusters >= 1, "bad sizes");
  if (n_sub == 0) return FESR_OK;
  FESR_CHECK_ARG(latent && scaler_mean && scaler_scale && centroids && labels, "NULL pointer");
  cluster_kernel<<<(unsigned)ceil_div(n_sub, 128), 128, 0, as_stream(stream_)>>>(latent, n_sub, n_comp, scaler_mean,
                                                                                scaler_scale, centroids, n_clusters,
                                                                                labels);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}
