// Wall shear stress of the stitched velocity field -- the step AFTER the path (reference compute_wss.py:5-120):
//   vtkGradientFilter            point gradient of the velocity = mean of the (constant) gradients of the tets at the point
//   vtkDataSetSurfaceFilter      boundary faces = tet faces that belong to exactly one cell
//   vtkPolyDataNormals           point normal = normalised sum of the unit normals of the boundary faces at the point
//   tau = mu (grad u + grad u^T) n,  tau_wall = tau - (tau . n) n,  |tau_wall|            (compute_wss.py:86-99)
// Where this differs from VTK's filters (vtk==9.4.1, not vendored; PARITY UNPINNED): faces are oriented OUTWARD (VTK
// leaves the sign to its traversal order; the wall-shear VECTOR flips with it, the magnitude does not) and points on
// sharp edges are not split at the 30-degree feature angle -- they keep one normal, the normalised average of the
// walls that meet there.
//
// All reductions (cells of a node, faces of a node) run over sorted incidence lists in a fixed order: deterministic.
#include "common.cuh"
#include "sortutil.cuh"

namespace fesr {

// ---------------------------------------------------------------- per-cell gradient: G[i][j] = d u_i / d x_j
__global__ void tet_gradient_kernel(const float* __restrict__ pos, const int32_t* __restrict__ cells,
                                    const float* __restrict__ u, int64_t C, float* __restrict__ grad) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int4 v = reinterpret_cast<const int4*>(cells)[c];
  const int id[4] = {v.x, v.y, v.z, v.w};
  float p[4][3], f[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      p[k][d] = pos[(int64_t)id[k] * 3 + d];
      f[k][d] = u[(int64_t)id[k] * 3 + d];
    }
  // E rows = edge vectors from vertex 0; solve E g_i = (u_i(v_k) - u_i(v_0))_k for every component i
  float e[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) e[k][d] = p[k + 1][d] - p[0][d];
  const float c00 = e[1][1] * e[2][2] - e[1][2] * e[2][1], c01 = e[1][2] * e[2][0] - e[1][0] * e[2][2],
              c02 = e[1][0] * e[2][1] - e[1][1] * e[2][0];
  const float det = e[0][0] * c00 + e[0][1] * c01 + e[0][2] * c02;
  const float inv = det != 0.f ? 1.0f / det : 0.f;          // degenerate tet: zero gradient
  // inverse of E (adjugate / det): inv_e[d][k]
  float ie[3][3];
  ie[0][0] = c00 * inv;
  ie[1][0] = c01 * inv;
  ie[2][0] = c02 * inv;
  ie[0][1] = (e[0][2] * e[2][1] - e[0][1] * e[2][2]) * inv;
  ie[1][1] = (e[0][0] * e[2][2] - e[0][2] * e[2][0]) * inv;
  ie[2][1] = (e[0][1] * e[2][0] - e[0][0] * e[2][1]) * inv;
  ie[0][2] = (e[0][1] * e[1][2] - e[0][2] * e[1][1]) * inv;
  ie[1][2] = (e[0][2] * e[1][0] - e[0][0] * e[1][2]) * inv;
  ie[2][2] = (e[0][0] * e[1][1] - e[0][1] * e[1][0]) * inv;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float d1 = f[1][i] - f[0][i], d2 = f[2][i] - f[0][i], d3 = f[3][i] - f[0][i];
#pragma unroll
    for (int j = 0; j < 3; ++j) grad[c * 9 + i * 3 + j] = ie[j][0] * d1 + ie[j][1] * d2 + ie[j][2] * d3;
  }
}

// point value = mean over the incident entries; incidence list from fesr_occurrence_build over a flattened
// [M, verts] connectivity: entry j belongs to item occ_idx[j] / verts
template <int W>
__global__ void node_mean_kernel(const float* __restrict__ item_val, const int32_t* __restrict__ occ_ptr,
                                 const int32_t* __restrict__ occ_idx, int verts, int64_t N, int normalise,
                                 float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float acc[W];
#pragma unroll
  for (int k = 0; k < W; ++k) acc[k] = 0.f;
  const int b = occ_ptr[i], e = occ_ptr[i + 1];
  for (int j = b; j < e; ++j) {
    const int64_t item = occ_idx[j] / verts;
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k] += item_val[item * W + k];
  }
  float s = e > b ? 1.0f / (float)(e - b) : 0.f;
  if (normalise) {             // unit vector of the sum (point normals); W == 3
    const float n2 = acc[0] * acc[0] + acc[1] * acc[1] + acc[2 % W] * acc[2 % W];
    s = n2 > 0.f ? rsqrtf(n2) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < W; ++k) out[i * W + k] = acc[k] * s;
}

// ---------------------------------------------------------------- boundary faces
// face f of a tet = the three vertices other than vertex f; key = its sorted vertex triple (3 x 21 bits)
__global__ void face_keys_kernel(const int32_t* __restrict__ cells, int64_t C, uint64_t* __restrict__ keys,
                                 int32_t* __restrict__ vals) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 4 * C) return;
  const int64_t c = t >> 2;
  const int f = (int)(t & 3);
  int a[3], k = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q != f) a[k++] = cells[c * 4 + q];
  if (a[0] > a[1]) { const int s = a[0]; a[0] = a[1]; a[1] = s; }
  if (a[1] > a[2]) { const int s = a[1]; a[1] = a[2]; a[2] = s; }
  if (a[0] > a[1]) { const int s = a[0]; a[0] = a[1]; a[1] = s; }
  keys[t] = ((uint64_t)a[0] << 42) | ((uint64_t)a[1] << 21) | (uint64_t)a[2];
  vals[t] = (int32_t)t;
}

__global__ void face_unique_flags_kernel(const uint64_t* __restrict__ keys, int64_t m, int32_t* __restrict__ flags) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  const bool same_prev = j > 0 && keys[j - 1] == keys[j], same_next = j + 1 < m && keys[j + 1] == keys[j];
  flags[j] = (same_prev || same_next) ? 0 : 1;
}

// boundary face slot -> oriented vertex triple (outward: away from the tet's fourth vertex), unit normal, owner cell
__global__ void face_emit_kernel(const float* __restrict__ pos, const int32_t* __restrict__ cells,
                                 const int32_t* __restrict__ sorted_vals, const int32_t* __restrict__ flags,
                                 const int32_t* __restrict__ scan, int64_t m, int32_t* __restrict__ faces,
                                 int32_t* __restrict__ face_cell, float* __restrict__ face_normal) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m || !flags[j]) return;
  const int64_t t = sorted_vals[j], c = t >> 2;
  const int f = (int)(t & 3);
  int a[3], k = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q != f) a[k++] = cells[c * 4 + q];
  const int opp = cells[c * 4 + f];
  float p[3][3], po[3];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int d = 0; d < 3; ++d) p[q][d] = pos[(int64_t)a[q] * 3 + d];
#pragma unroll
  for (int d = 0; d < 3; ++d) po[d] = pos[(int64_t)opp * 3 + d];
  const float ux = p[1][0] - p[0][0], uy = p[1][1] - p[0][1], uz = p[1][2] - p[0][2];
  const float vx = p[2][0] - p[0][0], vy = p[2][1] - p[0][1], vz = p[2][2] - p[0][2];
  float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
  const float side = nx * (p[0][0] - po[0]) + ny * (p[0][1] - po[1]) + nz * (p[0][2] - po[2]);
  if (side < 0.f) {
    const int s = a[1]; a[1] = a[2]; a[2] = s;
    nx = -nx; ny = -ny; nz = -nz;
  }
  const float n2 = nx * nx + ny * ny + nz * nz;
  const float r = n2 > 0.f ? rsqrtf(n2) : 0.f;
  const int64_t o = scan[j];
  faces[o * 3] = a[0];
  faces[o * 3 + 1] = a[1];
  faces[o * 3 + 2] = a[2];
  face_cell[o] = (int32_t)c;
  face_normal[o * 3] = nx * r;
  face_normal[o * 3 + 1] = ny * r;
  face_normal[o * 3 + 2] = nz * r;
}

// ---------------------------------------------------------------- wall shear stress at the surface points
__global__ void wss_kernel(const float* __restrict__ grad_pt, const float* __restrict__ normal_pt,
                           const int32_t* __restrict__ surf_nodes, int64_t M, float mu, float* __restrict__ tau,
                           float* __restrict__ mag) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int64_t i = surf_nodes ? surf_nodes[s] : s;
  float g[9], n[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) g[k] = grad_pt[i * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) n[k] = normal_pt[i * 3 + k];
  float t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) t[r] = mu * ((g[r * 3 + 0] + g[0 * 3 + r]) * n[0] + (g[r * 3 + 1] + g[1 * 3 + r]) * n[1] +
                                          (g[r * 3 + 2] + g[2 * 3 + r]) * n[2]);
  const float tn = t[0] * n[0] + t[1] * n[1] + t[2] * n[2];
  float w[3], m2 = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    w[r] = t[r] - tn * n[r];
    m2 += w[r] * w[r];
    tau[s * 3 + r] = w[r];
  }
  mag[s] = sqrtf(m2);
}

}  // namespace fesr

using namespace fesr;

extern "C" {

int fesr_tet_gradient(const float* pos, const int32_t* cells, const float* field, int64_t C, float* grad, void* stream_) {
  FESR_CHECK_ARG(C >= 0 && C < (1ll << 29), "cell count out of range");
  if (C == 0) return FESR_OK;
  FESR_CHECK_ARG(pos && cells && field && grad, "NULL pointer");
  FESR_CHECK_ARG((reinterpret_cast<uintptr_t>(cells) & 15) == 0, "cells must be 16-byte aligned");
  tet_gradient_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, as_stream(stream_)>>>(pos, cells, field, C, grad);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

int fesr_incident_mean(const float* item_val, int32_t width, const int32_t* occ_ptr, const int32_t* occ_idx,
                       int32_t verts, int64_t N, int32_t unit_vector, float* out, void* stream_) {
  FESR_CHECK_ARG(width == 3 || width == 9, "width must be 3 or 9, got %d", width);
  FESR_CHECK_ARG(verts >= 1 && N >= 0 && (!unit_vector || width == 3), "bad arguments");
  if (N == 0) return FESR_OK;
  FESR_CHECK_ARG(item_val && occ_ptr && occ_idx && out, "NULL pointer");
  cudaStream_t s = as_stream(stream_);
  const unsigned g = (unsigned)ceil_div(N, 256);
  if (width == 9) node_mean_kernel<9><<<g, 256, 0, s>>>(item_val, occ_ptr, occ_idx, verts, N, 0, out);
  else node_mean_kernel<3><<<g, 256, 0, s>>>(item_val, occ_ptr, occ_idx, verts, N, unit_vector, out);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

size_t fesr_boundary_faces_workspace_bytes(int64_t C) {
  if (C < 0) return 0;
  const int64_t m = 4 * (C > 0 ? C : 1);
  return ((sort_pairs_u64_bytes(m) + 255) & ~(size_t)255) + 2 * (((size_t)(m + 1) * sizeof(int32_t) + 255) & ~(size_t)255) +
         ((scan_temp_bytes(m + 1) + 255) & ~(size_t)255);
}

// faces [4C, 3], face_cell [4C], face_normal [4C, 3] are CAPACITY buffers; the number of boundary faces is written
// to host_count (the call synchronises the stream for that, like the *_count calls of the assembly)
int fesr_boundary_faces(const float* pos, const int32_t* cells, int64_t N, int64_t C, int32_t* faces, int32_t* face_cell,
                        float* face_normal, int64_t* host_count, void* workspace, size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(C >= 0 && C < (1ll << 29) && N >= 0 && N < (1ll << 21), "sizes out of range (node ids are packed in 21 bits)");
  FESR_CHECK_ARG(host_count != nullptr, "host_count is NULL");
  *host_count = 0;
  if (C == 0) return FESR_OK;
  FESR_CHECK_ARG(pos && cells && faces && face_cell && face_normal, "NULL pointer");
  FESR_CHECK_ARG(workspace && workspace_bytes >= fesr_boundary_faces_workspace_bytes(C), "workspace too small");
  cudaStream_t s = as_stream(stream_);
  const int64_t m = 4 * C;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  SortBuffers sb = carve_sort_buffers(base, m);
  size_t off = (sort_pairs_u64_bytes(m) + 255) & ~(size_t)255;
  int32_t* flags = reinterpret_cast<int32_t*>(base + off);
  off += ((size_t)(m + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
  int32_t* scan = reinterpret_cast<int32_t*>(base + off);
  off += ((size_t)(m + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
  void* scan_tmp = base + off;
  const int T = 256;
  face_keys_kernel<<<(unsigned)ceil_div(m, T), T, 0, s>>>(cells, C, sb.keys_in, sb.vals_in);
  FESR_LAUNCH_CHECK();
  int rc = sort_pairs_u64(sb, m, 0, 63, s);
  if (rc) return rc;
  face_unique_flags_kernel<<<(unsigned)ceil_div(m, T), T, 0, s>>>(sb.keys_out, m, flags);
  FESR_LAUNCH_CHECK();
  if ((rc = exclusive_scan_i32(flags, scan, m, scan_tmp, scan_temp_bytes(m + 1), s))) return rc;
  if ((rc = launch_scan_total(flags, scan, m, s))) return rc;
  face_emit_kernel<<<(unsigned)ceil_div(m, T), T, 0, s>>>(pos, cells, sb.vals_out, flags, scan, m, faces, face_cell, face_normal);
  FESR_LAUNCH_CHECK();
  int32_t total = 0;
  FESR_CUDA(cudaMemcpyAsync(&total, scan + m, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FESR_CUDA(cudaStreamSynchronize(s));
  *host_count = total;
  return FESR_OK;
}

int fesr_wall_shear_stress(const float* grad_pt, const float* normal_pt, const int32_t* surf_nodes, int64_t M, float mu,
                           float* tau, float* mag, void* stream_) {
  FESR_CHECK_ARG(M >= 0, "negative count");
  if (M == 0) return FESR_OK;
  FESR_CHECK_ARG(grad_pt && normal_pt && tau && mag, "NULL pointer");
  wss_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream_)>>>(grad_pt, normal_pt, surf_nodes, M, mu, tau, mag);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // extern "C"
