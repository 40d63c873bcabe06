// kd decomposition of the mesh cells on the GPU (bit-exact with oracle/graph.py kd_build/kd_assign).
//
// Stands in for vtkRedistributeDataSetFilter (reference dataset/GraphDataset.py:1208-1230):
// exact-median bisection of the fp32 cell centroids, level by level; one stable radix sort of
// (region, coordinate) per level gives every region's median at once.  Halo cells are then
// assigned to every leaf whose box their bounding box touches (AssignToAllIntersectingRegions,
// :1219) or to the centroid's leaf only (AssignToOneRegion, :565).
#include "common.cuh"
#include "sortutil.cuh"

namespace fesr {

__device__ __forceinline__ uint32_t f2ord(float f) {
  if (f == 0.f) f = 0.f;                       // -0 -> +0 so that ties compare equal, as numpy does
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// centroid = ((p0+p1)+(p2+p3))*0.25, SoA [3][C]
__global__ void centroid_kernel(const float* __restrict__ pos, const int32_t* __restrict__ cells, int64_t C,
                                float* __restrict__ cent) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int4 v = reinterpret_cast<const int4*>(cells)[c];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float s01 = __fadd_rn(pos[(int64_t)v.x * 3 + a], pos[(int64_t)v.y * 3 + a]);
    const float s23 = __fadd_rn(pos[(int64_t)v.z * 3 + a], pos[(int64_t)v.w * 3 + a]);
    cent[a * C + c] = __fmul_rn(__fadd_rn(s01, s23), 0.25f);
  }
}

__global__ void fill_u32_kernel(uint32_t* p, int64_t n, uint32_t v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// per-region bounding box of the centroids (ordered-uint min / max; order independent => exact)
__global__ void region_bbox_kernel(const float* __restrict__ cent, const int32_t* __restrict__ region, int64_t C,
                                   int R, uint32_t* __restrict__ bbmin, uint32_t* __restrict__ bbmax) {
  constexpr int SR = 32;
  __shared__ uint32_t smin[SR * 3], smax[SR * 3];
  const bool use_smem = R <= SR;
  if (use_smem) {
    for (int i = threadIdx.x; i < R * 3; i += blockDim.x) {
      smin[i] = 0xffffffffu;
      smax[i] = 0u;
    }
    __syncthreads();
  }
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x) {
    const int r = region[c];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const uint32_t o = f2ord(cent[a * C + c]);
      if (use_smem) {
        atomicMin(&smin[r * 3 + a], o);
        atomicMax(&smax[r * 3 + a], o);
      } else {
        atomicMin(&bbmin[r * 3 + a], o);
        atomicMax(&bbmax[r * 3 + a], o);
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < R * 3; i += blockDim.x) {
      if (smin[i] != 0xffffffffu) atomicMin(&bbmin[i], smin[i]);
      if (smax[i] != 0u) atomicMax(&bbmax[i], smax[i]);
    }
  }
}

__global__ void region_axis_kernel(const uint32_t* __restrict__ bbmin, const uint32_t* __restrict__ bbmax, int R,
                                   int node0, int32_t* __restrict__ axis_of, int32_t* __restrict__ tree_axis) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int axis = 0;
  if (bbmin[r * 3] <= bbmax[r * 3]) {   // non-empty
    float ext[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ext[a] = __fsub_rn(ord2f(bbmax[r * 3 + a]), ord2f(bbmin[r * 3 + a]));
    axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
  }
  axis_of[r] = axis;
  tree_axis[node0 + r] = axis;
}

__global__ void level_keys_kernel(const float* __restrict__ cent, const int32_t* __restrict__ region,
                                  const int32_t* __restrict__ axis_of, int64_t C, uint64_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r = region[c];
  keys[c] = ((uint64_t)r << 32) | f2ord(cent[(int64_t)axis_of[r] * C + c]);
  vals[c] = (int32_t)c;
}

__global__ void region_split_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ starts, int R,
                                    int node0, float* __restrict__ tree_split) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int b = starts[r], m = starts[r + 1] - b;
  tree_split[node0 + r] = (m > 0) ? ord2f((uint32_t)(keys[b + m / 2] & 0xffffffffull)) : INFINITY;
}

__global__ void relabel_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                               const int32_t* __restrict__ starts, int64_t C, int32_t* __restrict__ region) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= C) return;
  const int r = (int)(keys[j] >> 32);
  const int b = starts[r], m = starts[r + 1] - b;
  region[vals[j]] = 2 * r + (((int)j - b) >= m / 2 ? 1 : 0);
}

// relabel + the bounding boxes of the NEXT level's regions, in one pass over the level's sorted order.  Position j
// of the sorted array belongs to child 2r + (j - b >= m / 2) of its region r, and that child id is non-decreasing in
// j: a warp sees one run (rarely two or three), so the per-region min / max is a segmented warp scan followed by one
// atomic per run and axis from the run's last lane -- ~6 atomics per warp instead of the 192 a per-cell atomic pass
// over unsorted cells needs (which serialised on 6 R addresses and cost 130 us per level at 527 k cells).
// Ordered-uint min / max is order independent, so the boxes are exact.
__global__ void relabel_bbox_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                    const int32_t* __restrict__ starts, const float* __restrict__ cent, int64_t C,
                                    int32_t* __restrict__ region, uint32_t* __restrict__ bbmin,
                                    uint32_t* __restrict__ bbmax) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int nr = 0x7fffffff;
  uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  if (j < C) {
    const int r = (int)(keys[j] >> 32);
    const int b = starts[r], m = starts[r + 1] - b;
    const int cell = vals[j];
    nr = 2 * r + (((int)j - b) >= m / 2 ? 1 : 0);
    region[cell] = nr;
    if (bbmin != nullptr) {
#pragma unroll
      for (int a = 0; a < 3; ++a) lo[a] = hi[a] = f2ord(cent[a * C + cell]);
    }
  }
  if (bbmin == nullptr) return;          // last level: nothing splits the leaves further
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int onr = __shfl_up_sync(FULL, nr, off);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const uint32_t ol = __shfl_up_sync(FULL, lo[a], off), oh = __shfl_up_sync(FULL, hi[a], off);
      if (lane >= off && onr == nr) {
        lo[a] = min(lo[a], ol);
        hi[a] = max(hi[a], oh);
      }
    }
  }
  const int nnr = __shfl_down_sync(FULL, nr, 1);
  if (j < C && (lane == 31 || nnr != nr)) {          // last lane of a run holds the run's box
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bbmin[nr * 3 + a], lo[a]);
      atomicMax(&bbmax[nr * 3 + a], hi[a]);
    }
  }
}

// ------------------------------------------------------------------------------- assignment
struct AssignArgs {
  const float* pos;
  const int32_t* cells;
  const int32_t* home_leaf;
  const int32_t* tree_axis;
  const float* tree_split;
  int64_t C;
  int levels;
  int mode;
};

// Visits the leaves of cell c in ascending order, calling f(leaf).
template <typename F>
__device__ __forceinline__ void visit_leaves(const AssignArgs& a, int64_t c, F&& f) {
  const int home = a.home_leaf[c];
  if (a.mode == FESR_ONE_REGION || a.levels == 0) {
    f(home);
    return;
  }
  const int4 v = reinterpret_cast<const int4*>(a.cells)[c];
  float lo[3], hi[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float p0 = a.pos[(int64_t)v.x * 3 + ax], p1 = a.pos[(int64_t)v.y * 3 + ax];
    const float p2 = a.pos[(int64_t)v.z * 3 + ax], p3 = a.pos[(int64_t)v.w * 3 + ax];
    lo[ax] = fminf(fminf(p0, p1), fminf(p2, p3));
    hi[ax] = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
  }
  // iterative DFS, left first; stack entries: within-level index and depth
  int st_node[32];
  int st_depth[32];
  int sp = 0;
  st_node[0] = 0;
  st_depth[0] = 0;
  sp = 1;
  while (sp > 0) {
    --sp;
    const int j = st_node[sp], d = st_depth[sp];
    if (d == a.levels) {
      f(j);
      continue;
    }
    const int heap = (1 << d) - 1 + j;
    const int ax = a.tree_axis[heap];
    const float split = a.tree_split[heap];
    const bool home_here = (home >> (a.levels - d)) == j;
    const bool home_right = ((home >> (a.levels - d - 1)) & 1) == 1;
    const float l = ax == 0 ? lo[0] : (ax == 1 ? lo[1] : lo[2]);
    const float h = ax == 0 ? hi[0] : (ax == 1 ? hi[1] : hi[2]);
    const bool go_left = (l < split) || (home_here && !home_right);
    const bool go_right = (h >= split) || (home_here && home_right);
    if (go_right) {               // push right first so that left is visited first
      st_node[sp] = 2 * j + 1;
      st_depth[sp] = d + 1;
      ++sp;
    }
    if (go_left) {
      st_node[sp] = 2 * j;
      st_depth[sp] = d + 1;
      ++sp;
    }
  }
}

__global__ void assign_count_kernel(AssignArgs a, int32_t* __restrict__ cell_count, int32_t* __restrict__ leaf_count) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  int n = 0;
  visit_leaves(a, c, [&](int leaf) {
    ++n;
    if (leaf_count) atomicAdd(&leaf_count[leaf], 1);
  });
  cell_count[c] = n;
}

__global__ void assign_emit_kernel(AssignArgs a, const int32_t* __restrict__ cell_off, uint64_t* __restrict__ keys,
                                   int32_t* __restrict__ vals) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  int o = cell_off[c];
  visit_leaves(a, c, [&](int leaf) {
    keys[o] = (uint64_t)leaf;
    vals[o] = (int32_t)c;
    ++o;
  });
}

struct PartWs {
  float* cent;
  uint32_t* bbmin;
  uint32_t* bbmax;
  int32_t* axis_of;
  int32_t* starts;
  SortBuffers sb;
  size_t bytes;
};

static PartWs carve_part(void* base, int64_t C, int levels) {
  Carver c(base);
  PartWs w;
  const int64_t R = 1ll << (levels > 0 ? levels - 1 : 0);
  w.cent = c.take<float>(3 * (size_t)(C > 0 ? C : 1));
  w.bbmin = c.take<uint32_t>(3 * R);
  w.bbmax = c.take<uint32_t>(3 * R);
  w.axis_of = c.take<int32_t>(R);
  w.starts = c.take<int32_t>(R + 1);
  const size_t sort_off = c.used();
  w.sb = carve_sort_buffers(base ? static_cast<char*>(base) + sort_off : nullptr, C);
  w.bytes = sort_off + sort_pairs_u64_bytes(C);
  return w;
}

struct AssignWs {
  int32_t* cell_count;
  int32_t* cell_off;
  int32_t* leaf_count;
  void* scan_temp;
  size_t scan_bytes;
  SortBuffers sb;
  size_t bytes;
};

static AssignWs carve_assign(void* base, int64_t C, int levels, int64_t total) {
  Carver c(base);
  AssignWs w;
  const int64_t S = 1ll << levels;
  const size_t cc = (size_t)(C > 0 ? C : 1);
  w.cell_count = c.take<int32_t>(cc + 1);
  w.cell_off = c.take<int32_t>(cc + 1);
  w.leaf_count = c.take<int32_t>(S + 1);
  w.scan_bytes = scan_temp_bytes((int64_t)cc + 1 > S + 1 ? (int64_t)cc + 1 : S + 1);
  w.scan_temp = c.take<char>(w.scan_bytes);
  const size_t sort_off = c.used();
  w.sb = carve_sort_buffers(base ? static_cast<char*>(base) + sort_off : nullptr, total);
  w.bytes = sort_off + sort_pairs_u64_bytes(total);
  return w;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

size_t fesr_partition_workspace_bytes(int64_t C, int32_t levels) {
  if (C < 0 || levels < 0 || levels > 20) return 0;
  return carve_part(nullptr, C, levels).bytes;
}

int fesr_partition_cells(const float* pos, const int32_t* cells, int64_t N, int64_t C, int32_t levels,
                         int32_t* home_leaf, int32_t* tree_axis, float* tree_split, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(levels >= 0 && levels <= 20, "levels must be in [0, 20]");
  FESR_CHECK_ARG(C >= 0 && C < (1ll << 31) && N >= 0 && N < (1ll << 31), "C/N out of int32 range");
  cudaStream_t s = as_stream(stream_);
  if (C == 0 && levels == 0) return FESR_OK;
  FESR_CHECK_ARG(C == 0 || (pos && cells && home_leaf), "NULL pointer");
  FESR_CHECK_ARG(levels == 0 || (tree_axis && tree_split), "NULL tree arrays");
  if (C > 0) FESR_CUDA(cudaMemsetAsync(home_leaf, 0, (size_t)C * sizeof(int32_t), s));
  if (levels == 0) return FESR_OK;
  PartWs w = carve_part(workspace, C, levels);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("partition workspace too small: need %zu, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  const int T = 256;
  const unsigned gC = (unsigned)ceil_div(C > 0 ? C : 1, T);
  if (C > 0) {
    centroid_kernel<<<gC, T, 0, s>>>(pos, cells, C, w.cent);
    FESR_LAUNCH_CHECK();
  }
  for (int d = 0; d < levels; ++d) {
    const int R = 1 << d, node0 = R - 1;
    if (d == 0) {      // the root's box: one region (shared-memory path); deeper levels get theirs from relabel_bbox_kernel
      fill_u32_kernel<<<1, T, 0, s>>>(w.bbmin, 3, 0xffffffffu);
      fill_u32_kernel<<<1, T, 0, s>>>(w.bbmax, 3, 0u);
      FESR_LAUNCH_CHECK();
      if (C > 0) {
        const unsigned gb = gC < 4u * num_sms() ? gC : 4u * num_sms();
        region_bbox_kernel<<<gb, T, 0, s>>>(w.cent, home_leaf, C, R, w.bbmin, w.bbmax);
        FESR_LAUNCH_CHECK();
      }
    }
    region_axis_kernel<<<(unsigned)ceil_div(R, T), T, 0, s>>>(w.bbmin, w.bbmax, R, node0, w.axis_of, tree_axis);
    FESR_LAUNCH_CHECK();
    if (C > 0) {
      level_keys_kernel<<<gC, T, 0, s>>>(w.cent, home_leaf, w.axis_of, C, w.sb.keys_in, w.sb.vals_in);
      FESR_LAUNCH_CHECK();
      int rc = sort_pairs_u64(w.sb, C, 0, 32 + d, s);
      if (rc) return rc;
    }
    int rc = launch_ptr_from_sorted(w.sb.keys_out, C, 32, R, w.starts, s);
    if (rc) return rc;
    region_split_kernel<<<(unsigned)ceil_div(R, T), T, 0, s>>>(w.sb.keys_out, w.starts, R, node0, tree_split);
    FESR_LAUNCH_CHECK();
    const bool more = d + 1 < levels;
    if (more) {        // (region_axis_kernel of this level has consumed the boxes: the arrays take the next level's)
      fill_u32_kernel<<<(unsigned)ceil_div(6 * R, T), T, 0, s>>>(w.bbmin, 6 * R, 0xffffffffu);
      fill_u32_kernel<<<(unsigned)ceil_div(6 * R, T), T, 0, s>>>(w.bbmax, 6 * R, 0u);
      FESR_LAUNCH_CHECK();
    }
    if (C > 0) {
      relabel_bbox_kernel<<<gC, T, 0, s>>>(w.sb.keys_out, w.sb.vals_out, w.starts, w.cent, C, home_leaf,
                                         more ? w.bbmin : nullptr, more ? w.bbmax : nullptr);
      FESR_LAUNCH_CHECK();
    }
  }
  return FESR_OK;
}

size_t fesr_assign_workspace_bytes(int64_t C, int32_t levels, int64_t total_pairs) {
  if (C < 0 || levels < 0 || levels > 20 || total_pairs < 0) return 0;
  return carve_assign(nullptr, C, levels, total_pairs).bytes;
}

static int assign_common(const AssignArgs& a, AssignWs& w, bool with_leaf_count, cudaStream_t s) {
  const int T = 256;
  const int64_t S = 1ll << a.levels;
  if (with_leaf_count) FESR_CUDA(cudaMemsetAsync(w.leaf_count, 0, (size_t)(S + 1) * sizeof(int32_t), s));
  FESR_CUDA(cudaMemsetAsync(w.cell_count, 0, (size_t)(a.C + 1) * sizeof(int32_t), s));
  if (a.C > 0) {
    assign_count_kernel<<<(unsigned)ceil_div(a.C, T), T, 0, s>>>(a, w.cell_count, with_leaf_count ? w.leaf_count : nullptr);
    FESR_LAUNCH_CHECK();
  }
  return FESR_OK;
}

int fesr_assign_count(const float* pos, const int32_t* cells, int64_t C, int32_t levels, int32_t mode,
                      const int32_t* home_leaf, const int32_t* tree_axis, const float* tree_split,
                      int32_t* leaf_ptr, int64_t* host_total, void* workspace, size_t workspace_bytes,
                      void* stream_) {
  FESR_CHECK_ARG(levels >= 0 && levels <= 20 && C >= 0 && C < (1ll << 31), "bad C/levels");
  FESR_CHECK_ARG(mode == FESR_ONE_REGION || mode == FESR_ALL_INTERSECTING, "bad mode");
  FESR_CHECK_ARG(leaf_ptr && host_total && (C == 0 || (pos && cells && home_leaf)), "NULL pointer");
  cudaStream_t s = as_stream(stream_);
  AssignWs w = carve_assign(workspace, C, levels, 0);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("assign workspace too small: need %zu, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  AssignArgs a{pos, cells, home_leaf, tree_axis, tree_split, C, levels, mode};
  int rc = assign_common(a, w, true, s);
  if (rc) return rc;
  const int64_t S = 1ll << levels;
  rc = exclusive_scan_i32(w.leaf_count, leaf_ptr, S + 1, w.scan_temp, w.scan_bytes, s);
  if (rc) return rc;
  int32_t total32 = 0;
  FESR_CUDA(cudaMemcpyAsync(&total32, leaf_ptr + S, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FESR_CUDA(cudaStreamSynchronize(s));
  *host_total = total32;
  return FESR_OK;
}

int fesr_assign_fill(const float* pos, const int32_t* cells, int64_t C, int32_t levels, int32_t mode,
                     const int32_t* home_leaf, const int32_t* tree_axis, const float* tree_split,
                     const int32_t* leaf_ptr, int64_t total_pairs, int32_t* leaf_cells, void* workspace,
                     size_t workspace_bytes, void* stream_) {
  FESR_CHECK_ARG(levels >= 0 && levels <= 20 && C >= 0 && C < (1ll << 31), "bad C/levels");
  FESR_CHECK_ARG(total_pairs >= 0 && total_pairs < (1ll << 31), "bad total_pairs");
  (void)leaf_ptr;
  if (total_pairs == 0) return FESR_OK;
  FESR_CHECK_ARG(leaf_cells && pos && cells && home_leaf, "NULL pointer");
  cudaStream_t s = as_stream(stream_);
  AssignWs w = carve_assign(workspace, C, levels, total_pairs);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("assign workspace too small: need %zu, got %zu", w.bytes, workspace_bytes);
    return FESR_EWORKSPACE;
  }
  AssignArgs a{pos, cells, home_leaf, tree_axis, tree_split, C, levels, mode};
  int rc = assign_common(a, w, false, s);
  if (rc) return rc;
  rc = exclusive_scan_i32(w.cell_count, w.cell_off, C + 1, w.scan_temp, w.scan_bytes, s);
  if (rc) return rc;
  assign_emit_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, s>>>(a, w.cell_off, w.sb.keys_in, w.sb.vals_in);
  FESR_LAUNCH_CHECK();
  rc = sort_pairs_u64(w.sb, total_pairs, 0, levels > 0 ? levels : 1, s);
  if (rc) return rc;
  FESR_CUDA(cudaMemcpyAsync(leaf_cells, w.sb.vals_out, (size_t)total_pairs * sizeof(int32_t),
                            cudaMemcpyDeviceToDevice, s));
  return FESR_OK;
}

}  // extern "C"
