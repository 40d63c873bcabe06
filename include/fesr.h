/*
 * fesr.h -- C ABI of libfesr.so: the B200-native (sm_100a only) hot path of
 * cmudrc/fast-eng-super-resolution.
 *
 * The reference has no FFI: its boundary for this path is Python.
 * Each entry point below names the reference interface (file:line under /root/reference)
 * whose device work it replaces; INTEGRATION.md shows the ctypes stubs a maintainer adds
 * on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless named host_*;
 *     the library never allocates, frees or retains caller memory;
 *   - workspace is sized by the matching *_workspace_bytes() query and passed in;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *     the library never calls cudaDeviceSynchronize;
 *   - return value: 0 on success, a negative FESR_E* code otherwise, message from
 *     fesr_last_error() (thread local).  No C++ exception crosses this boundary;
 *   - there is NO CPU fallback and no other GPU target: a device that is not
 *     compute capability 10.x is FESR_EDEVICE.
 */
#ifndef FESR_H_
#define FESR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FESR_VERSION 100

enum {
  FESR_OK = 0,
  FESR_EINVAL = -1,   /* bad shape / null pointer / unsupported width */
  FESR_EDEVICE = -2,  /* not an sm_100 device */
  FESR_ECUDA = -3,    /* a CUDA runtime call failed */
  FESR_EWORKSPACE = -4
};

enum { FESR_KERNELNN = 0, FESR_TEECNET = 1 };
/* arithmetic of the node contraction Z x T' (everything else is fp32 CUDA-core):
 *   FP32   fp32 FFMA, rel-L2 <= 1e-5 vs the reference's fp32 CPU result
 *   TF32   tcgen05.mma kind::tf32, fp32 accumulate in TMEM, rel-L2 <= 1e-3
 *   F16    tcgen05.mma kind::f16 on fp16 Z / T' (11-bit mantissa like TF32, half the HBM bytes of
 *          the [n, zk] intermediate; values saturate at +-65504), fp32 accumulate, rel-L2 <= 1e-3 */
enum { FESR_PREC_FP32 = 0, FESR_PREC_TF32 = 1, FESR_PREC_TF32X3 = 2 /* reserved */, FESR_PREC_F16 = 3 };
enum { FESR_ONE_REGION = 0, FESR_ALL_INTERSECTING = 1 };

int fesr_version(void);
const char* fesr_last_error(void);
/* 0 if the current device is sm_100; FESR_EDEVICE otherwise. */
int fesr_device_check(void);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long fesr_launch_count(void);
/* Optional CUDA-event timing per kernel class, recorded on the launching stream (bench.py's
 * roofline figures).  Classes: 0 prepare, 1 edge_hidden, 2 fc_in, 3 zbuild, 4 node_gemm,
 * 5 fc_out, 6 node_weight, 7 stitch, 8 graph, 9 backward.  _collect synchronises the recorded
 * events, sums milliseconds / scopes per class and clears the record. */
#define FESR_PROF_NKINDS 11
int fesr_profile_enable(int on);
int fesr_profile_collect(double* ms_by_kind, long long* launches_by_kind, int nkinds);

/* ------------------------------------------------------------------------------------
 * Model geometry.  Filled by fesr_model_dims_init from (kind, width, channels, layers).
 * Replaces nothing in the reference; it fixes the padded HBM layouts every kernel uses.
 * ---------------------------------------------------------------------------------- */
typedef struct fesr_model_dims {
  int32_t kind;        /* FESR_KERNELNN | FESR_TEECNET */
  int32_t w;           /* hidden width (43 in the shipped checkpoints, 48 default config) */
  int32_t wp;          /* padded width: multiple of 16, >= w (+1 constant column for TEECNet) */
  int32_t in_ch;       /* 4 */
  int32_t out_ch;      /* 4 */
  int32_t layers;      /* conv applications (shared weights), 5 */
  int32_t n_hidden;    /* hidden Linear layers of the edge MLP: 2 (KernelNN) / 3 (TEECNet) */
  int32_t hidden[4];   /* their output sizes: {w,w} / {32,64,128} */
  int32_t k1;          /* edge-feature channels incl. the constant-1 channel: hidden[last]+1 */
  int32_t kt;          /* channels per lane group in the outer-product kernel */
  int32_t ktp;         /* kt rounded up to 4 (16-byte rows) */
  int32_t passes;      /* passes of 4*kt channels */
  int32_t kp;          /* row stride of g (floats) = passes*4*ktp */
  int32_t k1p;         /* padded channel count = passes*4*kt */
  int32_t zk_main;     /* k1p*wp */
  int32_t zk;          /* row stride of Z (elements): zk_main + wp rounded up to 64 */
  int32_t leaky;       /* 0: ReLU edge MLP + ReLU between layers; 1: LeakyReLU(0.01), none between */
} fesr_model_dims;

int fesr_model_dims_init(int kind, int w, int in_ch, int out_ch, int layers, fesr_model_dims* out);

/* Device pointers to the model's parameters, named as in the reference's state_dict
 * (models/model.py:543-554 KernelNN / :269-276 + :395-410 TEECNet).  fp32, contiguous. */
typedef struct fesr_params {
  const float* fc1_w;      /* [w, in_ch]            fc1.weight */
  const float* fc1_b;      /* [w]                   fc1.bias */
  const float* mlp_w[4];   /* edge MLP Linear weights in order; the last one is [w*w, hidden[last]] */
  const float* mlp_b[4];   /* and biases              (conv1.nn.layers.{0,2,4} / kernel.operator_kernel.layers.{0,2,4,6}) */
  const float* lin_w;      /* [w, w]  kernel.linear.weight (TEECNet) or NULL */
  const float* lin_b;      /* [w]     kernel.linear.bias   (TEECNet) or NULL */
  const float* root;       /* [w(in), w(out)]       conv1.root / kernel.root_param */
  const float* bias;       /* [w]                   conv1.bias / kernel.bias */
  const float* fc2_w;      /* [out_ch, w]           fc2.weight / fc_out.weight */
  const float* fc2_b;      /* [out_ch]              fc2.bias   / fc_out.bias */
} fesr_params;

/* Same shapes, gradient accumulators (fesr_nnconv_backward ADDS into them). */
typedef struct fesr_param_grads {
  float* fc1_w; float* fc1_b;
  float* mlp_w[4]; float* mlp_b[4];
  float* lin_w; float* lin_b;
  float* root; float* bias;
  float* fc2_w; float* fc2_b;
} fesr_param_grads;

/* ------------------------------------------------------------------------------------
 * Graph: CSR of edge_index sorted by destination.
 * Replaces the index_select / scatter_add_ pair inside torch_geometric's
 * MessagePassing.propagate (call sites models/model.py:424,525) with a deterministic layout.
 *   edge_index : [2, E] int64 row-major (row 0 = source j, row 1 = destination i), any order
 *   rowptr     : [n+1] int32   edges of destination i are [rowptr[i], rowptr[i+1])
 *   src_sorted : [E]   int32   source of the e-th CSR edge; order is (dst, src, original id)
 *   perm       : [E]   int32   original edge id of the e-th CSR edge
 * ---------------------------------------------------------------------------------- */
size_t fesr_csr_workspace_bytes(int64_t n, int64_t E);
int fesr_csr_build(const int64_t* edge_index, int64_t E, int64_t n,
                   int32_t* rowptr, int32_t* src_sorted, int32_t* perm,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Message passing forward.
 * Replaces KernelNN.forward / NNConv_old.{forward,message,update} / DenseNet.forward
 * (models/model.py:555-561, 521-536, 311-315), TEECNet.forward / KernelConv.*
 * (models/model.py:278-286, 421-445) and PyG's mean aggregation, for one block-diagonal
 * batch of subdomains (what models/scheduler_gnn.py:217-226 loops over one by one).
 *
 *   x          : [n, in_ch] fp32
 *   edge_attr  : [E] fp32 in ORIGINAL edge order (perm maps CSR slot -> original id);
 *                perm == NULL means edge_attr is already in CSR order
 *   y          : [n, out_ch] fp32
 *   workspace  : fesr_forward_workspace_bytes(dims, n, E, keep) bytes.  With FESR_FWD_KEEP the
 *                per-layer activations needed by fesr_nnconv_backward stay in it.
 *   keep_for_backward : bit flags.  FESR_FWD_KEEP (1) as above.  FESR_FWD_WEIGHTS_PREPARED (2): the
 *                head of this workspace still holds the padded / permuted weight copies that an
 *                earlier call made from the SAME parameter values (predict loops: the weights do
 *                not change between calls), so the preparation kernels are skipped.
 * ---------------------------------------------------------------------------------- */
#define FESR_FWD_KEEP 1
#define FESR_FWD_WEIGHTS_PREPARED 2
/* Two-phase call, so that the caller can overlap the host -> device copy of x with the part of the pass that does
 * not read x: FESR_FWD_EDGE_ONLY (4) runs the weight preparation and the edge MLP (x and y are not touched and may
 * be NULL) and leaves the edge features in the workspace; a following call with FESR_FWD_EDGE_DONE (8) on the SAME
 * workspace, dims, parameters, graph, edge_attr and precision runs the rest (fc1, the layers, fc2). */
#define FESR_FWD_EDGE_ONLY 4
#define FESR_FWD_EDGE_DONE 8
/* Size query only (fesr_forward_workspace_bytes): with FESR_FWD_KEEP and the FESR_PREC_TF32 arm the kept Z stash is
 * fp16 (half the bytes per layer); a caller that will run that arm may pass FESR_FWD_KEEP | FESR_FWD_KEEP_Z16 to get
 * the smaller size.  Without it the query returns the fp32-stash size, which is always enough. */
#define FESR_FWD_KEEP_Z16 16
size_t fesr_forward_workspace_bytes(const fesr_model_dims* dims, int64_t n, int64_t E, int keep_for_backward);
/* fp16 range guard of the reduced-precision arms: byte offset, inside ANY forward workspace of `dims`, of one int32
 * that the pass zeroes when it starts and sets to 1 if a value it packs into fp16 (edge features g, node features h,
 * the Z stash) exceeds 65504 or is not a number.  When the flag is set y is filled with NaN instead of a plausible
 * field built from clipped intermediates; the caller reads the flag (after the stream has reached the end of the
 * pass) to tell "overflow: use FESR_PREC_TF32 / FP32" from NaN inputs. */
size_t fesr_forward_overflow_offset(const fesr_model_dims* dims);
int fesr_nnconv_forward(const fesr_model_dims* dims, const fesr_params* params,
                        const float* x, const int32_t* rowptr, const int32_t* src_sorted,
                        const int32_t* perm, const float* edge_attr,
                        int64_t n, int64_t E, int precision, int keep_for_backward,
                        float* y, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above for the train step (models/scheduler_gnn.py:398-408: MSELoss ->
 * loss.backward()).  Needs the workspace of a forward run with keep_for_backward = 1 and the CSR
 * of the REVERSED graph built from the forward CSR slots (fesr_csr_build on the [2,E] array
 * {dst of slot e ; src of slot e}): rowptr_t [n+1] groups the edges by SOURCE node, src_t [E] is
 * the original destination of each reversed slot and rev_to_fwd [E] its forward CSR slot.
 * grad_y [n, out_ch] in; parameter gradients are ADDED into *grads; grad_x [n, in_ch] may be NULL.
 * The workspace of the tf32 arm keeps one bf16 dZ [n, zk] per layer (2 * n * zk bytes each; layers <= 8) so that the
 * edge gradient runs once over all layers; FESR_EDGE_GRAD_LAYERS=0 in the environment (read once per process, by the
 * size query as well) drops those buffers and runs the edge gradient layer by layer. */
size_t fesr_backward_workspace_bytes(const fesr_model_dims* dims, int64_t n, int64_t E);
int fesr_nnconv_backward(const fesr_model_dims* dims, const fesr_params* params,
                         const float* x, const int32_t* rowptr, const int32_t* src_sorted,
                         const int32_t* perm, const int32_t* rowptr_t, const int32_t* src_t,
                         const int32_t* rev_to_fwd, const float* edge_attr,
                         int64_t n, int64_t E, int precision, const float* grad_y,
                         const void* forward_workspace, fesr_param_grads* grads, float* grad_x,
                         void* workspace, size_t workspace_bytes, void* stream);

/* MSE loss + its gradient (torch.nn.MSELoss, models/scheduler_gnn.py:390,406):
 * loss[0] = mean((pred-target)^2) over n*c elements; grad = 2*(pred-target)/(n*c). */
#define FESR_REDUCE_WS_BYTES 8192
int fesr_mse_loss(const float* pred, const float* target, int64_t count,
                  float* loss, float* grad, void* workspace /* FESR_REDUCE_WS_BYTES */, void* stream);

/* One Adam step on a flat fp32 buffer (torch.optim.Adam defaults, scheduler_gnn.py:391):
 * step is the 1-based step count. */
int fesr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t count, float lr, float beta1, float beta2, float eps,
                   int64_t step, void* stream);

/* ------------------------------------------------------------------------------------
 * Node weight.  Replaces GradientbasedLoss.compute_node_weight
 * (models/scheduler_gnn.py:503-514): s = sum_e max_c[((p_src-p_dst) - (y_src-y_dst))/d_e],
 * computed per subdomain of a block-diagonal batch.
 *   node_ptr     : [S+1] int32 node range of every subdomain (NULL: one subdomain = all n nodes)
 *   out          : [S] fp32 (the reference broadcasts it to [n_s]; the host wrapper does that)
 *   node_scratch : [n] fp32 scratch (per-destination partial sums; reduced in a fixed order)
 *   clamp_max    : per-node clamp applied before the sum (GradientbasedLoss.forward, :491-495,
 *                  which scatters by destination); pass +INFINITY for compute_node_weight
 * ---------------------------------------------------------------------------------- */
int fesr_node_weight(const float* pred, const float* target, int32_t channels,
                     const int32_t* rowptr, const int32_t* src_sorted, const int32_t* perm,
                     const float* edge_attr, const int32_t* node_ptr, int32_t n_sub,
                     int64_t n, int64_t E, float clamp_max, float* out, float* node_scratch, void* stream);

/* ------------------------------------------------------------------------------------
 * Overlap stitch.  Replaces AnsysDataset.reconstruct_from_partition's averaging loop
 * (dataset/GraphDataset.py:1370-1400): mean over all subdomain copies of every global node.
 *   fesr_occurrence_build: occ_ptr[N+1], occ_idx[n_tot] from global_ids[n_tot]
 *                          (positions ascending inside every global node)
 *   fesr_stitch_mean:      field[N, c] = mean_j values[occ_idx[j], :]; count[N] = copies;
 *                          merged[n_tot, c] (optional) = field[global_ids] -- the per-copy
 *                          array the reference writes back into the appended grid.
 * ---------------------------------------------------------------------------------- */
size_t fesr_occurrence_workspace_bytes(int64_t n_tot, int64_t N);
int fesr_occurrence_build(const int64_t* global_ids, int64_t n_tot, int64_t N,
                          int32_t* occ_ptr, int32_t* occ_idx,
                          void* workspace, size_t workspace_bytes, void* stream);
int fesr_stitch_mean(const float* values, int32_t channels, const int32_t* occ_ptr,
                     const int32_t* occ_idx, const int64_t* global_ids, int64_t n_tot, int64_t N,
                     float* field, int32_t* count, float* merged, void* stream);

/* ------------------------------------------------------------------------------------
 * Subdomain assembly.  Replaces AnsysDataset._get_partition_domain + vtk_to_pyg
 * (dataset/GraphDataset.py:1183-1306, 838-869; Duct twin :529-642): kd decomposition of the
 * cells into 2^levels regions (halo = cells assigned to every region they touch), then per
 * region node compaction, edge de-duplication, edge lengths and the destination CSR.
 *
 * Two-phase because output sizes are data dependent: *_count writes the totals to
 * host_totals (it synchronises the stream), the caller allocates, *_fill writes the arrays.
 * ---------------------------------------------------------------------------------- */
size_t fesr_partition_workspace_bytes(int64_t C, int32_t levels);
size_t fesr_assign_workspace_bytes(int64_t C, int32_t levels, int64_t total_pairs /* 0 for _count */);
/* home_leaf[C] int32, tree_axis[2^levels-1] int32, tree_split[2^levels-1] fp32 */
int fesr_partition_cells(const float* pos, const int32_t* cells, int64_t N, int64_t C,
                         int32_t levels, int32_t* home_leaf, int32_t* tree_axis, float* tree_split,
                         void* workspace, size_t workspace_bytes, void* stream);
/* leaf_ptr[2^levels+1] int32 is written by _count; host_total = number of (leaf, cell) pairs */
int fesr_assign_count(const float* pos, const int32_t* cells, int64_t C, int32_t levels, int32_t mode,
                      const int32_t* home_leaf, const int32_t* tree_axis, const float* tree_split,
                      int32_t* leaf_ptr, int64_t* host_total, void* workspace, size_t workspace_bytes,
                      void* stream);
/* leaf_cells[total] int32: cells of every leaf, ascending cell id */
int fesr_assign_fill(const float* pos, const int32_t* cells, int64_t C, int32_t levels, int32_t mode,
                     const int32_t* home_leaf, const int32_t* tree_axis, const float* tree_split,
                     const int32_t* leaf_ptr, int64_t total_pairs, int32_t* leaf_cells,
                     void* workspace, size_t workspace_bytes, void* stream);

size_t fesr_subdomain_workspace_bytes(int64_t total_pairs, int64_t N, int32_t n_sub);
/* node_ptr[S+1], edge_ptr[S+1] int32 written; host_totals[0] = sum n_s, [1] = sum E_s.
 * The workspace keeps intermediate state for _fill and must not be touched in between. */
int fesr_subdomain_count(const int32_t* cells, const int32_t* leaf_ptr, const int32_t* leaf_cells,
                         int32_t n_sub, int64_t total_pairs, int64_t N,
                         int32_t* node_ptr, int32_t* edge_ptr, int64_t* host_totals,
                         void* workspace, size_t workspace_bytes, void* stream);
/* global_ids[sum n] int64 ascending per subdomain; edge_src/edge_dst[sum E] int32 batch-level
 * indices in (subdomain, dst, src) order; edge_attr[sum E] fp32 = |pos[src]-pos[dst]|;
 * rowptr[sum n + 1] int32. */
int fesr_subdomain_fill(const float* pos, const int32_t* node_ptr, const int32_t* edge_ptr,
                        int32_t n_sub, int64_t total_pairs, int64_t n_tot, int64_t e_tot,
                        int64_t* global_ids, int32_t* edge_src, int32_t* edge_dst, float* edge_attr,
                        int32_t* rowptr, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * ALDS routing.  Replaces PCAEncoder.get_latent_space + KMeansClassifier.cluster
 * (models/encoder.py:143-157, models/classifier.py:48-50) for a block-diagonal batch:
 * feature row s = x[node_ptr[s] : node_ptr[s]+rows, :] flattened (rows = 280),
 * z = (f - pca_mean) @ comp^T, zs = (z - sc_mean) / sc_scale, label = argmin ||zs - c||^2.
 * All model arrays fp64 (sklearn's dtype), labels int32, latent fp64 [S, n_comp] (optional).
 * ---------------------------------------------------------------------------------- */
int fesr_route(const float* x, int32_t channels, const int32_t* node_ptr, int32_t n_sub, int32_t rows,
               const double* pca_mean, const double* pca_components, int32_t n_comp,
               const double* scaler_mean, const double* scaler_scale,
               const double* centroids, int32_t n_clusters,
               int32_t* labels, double* latent, void* stream);

/* KMeansClassifier.cluster alone (models/classifier.py:48-50) on a latent array [S, n_comp] fp64. */
int fesr_cluster(const double* latent, int32_t n_sub, int32_t n_comp, const double* scaler_mean,
                 const double* scaler_scale, const double* centroids, int32_t n_clusters,
                 int32_t* labels, void* stream);

/* Row selection of the routed predict.  Replaces `x_in = [x[j] for j in idx]` (models/scheduler_gnn.py:240-251:
 * the rows of the subdomains of one cluster, gathered into that cluster's block-diagonal sub-batch) and
 * reorder_predictions (:302-309: the per-cluster results written back in subdomain order).
 *   gather : dst[r, :] = src[index[r], :]      scatter : dst[index[r], :] = src[r, :]      r < rows
 * rows of row_floats fp32 (a multiple of 4: 16-byte rows and pointers), index int64 on the device; the indices of a
 * scatter must be distinct. */
int fesr_gather_rows(const float* src, const int64_t* index, int64_t rows, int32_t row_floats, float* dst, void* stream);
int fesr_scatter_rows(const float* src, const int64_t* index, int64_t rows, int32_t row_floats, float* dst, void* stream);

/* ------------------------------------------------------------------------------------
 * Low-resolution -> high-resolution field transfer (the step BEFORE the path).  Replaces
 * AnsysDataset._lagrangian_interpolation (dataset/GraphDataset.py:1041-1105), i.e.
 * vtkPointInterpolator + vtkGaussianKernel(radius = 3 * mesh_spacing, sharpness = 2):
 *   out[t, :] = sum_j w_j src_val[j, :] / sum_j w_j  over source points with |p_j - q_t| <= radius,
 *   w_j = exp(-(sharpness / radius)^2 |p_j - q_t|^2); no source point in range -> null_value (VTK: 0).
 *   src_pos [n_src, 3], src_val [n_src, channels], dst_pos [n_dst, 3], out [n_dst, channels] fp32;
 *   channels 1, 3 or 4; count [n_dst] int32 (optional) = source points used per target.
 * ---------------------------------------------------------------------------------- */
size_t fesr_interp_workspace_bytes(int64_t n_src, int32_t channels);
int fesr_interp_gaussian(const float* src_pos, const float* src_val, int32_t channels, int64_t n_src,
                         const float* dst_pos, int64_t n_dst, float radius, float sharpness, float null_value,
                         float* out, int32_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Wall shear stress of the (stitched) velocity field -- the step AFTER the path.  Replaces
 * compute_wss.py:5-120 (vtkGradientFilter -> vtkDataSetSurfaceFilter -> vtkPolyDataNormals ->
 * tau_wall = tau - (tau . n) n with tau = mu (grad u + grad u^T) n, :86-99) on a tetrahedral mesh:
 *   fesr_tet_gradient      grad [C, 9]: the constant gradient of every tet, grad[3 i + j] = d u_i / d x_j
 *   fesr_incident_mean     out[i, :] = mean (or, unit_vector = 1, normalised sum) of item_val[item, :] over the
 *                          items incident to node i; the incidence list is fesr_occurrence_build over the
 *                          flattened [items, verts] connectivity (entry j belongs to item occ_idx[j] / verts).
 *                          width 9: point gradients from tet gradients (verts = 4); width 3 + unit_vector: point
 *                          normals from boundary-face normals (verts = 3)
 *   fesr_boundary_faces    tet faces owned by exactly one cell, oriented outward, with unit normals; outputs are
 *                          capacity buffers of 4 C rows, the count goes to host_count (synchronises the stream);
 *                          node ids must fit 21 bits
 *   fesr_wall_shear_stress tau [M, 3], mag [M] at the nodes surf_nodes[M] (NULL: nodes 0..M-1)
 * VTK differences (parity unpinned): outward orientation instead of traversal order, no feature-edge splitting.
 * ---------------------------------------------------------------------------------- */
int fesr_tet_gradient(const float* pos, const int32_t* cells, const float* field, int64_t C, float* grad, void* stream);
int fesr_incident_mean(const float* item_val, int32_t width, const int32_t* occ_ptr, const int32_t* occ_idx,
                       int32_t verts, int64_t N, int32_t unit_vector, float* out, void* stream);
size_t fesr_boundary_faces_workspace_bytes(int64_t C);
int fesr_boundary_faces(const float* pos, const int32_t* cells, int64_t N, int64_t C, int32_t* faces, int32_t* face_cell,
                        float* face_normal, int64_t* host_count, void* workspace, size_t workspace_bytes, void* stream);
int fesr_wall_shear_stress(const float* grad_pt, const float* normal_pt, const int32_t* surf_nodes, int64_t M, float mu,
                           float* tau, float* mag, void* stream);

/* ------------------------------------------------------------------------------------
 * Collectives of the sharded path, on the caller's stream (NCCL over NVLink / NVSwitch; the
 * library resolves libnccl.so.2 at run time -- the copy the host process already uses).
 * Replace the reference's multi-GPU fan-out / fan-in through mp.Process + Manager().dict()
 * (models/scheduler_gnn.py:254-291, results pickled back to the parent) and the bucket
 * all-reduce of DistributedDataParallel (models/scheduler_gnn.py:386).  One communicator per
 * process (= per GPU).
 *   fesr_comm_unique_id   rank 0 fills host_id[128]; the caller ships it to the other ranks
 *                         (any side channel: torch.distributed store, MPI, a file)
 *   fesr_comm_init        collective over all ranks; binds to the current CUDA device
 *   fesr_allgatherv_pred  all-gather of per-rank blocks of different lengths through a padded,
 *                         persistent layout: slots is [world, slot_elems] fp32, rank r has written its
 *                         block (predictions [rows_r, c] | reference rows | subdomain weights ...) at the
 *                         start of slot r; after the call every rank holds every slot.  In place (no send
 *                         copy); the stitch reads the padded buffer through a remapped occurrence index.
 *   fesr_allreduce_grads  flat[i] = mean over ranks of flat[i]   (sum, then * 1/world)
 * ---------------------------------------------------------------------------------- */
int fesr_comm_unique_id(void* host_id /* 128 bytes */);
int fesr_comm_init(const void* host_id /* 128 bytes */, int rank, int world);
int fesr_comm_destroy(void);
int fesr_comm_rank(void);   /* -1 without a communicator */
int fesr_comm_world(void);  /* 0 without a communicator */
int fesr_allgatherv_pred(float* slots, int64_t slot_elems, void* stream);
int fesr_allreduce_grads(float* flat, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FESR_H_ */
