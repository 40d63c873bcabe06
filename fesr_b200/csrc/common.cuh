// Shared helpers for libfesr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fesr.h"

namespace fesr {

void set_error(const char* fmt, ...);

#define FESR_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::fesr::set_error(__VA_ARGS__);                              \
      return FESR_EINVAL;                                          \
    }                                                              \
  } while (0)

#define FESR_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      ::fesr::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return FESR_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

void count_launch();

// fp16 range guard of the f16 arms.  Every kernel that packs fp32 results into fp16 notes the magnitudes it packs
// and, if any exceeds the largest finite fp16 (65504) or is not a number, raises the forward's overflow flag (one
// int32 in the workspace, zeroed at the start of the pass).  fc_out then poisons the output with NaN and the
// host wrapper raises: a clipped intermediate can never come back as a plausible field.
// The flag of the pass being issued is a host-side thread-local (cur_ovf) so that launch wrappers shared with the
// backward (where it is NULL: no check) need no extra parameter.
int* cur_ovf();
void set_cur_ovf(int* flag);
struct OvfScope {
  explicit OvfScope(int* flag) { set_cur_ovf(flag); }
  ~OvfScope() { set_cur_ovf(nullptr); }
};
// Centring of the edge features in the fused f16 predict arm: the planar g rows the fused
// layer kernel consumes hold g_e - g(0) (and the constant FESR_LO_SCALE in one padding slot); cur_gcenter() is the
// per-slot vector the edge-MLP kernels subtract (NULL: plain g), set like cur_ovf by the pass being issued.
const float* cur_gcenter();
void set_cur_gcenter(const float* gc);
struct GcenterScope {
  explicit GcenterScope(const float* gc) { set_cur_gcenter(gc); }
  ~GcenterScope() { set_cur_gcenter(nullptr); }
};
#define FESR_LO_SCALE 0.00390625f      /* 2^-8: scale of the low-order weight terms carried by a padding slot */
#ifdef __CUDACC__
struct F16Guard {
  unsigned m;
  __device__ __forceinline__ F16Guard() : m(0u) {}
  __device__ __forceinline__ void note(float a, float b) {
    m = max(m, max(__float_as_uint(a) & 0x7fffffffu, __float_as_uint(b) & 0x7fffffffu));
  }
  __device__ __forceinline__ void flush(int* flag) const {
    if (flag != nullptr && m > 0x477fe000u) *flag = 1;      // 0x477fe000 = 65504.0f; inf / NaN bit patterns are larger
  }
};
#endif
#define FESR_LAUNCH_CHECK()          \
  do {                               \
    ::fesr::count_launch();          \
    FESR_CUDA(cudaGetLastError());   \
  } while (0)

// Optional per-kernel-class CUDA-event timing (bench.py's roofline numbers); off by default.
enum ProfKind {
  PROF_PREPARE = 0, PROF_EDGE_HIDDEN, PROF_FC_IN, PROF_ZBUILD, PROF_NODE_GEMM, PROF_FC_OUT,
  PROF_NODE_WEIGHT, PROF_STITCH, PROF_GRAPH, PROF_BACKWARD, PROF_LAYER_FUSED, PROF_NKINDS
};
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(int kind, cudaStream_t s);
  ~ProfScope();
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();

// ---- prepared (padded / permuted) weights living at the head of the forward workspace ----
struct Prepared {
  float* tprime;     // [zk, wp]  row-major: T'[(k,a), b], root rows appended
  float* tprime_t;   // [wp, zk]  K-major copy for the tensor-core path (tf32-rounded hi part)
  float* tprime_t_lo;// [wp, zk]  lo part for TF32X3
  float* ttilde;     // [zk, wp]  T~[(k,b), a] = T'[(k,a), b] (every wp x wp block transposed) -- backward
  float* ttilde_t;   // [wp, zk]  K-major tf32 copy of T~
  float* ttilde_t_lo;// [wp, zk]  lo part of T~ for the 3xTF32 dh product of the fp32 backward
  void* tprime_t_h;  // [wp, zk]  K-major fp16 copy of T'   (FESR_PREC_F16)
  void* tprime_t_h_lo;  // [wp, zk]  fp16(T' - fp16(T')): low-order term for the two-term predict GEMM
  void* ttilde_t_h;  // [wp, zk]  K-major fp16 copy of T~
  void* tfused_h;    // [parts, 48, 832] fp16 T' in the fused layer kernel's K order (layer_fused.cu), or NULL
  float* bias_p;     // [wp]
  float* fc1_wp;     // [in_ch, wp] transposed + padded
  float* fc1_bp;     // [wp] (TEECNet: constant-1 column set here)
  int* ovf;          // [1] fp16 overflow flag of the pass (F16Guard)
  float* gcenter;    // [kp] slot layout: g(0) of every channel slot (0 for the constant-1 slot), -FESR_LO_SCALE in the
                     //      lo slot (first padding slot), 0 elsewhere -- subtracted from g by the fused arm's edge MLP
  float* mfull;      // [wp, wp] fp32: T'[(K, a), b] + sum_k g(0)[k] T'[(k, a), b], the weights of the constant-1 slot
                     //      once g is centred
};

}  // namespace fesr
