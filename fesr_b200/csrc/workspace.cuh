// Layout of the forward workspace (shared by forward.cu and backward.cu).
#pragma once
#include <stdlib.h>

#include "common.cuh"

#define FESR_MAX_LAYERS 64

namespace fesr {

struct ForwardWs {
  Prepared prep;
  float* g;                        // [E, kp]
  float* h[FESR_MAX_LAYERS + 1];   // keep: h[0..layers]; else ping-pong h[0], h[1]
  float* Z[FESR_MAX_LAYERS];       // keep: one per layer; else Z[0]
  int n_h, n_z;
  size_t bytes;
};

// tf32 arm, training: the Z stash of the forward (read again by the node GEMM and by the weight-gradient GEMM of the
// backward) is stored as fp16 -- the same 10-bit mantissa the tf32 tensor path would round it to, half the HBM bytes
// of the three kernels that touch it and half the stash.  FESR_Z16=0 keeps it fp32 (A/B switch).
inline bool z_stash_half(int precision) {
  static const bool on = !(getenv("FESR_Z16") && atoi(getenv("FESR_Z16")) == 0);
  return on && precision == FESR_PREC_TF32;
}

// z_half: the kept Z stash holds fp16 rows (z_stash_half) -- half the bytes per layer
ForwardWs carve_forward(void* base, const fesr_model_dims& d, int64_t n, int64_t E, int keep, int z_half = 0);

}  // namespace fesr
