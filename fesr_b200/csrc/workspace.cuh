// Layout of the forward workspace (shared by forward.cu and backward.cu).
#pragma once
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

ForwardWs carve_forward(void* base, const fesr_model_dims& d, int64_t n, int64_t E, int keep);

}  // namespace fesr
