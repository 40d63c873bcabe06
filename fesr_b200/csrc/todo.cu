// Entry points declared in include/fesr.h whose kernels are not written yet.  They fail
// loudly (no fallback of any kind); each moves to its own file when implemented.
#include "common.cuh"

using namespace fesr;

#define NOT_BUILT(name)                        \
  do {                                         \
    set_error(name " is not implemented yet"); \
    return FESR_EINVAL;                        \
  } while (0)

extern "C" {

size_t fesr_backward_workspace_bytes(const fesr_model_dims*, int64_t, int64_t) { return 0; }
int fesr_nnconv_backward(const fesr_model_dims*, const fesr_params*, const float*, const int32_t*, const int32_t*,
                         const int32_t*, const int32_t*, const int32_t*, const int32_t*, const float*, int64_t,
                         int64_t, int, const float*, const void*, fesr_param_grads*, float*, void*, size_t, void*) {
  NOT_BUILT("fesr_nnconv_backward");
}

}  // extern "C"
