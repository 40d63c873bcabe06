// libfesr.so: version / error / device / model geometry entry points.
#include <stdarg.h>

#include "common.cuh"

namespace fesr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace fesr

extern "C" {

int fesr_version(void) { return FESR_VERSION; }

const char* fesr_last_error(void) { return fesr::g_err; }

int fesr_device_check(void) {
  int dev = 0;
  FESR_CUDA(cudaGetDevice(&dev));
  int major = 0;
  FESR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    fesr::set_error("libfesr.so is built for sm_100a only; device %d has compute capability %d.x", dev, major);
    return FESR_EDEVICE;
  }
  return FESR_OK;
}

int fesr_model_dims_init(int kind, int w, int in_ch, int out_ch, int layers, fesr_model_dims* d) {
  FESR_CHECK_ARG(d != nullptr, "dims is NULL");
  FESR_CHECK_ARG(kind == FESR_KERNELNN || kind == FESR_TEECNET, "unknown model kind %d", kind);
  FESR_CHECK_ARG(in_ch >= 1 && in_ch <= 16 && out_ch >= 1 && out_ch <= 16, "in/out channels must be in [1,16]");
  FESR_CHECK_ARG(layers >= 1 && layers <= 64, "layers must be in [1,64]");
  const int need = (kind == FESR_TEECNET) ? w + 1 : w;   // TEECNet keeps a constant-1 column
  FESR_CHECK_ARG(w >= 1 && need <= 64, "width %d unsupported (padded width must be <= 64)", w);
  memset(d, 0, sizeof(*d));
  d->kind = kind;
  d->w = w;
  d->wp = (need + 15) / 16 * 16;
  d->in_ch = in_ch;
  d->out_ch = out_ch;
  d->layers = layers;
  if (kind == FESR_KERNELNN) {          // DenseNet([1, w, w, w*w], ReLU)   models/model.py:550
    d->n_hidden = 2;
    d->hidden[0] = w;
    d->hidden[1] = w;
    d->leaky = 0;
  } else {                              // DenseNet([1, 32, 64, 128, w*w], LeakyReLU)  models/model.py:403
    d->n_hidden = 3;
    d->hidden[0] = 32;
    d->hidden[1] = 64;
    d->hidden[2] = 128;
    d->leaky = 1;
  }
  d->k1 = d->hidden[d->n_hidden - 1] + 1;
  // pick (kt, passes): minimise padded channels 4*kt*passes, then passes
  const int cand[4] = {4, 8, 11, 13};
  int best_kt = 13, best_p = (d->k1 + 51) / 52;
  for (int c = 0; c < 4; ++c) {
    int p = (d->k1 + 4 * cand[c] - 1) / (4 * cand[c]);
    int tot = 4 * cand[c] * p, btot = 4 * best_kt * best_p;
    if (tot < btot || (tot == btot && p < best_p)) {
      best_kt = cand[c];
      best_p = p;
    }
  }
  d->kt = best_kt;
  d->passes = best_p;
  d->ktp = (d->kt + 3) / 4 * 4;
  d->kp = d->passes * 4 * d->ktp;
  d->k1p = d->passes * 4 * d->kt;
  d->zk_main = d->k1p * d->wp;
  d->zk = (d->zk_main + d->wp + 31) / 32 * 32;
  return FESR_OK;
}

}  // extern "C"
