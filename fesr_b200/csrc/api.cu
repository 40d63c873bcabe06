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

static thread_local const float* g_cur_gcenter = nullptr;
const float* cur_gcenter() { return g_cur_gcenter; }
void set_cur_gcenter(const float* gc) { g_cur_gcenter = gc; }

static thread_local int* g_cur_ovf = nullptr;
int* cur_ovf() { return g_cur_ovf; }
void set_cur_ovf(int* flag) { g_cur_ovf = flag; }

static long long g_launches = 0;
void count_launch() { ++g_launches; }

// ---- profiling: pairs of events per kernel class, recorded on the launching stream
static bool g_prof_on = false;
struct ProfPair { cudaEvent_t a, b; int kind; };
static ProfPair g_pairs[8192];
static int g_npairs = 0, g_nalloc = 0;

ProfScope::ProfScope(int kind, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on || g_npairs >= 8192) return;
  if (g_npairs >= g_nalloc) {
    if (cudaEventCreate(&g_pairs[g_nalloc].a) != cudaSuccess || cudaEventCreate(&g_pairs[g_nalloc].b) != cudaSuccess) return;
    ++g_nalloc;
  }
  slot = g_npairs++;
  g_pairs[slot].kind = kind;
  cudaEventRecord(g_pairs[slot].a, stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_pairs[slot].b, stream);
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

long long fesr_launch_count(void) { return fesr::g_launches; }

int fesr_profile_enable(int on) {
  fesr::g_prof_on = on != 0;
  fesr::g_npairs = 0;
  return FESR_OK;
}

int fesr_profile_collect(double* ms_by_kind, long long* launches_by_kind, int nkinds) {
  FESR_CHECK_ARG(ms_by_kind && launches_by_kind && nkinds >= fesr::PROF_NKINDS, "need %d kinds", (int)fesr::PROF_NKINDS);
  for (int k = 0; k < nkinds; ++k) {
    ms_by_kind[k] = 0.0;
    launches_by_kind[k] = 0;
  }
  for (int i = 0; i < fesr::g_npairs; ++i) {
    FESR_CUDA(cudaEventSynchronize(fesr::g_pairs[i].b));
    float ms = 0.f;
    FESR_CUDA(cudaEventElapsedTime(&ms, fesr::g_pairs[i].a, fesr::g_pairs[i].b));
    ms_by_kind[fesr::g_pairs[i].kind] += ms;
    launches_by_kind[fesr::g_pairs[i].kind] += 1;
  }
  fesr::g_npairs = 0;
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
  d->zk = (d->zk_main + d->wp + 63) / 64 * 64;   // 64: one 128-byte TMA row of 16-bit elements
  return FESR_OK;
}

}  // extern "C"
