// tools/dev: probe of TMA tile::gather4 on sm_100a (layout of the 4 gathered rows in shared memory for a 96-byte row,
// per swizzle mode and box shape) and its issue cost against 16-byte cp.async gathers of the same rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather4_probe tools/dev/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e__), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

__global__ void layout_kernel(const __grid_constant__ CUtensorMap tm, int bytes, uint16_t* out) {
  __shared__ __align__(1024) uint16_t buf[1024];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = 0xffff;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    gather4(smem_u32(buf), &tm, smem_u32(&bar), 0, 5, 77, 300, 1023);
  }
  mbar_wait(smem_u32(&bar), 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}

// ---- issue-cost comparison: NW producer warps per CTA, each stages whole segments of CAP rows into its own two slots
constexpr int CAP = 112;
constexpr int NW = 3;
template <int MODE>   // 0: gather4, 1: cp.async 16 B
__global__ void __launch_bounds__(NW * 32, 1) cost_kernel(const __grid_constant__ CUtensorMap tm, const uint4* h16, const int* src, int n_seg_per_cta,
                                                          long long* cycles, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);              // [NW * 2]
  uint8_t* slots = sm + 1024;                                    // [NW * 2][CAP * 96 (+pad to 128)]
  constexpr int SLOT = (CAP * 96 + 127) / 128 * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(MODE == 0 ? 1 : 32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  unsigned acc = 0;
  int k = 0;
  for (int s = warp; s < n_seg_per_cta; s += NW, ++k) {
    const int slot = warp * 2 + (k & 1);
    const uint32_t bar = smem_u32(&bars[slot]);
    const uint32_t dst = smem_u32(slots + (size_t)slot * SLOT);
    if (k >= 2) {
      mbar_wait(bar, (uint32_t)(((k - 2) >> 1) & 1));     // the slot's previous fill has landed ("consumed" at once)
      acc += *reinterpret_cast<volatile unsigned*>(slots + (size_t)slot * SLOT + lane * 4);
    }
    const int* sp = src + ((size_t)blockIdx.x * n_seg_per_cta + s) * CAP;
    if (MODE == 0) {
      if (lane < CAP / 4) {
        const int4 r = *reinterpret_cast<const int4*>(sp + 4 * lane);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CAP * 96) : "memory");
        __syncwarp(0x0fffffff);
        gather4(dst + lane * 384, &tm, bar, 0, r.x, r.y, r.z, r.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < CAP * 6 / 32; ++i) {
        const int t = i * 32 + lane;
        const int row = t / 6, c = t % 6;
        const int sidx = __ldg(sp + row);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(row * 96 + ((c ^ ((row >> 2) & 1)) << 4))), "l"(h16 + (size_t)sidx * 6 + c) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
    }
  }
  if (MODE == 1) asm volatile("cp.async.wait_all;" ::: "memory");
  // drain: wait for the last fills of both slots
  for (int q = 0; q < 2; ++q) {
    const int kk = k - 1 - q;
    if (kk >= 0) mbar_wait(smem_u32(&bars[warp * 2 + (kk & 1)]), (uint32_t)((kk >> 1) & 1));
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * NW + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  encode_fn_t encode = (encode_fn_t)fn;
  const int n = 160000, w = 48;
  std::vector<uint16_t> h((size_t)n * w);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < w; ++c) h[(size_t)r * w + c] = (uint16_t)(((r & 1023) << 6) | c);
  uint16_t* dh;
  CK(cudaMalloc(&dh, h.size() * 2));
  CK(cudaMemcpy(dh, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  uint16_t* dout;
  CK(cudaMalloc(&dout, 2048));
  const CUtensorMapSwizzle sw[3] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  const char* swn[3] = {"none", "64B", "128B"};
  CUtensorMap good;
  bool have_good = false;
  for (int bo = 1; bo <= 1; bo += 3)
    for (int s = 0; s < 3; ++s) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)n};
      cuuint64_t strides[1] = {(cuuint64_t)w * 2};
      cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)bo};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dh, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw[s],
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== box {48,%d} swizzle %s: encode rc %d\n", bo, swn[s], (int)r);
      if (r != CUDA_SUCCESS) continue;
      layout_kernel<<<1, 128>>>(tm, 4 * 96, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("   kernel error: %s\n", cudaGetErrorString(e));
        return 2;      // context is gone
      }
      std::vector<uint16_t> o(1024);
      CK(cudaMemcpy(o.data(), dout, 2048, cudaMemcpyDeviceToHost));
      // print per 16-byte chunk: (row, first column) or '.' for untouched
      for (int ch = 0; ch < 48; ++ch) {
        const uint16_t v = o[ch * 8];
        if (v == 0xffff) printf(" ....");
        else printf(" %4d:%02d", v >> 6, v & 63);
        if (ch % 8 == 7) printf("\n");
      }
      if (bo == 1 && s == 0) { good = tm; have_good = true; }
    }
  if (!have_good) return 0;
  // ---- cost
  const int nseg = 300, grid = 148;
  std::vector<int> src((size_t)grid * nseg * CAP);
  uint32_t st = 12345;
  for (int b = 0; b < grid; ++b)
    for (int s = 0; s < nseg; ++s)
      for (int e = 0; e < CAP; ++e) {
        st = st * 1664525u + 1013904223u;
        const int centre = (int)(((size_t)b * nseg + s) * 16 % (n - 4000)) + 2000;     // 16 nodes per segment pair, local neighbourhood
        src[((size_t)b * nseg + s) * CAP + e] = centre + (int)((st >> 8) % 1500) - 750;
      }
  int* dsrc;
  CK(cudaMalloc(&dsrc, src.size() * 4));
  CK(cudaMemcpy(dsrc, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
  long long* dcy;
  CK(cudaMalloc(&dcy, grid * NW * 8));
  unsigned* dsink;
  CK(cudaMalloc(&dsink, 4));
  const size_t smem = 1024 + (size_t)NW * 2 * ((CAP * 96 + 127) / 128 * 128);
  CK(cudaFuncSetAttribute(cost_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(cost_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) cost_kernel<0><<<grid, NW * 32, smem>>>(good, (const uint4*)dh, dsrc, nseg, dcy, dsink);
      else cost_kernel<1><<<grid, NW * 32, smem>>>(good, (const uint4*)dh, dsrc, nseg, dcy, dsink);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cost kernel mode %d error: %s\n", mode, cudaGetErrorString(e)); return 3; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cy(grid * NW);
      CK(cudaMemcpy(cy.data(), dcy, cy.size() * 8, cudaMemcpyDeviceToHost));
      double mean = 0;
      for (auto c : cy) mean += (double)c;
      mean /= cy.size();
      printf("mode %s rep %d: %.3f ms, %.0f cycles per warp for %d own segments -> %.0f cycles per segment per warp, %.0f per segment per CTA (3 warps)\n",
             mode == 0 ? "gather4" : "cp.async16", rep, ms, mean, nseg / NW, mean / (nseg / NW), mean / nseg);
    }
  return 0;
}
