// tools/dev: issue rate of mma.sync.m16n8k16 (f16 and f32 accumulate) on sm_100a, per SM sub-partition, for 1..4 warps
// per sub-partition with 18 independent accumulator sets each (the consumer loop of the fused layer kernel).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
template <int F32ACC>
__global__ void k(long long* out, uint32_t* sink, int iters) {
  uint32_t a[3][4], b[6][2];
  for (int i = 0; i < 3; ++i) for (int q = 0; q < 4; ++q) a[i][q] = threadIdx.x * 7 + i + q;
  for (int i = 0; i < 6; ++i) for (int q = 0; q < 2; ++q) b[i][q] = threadIdx.x * 3 + i + q;
  uint32_t c[18][2] = {};
  float f[18][4] = {};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int n = 0; n < 6; ++n) {
        if (F32ACC)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(f[m * 6 + n][0]), "+f"(f[m * 6 + n][1]), "+f"(f[m * 6 + n][2]), "+f"(f[m * 6 + n][3])
                       : "r"(a[m][0]), "r"(a[m][1]), "r"(a[m][2]), "r"(a[m][3]), "r"(b[n][0]), "r"(b[n][1]));
        else
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                       : "+r"(c[m * 6 + n][0]), "+r"(c[m * 6 + n][1])
                       : "r"(a[m][0]), "r"(a[m][1]), "r"(a[m][2]), "r"(a[m][3]), "r"(b[n][0]), "r"(b[n][1]));
      }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 18; ++i) s += c[i][0] + c[i][1] + __float_as_uint(f[i][0] + f[i][1] + f[i][2] + f[i][3]);
  if (s == 0x12345) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int acc = 0; acc < 2; ++acc)
    for (int warps = 4; warps <= 16; warps *= 2) {
      if (acc) k<1><<<148, warps * 32>>>(d, s, iters); else k<0><<<148, warps * 32>>>(d, s, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = (double)h[0] / iters / 18;
      printf("%s accumulate, %2d warps/SM (%d per sub-partition): %.2f cycles per HMMA per warp, %.2f per HMMA per sub-partition -> %.0f FMA/clk/SM\n",
             acc ? "f32" : "f16", warps, warps / 4, cyc, cyc / (warps / 4), 2048.0 * 4 / (cyc / (warps / 4)));
    }
  return 0;
}
