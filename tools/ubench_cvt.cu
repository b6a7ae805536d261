// Throughput of cvt.rn.bf16x2.f32 (F2FP) alone and next to MUFU.EX2:  nvcc -arch=sm_100a -O3 -o tools/_bin/ubench_cvt tools/ubench_cvt.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
  uint32_t acc[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { acc[i / 2] ^= cvt2(v[i], v[i + 1]); v[i] += 1.f; }                       // 1 cvt + 1 FADD + 1 LOP per pair
      else if (MODE == 1) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); acc[i / 2] ^= cvt2(v[i], v[i + 1]); }  // kernel ratio: 2 MUFU : 1 cvt
      else if (MODE == 2) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); acc[i / 2] ^= prmt(__float_as_uint(v[i]) + 0x8000u, __float_as_uint(v[i + 1]) + 0x8000u); }
      else { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); acc[i / 2] ^= __float_as_uint(v[i]); }  // MUFU only reference with the same glue
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(acc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[4] = {"cvt.rn.bf16x2 only (pairs/clk/SM)", "2 MUFU + 1 cvt per pair (pairs/clk/SM)", "2 MUFU + 2 IADD + PRMT per pair (pairs/clk/SM)", "2 MUFU per pair (pairs/clk/SM)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int threads : {128, 384, 512}) {
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
      else if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
      else if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
      else k<3><<<148, threads>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("mode %d %-48s threads %4d: %.2f\n", mode, names[mode], threads, 8.0 * iters * threads / c);
    }
  return 0;
}
