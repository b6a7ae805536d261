// Throughput of MUFU.EX2 and of an FMA-pipe polynomial exp2 per SM on this GPU:  nvcc -arch=sm_100a -O3 -o tools/_bin/ubench_mufu tools/ubench_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      else if (MODE == 1) {   // Cody-Waite + degree-3 polynomial on the FMA pipe
        float x = fmaxf(v[i], -126.f);
        float t = x + 12582912.f;
        float xf = x - (t - 12582912.f);
        float p = fmaf(fmaf(fmaf(0.0555041f, xf, 0.2402265f), xf, 0.6931472f), xf, 1.0f);
        v[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.5f;
      } else {               // half MUFU, half polynomial
        if (i & 1) v[i] = ex2(v[i]);
        else {
          float x = fmaxf(v[i], -126.f);
          float t = x + 12582912.f;
          float xf = x - (t - 12582912.f);
          float p = fmaf(fmaf(fmaf(0.0555041f, xf, 0.2402265f), xf, 0.6931472f), xf, 1.0f);
          v[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.5f;
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
      else if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
      else k<2><<<148, threads>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("mode %d (%s) threads %4d: %.2f exp2 / clk / SM\n", mode, mode == 0 ? "MUFU.EX2" : mode == 1 ? "FMA polynomial" : "half/half", threads,
             16.0 * iters * threads / c);
    }
  return 0;
}
