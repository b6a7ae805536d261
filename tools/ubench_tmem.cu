// Micro-benchmarks that size the tcgen05 window-attention design (tools only, not product code):
//   A  tcgen05.ld 32x32b.x32 throughput per SM with 4 / 8 warps
//   B  tcgen05.st 32x32b.x32 throughput per SM
//   C  two-pass TMEM softmax inner loop (ld + bias LDS + FADD2 + FMNMX3 + st ; ld + FADD2 + EX2 + FADD2 + cvt + st)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../lavt_rs_b200/csrc/common.cuh"

using namespace lavt;

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int MODE>   // 0 = ld, 1 = st, 2 = softmax two-pass
__global__ void __launch_bounds__(512, 1) bench_kernel(long long* cycles, float* sink, int iters, int ncols) {
  __shared__ uint32_t tmem_ptr;
  __shared__ float tab[4096];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = 0.001f * (i & 63);
  if (warp == 0) tmem_alloc(&tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // zero TMEM so the math runs on finite values
  {
    uint32_t z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = 0;
    if (warp < 4)
      for (int c = 0; c < 512; c += 32) tmem_st_32x32b_x32(tb + c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // warpgroup g handles columns [g*half, g*half + half)
  const int wg = warp >> 2;
  const int nwg = blockDim.x >> 7;
  const int half = ncols / nwg;
  const int c0 = wg * half;
  float acc = 0.f;
  const float* tq = tab + 2048 + (threadIdx.x & 31) * 5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      for (int c = 0; c < half; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tb + c0 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) acc += __uint_as_float(v[j]);
      }
    } else if (MODE == 1) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(acc + j);
      for (int c = 0; c < half; c += 32) tmem_st_32x32b_x32(tb + c0 + c, v);
      tmem_st_wait();
      acc += 1.f;
    } else {
      // pass 1: s += bias ; row max ; write back
      float m0 = -1e30f, m1 = -1e30f;
      for (int c = 0; c < half; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tb + c0 + c, v);
        tmem_ld_wait();
        const float* tp = tq - c - (it & 63);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
          add2(a0, a1, tp[-j], tp[-j - 1]);
          if (j & 2) m1 = max3(m1, a0, a1); else m0 = max3(m0, a0, a1);
          v[j] = __float_as_uint(a0);
          v[j + 1] = __float_as_uint(a1);
        }
        tmem_st_32x32b_x32(tb + c0 + c, v);
      }
      tmem_st_wait();
      const float nm = -fmaxf(m0, m1);
      // pass 2: p = exp2(s - m) ; sum ; pack bf16 ; write P over the first half of the S columns
      float l0 = 0.f, l1 = 0.f;
      for (int c = 0; c < half; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tb + c0 + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
          add2(a0, a1, nm, nm);
          a0 = ex2f(a0);
          a1 = ex2f(a1);
          add2(l0, l1, a0, a1);
          pk[j >> 1] = pack_bf16x2(a0, a1);
        }
        tmem_st_32x32b_x16(tb + c0 + (c >> 1), pk);
      }
      tmem_st_wait();
      acc += l0 + l1;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_ptr, 512);
  }
}

template <int MODE>
static void run(const char* name, int threads, int ncols, int iters) {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * sizeof(float));
  bench_kernel<MODE><<<148, threads>>>(cyc, sink, 2, ncols);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench_kernel<MODE><<<148, threads>>>(cyc, sink, iters, ncols);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double elems = 128.0 * ncols * iters;   // fp32 elements per SM
  printf("%-28s threads=%3d cols=%3d iters=%d  err=%d  %.1f cyc/iter  %.2f elem/clk/SM  %.1f B/clk/SM  (%.3f ms)\n", name,
         threads, ncols, iters, (int)err, avg / iters, elems / avg, 4.0 * elems / avg, ms);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  run<0>("tcgen05.ld x32", 128, 512, 2000);
  run<0>("tcgen05.ld x32", 256, 512, 2000);
  run<1>("tcgen05.st x32", 128, 512, 2000);
  run<1>("tcgen05.st x32", 256, 512, 2000);
  run<2>("softmax 2-pass", 128, 384, 2000);
  run<2>("softmax 2-pass", 256, 384, 2000);
  run<2>("softmax 2-pass", 256, 448, 2000);
  run<2>("softmax 2-pass", 384, 384, 2000);
  run<2>("softmax 2-pass", 512, 384, 2000);
  run<2>("softmax 2-pass", 512, 512, 2000);
  return 0;
}
