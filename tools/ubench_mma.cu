// Cost of SMALL tcgen05.mma instructions (the shapes of the window-attention kernels) issued by one thread of one CTA per SM:
//   per-MMA issue time (clock64 around the issue loop) and per-MMA completion time (until the commit's mbarrier flips), for
//   Q K^T pairs (SS, M128 N{64,128,256} K16 x 2 dependent k-steps) and P V chains (TS, M128 N32 K16 x 4 dependent k-steps),
//   with the accumulator chains dependent (one accumulator) or independent (alternating accumulators).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../lavt_rs_b200/csrc/common.cuh"
#include "../lavt_rs_b200/csrc/attn_tc_ptx.cuh"
using namespace lavt;

// mode 0: QK pairs N=n, alternating two S buffers     mode 1: PV 4-chains into ONE O     mode 2: PV chains alternating TWO O accumulators
// mode 3: the tc3 item: PV 4-chain + QK pair (N = 64), one commit per group            mode 4: like 3 but from THREE warps (three issuers)
__global__ void __launch_bounds__(128, 1) k(long long* out, int mode, int n, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], i == 0 ? (mode == 4 ? 3 : 1) : 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  const int nissuers = mode == 4 ? 3 : 1;
  if (warp < nissuers) {
    const uint32_t base = tb + warp * 160;
    const uint64_t dq = make_sw64_desc(smem_u32(smem)), dk = make_sw64_desc(smem_u32(smem) + 8192), dv = make_sw64_desc(smem_u32(smem) + 8192 + 16384);
    const uint32_t idq = make_idesc_bf16_f32(128, n), idp = make_idesc_bf16_f32(128, 32) | (1u << 16);
    __syncwarp();
    const long long t0 = clock64();
    if (elect_one_sync()) {
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) {
          const uint32_t ts = tb + (r & 1) * 256;
          umma_bf16_ss(ts, dq, dk, idq, 0);
          umma_bf16_ss(ts, dq + 2, dk + 2, idq, 1);
        } else if (mode == 1 || mode == 2) {
          const uint32_t to = tb + 448 + ((mode == 2) ? (r & 1) * 32 : 0);
          for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(to, tb + 8 * ks, dv + 64 * ks, idp, ks > 0);
        } else {
          const uint32_t ts = base + (r & 1) * 64;
          for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(base + 128, ts + 8 * ks, dv + 64 * ks, idp, ks > 0);
          umma_commit(&bar[1]);
          umma_bf16_ss(ts, dq, dk, idq, 0);
          umma_bf16_ss(ts, dq + 2, dk + 2, idq, 1);
          umma_commit(&bar[2]);
        }
      }
      umma_commit(&bar[0] + 0);
    }
    __syncwarp();
    const long long t1 = clock64();
    if (warp == 0) {
      mbar_wait(&bar[0], 0);     // one arrival per issuer
      const long long t2 = clock64();
      if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 256;
  struct { int mode, n; const char* what; int mmas; } cases[] = {
      {0, 64, "QK pair N=64 (2 dependent k-steps)", 2}, {0, 128, "QK pair N=128", 2}, {0, 256, "QK pair N=256", 2},
      {1, 64, "PV 4-chain N=32, one accumulator", 4}, {2, 64, "PV 4-chain N=32, two accumulators alternating", 4},
      {3, 64, "tc3 item: PV 4-chain + commit + QK pair + commit", 6}, {4, 64, "tc3 item from three issuer warps (per issuer)", 6}};
  for (auto& c : cases) {
    k<<<148, 128, 64 * 1024>>>(out, c.mode, c.n, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; } a /= 148; b /= 148;
    fflush(stdout); printf("%-52s issue %7.1f cycles / group (%5.1f / MMA)   complete %7.1f cycles / group (%5.1f / MMA)\n", c.what, a / reps, a / reps / c.mmas,
           b / reps, b / reps / c.mmas);
  }
  return 0;
}
