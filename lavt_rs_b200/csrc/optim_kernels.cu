// Multi-tensor AdamW step (reference train.py:688-692: torch.optim.AdamW over the backbone-no-decay / backbone / classifier / text
// encoder parameter groups): one launch updates every tensor of a group.  Block b finds its tensor by binary search over a
// prefix table of block counts; decoupled weight decay, bias correction and the optional AMSGrad maximum follow torch.optim.AdamW:
//     p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(max v) / sqrt(bc2) + eps)
#include "../../include/lavt_b200.h"
#include "kernels.cuh"

namespace lavt {

constexpr int ADAMW_CHUNK = 4096;      // elements per block: 256 threads x 4 x float4

__global__ void __launch_bounds__(256) adamw_kernel(const lavt_adamw_tensor_t* __restrict__ table, const int* __restrict__ prefix,
                                                    int n_tensors, float lr, float b1, float b2, float omb1, float omb2, float eps,
                                                    float wd) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_tensors - 1;
    const int b = blockIdx.x;
    while (lo < hi) {                     // last tensor whose first block is <= b
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= b) lo = mid; else hi = mid - 1;
    }
    s_t = lo;
  }
  __syncthreads();
  const lavt_adamw_tensor_t tt = table[s_t];
  const long long base = static_cast<long long>(blockIdx.x - prefix[s_t]) * ADAMW_CHUNK;
  const float decay = 1.0f - lr * wd;
  const float step = lr / tt.bc1;
  const float inv_sqrt_bc2 = rsqrtf(tt.bc2);
  const bool vec = ((reinterpret_cast<uintptr_t>(tt.p) | reinterpret_cast<uintptr_t>(tt.g) | reinterpret_cast<uintptr_t>(tt.m) |
                     reinterpret_cast<uintptr_t>(tt.v) | reinterpret_cast<uintptr_t>(tt.vmax)) & 15) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + (k * 256 + threadIdx.x) * 4;
    if (i >= tt.n) break;
    float p[4], g[4], m[4], v[4], vm[4];
    const int cnt = (tt.n - i >= 4) ? 4 : static_cast<int>(tt.n - i);
    if (vec && cnt == 4) {
      const float4 a = *reinterpret_cast<const float4*>(tt.p + i), b = *reinterpret_cast<const float4*>(tt.g + i);
      const float4 c = *reinterpret_cast<const float4*>(tt.m + i), d = *reinterpret_cast<const float4*>(tt.v + i);
      p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; g[0] = b.x; g[1] = b.y; g[2] = b.z; g[3] = b.w;
      m[0] = c.x; m[1] = c.y; m[2] = c.z; m[3] = c.w; v[0] = d.x; v[1] = d.y; v[2] = d.z; v[3] = d.w;
      if (tt.vmax) { const float4 e = *reinterpret_cast<const float4*>(tt.vmax + i); vm[0] = e.x; vm[1] = e.y; vm[2] = e.z; vm[3] = e.w; }
    } else {
      for (int j = 0; j < cnt; ++j) { p[j] = tt.p[i + j]; g[j] = tt.g[i + j]; m[j] = tt.m[i + j]; v[j] = tt.v[i + j]; if (tt.vmax) vm[j] = tt.vmax[i + j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= cnt) break;
      p[j] *= decay;
      m[j] = b1 * m[j] + omb1 * g[j];
      v[j] = b2 * v[j] + omb2 * g[j] * g[j];
      float vv = v[j];
      if (tt.vmax) { vm[j] = fmaxf(vm[j], v[j]); vv = vm[j]; }
      p[j] -= step * m[j] / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    }
    if (vec && cnt == 4) {
      *reinterpret_cast<float4*>(tt.p + i) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(tt.m + i) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(tt.v + i) = make_float4(v[0], v[1], v[2], v[3]);
      if (tt.vmax) *reinterpret_cast<float4*>(tt.vmax + i) = make_float4(vm[0], vm[1], vm[2], vm[3]);
    } else {
      for (int j = 0; j < cnt; ++j) { tt.p[i + j] = p[j]; tt.m[i + j] = m[j]; tt.v[i + j] = v[j]; if (tt.vmax) tt.vmax[i + j] = vm[j]; }
    }
  }
}

}  // namespace lavt

extern "C" int lavt_adamw_chunk_elems(void) { return lavt::ADAMW_CHUNK; }

extern "C" int lavt_adamw_step(const lavt_adamw_tensor_t* table_dev, const int32_t* block_prefix_dev, int32_t n_tensors, int32_t n_blocks,
                               float lr, double beta1_d, double beta2_d, float eps, float weight_decay, void* stream) {
  using namespace lavt;
  const float beta1 = static_cast<float>(beta1_d), beta2 = static_cast<float>(beta2_d);
  LAVT_REQUIRE(table_dev && block_prefix_dev && n_tensors > 0 && n_blocks > 0, "adamw: empty group");
  // 1 - beta is formed in double and rounded once, as torch does (1.0f - 0.999f is off by 1.3e-5 relative)
  adamw_kernel<<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, block_prefix_dev, n_tensors, lr, beta1, beta2,
                                                                          static_cast<float>(1.0 - static_cast<double>(beta1_d)),
                                                                          static_cast<float>(1.0 - static_cast<double>(beta2_d)), eps, weight_decay);
  LAVT_LAUNCH_CHECK("adamw_kernel");
  return LAVT_OK;
}
