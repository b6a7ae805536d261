// PWAM pixel-word attention core (reference lib/video_swin_transformer.py:975-1009, 2-D twin lib/backbone.py:1329-1372).
//
//   pwam_kv    k, v = (W l + b) * l_mask for the <= 77 words of a clip (768 -> C), fp32       (:985-988)
//   pwam_core  per pixel: q = InstanceNorm(q_pre); S = C^-0.5 q k^T + (1e4 m - 1e4); softmax over words;
//              o = P v                                                                           (:990-1003)
// The word side is tiny (Nl x C); the pixel side is one pass over q_pre with k, v served from L1/L2.
#include "kernels.cuh"

namespace lavt {

// grid (Nl, B, ceil(2C / 64)), block 256: smem holds the word vector l[b, :, j]; each warp produces 8 of the block's 64
// outputs (k channels first, then v channels)
__global__ void __launch_bounds__(256) pwam_kv_kernel(const float* __restrict__ l, const float* __restrict__ mask,
                                                      const float* __restrict__ wk, const float* __restrict__ bk,
                                                      const float* __restrict__ wv, const float* __restrict__ bv,
                                                      float* __restrict__ k, float* __restrict__ v, int Nl, int Lin, int C) {
  extern __shared__ float lv[];
  const int j = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < Lin; i += blockDim.x) lv[i] = l[(static_cast<long long>(b) * Lin + i) * Nl + j];
  __syncthreads();
  const float m = mask[b * Nl + j];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int cbeg = blockIdx.z * 64;
  const int cend = min(2 * C, cbeg + 64);
  for (int c = cbeg + warp; c < cend; c += nw) {
    const bool is_v = c >= C;
    const int cc = is_v ? c - C : c;
    const float* w = (is_v ? wv : wk) + static_cast<long long>(cc) * Lin;
    float acc = 0.f;
    for (int i = lane; i < Lin; i += 32) acc = fmaf(__ldg(w + i), lv[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float bias = is_v ? bv[cc] : bk[cc];
      (is_v ? v : k)[(static_cast<long long>(b) * Nl + j) * C + cc] = (acc + bias) * m;
    }
  }
}

int pwam_kv_dispatch(const float* l, const float* mask, const float* wk, const float* bk, const float* wv, const float* bv,
                     float* k, float* v, int B, int Nl, int Lin, int C, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Lin > 0 && C > 0, "pwam_kv: empty input");
  LAVT_REQUIRE(Lin * sizeof(float) <= 48 * 1024, "pwam_kv: language width %d too large", Lin);
  pwam_kv_kernel<<<dim3(Nl, B, (2 * C + 63) / 64), 256, Lin * sizeof(float), st>>>(l, mask, wk, bk, wv, bv, k, v, Nl, Lin, C);
  LAVT_LAUNCH_CHECK("pwam_kv_kernel");
  return LAVT_OK;
}

// One warp handles PIX pixels at a time; lane owns channels [lane*VPL*? ...] laid out so that each head's channels
// sit in a contiguous group of 32/heads lanes.  CPL = channels per lane = C / 32.
constexpr int PWAM_MAX_NL = 80;

template <int CPL, int PIX>
__global__ void __launch_bounds__(256) pwam_core_kernel(const float* __restrict__ qpre, const float* __restrict__ stats,
                                                        const float* __restrict__ k, const float* __restrict__ v,
                                                        const float* __restrict__ mask, __nv_bfloat16* __restrict__ o,
                                                        long long n, int Nl, int heads, float scale) {
  constexpr int C = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const long long warp_global = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long p0 = warp_global * PIX;
  if (p0 >= n) return;
  const int c0 = lane * CPL;                       // this lane's first channel
  const int gl = 32 / heads;                       // lanes per head (heads divides 32)

  float mu[CPL], rs[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    mu[i] = __ldg(stats + (static_cast<long long>(b) * 2) * C + c0 + i);
    rs[i] = __ldg(stats + (static_cast<long long>(b) * 2 + 1) * C + c0 + i);
  }
  float q[PIX][CPL];
#pragma unroll
  for (int pi = 0; pi < PIX; ++pi) {
    const long long pp = min(p0 + pi, n - 1);
    const float* src = qpre + (static_cast<long long>(b) * n + pp) * C + c0;
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(src + i));
      q[pi][i] = u.x; q[pi][i + 1] = u.y; q[pi][i + 2] = u.z; q[pi][i + 3] = u.w;
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) q[pi][i] = (q[pi][i] - mu[i]) * rs[i] * scale;
  }

  // online softmax over words (two sweeps would re-read k; one sweep with running max keeps k, v traffic minimal)
  float mx[PIX], den[PIX], acc[PIX][CPL];
#pragma unroll
  for (int pi = 0; pi < PIX; ++pi) {
    mx[pi] = -INFINITY;
    den[pi] = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[pi][i] = 0.f;
  }
  const float* kb = k + static_cast<long long>(b) * Nl * C + c0;
  const float* vb = v + static_cast<long long>(b) * Nl * C + c0;
  for (int j = 0; j < Nl; ++j) {
    float kk[CPL], vv[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(kb + static_cast<long long>(j) * C + i));
      kk[i] = a.x; kk[i + 1] = a.y; kk[i + 2] = a.z; kk[i + 3] = a.w;
      const float4 c = __ldg(reinterpret_cast<const float4*>(vb + static_cast<long long>(j) * C + i));
      vv[i] = c.x; vv[i + 1] = c.y; vv[i + 2] = c.z; vv[i + 3] = c.w;
    }
    const float madd = 1e4f * __ldg(mask + b * Nl + j) - 1e4f;
#pragma unroll
    for (int pi = 0; pi < PIX; ++pi) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) s = fmaf(q[pi][i], kk[i], s);
      // reduce inside the head's lane group
      for (int off = gl >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      s += madd;
      const float nm = fmaxf(mx[pi], s);
      const float corr = __expf(mx[pi] - nm);
      const float pj = __expf(s - nm);
      den[pi] = den[pi] * corr + pj;
#pragma unroll
      for (int i = 0; i < CPL; ++i) acc[pi][i] = fmaf(pj, vv[i], acc[pi][i] * corr);
      mx[pi] = nm;
    }
  }
#pragma unroll
  for (int pi = 0; pi < PIX; ++pi) {
    if (p0 + pi >= n) break;
    const float inv = 1.0f / den[pi];
    __nv_bfloat16* dst = o + (static_cast<long long>(b) * n + p0 + pi) * C + c0;
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(acc[pi][i] * inv, acc[pi][i + 1] * inv),
                                                      pack_bf16x2(acc[pi][i + 2] * inv, acc[pi][i + 3] * inv));
    }
  }
}

int pwam_core_dispatch(const float* qpre, const float* stats, const float* k, const float* v, const float* mask,
                       __nv_bfloat16* o, int B, long long n, int C, int Nl, int heads, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && n > 0, "pwam_core: empty input");
  LAVT_REQUIRE(Nl >= 1 && Nl <= 4096, "pwam_core: Nl=%d out of range", Nl);
  LAVT_REQUIRE(heads >= 1 && heads <= 32 && (32 % heads) == 0, "pwam_core: fusion heads=%d must divide 32", heads);
  const float scale = 1.0f / sqrtf(static_cast<float>(C));
#define LAVT_PWAM_CASE(cpl, pix)                                                                              \
  case cpl * 32: {                                                                                            \
    const long long warps = (n + pix - 1) / pix;                                                              \
    dim3 grid(static_cast<unsigned>((warps + 7) / 8), B);                                                     \
    pwam_core_kernel<cpl, pix><<<grid, 256, 0, st>>>(qpre, stats, k, v, mask, o, n, Nl, heads, scale);      \
    break;                                                                                                    \
  }
  // pixels per warp pass chosen so that q + accumulators stay in registers
  switch (C) {
    LAVT_PWAM_CASE(4, 4)
    LAVT_PWAM_CASE(8, 4)
    LAVT_PWAM_CASE(12, 2)
    LAVT_PWAM_CASE(16, 2)
    LAVT_PWAM_CASE(24, 1)
    LAVT_PWAM_CASE(32, 1)
    default:
      set_last_error("pwam_core: C=%d unsupported (need 128/256/384/512/768/1024)", C);
      return LAVT_ERR_SHAPE;
  }
#undef LAVT_PWAM_CASE
  LAVT_LAUNCH_CHECK("pwam_core_kernel");
  return LAVT_OK;
}

}  // namespace lavt
