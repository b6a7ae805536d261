// Bandwidth-bound pieces of SimpleDecoding (reference lib/mask_predictor.py:56-99) and the final logits
// upsample (lib/_utils.py:106).  The conv3x3+BN+ReLU stack itself runs on the tcgen05 implicit-GEMM kernel.
//
//   upsample_concat   cat[ bilinear(prev -> skip size, align_corners=True), skip ]  -> NHWC bf16 conv input
//   conv1x1_logits    512 -> 2 pointwise conv + bias (one warp per pixel)          -> (pix, 2) fp32
//   upsample_logits   (n,h,w,2) fp32 -> (n,2,H,W) fp32 NCHW, bilinear align_corners=True
//   nhwc_to_nchw      fp32 NHWC -> NCHW (API-compat output layout of backbone.forward) via smem transpose
//   nchw_to_nhwc_bf16 fp32 NCHW -> NHWC bf16 (API-compat input layout of classifier.forward)
#include "kernels.cuh"

namespace lavt {

__device__ __forceinline__ void bilinear_taps(int o, int in, int out, int& i0, int& i1, float& f) {
  // align_corners=True: src = o * (in-1)/(out-1)
  const float s = (out > 1) ? static_cast<float>(o) * (static_cast<float>(in - 1) / static_cast<float>(out - 1)) : 0.f;
  i0 = min(static_cast<int>(s), in - 1);
  i1 = min(i0 + 1, in - 1);
  f = s - static_cast<float>(i0);
}

// thread per (pixel, 8-channel group): consecutive threads touch consecutive 16-byte chunks, so every load / store
// instruction of a warp covers 512 contiguous bytes (the bilinear taps of neighbouring pixels share cache lines)
__global__ void __launch_bounds__(256) upsample_concat_kernel(const __nv_bfloat16* __restrict__ prev, int ph, int pw, int C1,
                                                              const __nv_bfloat16* __restrict__ skip, int C2,
                                                              __nv_bfloat16* __restrict__ out, int n_img, int H, int W) {
  // grid = (chunks of one output row, H, images): no 64-bit index arithmetic on the per-thread path
  const int Ct = C1 + C2, g8 = Ct / 8;
  const int xi = blockIdx.x * blockDim.x + threadIdx.x;
  if (xi >= W * g8) return;
  const int w = xi / g8;
  const int c = (xi - w * g8) * 8;
  const int h = blockIdx.y;
  const long long img = blockIdx.z;
  const long long pix = (img * H + h) * W + w;
  uint4 r;
  if (c >= C1) {
    r = __ldg(reinterpret_cast<const uint4*>(skip + pix * C2 + (c - C1)));
  } else if (ph == H && pw == W) {
    r = __ldg(reinterpret_cast<const uint4*>(prev + pix * C1 + c));
  } else {
    int y0, y1, x0, x1;
    float fy, fx;
    bilinear_taps(h, ph, H, y0, y1, fy);
    bilinear_taps(w, pw, W, x0, x1, fx);
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    const __nv_bfloat16* pbase = prev + img * ph * pw * C1 + c;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(pbase + (static_cast<long long>(y0) * pw + x0) * C1));
    const uint4 bq = __ldg(reinterpret_cast<const uint4*>(pbase + (static_cast<long long>(y0) * pw + x1) * C1));
    const uint4 cc = __ldg(reinterpret_cast<const uint4*>(pbase + (static_cast<long long>(y1) * pw + x0) * C1));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(pbase + (static_cast<long long>(y1) * pw + x1) * C1));
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {bq.x, bq.y, bq.z, bq.w}, uc[4] = {cc.x, cc.y, cc.z, cc.w},
                   ud[4] = {d.x, d.y, d.z, d.w};
    uint32_t uo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]), fc = unpack_bf16x2(uc[j]), fd = unpack_bf16x2(ud[j]);
      uo[j] = pack_bf16x2(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x, w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y);
    }
    r = make_uint4(uo[0], uo[1], uo[2], uo[3]);
  }
  *reinterpret_cast<uint4*>(out + pix * Ct + c) = r;
}

int upsample_concat_dispatch(const __nv_bfloat16* prev, int ph, int pw, int C1, const __nv_bfloat16* skip, int C2,
                             __nv_bfloat16* out, int n_img, int H, int W, cudaStream_t st) {
  // C2 == 0 (skip may be NULL): plain bilinear upsample -- the nn.Upsample(x2) steps of ProgressiveDecoding (lib/vlt.py:437-452)
  LAVT_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || skip != nullptr), "upsample_concat: channels must be multiples of 8");
  LAVT_REQUIRE(ph <= H && pw <= W && ph > 0 && pw > 0, "upsample_concat: prev (%dx%d) larger than skip (%dx%d)", ph, pw, H, W);
  LAVT_REQUIRE(n_img > 0 && n_img < 65536 && H > 0 && H < 65536 && W > 0, "upsample_concat: empty or too large input");
  const int per_row = W * ((C1 + C2) / 8);
  upsample_concat_kernel<<<dim3((per_row + 255) / 256, H, n_img), 256, 0, st>>>(prev, ph, pw, C1, skip, C2, out, n_img, H, W);
  LAVT_LAUNCH_CHECK("upsample_concat_kernel");
  return LAVT_OK;
}

// 8 lanes per pixel (4 pixels per warp): logits[pix, o] = sum_c y[pix, c] * w[o, c] + b[o], o in {0, 1}; w staged in smem
__global__ void __launch_bounds__(256) conv1x1_logits_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                                                             const float* __restrict__ b, float* __restrict__ out, long long npix,
                                                             int C) {
  extern __shared__ float sw[];   // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const bool live = pix < npix;
  float a0 = 0.f, a1 = 0.f;
  if (live) {
    const __nv_bfloat16* row = y + pix * C;
    for (int c = sub * 8; c < C; c += 64) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + c));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uu[j]);
        a0 = fmaf(f.x, sw[c + 2 * j], fmaf(f.y, sw[c + 2 * j + 1], a0));
        a1 = fmaf(f.x, sw[C + c + 2 * j], fmaf(f.y, sw[C + c + 2 * j + 1], a1));
      }
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (live && sub == 0) *reinterpret_cast<float2*>(out + pix * 2) = make_float2(a0 + b[0], a1 + b[1]);
}

int conv1x1_logits_dispatch(const __nv_bfloat16* y, const float* w, const float* b, float* out, long long npix, int C,
                            cudaStream_t st) {
  LAVT_REQUIRE(npix > 0 && C % 8 == 0 && C <= 4096, "conv1x1: bad shape");
  const long long threads = npix * 8;
  conv1x1_logits_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 2 * C * sizeof(float), st>>>(y, w, b, out, npix, C);
  LAVT_LAUNCH_CHECK("conv1x1_logits_kernel");
  return LAVT_OK;
}

// thread per output pixel (both classes); writes NCHW
__global__ void __launch_bounds__(256) upsample_logits_kernel(const float* __restrict__ in, float* __restrict__ out, int n_img,
                                                              int h, int w, int H, int W) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_img) * H * W;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H);
  const long long img = idx / (static_cast<long long>(W) * H);
  int y0, y1, x0, x1; float fy, fx;
  bilinear_taps(y, h, H, y0, y1, fy);
  bilinear_taps(x, w, W, x0, x1, fx);
  const float2* base = reinterpret_cast<const float2*>(in) + img * h * w;
  const float2 a = __ldg(base + y0 * w + x0), b = __ldg(base + y0 * w + x1), c = __ldg(base + y1 * w + x0),
               d = __ldg(base + y1 * w + x1);
  const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
  const long long plane = static_cast<long long>(H) * W;
  float* o = out + img * 2 * plane + static_cast<long long>(y) * W + x;
  o[0] = w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
  o[plane] = w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
}

int upsample_logits_dispatch(const float* in, float* out, int n_img, int h, int w, int H, int W, cudaStream_t st) {
  const long long total = static_cast<long long>(n_img) * H * W;
  LAVT_REQUIRE(total > 0 && h > 0 && w > 0, "upsample_logits: empty input");
  upsample_logits_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, n_img, h, w, H, W);
  LAVT_LAUNCH_CHECK("upsample_logits_kernel");
  return LAVT_OK;
}

// (n, P, C) fp32 -> (n, C, P) fp32, 32x32 smem tiles.  grid (ceil(P/32), ceil(C/32), n), block (32, 8)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int C) {
  __shared__ float tile[32][33];
  const long long img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < P && c < C) ? in[(img * P + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < P && c < C) out[(img * C + c) * P + p] = tile[threadIdx.x][i];
  }
}

int nhwc_to_nchw_dispatch(const float* in, float* out, int n_img, int P, int C, cudaStream_t st) {
  LAVT_REQUIRE(n_img > 0 && P > 0 && C > 0 && n_img < 65536, "nhwc_to_nchw: bad shape");
  nhwc_to_nchw_kernel<<<dim3((P + 31) / 32, (C + 31) / 32, n_img), dim3(32, 8), 0, st>>>(in, out, P, C);
  LAVT_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return LAVT_OK;
}

// (n, C, P) fp32 -> (n, P, C) bf16
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int P, int C) {
  __shared__ float tile[32][33];
  const long long img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < P && c < C) ? in[(img * C + c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < P && c < C) out[(img * P + p) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

int nchw_to_nhwc_bf16_dispatch(const float* in, __nv_bfloat16* out, int n_img, int P, int C, cudaStream_t st) {
  LAVT_REQUIRE(n_img > 0 && P > 0 && C > 0 && n_img < 65536, "nchw_to_nhwc: bad shape");
  nchw_to_nhwc_bf16_kernel<<<dim3((P + 31) / 32, (C + 31) / 32, n_img), dim3(32, 8), 0, st>>>(in, out, P, C);
  LAVT_LAUNCH_CHECK("nchw_to_nhwc_bf16_kernel");
  return LAVT_OK;
}

}  // namespace lavt
