// Input / output edges of the inference scripts:
//   normalize_u8    T.ToTensor() + T.Normalize(mean, std) of already resized frames (reference train.py:54-60, transforms.py:83-112):
//                   uint8 HWC frames -> fp32 (n, 3, H, W) planes, (v / 255 - mean[c]) / std[c]
//   logits_to_mask  F.interpolate(outputs, (origin_h, origin_w), bilinear, align_corners=True).argmax(1) and the 0 / 255 greyscale
//                   image that is saved (test_ytvos.py:249-253, 274-279): fp32 (n, 2, H, W) logits -> uint8 (n, oh, ow)
// (The PIL antialiased resize of T.Resize and the PNG encoder stay on the host.)
#include "kernels.cuh"

namespace lavt {

__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long npix_img,
                                                           long long total, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long img = idx / npix_img, px = idx - img * npix_img;
  const uint8_t* p = in + idx * 3;
  float* o = out + img * 3 * npix_img + px;
  o[0] = (static_cast<float>(p[0]) * (1.0f / 255.0f) - m0) / s0;
  o[npix_img] = (static_cast<float>(p[1]) * (1.0f / 255.0f) - m1) / s1;
  o[2 * npix_img] = (static_cast<float>(p[2]) * (1.0f / 255.0f) - m2) / s2;
}

int normalize_u8_dispatch(const uint8_t* in, float* out, int n_img, int H, int W, const float* mean, const float* stdv, cudaStream_t st) {
  const long long npix = 1LL * H * W, total = npix * n_img;
  LAVT_REQUIRE(total > 0, "normalize: empty input");
  normalize_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, npix, total, mean[0], mean[1], mean[2], stdv[0],
                                                                                stdv[1], stdv[2]);
  LAVT_LAUNCH_CHECK("normalize_u8_kernel");
  return LAVT_OK;
}

__device__ __forceinline__ void edge_taps(int o, int in, int out, int& i0, int& i1, float& f) {
  // align_corners=True, same arithmetic as torch's area_pixel_compute_source_index: src = o * (in-1)/(out-1)
  const float s = (out > 1) ? static_cast<float>(o) * (static_cast<float>(in - 1) / static_cast<float>(out - 1)) : 0.f;
  i0 = min(static_cast<int>(s), in - 1);
  i1 = min(i0 + 1, in - 1);
  f = s - static_cast<float>(i0);
}

__global__ void __launch_bounds__(256) logits_to_mask_kernel(const float* __restrict__ logits, uint8_t* __restrict__ mask, int n_img, int H, int W,
                                                             int oh, int ow) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_img) * oh * ow;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % ow), y = static_cast<int>((idx / ow) % oh);
  const long long img = idx / (static_cast<long long>(ow) * oh);
  int y0, y1, x0, x1;
  float fy, fx;
  edge_taps(y, H, oh, y0, y1, fy);
  edge_taps(x, W, ow, x0, x1, fx);
  const long long plane = static_cast<long long>(H) * W;
  const float* b = logits + img * 2 * plane;
  float v[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* q = b + c * plane;
    const float top = q[static_cast<long long>(y0) * W + x0] * (1.f - fx) + q[static_cast<long long>(y0) * W + x1] * fx;
    const float bot = q[static_cast<long long>(y1) * W + x0] * (1.f - fx) + q[static_cast<long long>(y1) * W + x1] * fx;
    v[c] = top * (1.f - fy) + bot * fy;
  }
  mask[idx] = (v[1] > v[0]) ? 255 : 0;        // argmax(1): ties go to class 0
}

int logits_to_mask_dispatch(const float* logits, uint8_t* mask, int n_img, int H, int W, int oh, int ow, cudaStream_t st) {
  const long long total = 1LL * n_img * oh * ow;
  LAVT_REQUIRE(total > 0 && H > 0 && W > 0, "logits_to_mask: empty input");
  logits_to_mask_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(logits, mask, n_img, H, W, oh, ow);
  LAVT_LAUNCH_CHECK("logits_to_mask_kernel");
  return LAVT_OK;
}

}  // namespace lavt

extern "C" int lavt_normalize_u8(const uint8_t* frames_hwc, float* out_nchw, int32_t n_img, int32_t H, int32_t W, const float* mean3_host,
                                 const float* std3_host, void* stream) {
  using namespace lavt;
  LAVT_REQUIRE(mean3_host && std3_host, "normalize: mean / std are NULL");
  return normalize_u8_dispatch(frames_hwc, out_nchw, n_img, H, W, mean3_host, std3_host, static_cast<cudaStream_t>(stream));
}

extern "C" int lavt_logits_to_mask(const float* logits_nchw, uint8_t* mask, int32_t n_img, int32_t H, int32_t W, int32_t out_h, int32_t out_w,
                                   void* stream) {
  return lavt::logits_to_mask_dispatch(logits_nchw, mask, n_img, H, W, out_h, out_w, static_cast<cudaStream_t>(stream));
}
