// Training-mode SimpleDecoding (reference lib/mask_predictor.py:56-99 with BatchNorm2d in train mode, i.e. batch statistics;
// train.py:589 converts them to SyncBatchNorm), its backward, the final x4 upsample (lib/_utils.py:106) backward and the
// [0.9, 1.1]-weighted cross-entropy (losses.py:7-11).  The conv3x3 forward / input-gradient run on the tcgen05 implicit-GEMM
// kernel, the weight gradients on the split-K GEMM over zero-padded transposed layouts; these are the bandwidth kernels around.
//
//   bn_relu_apply        t = relu((z - mean) * rstd * gamma + beta)            z fp32 conv output, t bf16 NHWC
//   bn_relu_bwd_reduce   sums[0] = sum dy, sums[1] = sum dy * z^  with dy = dt * [t > 0]   (= d beta, d gamma)
//   bn_relu_bwd_apply    dz = gamma * rstd * (dy - sums[0]/N - z^ sums[1]/N)   bf16
//   nhwc_pad_transpose   [n,H,W,C] bf16 -> [C, n*(H+2)*Wp] (zero border pre-set, Wp = W+2 rounded up to 8): operand layout of the conv
//                        weight gradient, in which the 3x3 tap (ky,kx) is the constant column offset (ky-1)*Wp + (kx-1).  TMA needs
//                        the innermost box coordinate 16-byte aligned (measured: an odd element offset raises an illegal-instruction
//                        fault), so the row part (ky-1)*Wp goes into the TMA coordinate and the +-1 column part into three copies of
//                        x^T written with a column shift
//   upsample_concat_bwd  gradient of the bilinear (align_corners=True) upsample half of cat[U(prev), skip] as a GATHER over
//                        the hat-function support (no atomics)
//   conv1x1_logits_bwd, upsample_logits_bwd, ce_loss_fwd / ce_loss_bwd
#include "kernels.cuh"

namespace lavt {

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const float4* __restrict__ z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            uint2* __restrict__ t, int C, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % (C / 4));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(stats) + c4), rs = __ldg(reinterpret_cast<const float4*>(stats + C) + c4);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  const float4 x = __ldg(z + i);
  const float o0 = fmaxf((x.x - mu.x) * rs.x * g.x + b.x, 0.f), o1 = fmaxf((x.y - mu.y) * rs.y * g.y + b.y, 0.f);
  const float o2 = fmaxf((x.z - mu.z) * rs.z * g.z + b.z, 0.f), o3 = fmaxf((x.w - mu.w) * rs.w * g.w + b.w, 0.f);
  t[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
}

int bn_relu_apply_dispatch(const float* z, const float* stats, const float* gamma, const float* beta, __nv_bfloat16* t, long long npix,
                           int C, cudaStream_t st) {
  LAVT_REQUIRE(npix > 0 && C % 4 == 0, "bn_relu: bad sizes");
  const long long total4 = npix * (C / 4);
  bn_relu_apply_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(z), stats, gamma, beta,
                                                                                 reinterpret_cast<uint2*>(t), C, total4);
  LAVT_LAUNCH_CHECK("bn_relu_apply_kernel");
  return LAVT_OK;
}

// grid = chunks of 256 pixels; thread -> 4 channels, row groups stride the chunk; block partials -> atomics
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dt, const __nv_bfloat16* __restrict__ t,
                                                                 const float* __restrict__ z, const float* __restrict__ stats,
                                                                 float* __restrict__ sums, long long npix, int C) {
  extern __shared__ float brr_sm[];
  const int tpr = C / 4, rg = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(stats) + tc), rs = __ldg(reinterpret_cast<const float4*>(stats + C) + tc);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const long long r0 = static_cast<long long>(blockIdx.x) * 256;
  const long long r1 = (r0 + 256 < npix) ? r0 + 256 : npix;
  if (tr < rg) {
    for (long long r = r0 + tr; r < r1; r += rg) {
      const long long e = r * C + tc * 4;
      const uint2 ud = __ldg(reinterpret_cast<const uint2*>(dt + e)), ut = __ldg(reinterpret_cast<const uint2*>(t + e));
      const float4 x = __ldg(reinterpret_cast<const float4*>(z + e));
      const float2 d0 = unpack_bf16x2(ud.x), d1 = unpack_bf16x2(ud.y), t0 = unpack_bf16x2(ut.x), t1 = unpack_bf16x2(ut.y);
      const float y0 = t0.x > 0.f ? d0.x : 0.f, y1 = t0.y > 0.f ? d0.y : 0.f, y2 = t1.x > 0.f ? d1.x : 0.f, y3 = t1.y > 0.f ? d1.y : 0.f;
      s1.x += y0; s1.y += y1; s1.z += y2; s1.w += y3;
      s2.x += y0 * (x.x - mu.x) * rs.x; s2.y += y1 * (x.y - mu.y) * rs.y; s2.z += y2 * (x.z - mu.z) * rs.z; s2.w += y3 * (x.w - mu.w) * rs.w;
    }
    float* o = brr_sm + (static_cast<long long>(tr) * C + tc * 4) * 2;
    o[0] = s1.x; o[1] = s2.x; o[2] = s1.y; o[3] = s2.y; o[4] = s1.z; o[5] = s2.z; o[6] = s1.w; o[7] = s2.w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int g = 0; g < rg; ++g) {
      a += brr_sm[(g * C + c) * 2 + 0];
      q += brr_sm[(g * C + c) * 2 + 1];
    }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, q);
  }
}

__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dt, const __nv_bfloat16* __restrict__ t,
                                                                const float* __restrict__ z, const float* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ sums,
                                                                __nv_bfloat16* __restrict__ dz, float inv_n, int C, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % (C / 4));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(stats) + c4), rs = __ldg(reinterpret_cast<const float4*>(stats + C) + c4);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
  const float4 S1 = __ldg(reinterpret_cast<const float4*>(sums) + c4), S2 = __ldg(reinterpret_cast<const float4*>(sums + C) + c4);
  const uint2 ud = __ldg(reinterpret_cast<const uint2*>(dt) + i), ut = __ldg(reinterpret_cast<const uint2*>(t) + i);
  const float4 x = __ldg(reinterpret_cast<const float4*>(z) + i);
  const float2 d0 = unpack_bf16x2(ud.x), d1 = unpack_bf16x2(ud.y), t0 = unpack_bf16x2(ut.x), t1 = unpack_bf16x2(ut.y);
  const float y0 = t0.x > 0.f ? d0.x : 0.f, y1 = t0.y > 0.f ? d0.y : 0.f, y2 = t1.x > 0.f ? d1.x : 0.f, y3 = t1.y > 0.f ? d1.y : 0.f;
  const float h0 = (x.x - mu.x) * rs.x, h1 = (x.y - mu.y) * rs.y, h2 = (x.z - mu.z) * rs.z, h3 = (x.w - mu.w) * rs.w;
  const float o0 = g.x * rs.x * (y0 - S1.x * inv_n - h0 * S2.x * inv_n), o1 = g.y * rs.y * (y1 - S1.y * inv_n - h1 * S2.y * inv_n);
  const float o2 = g.z * rs.z * (y2 - S1.z * inv_n - h2 * S2.z * inv_n), o3 = g.w * rs.w * (y3 - S1.w * inv_n - h3 * S2.w * inv_n);
  reinterpret_cast<uint2*>(dz)[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
}

int bn_relu_bwd_dispatch(const __nv_bfloat16* dt, const __nv_bfloat16* t, const float* z, const float* stats, const float* gamma,
                         float* sums, __nv_bfloat16* dz, long long npix, long long n_stat, int C, int phase, cudaStream_t st) {
  LAVT_REQUIRE(npix > 0 && C % 4 == 0 && C <= 1024, "bn backward: C=%d unsupported", C);
  if (phase == 0) {
    const int rg = 256 / (C / 4);
    bn_relu_bwd_reduce_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, static_cast<size_t>(rg) * C * 2 * sizeof(float), st>>>(
        dt, t, z, stats, sums, npix, C);
    LAVT_LAUNCH_CHECK("bn_relu_bwd_reduce_kernel");
  } else {
    const long long total4 = npix * (C / 4);
    bn_relu_bwd_apply_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(dt, t, z, stats, gamma, sums, dz,
                                                                                       1.0f / static_cast<float>(n_stat), C, total4);
    LAVT_LAUNCH_CHECK("bn_relu_bwd_apply_kernel");
  }
  return LAVT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// out[c, p' - dshift] = in[(img,h,w), c],  p' = (img*(H+2) + h+1)*Wp + w+1   (64 pixels x 64 channels per block through smem)
__global__ void __launch_bounds__(256) nhwc_pad_transpose_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                                 __nv_bfloat16* __restrict__ out, long long ldo, long long npix, int C, int H,
                                                                 int W, int Wp, int dshift, int D) {
  __shared__ unsigned short tile[64][66];
  __shared__ long long dcol[64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long m0 = static_cast<long long>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  if (threadIdx.x < 64) {
    const long long m = m0 + threadIdx.x;
    long long d = -1;
    if (m < npix) {
      const int w = static_cast<int>(m % W), h = static_cast<int>((m / W) % H);
      long long img = m / (static_cast<long long>(W) * H);
      if (D > 0) img = (img / D) * (D + 2) + img % D + 1;      // 3-D: frames of a clip padded with one zero frame on either side
      d = (img * (H + 2) + h + 1) * Wp + w + 1 - dshift;
    }
    dcol[threadIdx.x] = d;
  }
  for (int r = ty; r < 64; r += 8) {
    uint32_t v = 0;
    if (m0 + r < npix && n0 + 2 * tx < C) v = *reinterpret_cast<const uint32_t*>(in + (m0 + r) * ldi + n0 + 2 * tx);
    tile[r][2 * tx] = static_cast<unsigned short>(v & 0xffffu);
    tile[r][2 * tx + 1] = static_cast<unsigned short>(v >> 16);
  }
  __syncthreads();
  unsigned short* o16 = reinterpret_cast<unsigned short*>(out);
  for (int c = ty; c < 64; c += 8) {
    if (n0 + c >= C) continue;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = tx + 32 * k;
      const long long d = dcol[r];
      if (d >= 0) o16[static_cast<long long>(n0 + c) * ldo + d] = tile[r][c];
    }
  }
}

int nhwc_pad_transpose_dispatch(const __nv_bfloat16* in, long long ldi, __nv_bfloat16* out, long long ldo, int n_img, int H, int W, int C,
                                int Wp, int dshift, int D, cudaStream_t st) {
  const long long npix = 1LL * n_img * H * W;
  LAVT_REQUIRE(npix > 0 && C % 2 == 0 && ldi % 2 == 0, "pad transpose: bad sizes");
  LAVT_REQUIRE(Wp >= W + 2 && dshift >= -1 && dshift <= 1, "pad transpose: padded width %d / shift %d invalid", Wp, dshift);
  LAVT_REQUIRE(D == 0 || n_img % D == 0, "pad transpose: %d frames do not split into clips of %d", n_img, D);
  LAVT_REQUIRE(ldo >= 1LL * (D > 0 ? n_img / D * (D + 2) : n_img) * (H + 2) * Wp, "pad transpose: output pitch too small");
  nhwc_pad_transpose_kernel<<<dim3(static_cast<unsigned>((npix + 63) / 64), static_cast<unsigned>((C + 63) / 64)), 256, 0, st>>>(
      in, ldi, out, ldo, npix, C, H, W, Wp, dshift, D);
  LAVT_LAUNCH_CHECK("nhwc_pad_transpose_kernel");
  return LAVT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// dprev[img, sy, sx, c] = sum over output pixels (y, x) of hat(y*ry - sy) * hat(x*rx - sx) * dcat[img, y, x, c],  c < C1
// (adjoint of the align_corners=True bilinear upsample: the interpolation weight of source sy for output y is
//  max(0, 1 - |y*ry - sy|), ry = (ph-1)/(H-1)); thread per (source pixel, 8 channels)
__device__ __forceinline__ void hat_range(int s, int in, int out, float r, int& lo, int& hi) {
  if (out <= 1 || in <= 1) { lo = 0; hi = out - 1; return; }
  const float inv = 1.0f / r;
  lo = max(0, static_cast<int>(floorf((s - 1) * inv)) - 1);
  hi = min(out - 1, static_cast<int>(ceilf((s + 1) * inv)) + 1);
}
__device__ __forceinline__ float hat_weight(int o, int s, int in, int out) {
  int i0, i1;
  float f;
  // same arithmetic as the forward's bilinear_taps
  const float sc = (out > 1) ? static_cast<float>(o) * (static_cast<float>(in - 1) / static_cast<float>(out - 1)) : 0.f;
  i0 = min(static_cast<int>(sc), in - 1);
  i1 = min(i0 + 1, in - 1);
  f = sc - static_cast<float>(i0);
  float w = 0.f;
  if (s == i0) w += 1.f - f;
  if (s == i1) w += f;
  return w;
}

__global__ void __launch_bounds__(256) upsample_concat_bwd_kernel(const __nv_bfloat16* __restrict__ dcat, int Ct, __nv_bfloat16* __restrict__ dprev,
                                                                  int ph, int pw, int C1, int H, int W, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g8 = C1 / 8;
  const int c = static_cast<int>(idx % g8) * 8;
  const long long sp = idx / g8;
  const int sx = static_cast<int>(sp % pw), sy = static_cast<int>((sp / pw) % ph);
  const long long img = sp / (static_cast<long long>(pw) * ph);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ph == H && pw == W) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(dcat + ((img * H + sy) * W + sx) * Ct + c));
    *reinterpret_cast<uint4*>(dprev + sp * C1 + c) = u;
    return;
  }
  const float ry = (H > 1) ? static_cast<float>(ph - 1) / static_cast<float>(H - 1) : 1.f;
  const float rx = (W > 1) ? static_cast<float>(pw - 1) / static_cast<float>(W - 1) : 1.f;
  int ylo, yhi, xlo, xhi;
  hat_range(sy, ph, H, ry, ylo, yhi);
  hat_range(sx, pw, W, rx, xlo, xhi);
  for (int y = ylo; y <= yhi; ++y) {
    const float wy = hat_weight(y, sy, ph, H);
    if (wy == 0.f) continue;
    for (int x = xlo; x <= xhi; ++x) {
      const float wgt = wy * hat_weight(x, sx, pw, W);
      if (wgt == 0.f) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dcat + ((img * H + y) * W + x) * Ct + c));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uu[j]);
        acc[2 * j] = fmaf(wgt, f.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(wgt, f.y, acc[2 * j + 1]);
      }
    }
  }
  *reinterpret_cast<uint4*>(dprev + sp * C1 + c) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                              pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

int upsample_concat_bwd_dispatch(const __nv_bfloat16* dcat, int Ct, __nv_bfloat16* dprev, int ph, int pw, int C1, int n_img, int H, int W,
                                 cudaStream_t st) {
  LAVT_REQUIRE(C1 % 8 == 0 && Ct % 8 == 0 && C1 <= Ct && n_img > 0, "upsample backward: bad sizes");
  const long long total = 1LL * n_img * ph * pw * (C1 / 8);
  upsample_concat_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dcat, Ct, dprev, ph, pw, C1, H, W, total);
  LAVT_LAUNCH_CHECK("upsample_concat_bwd_kernel");
  return LAVT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// logits = y w^T + b (2 classes): dy[pix, c] = dl0 w[0,c] + dl1 w[1,c];  dw[k, c] += sum_pix dl_k y[pix, c];  db[k] += sum dl_k
// block = 256 pixels; thread -> 8 channels x pixel group
__global__ void __launch_bounds__(256) conv1x1_logits_bwd_kernel(const float* __restrict__ dlog, const __nv_bfloat16* __restrict__ y,
                                                                 const float* __restrict__ w, __nv_bfloat16* __restrict__ dy,
                                                                 float* __restrict__ dw, float* __restrict__ db, long long npix, int C) {
  extern __shared__ float clb_sm[];        // [2][C] accumulators
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) clb_sm[i] = 0.f;
  __syncthreads();
  const int tpr = C / 8, rg = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const long long r0 = static_cast<long long>(blockIdx.x) * 256;
  const long long r1 = (r0 + 256 < npix) ? r0 + 256 : npix;
  float a0[8], a1[8], w0[8], w1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a0[j] = a1[j] = 0.f;
    w0[j] = __ldg(w + tc * 8 + j);
    w1[j] = __ldg(w + C + tc * 8 + j);
  }
  float b0 = 0.f, b1 = 0.f;
  if (tr < rg) {
    for (long long r = r0 + tr; r < r1; r += rg) {
      const float2 dl = __ldg(reinterpret_cast<const float2*>(dlog) + r);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(y + r * C + tc * 8));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uu[j]);
        a0[2 * j] = fmaf(dl.x, f.x, a0[2 * j]); a0[2 * j + 1] = fmaf(dl.x, f.y, a0[2 * j + 1]);
        a1[2 * j] = fmaf(dl.y, f.x, a1[2 * j]); a1[2 * j + 1] = fmaf(dl.y, f.y, a1[2 * j + 1]);
        o[j] = pack_bf16x2(dl.x * w0[2 * j] + dl.y * w1[2 * j], dl.x * w0[2 * j + 1] + dl.y * w1[2 * j + 1]);
      }
      *reinterpret_cast<uint4*>(dy + r * C + tc * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      if (tc == 0) { b0 += dl.x; b1 += dl.y; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&clb_sm[tc * 8 + j], a0[j]);
      atomicAdd(&clb_sm[C + tc * 8 + j], a1[j]);
    }
    if (tc == 0) { atomicAdd(db, b0); atomicAdd(db + 1, b1); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(dw + i, clb_sm[i]);
}

int conv1x1_logits_bwd_dispatch(const float* dlog, const __nv_bfloat16* y, const float* w, __nv_bfloat16* dy, float* dw, float* db,
                                long long npix, int C, cudaStream_t st) {
  LAVT_REQUIRE(npix > 0 && C % 8 == 0 && C <= 2048, "conv1x1 backward: C=%d unsupported", C);
  conv1x1_logits_bwd_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 2 * C * sizeof(float), st>>>(dlog, y, w, dy, dw, db, npix, C);
  LAVT_LAUNCH_CHECK("conv1x1_logits_bwd_kernel");
  return LAVT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint of upsample_logits: dout (n,2,H,W) fp32 NCHW -> din (n,h,w,2) fp32; thread per source pixel
__global__ void __launch_bounds__(256) upsample_logits_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int n_img, int h,
                                                                  int w, int H, int W) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_img) * h * w;
  if (idx >= total) return;
  const int sx = static_cast<int>(idx % w), sy = static_cast<int>((idx / w) % h);
  const long long img = idx / (static_cast<long long>(w) * h);
  const float ry = (H > 1) ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 1.f;
  const float rx = (W > 1) ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 1.f;
  int ylo, yhi, xlo, xhi;
  hat_range(sy, h, H, ry, ylo, yhi);
  hat_range(sx, w, W, rx, xlo, xhi);
  const long long plane = static_cast<long long>(H) * W;
  const float* base = dout + img * 2 * plane;
  float a0 = 0.f, a1 = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    const float wy = hat_weight(y, sy, h, H);
    if (wy == 0.f) continue;
    for (int x = xlo; x <= xhi; ++x) {
      const float wgt = wy * hat_weight(x, sx, w, W);
      if (wgt == 0.f) continue;
      a0 = fmaf(wgt, __ldg(base + static_cast<long long>(y) * W + x), a0);
      a1 = fmaf(wgt, __ldg(base + plane + static_cast<long long>(y) * W + x), a1);
    }
  }
  reinterpret_cast<float2*>(din)[idx] = make_float2(a0, a1);
}

int upsample_logits_bwd_dispatch(const float* dout, float* din, int n_img, int h, int w, int H, int W, cudaStream_t st) {
  const long long total = static_cast<long long>(n_img) * h * w;
  LAVT_REQUIRE(total > 0 && H > 0 && W > 0, "upsample_logits backward: empty input");
  upsample_logits_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dout, din, n_img, h, w, H, W);
  LAVT_LAUNCH_CHECK("upsample_logits_bwd_kernel");
  return LAVT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// weighted 2-class cross-entropy (losses.py:7-11): loss = sum_i w[t_i] * -log softmax(x_i)[t_i] / sum_i w[t_i]
// acc[0] += sum w * nll, acc[1] += sum w   (phase 0);   dlogits = w[t] * (p - onehot(t)) / acc[1] * gscale   (phase 1)
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float w0, float w1,
                                                      float* __restrict__ acc, float* __restrict__ dlogits, float gscale, long long plane,
                                                      long long total, int phase) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float nll = 0.f, wt = 0.f;
  if (idx < total) {
    const long long img = idx / plane, px = idx - img * plane;
    const float x0 = __ldg(logits + img * 2 * plane + px), x1 = __ldg(logits + img * 2 * plane + plane + px);
    const int t = static_cast<int>(__ldg(target + idx));
    const float m = fmaxf(x0, x1);
    const float e0 = __expf(x0 - m), e1 = __expf(x1 - m);
    const float lse = m + __logf(e0 + e1);
    wt = t ? w1 : w0;
    nll = wt * (lse - (t ? x1 : x0));
    if (phase == 1) {
      const float inv = 1.0f / (e0 + e1);
      const float s = wt * gscale / acc[1];
      dlogits[img * 2 * plane + px] = s * (e0 * inv - (t ? 0.f : 1.f));
      dlogits[img * 2 * plane + plane + px] = s * (e1 * inv - (t ? 1.f : 0.f));
    }
  }
  if (phase == 0) {
    nll = warp_sum(nll);
    wt = warp_sum(wt);
    __shared__ float sa[8], sb[8];
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = nll; sb[threadIdx.x >> 5] = wt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int i = 0; i < 8; ++i) { a += sa[i]; b += sb[i]; }
      atomicAdd(acc, a);
      atomicAdd(acc + 1, b);
    }
  }
}

int ce_loss_dispatch(const float* logits, const long long* target, float w0, float w1, float* acc, float* dlogits, float gscale, int n_img,
                     int H, int W, int phase, cudaStream_t st) {
  const long long plane = 1LL * H * W, total = plane * n_img;
  LAVT_REQUIRE(total > 0 && acc != nullptr && (phase == 0 || dlogits != nullptr), "cross-entropy: bad arguments");
  ce_loss_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(logits, target, w0, w1, acc, dlogits, gscale, plane, total, phase);
  LAVT_LAUNCH_CHECK("ce_loss_kernel");
  return LAVT_OK;
}

}  // namespace lavt
