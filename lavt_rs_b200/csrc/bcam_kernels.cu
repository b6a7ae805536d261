// BCAM fusion (reference lib/bcam.py:8-75, the --bcam ablation of the 2-D image backbone, lib/backbone.py:573-577; from BRINet).
// Every contraction of the module is a lavt_gemm_bf16 call (tcgen05); these are the three bandwidth kernels between them:
//     lr   = lang_reduce(l^T)                      bcam_words_kernel        -> bf16 [B, Nlp, C] and its transpose [B, C, Nlp]
//     sim  = softmax_words(relu(vis_1 x) lr^T + (1e4 mask - 1e4))   GEMM + bcam_softmax_rows_kernel (warp per row)
//     out  = sim lr                                 GEMM (K = Nlp)
//     A    = tanh(out_1(out) + vis_2_2(relu(vis_2 x)))              ONE GEMM over the row-concatenated operand [out | q2], K = 2C
//     rel  = softmax_pixels(a_proj(A))              GEMM (N = hw) + bcam_softmax_rows_kernel (block per row, row staged in shared memory)
//     out2 = rel relu(vis_3 x)                      bcam_transpose_pad_kernel (q3 -> [C, hw] K-major) + GEMM (K = hw)
//     out3 = relu(out3_proj([out2 | out])) + relu(vis_4 x)          GEMM (K = 2C, ReLU, residual epilogue)
// Rows / columns added to reach the GEMM's N % 32 / K % 8 granules are written as exact zeros here.
//
// EFN fusion (lib/bcam.py:160-269, --efn) reuses the same pieces -- words (with GELU and the word mask), row softmax (the column softmax of
// the co-attention map is the row softmax of the transposed GEMM), transpose-pad -- plus
//     efn_sentence_bias_kernel   project's language half: Conv1d over cat[x, sentence] = W_x x + (W_l sentence + b) -> a per-image bias
//     efn_norm_pool_kernel       InstanceNorm (precomputed stats) + AvgPool2d(2) of the token axis read as a square image (:240-249)
//     efn_norm_upsample_kernel   InstanceNorm + bilinear x 2 (align_corners False) back to the full map (:263-266)
// and the k = 3 Conv1d over the flattened token axis (:231-233) runs as three row-shifted accumulating GEMMs.
#include "../../include/lavt_b200.h"
#include "kernels.cuh"

namespace lavt {

constexpr int BCAM_WCHUNK = 8;     // words per warp in bcam_words_kernel

// grid (C, ceil(Nlp / 8), B), block 32: one warp = one output channel x 8 words; lanes stride over the Lin input features
__global__ void __launch_bounds__(32) bcam_words_kernel(const float* __restrict__ l, const float* __restrict__ w, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ lr, __nv_bfloat16* __restrict__ lrT, int Nl, int Nlp, int Lin,
                                                        int C, const float* __restrict__ mask, int act) {
  const int c = blockIdx.x, j0 = blockIdx.y * BCAM_WCHUNK, b = blockIdx.z, lane = threadIdx.x;
  float acc[BCAM_WCHUNK];
#pragma unroll
  for (int j = 0; j < BCAM_WCHUNK; ++j) acc[j] = 0.f;
  const float* lb = l + static_cast<long long>(b) * Lin * Nl;
  for (int k = lane; k < Lin; k += 32) {
    const float wk = __ldg(w + static_cast<long long>(c) * Lin + k);
#pragma unroll
    for (int j = 0; j < BCAM_WCHUNK; ++j)
      if (j0 + j < Nl) acc[j] = fmaf(wk, __ldg(lb + static_cast<long long>(k) * Nl + j0 + j), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < BCAM_WCHUNK; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    const float bc = bias[c];
#pragma unroll
    for (int j = 0; j < BCAM_WCHUNK; ++j) {
      if (j0 + j >= Nlp) break;
      float f = 0.f;
      if (j0 + j < Nl) {
        f = acc[j] + bc;
        if (act == LAVT_ACT_GELU) f = gelu_erf(f);
        else if (act == LAVT_ACT_RELU) f = fmaxf(f, 0.f);          // VLT project_lang (lib/vlt.py:320-322)
        if (mask) f *= mask[b * Nl + j0 + j];
      }
      const __nv_bfloat16 v = __float2bfloat16(f);
      lr[(static_cast<long long>(b) * Nlp + j0 + j) * C + c] = v;
      lrT[(static_cast<long long>(b) * C + c) * Nlp + j0 + j] = v;
    }
  }
}

// p[r, 0:cols] = softmax(s[r, 0:cols] + (1e4 mask - 1e4)), p[r, cols:ldp] = 0.  THREADS = 32: one warp per row (8 rows per block, short
// rows: the word axis); THREADS = 256: one block per row with the row staged in shared memory (the pixel axis, up to 16384 columns).
template <int THREADS>
__global__ void __launch_bounds__(256) bcam_softmax_rows_kernel(const float* __restrict__ s, long long lds, const float* __restrict__ mask,
                                                                long long rows_per_mask, __nv_bfloat16* __restrict__ p, long long ldp,
                                                                long long rows, int cols) {
  extern __shared__ float bs_row[];
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (THREADS == 32) ? static_cast<long long>(blockIdx.x) * 8 + warp : blockIdx.x;
  if (r >= rows) return;
  const int t = (THREADS == 32) ? lane : threadIdx.x;
  const float* sr = s + r * lds;
  const float* mr = mask ? mask + (r / rows_per_mask) * cols : nullptr;
  __nv_bfloat16* pr = p + r * ldp;
  float* row = (THREADS == 32) ? nullptr : bs_row;

  float mx = -INFINITY;
  for (int c = t; c < cols; c += THREADS) {
    float v = sr[c];
    if (mr) v += 1e4f * mr[c] - 1e4f;
    if (row) row[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (THREADS != 32) {
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < THREADS / 32; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
  }
  float sum = 0.f;
  for (int c = t; c < cols; c += THREADS) {
    float v = row ? row[c] : (mr ? sr[c] + (1e4f * mr[c] - 1e4f) : sr[c]);
    v = __expf(v - mx);
    if (row) row[c] = v;
    sum += v;
  }
  sum = warp_sum(sum);
  if (THREADS != 32) {
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) sum += red[i];
  }
  const float inv = 1.0f / sum;
  for (int c = t; c < ldp; c += THREADS) {
    float v = 0.f;
    if (c < cols) v = (row ? row[c] : __expf((mr ? sr[c] + (1e4f * mr[c] - 1e4f) : sr[c]) - mx)) * inv;
    pr[c] = __float2bfloat16(v);
  }
}

// out[b, c, 0:n] = in[b * n + i, c] for i < n, out[b, c, n:ldo] = 0.  grid (ceil(ldo / 32), ceil(C / 32), B), block (32, 8)
__global__ void __launch_bounds__(256) bcam_transpose_pad_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                                 __nv_bfloat16* __restrict__ out, long long ldo, long long n, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int b = blockIdx.z;
  const long long i0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += 8) {
    const long long i = i0 + dy;
    const int c = c0 + threadIdx.x;
    tile[dy][threadIdx.x] = (i < n && c < C) ? in[(static_cast<long long>(b) * n + i) * ldi + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += 8) {
    const int c = c0 + dy;
    const long long i = i0 + threadIdx.x;
    if (c < C && i < ldo) out[(static_cast<long long>(b) * C + c) * ldo + i] = tile[threadIdx.x][dy];
  }
}

// sb[b, c] = bias[c] + sum_k w[c, k] sent[b, k], sent = masked mean of the word features (lib/bcam.py:179-180).  grid (C, B), block 32
__global__ void __launch_bounds__(32) efn_sentence_bias_kernel(const float* __restrict__ l, const float* __restrict__ mask,
                                                               const float* __restrict__ w, long long ldw, const float* __restrict__ bias,
                                                               float* __restrict__ sb, int Nl, int Lin, int C) {
  const int c = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  const float* lb = l + static_cast<long long>(b) * Lin * Nl;
  const float* mb = mask + static_cast<long long>(b) * Nl;
  float cnt = 0.f;
  for (int j = 0; j < Nl; ++j) cnt += mb[j];
  float acc = 0.f;
  for (int k = lane; k < Lin; k += 32) {
    float sk = 0.f;
    for (int j = 0; j < Nl; ++j) sk = fmaf(__ldg(lb + static_cast<long long>(k) * Nl + j), mb[j], sk);
    acc = fmaf(__ldg(w + static_cast<long long>(c) * ldw + k), sk / cnt, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) sb[static_cast<long long>(b) * C + c] = acc + bias[c];
}

// out[b, i', :] = mean over the 2 x 2 block of (pre - mean) * rstd  (pool != 0; the n = h * h tokens are a row-major square image), or the
// plain normalised row (pool == 0); rows n_out .. rows_out are zero.  One thread = 4 channels.
__global__ void __launch_bounds__(256) efn_norm_pool_kernel(const float4* __restrict__ pre, const float* __restrict__ stats,
                                                            uint2* __restrict__ out, long long n, int h, int pool, long long n_out,
                                                            long long rows_out, int C, long long total4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = C >> 2;
  const int cq = static_cast<int>(idx % c4);
  const long long row = idx / c4;
  const long long b = row / rows_out, i = row % rows_out;
  uint2 o = make_uint2(0u, 0u);
  if (i < n_out) {
    const float4* pb = pre + b * n * c4 + cq;
    float4 v;
    if (pool) {
      const int h2 = h >> 1;
      const long long y = i / h2, x = i % h2;
      const long long t00 = (2 * y) * h + 2 * x;
      const float4 a = pb[t00 * c4], bb = pb[(t00 + 1) * c4], cc = pb[(t00 + h) * c4], d = pb[(t00 + h + 1) * c4];
      v = make_float4(0.25f * (a.x + bb.x + cc.x + d.x), 0.25f * (a.y + bb.y + cc.y + d.y), 0.25f * (a.z + bb.z + cc.z + d.z),
                      0.25f * (a.w + bb.w + cc.w + d.w));
    } else {
      v = pb[i * c4];
    }
    const float4 mu = *reinterpret_cast<const float4*>(stats + (b * 2) * C + cq * 4);
    const float4 rs = *reinterpret_cast<const float4*>(stats + (b * 2 + 1) * C + cq * 4);
    o.x = pack_bf16x2((v.x - mu.x) * rs.x, (v.y - mu.y) * rs.y);
    o.y = pack_bf16x2((v.z - mu.z) * rs.z, (v.w - mu.w) * rs.w);
  }
  out[idx] = o;
}

// out[b, i, :] = bilinear x 2 (align_corners False: src = max(0, (dst + 0.5) / 2 - 0.5)) of the normalised (h/2) x (h/2) map (up != 0), or
// the plain normalised row (up == 0).  Interpolation weights sum to one, so normalising after interpolating is the same map.
__global__ void __launch_bounds__(256) efn_norm_upsample_kernel(const float4* __restrict__ pre, const float* __restrict__ stats,
                                                                float4* __restrict__ out_f32, uint2* __restrict__ out_bf16, long long n,
                                                                int h, int up, long long n_in, int C, long long total4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = C >> 2;
  const int cq = static_cast<int>(idx % c4);
  const long long row = idx / c4;
  const long long b = row / n, i = row % n;
  const float4* pb = pre + b * n_in * c4 + cq;
  float4 v;
  if (up) {
    const int h2 = h >> 1;
    const int y = static_cast<int>(i / h), x = static_cast<int>(i % h);
    const float sy = fmaxf(0.f, (y + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.f, (x + 0.5f) * 0.5f - 0.5f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, h2 - 1), x1 = min(x0 + 1, h2 - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float4 a = pb[(static_cast<long long>(y0) * h2 + x0) * c4], bb = pb[(static_cast<long long>(y0) * h2 + x1) * c4];
    const float4 cc = pb[(static_cast<long long>(y1) * h2 + x0) * c4], d = pb[(static_cast<long long>(y1) * h2 + x1) * c4];
    v = make_float4(w00 * a.x + w01 * bb.x + w10 * cc.x + w11 * d.x, w00 * a.y + w01 * bb.y + w10 * cc.y + w11 * d.y,
                    w00 * a.z + w01 * bb.z + w10 * cc.z + w11 * d.z, w00 * a.w + w01 * bb.w + w10 * cc.w + w11 * d.w);
  } else {
    v = pb[i * c4];
  }
  const float4 mu = *reinterpret_cast<const float4*>(stats + (b * 2) * C + cq * 4);
  const float4 rs = *reinterpret_cast<const float4*>(stats + (b * 2 + 1) * C + cq * 4);
  v = make_float4((v.x - mu.x) * rs.x, (v.y - mu.y) * rs.y, (v.z - mu.z) * rs.z, (v.w - mu.w) * rs.w);
  if (out_f32) out_f32[idx] = v;
  if (out_bf16) out_bf16[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// EFN's key side kept in fp32: k_pre[i, :] = f_key(softmax_words(score[i, :] + pad mask) lang^T) = sum_j p[i, j] g[b, j, :] with
// g = f_key(lang^T) (bias included: the p[i, :] sum to one).  At most a few dozen words carry the whole token-to-token variation of this
// tensor and an InstanceNorm follows, so neither p nor the attended features may be rounded to bf16 (cf. PWAM's fp32 q_pre / lang_pre).
// One warp per token; block 256 = 8 tokens; Nl <= 128.
__global__ void __launch_bounds__(256) efn_word_attend_kernel(const float* __restrict__ score, long long lds, const float* __restrict__ mask,
                                                              const float* __restrict__ g, long long g_rows, float* __restrict__ out,
                                                              long long rows, long long n, int Nl, int C) {
  __shared__ float prob[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (r >= rows) return;
  const long long b = r / n;
  const float* sr = score + r * lds;
  const float* mr = mask + b * Nl;
  float v[4];
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = lane + 32 * q;
    v[q] = (j < Nl) ? sr[j] + (1e4f * mr[j] - 1e4f) : -INFINITY;
    mx = fmaxf(mx, v[q]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[q] = (lane + 32 * q < Nl) ? __expf(v[q] - mx) : 0.f;
    sum += v[q];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int q = 0; q < 4; ++q) prob[warp][lane + 32 * q] = v[q] * inv;
  __syncwarp();
  const float4* gb = reinterpret_cast<const float4*>(g + b * g_rows * C);
  float4* orow = reinterpret_cast<float4*>(out + r * C);
  const int c4 = C >> 2;
  for (int c = lane; c < c4; c += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < Nl; ++j) {
      const float pj = prob[warp][j];
      const float4 gv = __ldg(gb + static_cast<long long>(j) * c4 + c);
      acc.x = fmaf(pj, gv.x, acc.x);
      acc.y = fmaf(pj, gv.y, acc.y);
      acc.z = fmaf(pj, gv.z, acc.z);
      acc.w = fmaf(pj, gv.w, acc.w);
    }
    orow[c] = acc;
  }
}

}  // namespace lavt

using namespace lavt;

extern "C" int lavt_bcam_words(const float* l, const float* w, const float* bias, const float* mask, int32_t act, void* lr_bf16,
                               void* lrT_bf16, int32_t B, int32_t Nl, int32_t Nlp, int32_t Lin, int32_t C, void* stream) {
  LAVT_REQUIRE(act == LAVT_ACT_NONE || act == LAVT_ACT_GELU || act == LAVT_ACT_RELU, "bcam_words: activation %d not supported", act);
  LAVT_REQUIRE(B > 0 && Nl > 0 && Nlp >= Nl && Lin > 0 && C > 0 && B < 65536, "bcam_words: bad sizes (B=%d Nl=%d Nlp=%d)", B, Nl, Nlp);
  LAVT_REQUIRE(l && w && bias && lr_bf16 && lrT_bf16, "bcam_words: missing buffers");
  bcam_words_kernel<<<dim3(C, (Nlp + BCAM_WCHUNK - 1) / BCAM_WCHUNK, B), 32, 0, static_cast<cudaStream_t>(stream)>>>(
      l, w, bias, static_cast<__nv_bfloat16*>(lr_bf16), static_cast<__nv_bfloat16*>(lrT_bf16), Nl, Nlp, Lin, C, mask, act);
  LAVT_LAUNCH_CHECK("bcam_words_kernel");
  return LAVT_OK;
}

extern "C" int lavt_bcam_softmax_rows(const float* s, int64_t lds, const float* mask, int64_t rows_per_mask, void* p_bf16, int64_t ldp,
                                      int64_t rows, int32_t cols, void* stream) {
  LAVT_REQUIRE(rows > 0 && cols > 0 && cols <= 16384 && lds >= cols && ldp >= cols, "bcam_softmax_rows: bad sizes (rows=%lld cols=%d)",
               static_cast<long long>(rows), cols);
  LAVT_REQUIRE(s && p_bf16 && (!mask || rows_per_mask > 0), "bcam_softmax_rows: missing buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cols <= 128) {
    LAVT_REQUIRE((rows + 7) / 8 < (1LL << 31), "bcam_softmax_rows: too many rows");
    bcam_softmax_rows_kernel<32><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(s, lds, mask, rows_per_mask,
                                                                                      static_cast<__nv_bfloat16*>(p_bf16), ldp, rows, cols);
  } else {
    LAVT_REQUIRE(rows < (1LL << 31), "bcam_softmax_rows: too many rows");
    static bool configured = false;
    if (!configured) {
      LAVT_CUDA(cudaFuncSetAttribute(bcam_softmax_rows_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
      configured = true;
    }
    bcam_softmax_rows_kernel<256><<<static_cast<unsigned>(rows), 256, static_cast<size_t>(cols) * sizeof(float), st>>>(
        s, lds, mask, rows_per_mask, static_cast<__nv_bfloat16*>(p_bf16), ldp, rows, cols);
  }
  LAVT_LAUNCH_CHECK("bcam_softmax_rows_kernel");
  return LAVT_OK;
}

extern "C" int lavt_bcam_transpose_pad(const void* in_bf16, int64_t ldi, void* out_bf16, int64_t ldo, int32_t B, int64_t n, int32_t C,
                                       void* stream) {
  LAVT_REQUIRE(B > 0 && B < 65536 && n > 0 && C > 0 && ldi >= C && ldo >= n && (ldo + 31) / 32 < (1LL << 31) && (C + 31) / 32 < 65536,
               "bcam_transpose_pad: bad sizes (B=%d n=%lld C=%d)", B, static_cast<long long>(n), C);
  LAVT_REQUIRE(in_bf16 && out_bf16, "bcam_transpose_pad: missing buffers");
  bcam_transpose_pad_kernel<<<dim3(static_cast<unsigned>((ldo + 31) / 32), (C + 31) / 32, B), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), ldi, static_cast<__nv_bfloat16*>(out_bf16), ldo, n, C);
  LAVT_LAUNCH_CHECK("bcam_transpose_pad_kernel");
  return LAVT_OK;
}

extern "C" int lavt_efn_sentence_bias(const float* l, const float* mask, const float* w, int64_t ldw, const float* bias, float* sb, int32_t B,
                                      int32_t Nl, int32_t Lin, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && B < 65536 && Nl > 0 && Lin > 0 && C > 0 && ldw >= Lin, "efn_sentence_bias: bad sizes");
  LAVT_REQUIRE(l && mask && w && bias && sb, "efn_sentence_bias: missing buffers");
  efn_sentence_bias_kernel<<<dim3(C, B), 32, 0, static_cast<cudaStream_t>(stream)>>>(l, mask, w, ldw, bias, sb, Nl, Lin, C);
  LAVT_LAUNCH_CHECK("efn_sentence_bias_kernel");
  return LAVT_OK;
}

extern "C" int lavt_efn_norm_pool(const float* pre, const float* stats, void* out_bf16, int32_t B, int64_t n, int32_t h, int32_t pool,
                                  int64_t rows_out, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && n > 0 && C > 0 && C % 4 == 0, "efn_norm_pool: bad sizes (C=%d)", C);
  LAVT_REQUIRE(!pool || (h > 0 && h % 2 == 0 && static_cast<int64_t>(h) * h == n), "efn_norm_pool: %lld tokens are not an even square map",
               static_cast<long long>(n));
  const long long n_out = pool ? n / 4 : n;
  LAVT_REQUIRE(rows_out >= n_out && pre && stats && out_bf16, "efn_norm_pool: bad output rows / missing buffers");
  const long long total4 = 1LL * B * rows_out * (C / 4);
  LAVT_REQUIRE((total4 + 255) / 256 < (1LL << 31), "efn_norm_pool: too large");
  efn_norm_pool_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(pre), stats, reinterpret_cast<uint2*>(out_bf16), n, h, pool, n_out, rows_out, C, total4);
  LAVT_LAUNCH_CHECK("efn_norm_pool_kernel");
  return LAVT_OK;
}

extern "C" int lavt_efn_norm_upsample(const float* pre, const float* stats, float* out_f32, void* out_bf16, int32_t B, int64_t n, int32_t h,
                                      int32_t up, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && n > 0 && C > 0 && C % 4 == 0, "efn_norm_upsample: bad sizes (C=%d)", C);
  LAVT_REQUIRE(!up || (h > 0 && h % 2 == 0 && static_cast<int64_t>(h) * h == n), "efn_norm_upsample: %lld tokens are not an even square map",
               static_cast<long long>(n));
  LAVT_REQUIRE(pre && stats && (out_f32 || out_bf16), "efn_norm_upsample: missing buffers");
  const long long total4 = 1LL * B * n * (C / 4);
  LAVT_REQUIRE((total4 + 255) / 256 < (1LL << 31), "efn_norm_upsample: too large");
  efn_norm_upsample_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(pre), stats, reinterpret_cast<float4*>(out_f32), reinterpret_cast<uint2*>(out_bf16), n, h, up,
      up ? n / 4 : n, C, total4);
  LAVT_LAUNCH_CHECK("efn_norm_upsample_kernel");
  return LAVT_OK;
}

extern "C" int lavt_efn_word_attend(const float* score, int64_t lds, const float* mask, const float* g, int64_t g_rows, float* out, int32_t B,
                                    int64_t n, int32_t Nl, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && n > 0 && Nl > 0 && Nl <= 128 && lds >= Nl && g_rows >= Nl && C > 0 && C % 4 == 0,
               "efn_word_attend: bad sizes (Nl=%d, C=%d)", Nl, C);
  LAVT_REQUIRE(score && mask && g && out, "efn_word_attend: missing buffers");
  const long long rows = 1LL * B * n;
  LAVT_REQUIRE((rows + 7) / 8 < (1LL << 31), "efn_word_attend: too many rows");
  efn_word_attend_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(score, lds, mask, g, g_rows, out,
                                                                                                              rows, n, Nl, C);
  LAVT_LAUNCH_CHECK("efn_word_attend_kernel");
  return LAVT_OK;
}
