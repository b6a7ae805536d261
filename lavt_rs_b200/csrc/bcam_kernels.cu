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
#include "../../include/lavt_b200.h"
#include "kernels.cuh"

namespace lavt {

constexpr int BCAM_WCHUNK = 8;     // words per warp in bcam_words_kernel

// grid (C, ceil(Nlp / 8), B), block 32: one warp = one output channel x 8 words; lanes stride over the Lin input features
__global__ void __launch_bounds__(32) bcam_words_kernel(const float* __restrict__ l, const float* __restrict__ w, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ lr, __nv_bfloat16* __restrict__ lrT, int Nl, int Nlp, int Lin,
                                                        int C) {
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
      const __nv_bfloat16 v = __float2bfloat16((j0 + j < Nl) ? acc[j] + bc : 0.f);
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

}  // namespace lavt

using namespace lavt;

extern "C" int lavt_bcam_words(const float* l, const float* w, const float* bias, void* lr_bf16, void* lrT_bf16, int32_t B, int32_t Nl,
                               int32_t Nlp, int32_t Lin, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Nlp >= Nl && Lin > 0 && C > 0 && B < 65536, "bcam_words: bad sizes (B=%d Nl=%d Nlp=%d)", B, Nl, Nlp);
  LAVT_REQUIRE(l && w && bias && lr_bf16 && lrT_bf16, "bcam_words: missing buffers");
  bcam_words_kernel<<<dim3(C, (Nlp + BCAM_WCHUNK - 1) / BCAM_WCHUNK, B), 32, 0, static_cast<cudaStream_t>(stream)>>>(
      l, w, bias, static_cast<__nv_bfloat16*>(lr_bf16), static_cast<__nv_bfloat16*>(lrT_bf16), Nl, Nlp, Lin, C);
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
