// Bandwidth-bound pieces of the VLT fuse-and-classify head (reference lib/vlt.py) that sit between the tcgen05 GEMM / conv launches:
//   rows_affine_act     (x + add) * v[image, c] * s[c] + t[c] -> activation     joint_threshold(x_c4 * sentence) (:141-145), ReLU after lang_proj
//   avgpool2_nhwc       nn.AvgPool2d(2) over NHWC                         self.down (:151)
//   append_coords       vlt_concat_coords (:267-292) into a channel-padded NHWC operand of the first 3x3 conv of project_1
//   rows_add_table      x[r, :] + table[r % period, :]                    PositionalEncoding (:204-222) on batch-major rows
//   mha_small           softmax(q k^T / sqrt(32) + key padding mask) v    nn.MultiheadAttention cores: 16 x Nl, 900 x 900, 16 x 16, 16 x 900
//   gate_transpose      gate[b, q] * x[b, q, s] -> NHWC [b, s, q]         QueryBalancingModule's final product (:405) + q_to_spatial's view (:181)
// All of them move well under 100 MB per image batch; the contractions of the head run on gemm_bf16_tc_kernel.
#include "kernels.cuh"
#include "../../include/lavt_b200.h"

namespace lavt {

__device__ __forceinline__ float vlt_act(float x, int act) {
  if (act == LAVT_ACT_RELU) return fmaxf(x, 0.f);
  if (act == LAVT_ACT_TANH) return tanhf(x);
  if (act == LAVT_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-x));
  if (act == LAVT_ACT_GELU) return gelu_erf(x);
  return x;
}

// one thread per 8 channels of one row
template <bool IN_BF16>
__global__ void __launch_bounds__(256) rows_affine_act_kernel(const void* __restrict__ xin, long long ldx, const __nv_bfloat16* __restrict__ add,
                                                              long long lda, const float* __restrict__ v,
                                                              long long rows_per_image, const float* __restrict__ s,
                                                              const float* __restrict__ t, int act, __nv_bfloat16* __restrict__ ob,
                                                              float* __restrict__ of, long long ldo, long long rows, int C) {
  const int g8 = C / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * g8) return;
  const long long r = idx / g8;
  const int c = static_cast<int>(idx - r * g8) * 8;
  float x[8];
  if (IN_BF16) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(xin) + r * ldx + c));
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(uu[j]);
      x[2 * j] = f.x;
      x[2 * j + 1] = f.y;
    }
  } else {
    const float* p = static_cast<const float*>(xin) + r * ldx + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
  if (add) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(add + r * lda + c));
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(uu[j]);
      x[2 * j] += f.x;
      x[2 * j + 1] += f.y;
    }
  }
  const float* vr = v ? v + (r / rows_per_image) * C + c : nullptr;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = x[j];
    if (vr) y *= vr[j];
    if (s) y *= s[c + j];
    if (t) y += t[c + j];
    x[j] = vlt_act(y, act);
  }
  if (ob) {
    uint4 o;
    o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]); o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
    *reinterpret_cast<uint4*>(ob + r * ldo + c) = o;
  }
  if (of) {
    float* p = of + r * ldo + c;
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
  }
}

__global__ void __launch_bounds__(256) avgpool2_nhwc_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out,
                                                            long long ldo, int n_img, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, g8 = C / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_img) * Ho * Wo * g8) return;
  const int c = static_cast<int>(idx % g8) * 8;
  const long long pix = idx / g8;
  const int wo = static_cast<int>(pix % Wo), ho = static_cast<int>((pix / Wo) % Ho);
  const long long img = pix / (static_cast<long long>(Wo) * Ho);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long src = (img * H + 2 * ho + dy) * W + 2 * wo + dx;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + src * ldi + c));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uu[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  uint4 o;
  o.x = pack_bf16x2(0.25f * acc[0], 0.25f * acc[1]); o.y = pack_bf16x2(0.25f * acc[2], 0.25f * acc[3]);
  o.z = pack_bf16x2(0.25f * acc[4], 0.25f * acc[5]); o.w = pack_bf16x2(0.25f * acc[6], 0.25f * acc[7]);
  *reinterpret_cast<uint4*>(out + pix * ldo + c) = o;
}

// out[pix, 0:C] = in[pix, 0:C]; out[pix, C..C+2] = x in [-1, 1]; out[pix, C+3..C+5] = y in [-1, 1]; out[pix, C+6..C+7] = 0
__global__ void __launch_bounds__(256) append_coords_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out,
                                                            int n_img, int H, int W, int C) {
  const int g8 = C / 8 + 1;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_img) * H * W * g8) return;
  const int gi = static_cast<int>(idx % g8);
  const long long pix = idx / g8;
  uint4 o;
  if (gi < C / 8) {
    o = __ldg(reinterpret_cast<const uint4*>(in + pix * ldi + gi * 8));
  } else {
    const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
    const float xs = 2.0f * w / (W - 1.0f) - 1.0f, ys = 2.0f * h / (H - 1.0f) - 1.0f;
    o.x = pack_bf16x2(xs, xs); o.y = pack_bf16x2(xs, ys); o.z = pack_bf16x2(ys, ys); o.w = 0u;
  }
  *reinterpret_cast<uint4*>(out + pix * (C + 8) + gi * 8) = o;
}

template <bool IN_BF16>
__global__ void __launch_bounds__(256) rows_add_table_kernel(const void* __restrict__ xin, long long ldx, const float* __restrict__ tab,
                                                             long long period, __nv_bfloat16* __restrict__ ob, float* __restrict__ of,
                                                             long long ldo, long long rows, int C) {
  const int g4 = C / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * g4) return;
  const long long r = idx / g4;
  const int c = static_cast<int>(idx - r * g4) * 4;
  float x[4];
  if (IN_BF16) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(xin) + r * ldx + c));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(xin) + r * ldx + c));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  }
  const float4 p = __ldg(reinterpret_cast<const float4*>(tab + (r % period) * C + c));
  x[0] += p.x; x[1] += p.y; x[2] += p.z; x[3] += p.w;
  if (ob) *reinterpret_cast<uint2*>(ob + r * ldo + c) = make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
  if (of) *reinterpret_cast<float4*>(of + r * ldo + c) = make_float4(x[0], x[1], x[2], x[3]);
}

// One warp per (image, head, query): lane l scores the keys l, l + 32, ... (head_dim 32: a key row is 64 B = four 16-byte loads), warp
// softmax, every lane accumulates p.v over its keys and the 32 partial vectors are summed with shuffles.  S <= 1024 keys.
constexpr int MHA_MAX_PER_LANE = 32;
__global__ void __launch_bounds__(128) mha_small_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                                                        long long ldk, const __nv_bfloat16* __restrict__ v, long long ldv,
                                                        const float* __restrict__ mask, __nv_bfloat16* __restrict__ out, long long ldo, int B,
                                                        int Lq, int S, int heads) {
  const int lane = threadIdx.x & 31;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<long long>(B) * heads * Lq) return;
  const int iq = static_cast<int>(wid % Lq), h = static_cast<int>((wid / Lq) % heads), b = static_cast<int>(wid / (static_cast<long long>(Lq) * heads));
  float qf[32];
  {
    const __nv_bfloat16* qr = q + (static_cast<long long>(b) * Lq + iq) * ldq + h * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(qr + 8 * j));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(uu[e]);
        qf[8 * j + 2 * e] = f.x * 0.17677669529663687f;       // 32^-0.5
        qf[8 * j + 2 * e + 1] = f.y * 0.17677669529663687f;
      }
    }
  }
  float sc[MHA_MAX_PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < MHA_MAX_PER_LANE; ++t) {
    const int j = lane + 32 * t;
    float s = -INFINITY;
    if (j < S && (!mask || mask[static_cast<long long>(b) * S + j] != 0.f)) {
      const __nv_bfloat16* kr = k + (static_cast<long long>(b) * S + j) * ldk + h * 32;
      s = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(kr + 8 * jj));
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(uu[e]);
          s = fmaf(qf[8 * jj + 2 * e], f.x, fmaf(qf[8 * jj + 2 * e + 1], f.y, s));
        }
      }
    }
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (mx == -INFINITY) mx = 0.f;                  // every key masked: torch would give NaN; we give zeros
  float sum = 0.f;
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
#pragma unroll
  for (int t = 0; t < MHA_MAX_PER_LANE; ++t) {
    const int j = lane + 32 * t;
    if (j < S && sc[t] != -INFINITY) {
      const float p = __expf(sc[t] - mx);
      sum += p;
      const __nv_bfloat16* vr = v + (static_cast<long long>(b) * S + j) * ldv + h * 32;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(vr + 8 * jj));
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(uu[e]);
          acc[8 * jj + 2 * e] = fmaf(p, f.x, acc[8 * jj + 2 * e]);
          acc[8 * jj + 2 * e + 1] = fmaf(p, f.y, acc[8 * jj + 2 * e + 1]);
        }
      }
    }
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  // transpose-reduce: after the loop lane d holds the total of dimension d
  float mine = 0.f;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const float tot = warp_sum(acc[d]);
    if (lane == d) mine = tot;
  }
  out[(static_cast<long long>(b) * Lq + iq) * ldo + h * 32 + lane] = __float2bfloat16(mine * inv);
}

// out[b, s, q] = gate[b * Q + q] * x[(b * Q + q), s]   (x fp32 [B*Q, ldx], out bf16 NHWC with Q channels)
__global__ void __launch_bounds__(256) gate_transpose_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gate,
                                                             long long ldg, __nv_bfloat16* __restrict__ out, int B, int Q, int S) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * S * Q) return;
  const int qq = static_cast<int>(idx % Q);
  const long long bs = idx / Q;
  const int s = static_cast<int>(bs % S);
  const long long b = bs / S;
  const long long r = b * Q + qq;
  out[idx] = __float2bfloat16(gate[r * ldg] * x[r * ldx + s]);
}

}  // namespace lavt

using namespace lavt;

static inline unsigned vlt_blocks(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

extern "C" int lavt_rows_affine_act(const void* x, int32_t x_is_bf16, int64_t ldx, const void* add_bf16, int64_t lda, const float* v,
                                    int64_t rows_per_image, const float* s,
                                    const float* t, int32_t act, void* out_bf16, float* out_f32, int64_t ldo, int64_t rows, int32_t C,
                                    void* stream) {
  LAVT_REQUIRE(x && (out_bf16 || out_f32) && rows > 0 && C > 0 && C % 8 == 0 && ldx >= C && ldo >= C, "rows_affine_act: bad arguments");
  LAVT_REQUIRE(!v || rows_per_image > 0, "rows_affine_act: per-image vector needs rows_per_image");
  LAVT_REQUIRE(!add_bf16 || lda >= C, "rows_affine_act: addend pitch");
  LAVT_REQUIRE(act >= LAVT_ACT_NONE && act <= LAVT_ACT_SIGMOID, "rows_affine_act: activation %d not supported", act);
  const long long n = rows * (C / 8);
  LAVT_REQUIRE(n < (1LL << 39), "rows_affine_act: too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_is_bf16)
    rows_affine_act_kernel<true><<<vlt_blocks(n, 256), 256, 0, st>>>(x, ldx, static_cast<const __nv_bfloat16*>(add_bf16), lda, v, rows_per_image, s, t, act, static_cast<__nv_bfloat16*>(out_bf16),
                                                                     out_f32, ldo, rows, C);
  else
    rows_affine_act_kernel<false><<<vlt_blocks(n, 256), 256, 0, st>>>(x, ldx, static_cast<const __nv_bfloat16*>(add_bf16), lda, v, rows_per_image, s, t, act, static_cast<__nv_bfloat16*>(out_bf16),
                                                                      out_f32, ldo, rows, C);
  LAVT_LAUNCH_CHECK("rows_affine_act_kernel");
  return LAVT_OK;
}

extern "C" int lavt_avgpool2_nhwc(const void* in_bf16, int64_t ldi, void* out_bf16, int64_t ldo, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                  void* stream) {
  LAVT_REQUIRE(in_bf16 && out_bf16 && n_img > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0 && ldi >= C && ldo >= C,
               "avgpool2: needs even H, W (%dx%d) and C %% 8 == 0 (nn.AvgPool2d(2) of lib/vlt.py:61 assumes even maps)", H, W);
  const long long n = 1LL * n_img * (H / 2) * (W / 2) * (C / 8);
  avgpool2_nhwc_kernel<<<vlt_blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), ldi, static_cast<__nv_bfloat16*>(out_bf16), ldo, n_img, H, W, C);
  LAVT_LAUNCH_CHECK("avgpool2_nhwc_kernel");
  return LAVT_OK;
}

extern "C" int lavt_append_coords(const void* in_bf16, int64_t ldi, void* out_bf16, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                  void* stream) {
  LAVT_REQUIRE(in_bf16 && out_bf16 && n_img > 0 && H > 1 && W > 1 && C > 0 && C % 8 == 0 && ldi >= C, "append_coords: bad arguments");
  const long long n = 1LL * n_img * H * W * (C / 8 + 1);
  append_coords_kernel<<<vlt_blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in_bf16), ldi,
                                                                                         static_cast<__nv_bfloat16*>(out_bf16), n_img, H, W, C);
  LAVT_LAUNCH_CHECK("append_coords_kernel");
  return LAVT_OK;
}

extern "C" int lavt_rows_add_table(const void* x, int32_t x_is_bf16, int64_t ldx, const float* table, int64_t period, void* out_bf16,
                                   float* out_f32, int64_t ldo, int64_t rows, int32_t C, void* stream) {
  LAVT_REQUIRE(x && table && (out_bf16 || out_f32) && rows > 0 && period > 0 && C > 0 && C % 4 == 0 && ldx >= C && ldo >= C,
               "rows_add_table: bad arguments");
  const long long n = rows * (C / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_is_bf16)
    rows_add_table_kernel<true><<<vlt_blocks(n, 256), 256, 0, st>>>(x, ldx, table, period, static_cast<__nv_bfloat16*>(out_bf16), out_f32, ldo,
                                                                    rows, C);
  else
    rows_add_table_kernel<false><<<vlt_blocks(n, 256), 256, 0, st>>>(x, ldx, table, period, static_cast<__nv_bfloat16*>(out_bf16), out_f32, ldo,
                                                                     rows, C);
  LAVT_LAUNCH_CHECK("rows_add_table_kernel");
  return LAVT_OK;
}

extern "C" int lavt_mha_small(const void* q_bf16, int64_t ldq, const void* k_bf16, int64_t ldk, const void* v_bf16, int64_t ldv,
                              const float* key_mask, void* out_bf16, int64_t ldo, int32_t B, int32_t Lq, int32_t S, int32_t heads,
                              void* stream) {
  LAVT_REQUIRE(q_bf16 && k_bf16 && v_bf16 && out_bf16 && B > 0 && Lq > 0 && S > 0 && heads > 0, "mha_small: bad arguments");
  LAVT_REQUIRE(S <= 32 * MHA_MAX_PER_LANE, "mha_small: at most %d keys (got %d)", 32 * MHA_MAX_PER_LANE, S);
  LAVT_REQUIRE(ldq >= heads * 32 && ldk >= heads * 32 && ldv >= heads * 32 && ldo >= heads * 32 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0,
               "mha_small: head_dim is 32; row pitches must cover heads * 32 and be multiples of 8 elements");
  const long long warps = 1LL * B * heads * Lq;
  mha_small_kernel<<<vlt_blocks(warps, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(q_bf16), ldq, static_cast<const __nv_bfloat16*>(k_bf16), ldk, static_cast<const __nv_bfloat16*>(v_bf16),
      ldv, key_mask, static_cast<__nv_bfloat16*>(out_bf16), ldo, B, Lq, S, heads);
  LAVT_LAUNCH_CHECK("mha_small_kernel");
  return LAVT_OK;
}

extern "C" int lavt_gate_transpose(const float* x, int64_t ldx, const float* gate, int64_t ldg, void* out_bf16, int32_t B, int32_t Q, int32_t S,
                                   void* stream) {
  LAVT_REQUIRE(x && gate && out_bf16 && B > 0 && Q > 0 && S > 0 && ldx >= S && ldg >= 1, "gate_transpose: bad arguments");
  const long long n = 1LL * B * S * Q;
  gate_transpose_kernel<<<vlt_blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, gate, ldg, static_cast<__nv_bfloat16*>(out_bf16),
                                                                                          B, Q, S);
  LAVT_LAUNCH_CHECK("gate_transpose_kernel");
  return LAVT_OK;
}
