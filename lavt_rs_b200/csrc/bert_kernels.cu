// Text side of LAVTVideo / LAVTOne.forward (reference lib/_utils.py:52-54, 98-100: BertModel(text, attention_mask)[0];
// the reference's bert/ package is a copy of HuggingFace Transformers v3.0.2 modeling_bert.py, README.md:9-13).
// The dense layers run on the tcgen05 GEMM kernel and the LayerNorms on ln_rows; this file holds what is left:
//   bert_embed      word + position + token-type embedding gather (BertEmbeddings.forward) -> fp32 rows (LayerNorm follows)
//   bert_attention  per (sentence, head): softmax(q k^T / sqrt(64) + (1 - mask) * -10000) v  over <= 128 tokens
//                   (BertSelfAttention.forward); q arrives pre-scaled by 64^-0.5 * log2(e) (qkv GEMM epilogue)
//   rows_to_cf      (B, Nl, C) fp32 -> (B, C, Nl) fp32: the .permute(0, 2, 1) the PWAM kernels expect (lib/_utils.py:54)
#include "kernels.cuh"

namespace lavt {

constexpr int BERT_HD = 64;
constexpr int BERT_MAX_NL = 128;

__global__ void __launch_bounds__(192) bert_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                                         const float* __restrict__ pos, const float* __restrict__ type0,
                                                         float* __restrict__ out, int Nl, int H, int vocab) {
  // grid: (B * Nl); thread -> float4 of the hidden vector
  const int row = blockIdx.x;
  const int t = row % Nl;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* w = reinterpret_cast<const float4*>(word + id * H);
  const float4* p = reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * H);
  const float4* ty = reinterpret_cast<const float4*>(type0);
  float4* o = reinterpret_cast<float4*>(out + static_cast<long long>(row) * H);
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) {
    const float4 a = __ldg(w + i), b = __ldg(p + i), c = __ldg(ty + i);
    o[i] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
  }
}

int bert_embed_dispatch(const long long* ids, const float* word, const float* pos, const float* type0, float* out, int B, int Nl,
                        int H, int vocab, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && H % 4 == 0 && vocab > 0, "bert_embed: bad shape");
  bert_embed_kernel<<<B * Nl, 192, 0, st>>>(ids, word, pos, type0, out, Nl, H, vocab);
  LAVT_LAUNCH_CHECK("bert_embed_kernel");
  return LAVT_OK;
}

// grid (heads, B), 128 threads; thread i owns query token i.  K / V of the head are staged in shared memory (bf16 pairs) and read
// with warp-uniform (broadcast) addresses.
__global__ void __launch_bounds__(BERT_MAX_NL) bert_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ mask,
                                                                     __nv_bfloat16* __restrict__ out, int Nl, int H) {
  __shared__ uint32_t sk[BERT_MAX_NL][BERT_HD / 2 + 1];     // bf16 pairs
  __shared__ uint32_t sv[BERT_MAX_NL][BERT_HD / 2 + 1];
  __shared__ float smask[BERT_MAX_NL];
  const int head = blockIdx.x, b = blockIdx.y;
  const int ld = 3 * H;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * Nl * ld + head * BERT_HD;
  for (int i = threadIdx.x; i < Nl * (BERT_HD / 2); i += blockDim.x) {
    const int j = i / (BERT_HD / 2), c = (i % (BERT_HD / 2)) * 2;
    sk[j][c >> 1] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * ld + H + c);
    sv[j][c >> 1] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(j) * ld + 2 * H + c);
  }
  for (int j = threadIdx.x; j < Nl; j += blockDim.x)
    smask[j] = (1.0f - __ldg(mask + b * Nl + j)) * (-10000.0f * 1.4426950408889634f);   // extended attention mask, base-2 domain
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= Nl) return;
  float q[BERT_HD];
#pragma unroll
  for (int c = 0; c < BERT_HD; c += 2) {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + static_cast<long long>(i) * ld + c));
    q[c] = x.x; q[c + 1] = x.y;
  }
  // pass 1: scores and their maximum (kept in registers would need Nl floats: recompute in pass 2 instead)
  float mx = -INFINITY;
  for (int j = 0; j < Nl; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < BERT_HD; c += 2) {
      const float2 kk = unpack_bf16x2(sk[j][c >> 1]);
      s = fmaf(q[c], kk.x, fmaf(q[c + 1], kk.y, s));
    }
    mx = fmaxf(mx, s + smask[j]);
  }
  float acc[BERT_HD];
#pragma unroll
  for (int c = 0; c < BERT_HD; ++c) acc[c] = 0.f;
  float den = 0.f;
  for (int j = 0; j < Nl; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < BERT_HD; c += 2) {
      const float2 kk = unpack_bf16x2(sk[j][c >> 1]);
      s = fmaf(q[c], kk.x, fmaf(q[c + 1], kk.y, s));
    }
    const float pj = exp2f(s + smask[j] - mx);
    den += pj;
#pragma unroll
    for (int c = 0; c < BERT_HD; c += 2) {
      const float2 vv = unpack_bf16x2(sv[j][c >> 1]);
      acc[c] = fmaf(pj, vv.x, acc[c]);
      acc[c + 1] = fmaf(pj, vv.y, acc[c + 1]);
    }
  }
  const float inv = 1.0f / den;
  __nv_bfloat16* o = out + (static_cast<long long>(b) * Nl + i) * H + head * BERT_HD;
#pragma unroll
  for (int c = 0; c < BERT_HD; c += 2) *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2(acc[c] * inv, acc[c + 1] * inv);
}

int bert_attention_dispatch(const __nv_bfloat16* qkv, const float* mask, __nv_bfloat16* out, int B, int Nl, int H, int heads,
                            cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Nl <= BERT_MAX_NL, "bert_attention: sentence length %d out of range (1..%d)", Nl, BERT_MAX_NL);
  LAVT_REQUIRE(H == heads * BERT_HD, "bert_attention: head_dim must be 64 (hidden %d, heads %d)", H, heads);
  bert_attention_kernel<<<dim3(heads, B), BERT_MAX_NL, 0, st>>>(qkv, mask, out, Nl, H);
  LAVT_LAUNCH_CHECK("bert_attention_kernel");
  return LAVT_OK;
}

// ---- split-precision operands: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together).  One GEMM over the
// K-concatenated operands  [hi | lo | hi] x [W_hi | W_hi | W_lo]^T  = hi W_hi + lo W_hi + hi W_lo  accumulates all three products in
// fp32 on the tensor cores: relative error ~1e-5 instead of bf16's 4e-3 at three times the (tiny) FLOPs of the text encoder.
__global__ void __launch_bounds__(256) split3_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long M,
                                                          int Kd) {
  const int g4 = Kd / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * g4) return;
  const long long r = idx / g4;
  const int c = static_cast<int>(idx - r * g4) * 4;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
  const float f[4] = {v.x, v.y, v.z, v.w};
  float lo[4];
  uint32_t hi2[2], lo2[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) lo[j] = f[j] - __bfloat162float(__float2bfloat16(f[j]));
  hi2[0] = pack_bf16x2(f[0], f[1]); hi2[1] = pack_bf16x2(f[2], f[3]);
  lo2[0] = pack_bf16x2(lo[0], lo[1]); lo2[1] = pack_bf16x2(lo[2], lo[3]);
  __nv_bfloat16* o = out + r * 3 * Kd + c;
  *reinterpret_cast<uint2*>(o) = make_uint2(hi2[0], hi2[1]);
  *reinterpret_cast<uint2*>(o + Kd) = make_uint2(lo2[0], lo2[1]);
  *reinterpret_cast<uint2*>(o + 2 * Kd) = make_uint2(hi2[0], hi2[1]);
}

int split3_bf16_dispatch(const float* x, long long ldx, __nv_bfloat16* out, long long M, int Kd, cudaStream_t st) {
  LAVT_REQUIRE(x && out && M > 0 && Kd > 0 && Kd % 4 == 0 && ldx >= Kd, "split3: bad shape");
  const long long n = M * (Kd / 4);
  split3_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, ldx, out, M, Kd);
  LAVT_LAUNCH_CHECK("split3_bf16_kernel");
  return LAVT_OK;
}

// fp32 twin of bert_attention for the split-precision path: qkv fp32 [B*Nl, 3H], out fp32 [B*Nl, H]; K / V staged in dynamic shared memory
__global__ void __launch_bounds__(BERT_MAX_NL) bert_attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                         float* __restrict__ out, int Nl, int H) {
  extern __shared__ float sm_attn[];
  float* sk = sm_attn;                               // [Nl][65]
  float* sv = sk + Nl * (BERT_HD + 1);
  float* smask = sv + Nl * (BERT_HD + 1);
  const int head = blockIdx.x, b = blockIdx.y;
  const int ld = 3 * H;
  const float* base = qkv + static_cast<long long>(b) * Nl * ld + head * BERT_HD;
  for (int i = threadIdx.x; i < Nl * BERT_HD; i += blockDim.x) {
    const int j = i / BERT_HD, c = i % BERT_HD;
    sk[j * (BERT_HD + 1) + c] = base[static_cast<long long>(j) * ld + H + c];
    sv[j * (BERT_HD + 1) + c] = base[static_cast<long long>(j) * ld + 2 * H + c];
  }
  for (int j = threadIdx.x; j < Nl; j += blockDim.x) smask[j] = (1.0f - __ldg(mask + b * Nl + j)) * (-10000.0f * 1.4426950408889634f);
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= Nl) return;
  float q[BERT_HD];
#pragma unroll
  for (int c = 0; c < BERT_HD; ++c) q[c] = base[static_cast<long long>(i) * ld + c];
  float mx = -INFINITY;
  for (int j = 0; j < Nl; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < BERT_HD; ++c) s = fmaf(q[c], sk[j * (BERT_HD + 1) + c], s);
    mx = fmaxf(mx, s + smask[j]);
  }
  float acc[BERT_HD];
#pragma unroll
  for (int c = 0; c < BERT_HD; ++c) acc[c] = 0.f;
  float den = 0.f;
  for (int j = 0; j < Nl; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < BERT_HD; ++c) s = fmaf(q[c], sk[j * (BERT_HD + 1) + c], s);
    const float pj = exp2f(s + smask[j] - mx);
    den += pj;
#pragma unroll
    for (int c = 0; c < BERT_HD; ++c) acc[c] = fmaf(pj, sv[j * (BERT_HD + 1) + c], acc[c]);
  }
  const float inv = 1.0f / den;
  float* o = out + (static_cast<long long>(b) * Nl + i) * H + head * BERT_HD;
#pragma unroll
  for (int c = 0; c < BERT_HD; ++c) o[c] = acc[c] * inv;
}

int bert_attention_f32_dispatch(const float* qkv, const float* mask, float* out, int B, int Nl, int H, int heads, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Nl <= BERT_MAX_NL, "bert_attention: sentence length %d out of range (1..%d)", Nl, BERT_MAX_NL);
  LAVT_REQUIRE(H == heads * BERT_HD, "bert_attention: head_dim must be 64 (hidden %d, heads %d)", H, heads);
  const int smem = (2 * Nl * (BERT_HD + 1) + Nl) * 4;
  static int configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(bert_attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  bert_attention_f32_kernel<<<dim3(heads, B), BERT_MAX_NL, smem, st>>>(qkv, mask, out, Nl, H);
  LAVT_LAUNCH_CHECK("bert_attention_f32_kernel");
  return LAVT_OK;
}

__global__ void __launch_bounds__(256) rows_to_cf_kernel(const float* __restrict__ in, float* __restrict__ out, int Nl, int C) {
  // grid (ceil(C / 32), B); 32 x Nl tile through shared memory
  __shared__ float tile[BERT_MAX_NL][33];
  const int b = blockIdx.y, c0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < Nl * 32; i += blockDim.x) {
    const int j = i >> 5, c = i & 31;
    if (c0 + c < C) tile[j][c] = in[(static_cast<long long>(b) * Nl + j) * C + c0 + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Nl * 32; i += blockDim.x) {
    const int c = i / Nl, j = i % Nl;
    if (c0 + c < C) out[(static_cast<long long>(b) * C + c0 + c) * Nl + j] = tile[j][c];
  }
}

int rows_to_cf_dispatch(const float* in, float* out, int B, int Nl, int C, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Nl <= BERT_MAX_NL && C > 0, "rows_to_cf: bad shape");
  rows_to_cf_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(in, out, Nl, C);
  LAVT_LAUNCH_CHECK("rows_to_cf_kernel");
  return LAVT_OK;
}

}  // namespace lavt
