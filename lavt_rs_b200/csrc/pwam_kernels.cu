// PWAM pixel-word attention core (reference lib/video_swin_transformer.py:975-1009, 2-D twin lib/backbone.py:1329-1372).
//
//   pwam_kv    k, v = (W l + b) * l_mask for the <= 77 words of a clip (768 -> C), fp32       (:985-988)
//   pwam_core  per pixel: q = InstanceNorm(q_pre); S = C^-0.5 q k^T + (1e4 m - 1e4); softmax over words;
//              o = P v                                                                           (:990-1003)
// The word side is tiny (Nl x C); the pixel side is one pass over q_pre with k, v served from L1/L2.
#include "kernels.cuh"

namespace lavt {

// grid (ceil(B * Nl / 8), ceil(2C / 64)), block 256: smem holds the word vectors l[b, :, j] of EIGHT (clip, word) pairs; each warp produces
// 8 of the block's 64 outputs (k channels first, then v channels) for all eight pairs, so a weight row is read once per 8 pairs (one pair per
// block re-read the 2C x 768 weights 160 times: 121 us at C = 1024).
constexpr int KV_PAIRS = 8;
__global__ void __launch_bounds__(256) pwam_kv_kernel(const float* __restrict__ l, const float* __restrict__ mask,
                                                      const float* __restrict__ wk, const float* __restrict__ bk,
                                                      const float* __restrict__ wv, const float* __restrict__ bv,
                                                      float* __restrict__ k, float* __restrict__ v, int B, int Nl, int Lin, int C) {
  extern __shared__ float lv[];                 // [KV_PAIRS][Lin]
  const int p0 = blockIdx.x * KV_PAIRS, npairs = B * Nl;
  for (int i = threadIdx.x; i < KV_PAIRS * Lin; i += blockDim.x) {
    const int q = i / Lin, ii = i - q * Lin;
    const int pr = min(p0 + q, npairs - 1);
    const int b = pr / Nl, j = pr - b * Nl;
    lv[i] = l[(static_cast<long long>(b) * Lin + ii) * Nl + j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int cbeg = blockIdx.y * 64;
  const int cend = min(2 * C, cbeg + 64);
  for (int c = cbeg + warp; c < cend; c += nw) {
    const bool is_v = c >= C;
    const int cc = is_v ? c - C : c;
    const float* w = (is_v ? wv : wk) + static_cast<long long>(cc) * Lin;
    float acc[KV_PAIRS];
#pragma unroll
    for (int q = 0; q < KV_PAIRS; ++q) acc[q] = 0.f;
    for (int i = lane; i < Lin; i += 32) {
      const float wi = __ldg(w + i);
#pragma unroll
      for (int q = 0; q < KV_PAIRS; ++q) acc[q] = fmaf(wi, lv[q * Lin + i], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < KV_PAIRS; ++q) acc[q] = warp_sum(acc[q]);
    if (lane < KV_PAIRS && p0 + lane < npairs) {
      float a = acc[0];
#pragma unroll
      for (int q = 1; q < KV_PAIRS; ++q) a = lane == q ? acc[q] : a;
      const int pr = p0 + lane;
      const float bias = is_v ? bv[cc] : bk[cc];
      (is_v ? v : k)[static_cast<long long>(pr) * C + cc] = (a + bias) * mask[pr];
    }
  }
}

int pwam_kv_dispatch(const float* l, const float* mask, const float* wk, const float* bk, const float* wv, const float* bv,
                     float* k, float* v, int B, int Nl, int Lin, int C, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Lin > 0 && C > 0, "pwam_kv: empty input");
  LAVT_REQUIRE(KV_PAIRS * Lin * sizeof(float) <= 48 * 1024, "pwam_kv: language width %d too large", Lin);
  pwam_kv_kernel<<<dim3((B * Nl + KV_PAIRS - 1) / KV_PAIRS, (2 * C + 63) / 64), 256, KV_PAIRS * Lin * sizeof(float), st>>>(l, mask, wk, bk, wv, bv,
                                                                                                                     k, v, B, Nl, Lin, C);
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

// ---------------------------------------------------------------------------------------------------------------
// LangProject (reference lib/video_swin_transformer.py:1012-1039, the --fuse simple ablation): per clip, masked mean of the word
// features -> Linear(768 -> C) -> ReLU -> Linear(C -> C).  The result is written NEGATED as the "mean" row of an InstanceNorm
// statistics block (mean = -lang, rstd = 1) so that pwam_mul over an all-zero lang_pre tensor yields vis * lang.
// Two mat-vec launches with one warp per output row and 8 rows per CTA, grid (C / 8, B) (a single CTA per clip took 0.27 ms at C = 1024:
// profiles/r1_ncu_launches_image_gacd.txt); the hidden vector travels in the block's second row, which the second launch's successor
// then sets to ones.
__global__ void __launch_bounds__(256) lang_project_hidden_kernel(const float* __restrict__ l, const float* __restrict__ mask,
                                                                  const float* __restrict__ w0, const float* __restrict__ b0,
                                                                  float* __restrict__ stats, int Nl, int Lin, int C) {
  extern __shared__ float lp_sm[];      // [Lin] pooled sentence vector (recomputed by every CTA: Lin * Nl L2 reads)
  const int b = blockIdx.y;
  float cnt = 0.f;
  for (int j = 0; j < Nl; ++j) cnt += mask[b * Nl + j];
  for (int i = threadIdx.x; i < Lin; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Nl; ++j) s += l[(static_cast<long long>(b) * Lin + i) * Nl + j] * mask[b * Nl + j];
    lp_sm[i] = s / cnt;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  float acc = 0.f;
  for (int i = lane; i < Lin; i += 32) acc = fmaf(__ldg(w0 + static_cast<long long>(c) * Lin + i), lp_sm[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) stats[(static_cast<long long>(b) * 2 + 1) * C + c] = fmaxf(acc + b0[c], 0.f);      // hidden vector, parked in row 1
}

__global__ void __launch_bounds__(256) lang_project_out_kernel(const float* __restrict__ w2, const float* __restrict__ b2,
                                                               float* __restrict__ stats, int C) {
  extern __shared__ float lp_sm[];      // [C] hidden
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C; i += blockDim.x) lp_sm[i] = stats[(static_cast<long long>(b) * 2 + 1) * C + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  float acc = 0.f;
  for (int i = lane; i < C; i += 32) acc = fmaf(__ldg(w2 + static_cast<long long>(c) * C + i), lp_sm[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) stats[(static_cast<long long>(b) * 2) * C + c] = -(acc + b2[c]);
}

__global__ void __launch_bounds__(256) lang_project_ones_kernel(float* __restrict__ stats, int C, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // over B * C
  if (i < total) stats[(static_cast<long long>(i / C) * 2 + 1) * C + i % C] = 1.0f;
}

int lang_project_dispatch(const float* l, const float* mask, const float* w0, const float* b0, const float* w2, const float* b2,
                          float* stats, int B, int Nl, int Lin, int C, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && B < 65536 && Nl > 0 && Lin > 0 && C > 0, "lang_project: empty input");
  LAVT_REQUIRE(static_cast<size_t>(Lin) * sizeof(float) <= 48 * 1024 && static_cast<size_t>(C) * sizeof(float) <= 48 * 1024,
               "lang_project: widths %d / %d too large", Lin, C);
  const dim3 grid((C + 7) / 8, B);
  lang_project_hidden_kernel<<<grid, 256, static_cast<size_t>(Lin) * sizeof(float), st>>>(l, mask, w0, b0, stats, Nl, Lin, C);
  LAVT_LAUNCH_CHECK("lang_project_hidden_kernel");
  lang_project_out_kernel<<<grid, 256, static_cast<size_t>(C) * sizeof(float), st>>>(w2, b2, stats, C);
  LAVT_LAUNCH_CHECK("lang_project_out_kernel");
  lang_project_ones_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(stats, C, B * C);
  LAVT_LAUNCH_CHECK("lang_project_ones_kernel");
  return LAVT_OK;
}

// Tensor-core version (mma.sync m16n8k16, bf16 operands, fp32 accumulate) used whenever the word matrices fit in
// shared memory: 16 pixels per warp, scores for all (padded) words at once -> exact softmax, no online rescaling.
//   S = (IN(q_pre) * C^-0.5) k^T + (1e4 m - 1e4)      A = normalised q rows built in registers, B = k rows (ldmatrix)
//   O = softmax(S) v                                   A = P fragments (registers), B = v rows (ldmatrix.trans)
// NT = padded word count / 8.
// ---------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) pwam_core_mma_kernel(const float* __restrict__ qpre, const float* __restrict__ stats,
                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                            const float* __restrict__ mask, __nv_bfloat16* __restrict__ o,
                                                            long long n, int C, int Nl, int heads, float scale) {
  constexpr int NLP = NT * 8;
  constexpr int KS = NT / 2;                        // k16 steps over the words for P.V
  extern __shared__ __align__(16) uint8_t smem[];
  const int pitch = C * 2 + 16;                     // bytes per word row (odd multiple of 16 -> conflict-free ldmatrix)
  uint8_t* sk = smem;                               // [NLP][pitch]
  uint8_t* sv = sk + NLP * pitch;
  float* s_mu = reinterpret_cast<float*>(sv + NLP * pitch);    // [C]
  float* s_rs = s_mu + C;                           // [C] rstd * scale
  float* s_madd = s_rs + C;                         // [NLP] additive word mask (log-domain), -inf for padded words

  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // stage k, v (fp32 -> bf16), statistics and the word mask
  for (int i = threadIdx.x; i < NLP * (C / 2); i += blockDim.x) {
    const int j = i / (C / 2), c = (i - j * (C / 2)) * 2;
    float2 kk = make_float2(0.f, 0.f), vv = kk;
    if (j < Nl) {
      kk = __ldg(reinterpret_cast<const float2*>(k + (static_cast<long long>(b) * Nl + j) * C + c));
      vv = __ldg(reinterpret_cast<const float2*>(v + (static_cast<long long>(b) * Nl + j) * C + c));
    }
    *reinterpret_cast<uint32_t*>(sk + j * pitch + c * 2) = pack_bf16x2(kk.x, kk.y);
    *reinterpret_cast<uint32_t*>(sv + j * pitch + c * 2) = pack_bf16x2(vv.x, vv.y);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_mu[i] = __ldg(stats + (static_cast<long long>(b) * 2) * C + i);
    s_rs[i] = __ldg(stats + (static_cast<long long>(b) * 2 + 1) * C + i) * scale;
  }
  for (int i = threadIdx.x; i < NLP; i += blockDim.x) s_madd[i] = (i < Nl) ? (1e4f * __ldg(mask + b * Nl + i) - 1e4f) : -INFINITY;
  __syncthreads();

  const long long p0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp) * 16;
  if (p0 >= n) return;
  const long long r0 = min(p0 + g, n - 1), r1 = min(p0 + g + 8, n - 1);
  const float* q0 = qpre + (static_cast<long long>(b) * n + r0) * C;
  const float* q1 = qpre + (static_cast<long long>(b) * n + r1) * C;
  __nv_bfloat16* o0 = o + (static_cast<long long>(b) * n + r0) * C;
  __nv_bfloat16* o1 = o + (static_cast<long long>(b) * n + r1) * C;
  const bool w0 = (p0 + g) < n, w1 = (p0 + g + 8) < n;
  const int ch = C / heads;

  for (int h = 0; h < heads; ++h) {
    const int cb = h * ch;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    for (int kc = 0; kc < ch; kc += 16) {
      const int c0 = cb + kc + 2 * t;
      uint32_t a[4];
      {
        const float2 x00 = __ldg(reinterpret_cast<const float2*>(q0 + c0)), x01 = __ldg(reinterpret_cast<const float2*>(q0 + c0 + 8));
        const float2 x10 = __ldg(reinterpret_cast<const float2*>(q1 + c0)), x11 = __ldg(reinterpret_cast<const float2*>(q1 + c0 + 8));
        const float m0 = s_mu[c0], m1 = s_mu[c0 + 1], m8 = s_mu[c0 + 8], m9 = s_mu[c0 + 9];
        const float e0 = s_rs[c0], e1 = s_rs[c0 + 1], e8 = s_rs[c0 + 8], e9 = s_rs[c0 + 9];
        a[0] = pack_bf16x2((x00.x - m0) * e0, (x00.y - m1) * e1);
        a[1] = pack_bf16x2((x10.x - m0) * e0, (x10.y - m1) * e1);
        a[2] = pack_bf16x2((x01.x - m8) * e8, (x01.y - m9) * e9);
        a[3] = pack_bf16x2((x11.x - m8) * e8, (x11.y - m9) * e9);
      }
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t kf[4];
        // matrices: (words np*16 + 0..7, chunk 0), (same words, chunk 1), (words +8, chunk 0), (words +8, chunk 1)
        const int word = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = (lane >> 3) & 1;
        ldmatrix_x4(kf, sk + word * pitch + (cb + kc + chunk * 8) * 2);
        mma_bf16_16816(s[np * 2 + 0], a, kf[0], kf[1]);
        mma_bf16_16816(s[np * 2 + 1], a, kf[2], kf[3]);
      }
    }
    // masked softmax over the words (exact: all words are in registers)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float ma = s_madd[nt * 8 + 2 * t], mb = s_madd[nt * 8 + 2 * t + 1];
      s[nt][0] += ma; s[nt][1] += mb; s[nt][2] += ma; s[nt][3] += mb;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[KS][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float p0v = __expf(s[nt][0] - mx0), p1v = __expf(s[nt][1] - mx0);
      const float p2v = __expf(s[nt][2] - mx1), p3v = __expf(s[nt][3] - mx1);
      l0 += p0v + p1v;
      l1 += p2v + p3v;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0v, p1v);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2v, p3v);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // O = P v for this head's channels, 16 channels (two n8 tiles) at a time
    for (int oc = 0; oc < ch; oc += 16) {
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        uint32_t vf[4];
        // matrices: (words kk*16 + 0..7, chan chunk 0), (words +8, chunk 0), (words 0..7, chunk 1), (words +8, chunk 1)
        const int word = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldmatrix_x4_trans(vf, sv + word * pitch + (cb + oc + (lane >> 4) * 8) * 2);
        mma_bf16_16816(acc0, pf[kk], vf[0], vf[1]);
        mma_bf16_16816(acc1, pf[kk], vf[2], vf[3]);
      }
      const int c = cb + oc + 2 * t;
      if (w0) {
        *reinterpret_cast<uint32_t*>(o0 + c) = pack_bf16x2(acc0[0] * i0, acc0[1] * i0);
        *reinterpret_cast<uint32_t*>(o0 + c + 8) = pack_bf16x2(acc1[0] * i0, acc1[1] * i0);
      }
      if (w1) {
        *reinterpret_cast<uint32_t*>(o1 + c) = pack_bf16x2(acc0[2] * i1, acc0[3] * i1);
        *reinterpret_cast<uint32_t*>(o1 + c + 8) = pack_bf16x2(acc1[2] * i1, acc1[3] * i1);
      }
    }
  }
}

template <int NT>
static int launch_pwam_mma(const float* qpre, const float* stats, const float* k, const float* v, const float* mask,
                           __nv_bfloat16* o, int B, long long n, int C, int Nl, int heads, float scale, size_t smem, cudaStream_t st) {
  auto kfn = pwam_core_mma_kernel<NT>;
  static size_t configured = 0;
  if (smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  // 16 pixels per warp; 8 warps per block while that still gives every SM two blocks, fewer for short clips (stage 3 of the bench: 1152
  // tokens per clip ran as 72 blocks of 128 pixels, 77 us)
  int warps = 8;
  while (warps > 2 && static_cast<long long>(B) * ((n + 16 * warps - 1) / (16 * warps)) < 296) warps >>= 1;
  const long long blocks = (n + 16 * warps - 1) / (16 * warps);
  kfn<<<dim3(static_cast<unsigned>(blocks), B), 32 * warps, smem, st>>>(qpre, stats, k, v, mask, o, n, C, Nl, heads, scale);
  LAVT_LAUNCH_CHECK("pwam_core_mma_kernel");
  return LAVT_OK;
}

int pwam_core_dispatch(const float* qpre, const float* stats, const float* k, const float* v, const float* mask,
                       __nv_bfloat16* o, int B, long long n, int C, int Nl, int heads, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && n > 0, "pwam_core: empty input");
  LAVT_REQUIRE(Nl >= 1 && Nl <= 4096, "pwam_core: Nl=%d out of range", Nl);
  LAVT_REQUIRE(heads >= 1 && heads <= 32 && (32 % heads) == 0, "pwam_core: fusion heads=%d must divide 32", heads);
  const float scale = 1.0f / sqrtf(static_cast<float>(C));
  {
    // tensor-core path when the words fit in shared memory and the head width is MMA-friendly
    const int nlp = Nl <= 32 ? 32 : (Nl <= 48 ? 48 : (Nl <= 80 ? 80 : (Nl <= 128 ? 128 : 0)));
    const size_t smem = nlp ? (2 * static_cast<size_t>(nlp) * (C * 2 + 16) + 2 * C * sizeof(float) + nlp * sizeof(float)) : 0;
    if (nlp && smem <= 160 * 1024 && (C / heads) % 16 == 0 && C % 16 == 0) {
      switch (nlp) {
        case 32: return launch_pwam_mma<4>(qpre, stats, k, v, mask, o, B, n, C, Nl, heads, scale, smem, st);
        case 48: return launch_pwam_mma<6>(qpre, stats, k, v, mask, o, B, n, C, Nl, heads, scale, smem, st);
        case 80: return launch_pwam_mma<10>(qpre, stats, k, v, mask, o, B, n, C, Nl, heads, scale, smem, st);
        default: return launch_pwam_mma<16>(qpre, stats, k, v, mask, o, B, n, C, Nl, heads, scale, smem, st);
      }
    }
  }
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
