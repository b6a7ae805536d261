// Backward of PWAM + LanguageGate (reference PWAM.forward lib/video_swin_transformer.py:919-934,
// SpatialImageLanguageAttention.forward :975-1009, res_gate :519-525 applied at :570; differentiated by autograd in the
// reference, train.py:330-360).  The dense parts (1x1 convs, gate linears) go through the tcgen05 GEMM (dgrad) and the split-K
// GEMM (wgrad); these kernels are the adjoints of everything in between:
//   pwam_attend_bwd   per pixel: recompute q^ = IN(q_pre), the masked softmax over words and dP = dO v^T; emits d q^, the bf16
//                     rows of P and dS (block-diagonal over clips / fusion heads) that turn dk = dS^T q^ and dv = P^T dO into
//                     ordinary weight-gradient GEMMs, and the two InstanceNorm reductions sum(dq^), sum(dq^ q^)
//   pwam_mul_bwd      a2 = vis * IN(lang_pre): d vis_pre (through GELU') and the InstanceNorm reductions of lang
//   instnorm_bwd      d x_pre = rstd (g - mean g - x^ mean(g x^)) from the accumulated reductions
//   pwam_kv_bwd       k, v = (W l + b) * mask: dW, db, d l
//   gate kernels      x' = x + tanh(.) * r: elementwise adjoints (tanh', relu mask), GELU with an fp32 copy / fp32 gradient in
#include "kernels.cuh"

#include <cstdlib>

namespace lavt {

__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// CPL consecutive elements per lane: 128-bit accesses when CPL is a multiple of 4 (C = 128 .. 1024), scalar otherwise (the Swin-T/S
// widths 96 / 192: CPL = 3 / 6)
template <int CPL>
__device__ __forceinline__ void ld_f32(float (&o)[CPL], const float* __restrict__ p) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(p + i));
      o[i] = u.x; o[i + 1] = u.y; o[i + 2] = u.z; o[i + 3] = u.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CPL; ++i) o[i] = __ldg(p + i);
  }
}
template <int CPL>
__device__ __forceinline__ void ld_bf16(float (&o)[CPL], const __nv_bfloat16* __restrict__ p) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p + i));
      const float2 a = unpack_bf16x2(w.x), c = unpack_bf16x2(w.y);
      o[i] = a.x; o[i + 1] = a.y; o[i + 2] = c.x; o[i + 3] = c.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CPL; ++i) o[i] = __bfloat162float(p[i]);
  }
}
template <int CPL>
__device__ __forceinline__ float dot_row(const float (&a)[CPL], const float* __restrict__ p) {
  float k[CPL];
  ld_f32<CPL>(k, p);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) s = fmaf(a[i], k[i], s);
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per pixel (grid-strided); lane owns CPL = C / 32 contiguous channels, each fusion head's channels sit in a group
// of 32 / heads lanes (same layout as pwam_core_kernel).  Per-warp shared memory: s, P, dP, dS for heads x NlPad words.
template <int CPL>
__global__ void __launch_bounds__(256) pwam_attend_bwd_kernel(const float* __restrict__ qpre, const float* __restrict__ stats,
                                                              const float* __restrict__ k, const float* __restrict__ v,
                                                              const float* __restrict__ mask, const __nv_bfloat16* __restrict__ dO,
                                                              float* __restrict__ dqhat, __nv_bfloat16* __restrict__ qs_out,
                                                              __nv_bfloat16* __restrict__ P_bd, __nv_bfloat16* __restrict__ dS_bd,
                                                              float* __restrict__ sums, int B, long long n, int Nl, int NlPad, int heads,
                                                              float scale) {
  extern __shared__ float pab_smem[];
  constexpr int C = CPL * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int HW = heads * NlPad;
  float* sS = pab_smem + warp * 4 * HW;
  float* sP = sS + HW;
  float* sdP = sP + HW;
  float* sdS = sdP + HW;
  for (int i = lane; i < 4 * HW; i += 32) sS[i] = 0.f;
  __syncwarp();
  const int gl = 32 / heads;
  const int myhead = lane / gl;
  const bool leader = (lane % gl) == 0;
  const int c0 = lane * CPL;
  const int Wd = B * HW;
  float mu[CPL], rs[CPL], s1[CPL], s2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    mu[i] = __ldg(stats + (static_cast<long long>(b) * 2) * C + c0 + i);
    rs[i] = __ldg(stats + (static_cast<long long>(b) * 2 + 1) * C + c0 + i);
    s1[i] = 0.f;
    s2[i] = 0.f;
  }
  const float* kb = k + static_cast<long long>(b) * Nl * C + c0;
  const float* vb = v + static_cast<long long>(b) * Nl * C + c0;
  const float* mb = mask + b * Nl;
  float* hs = sS + myhead * NlPad;
  float* hp = sP + myhead * NlPad;
  float* hdp = sdP + myhead * NlPad;
  float* hds = sdS + myhead * NlPad;
  const long long total_warps = static_cast<long long>(gridDim.x) * 8;
  for (long long p = static_cast<long long>(blockIdx.x) * 8 + warp; p < n; p += total_warps) {
    const long long row = static_cast<long long>(b) * n + p;
    float qh[CPL], d[CPL], dq[CPL];
    ld_f32<CPL>(qh, qpre + row * C + c0);
    ld_bf16<CPL>(d, dO + row * C + c0);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      qh[i] = (qh[i] - mu[i]) * rs[i];
      dq[i] = 0.f;
    }
    // pass 1: scores (kept in shared memory), running max / denominator per head
    float mx = -INFINITY, den = 0.f;
    for (int j = 0; j < Nl; ++j) {
      float s = dot_row<CPL>(qh, kb + static_cast<long long>(j) * C);
      for (int off = gl >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      s = s * scale + (1e4f * __ldg(mb + j) - 1e4f);
      if (leader) hs[j] = s;
      const float nm = fmaxf(mx, s);
      den = den * __expf(mx - nm) + __expf(s - nm);
      mx = nm;
    }
    __syncwarp();
    const float inv = 1.0f / den;
    // pass 2: P, dP = dO . v_j, D = sum_j P dP
    float Dsum = 0.f;
    for (int j = 0; j < Nl; ++j) {
      float dp = dot_row<CPL>(d, vb + static_cast<long long>(j) * C);
      for (int off = gl >> 1; off > 0; off >>= 1) dp += __shfl_xor_sync(0xffffffffu, dp, off);
      const float pj = __expf(hs[j] - mx) * inv;
      Dsum = fmaf(pj, dp, Dsum);
      if (leader) { hp[j] = pj; hdp[j] = dp; }
    }
    __syncwarp();
    // pass 3: dS = P (dP - D), d q^ = scale * dS k
    for (int j = 0; j < Nl; ++j) {
      const float ds = hp[j] * (hdp[j] - Dsum);
      if (leader) hds[j] = ds;
      float kk[CPL];
      ld_f32<CPL>(kk, kb + static_cast<long long>(j) * C);
#pragma unroll
      for (int i = 0; i < CPL; ++i) dq[i] = fmaf(ds, kk[i], dq[i]);
    }
    __syncwarp();
    {
      float* dst = dqhat + row * C + c0;
      __nv_bfloat16* qdst = qs_out + row * C + c0;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const float o = dq[i] * scale;
        dq[i] = o;
        s1[i] += o;
        s2[i] = fmaf(o, qh[i], s2[i]);
      }
      if constexpr (CPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CPL; i += 4) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
          *reinterpret_cast<uint2*>(qdst + i) = make_uint2(pack_bf16x2(qh[i] * scale, qh[i + 1] * scale), pack_bf16x2(qh[i + 2] * scale, qh[i + 3] * scale));
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          dst[i] = dq[i];
          qdst[i] = __float2bfloat16(qh[i] * scale);
        }
      }
    }
    // block-diagonal rows: columns of clip b carry P / dS, all other clips zero
    for (int c2 = lane * 2; c2 < Wd; c2 += 64) {
      const int blk = c2 / HW, w = c2 - blk * HW;
      uint32_t pv = 0u, zv = 0u;
      if (blk == b) {
        pv = pack_bf16x2(sP[w], sP[w + 1]);
        zv = pack_bf16x2(sdS[w], sdS[w + 1]);
      }
      *reinterpret_cast<uint32_t*>(P_bd + row * Wd + c2) = pv;
      *reinterpret_cast<uint32_t*>(dS_bd + row * Wd + c2) = zv;
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    atomicAdd(sums + (static_cast<long long>(b) * 2) * C + c0 + i, s1[i]);
    atomicAdd(sums + (static_cast<long long>(b) * 2 + 1) * C + c0 + i, s2[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same adjoint for ONE fusion head and at most 32 (padded) words -- the default configuration (20-token expressions, --mha unset).
// The kernel above reduces every one of the Nl scores / dP values across the warp separately (2 x Nl x 5 shuffle steps per pixel: at
// C = 128 the shuffles and the per-word scalar code were 2/3 of its instructions).  Here every lane accumulates the partial dot
// products of ALL words over its own channels, and one transposing butterfly (31 shuffles for 32 values) leaves the total of word j
// in lane j; softmax, D = sum_j P_j dP_j and dS then are one value per lane with a warp max / sum.
template <int CPL>
__device__ __forceinline__ float words_to_lanes(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

template <int CPL, int NLP>
__global__ void __launch_bounds__(256, 2) pwam_attend_bwd_h1_kernel(const float* __restrict__ qpre, const float* __restrict__ stats,
                                                                 const float* __restrict__ k, const float* __restrict__ v,
                                                                 const float* __restrict__ mask, const __nv_bfloat16* __restrict__ dO,
                                                                 float* __restrict__ dqhat, __nv_bfloat16* __restrict__ qs_out,
                                                                 __nv_bfloat16* __restrict__ P_bd, __nv_bfloat16* __restrict__ dS_bd,
                                                                 float* __restrict__ sums, int B, long long n, int Nl, int NlPad,
                                                                 float scale) {
  __shared__ float sPZ[8][2][32];
  constexpr int C = CPL * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c0 = lane * CPL;
  const int Wd = B * NlPad;
  float mu[CPL], rs[CPL], s1[CPL], s2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    mu[i] = __ldg(stats + (static_cast<long long>(b) * 2) * C + c0 + i);
    rs[i] = __ldg(stats + (static_cast<long long>(b) * 2 + 1) * C + c0 + i);
    s1[i] = 0.f;
    s2[i] = 0.f;
  }
  const float* kb = k + static_cast<long long>(b) * Nl * C + c0;
  const float* vb = v + static_cast<long long>(b) * Nl * C + c0;
  const float mterm = (lane < Nl) ? (1e4f * __ldg(mask + b * Nl + lane) - 1e4f) : -INFINITY;      // lanes past the last word drop out
  float* sP = sPZ[warp][0];
  float* sZ = sPZ[warp][1];
  const long long total_warps = static_cast<long long>(gridDim.x) * 8;
  for (long long p = static_cast<long long>(blockIdx.x) * 8 + warp; p < n; p += total_warps) {
    const long long row = static_cast<long long>(b) * n + p;
    float qh[CPL], d[CPL], dq[CPL];
    ld_f32<CPL>(qh, qpre + row * C + c0);
    ld_bf16<CPL>(d, dO + row * C + c0);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      qh[i] = (qh[i] - mu[i]) * rs[i];
      dq[i] = 0.f;
    }
    float sj, dpj;
    {
      float ps[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ps[j] = (j < NLP && j < Nl) ? dot_row<CPL>(qh, kb + static_cast<long long>(j) * C) : 0.f;
      sj = words_to_lanes<CPL>(ps, lane) * scale + mterm;       // score of word `lane`
    }
    {
      float pd[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pd[j] = (j < NLP && j < Nl) ? dot_row<CPL>(d, vb + static_cast<long long>(j) * C) : 0.f;
      dpj = words_to_lanes<CPL>(pd, lane);                      // dP of word `lane`
    }
    float mx = sj;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float e = (lane < Nl) ? __expf(sj - mx) : 0.f;
    const float pj = e / warp_sum(e);
    const float Dsum = warp_sum(pj * dpj);
    const float dsj = pj * (dpj - Dsum);
    sP[lane] = pj;
    sZ[lane] = dsj;
    // d q^ = scale * sum_j dS_j k_j
#pragma unroll
    for (int j = 0; j < NLP; ++j) {
      if (j < Nl) {
        const float ds = __shfl_sync(0xffffffffu, dsj, j);
        float kk[CPL];
        ld_f32<CPL>(kk, kb + static_cast<long long>(j) * C);
#pragma unroll
        for (int i = 0; i < CPL; ++i) dq[i] = fmaf(ds, kk[i], dq[i]);
      }
    }
    {
      float* dst = dqhat + row * C + c0;
      __nv_bfloat16* qdst = qs_out + row * C + c0;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const float o = dq[i] * scale;
        dq[i] = o;
        s1[i] += o;
        s2[i] = fmaf(o, qh[i], s2[i]);
      }
      if constexpr (CPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CPL; i += 4) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
          *reinterpret_cast<uint2*>(qdst + i) = make_uint2(pack_bf16x2(qh[i] * scale, qh[i + 1] * scale), pack_bf16x2(qh[i + 2] * scale, qh[i + 3] * scale));
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          dst[i] = dq[i];
          qdst[i] = __float2bfloat16(qh[i] * scale);
        }
      }
    }
    __syncwarp();
    // block-diagonal rows: columns of clip b carry P / dS, all other clips (and the padded words) zero
    for (int c2 = lane * 2; c2 < Wd; c2 += 64) {
      const int blk = c2 / NlPad, w = c2 - blk * NlPad;
      uint32_t pv = 0u, zv = 0u;
      if (blk == b) {
        pv = pack_bf16x2(sP[w], sP[w + 1]);
        zv = pack_bf16x2(sZ[w], sZ[w + 1]);
      }
      *reinterpret_cast<uint32_t*>(P_bd + row * Wd + c2) = pv;
      *reinterpret_cast<uint32_t*>(dS_bd + row * Wd + c2) = zv;
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    atomicAdd(sums + (static_cast<long long>(b) * 2) * C + c0 + i, s1[i]);
    atomicAdd(sums + (static_cast<long long>(b) * 2 + 1) * C + c0 + i, s2[i]);
  }
}

int pwam_attend_bwd_dispatch(const float* qpre, const float* stats, const float* k, const float* v, const float* mask,
                             const __nv_bfloat16* dO, float* dqhat, __nv_bfloat16* qs_out, __nv_bfloat16* P_bd, __nv_bfloat16* dS_bd,
                             float* sums, int B, long long n, int C, int Nl, int NlPad, int heads, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && n > 0 && Nl > 0, "pwam backward: empty input");
  LAVT_REQUIRE(heads >= 1 && heads <= 32 && (32 % heads) == 0, "pwam backward: fusion heads=%d must divide 32", heads);
  LAVT_REQUIRE(NlPad >= Nl && NlPad % 8 == 0, "pwam backward: padded word count %d invalid for %d words", NlPad, Nl);
  const float scale = 1.0f / sqrtf(static_cast<float>(C));
  const size_t smem = static_cast<size_t>(8) * 4 * heads * NlPad * sizeof(float);
  LAVT_REQUIRE(smem <= 200 * 1024, "pwam backward: %d heads x %d words do not fit in shared memory", heads, NlPad);
  long long gx = (n + 7) / 8;
  if (gx > 148 * 2) gx = 148 * 2;
  dim3 grid(static_cast<unsigned>(gx), B);
  // one fusion head, <= 32 words, C = 128: word-per-lane kernel (measured per 4-clip step: stage 0 7.69 -> 7.11 ms; at C = 256 the
  // 128-register budget spills and stage 1 went 4.70 -> 5.10 ms, at C >= 512 the dot products dominate either way)
  static const bool h1_off = getenv("LAVT_PWAM_BWD_H1") && atoi(getenv("LAVT_PWAM_BWD_H1")) == 0;
  static const bool h1_wide = getenv("LAVT_PWAM_BWD_H1") && atoi(getenv("LAVT_PWAM_BWD_H1")) == 2;
  if (heads == 1 && NlPad <= 32 && (C == 128 || (C == 256 && h1_wide)) && !h1_off) {
    long long g1 = (n + 7) / 8;
    if (g1 > 148 * 4) g1 = 148 * 4;
    dim3 grid1(static_cast<unsigned>(g1), B);
#define LAVT_PAB_H1(cpl, nlp) pwam_attend_bwd_h1_kernel<cpl, nlp><<<grid1, 256, 0, st>>>(qpre, stats, k, v, mask, dO, dqhat, qs_out, P_bd, dS_bd, sums, B, n, Nl, NlPad, scale)
    if (C == 128) { if (NlPad <= 24) LAVT_PAB_H1(4, 24); else LAVT_PAB_H1(4, 32); }
    else { if (NlPad <= 24) LAVT_PAB_H1(8, 24); else LAVT_PAB_H1(8, 32); }
#undef LAVT_PAB_H1
    LAVT_LAUNCH_CHECK("pwam_attend_bwd_h1_kernel");
    return LAVT_OK;
  }
#define LAVT_PAB_CASE(cpl)                                                                                             \
  case cpl * 32: {                                                                                                     \
    if (smem > 48 * 1024)                                                                                              \
      LAVT_CUDA(cudaFuncSetAttribute(pwam_attend_bwd_kernel<cpl>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    pwam_attend_bwd_kernel<cpl><<<grid, 256, smem, st>>>(qpre, stats, k, v, mask, dO, dqhat, qs_out, P_bd, dS_bd, sums, B, n, Nl, NlPad, \
                                                         heads, scale);                                                \
    break;                                                                                                             \
  }
  switch (C) {
    LAVT_PAB_CASE(3)
    LAVT_PAB_CASE(4)
    LAVT_PAB_CASE(6)
    LAVT_PAB_CASE(8)
    LAVT_PAB_CASE(12)
    LAVT_PAB_CASE(16)
    LAVT_PAB_CASE(24)
    LAVT_PAB_CASE(32)
    default:
      set_last_error("pwam backward: C=%d unsupported (need 96/128/192/256/384/512/768/1024)", C);
      return LAVT_ERR_SHAPE;
  }
#undef LAVT_PAB_CASE
  LAVT_LAUNCH_CHECK("pwam_attend_bwd_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a2 = vis * IN(lang_pre), vis = GELU(vis_pre):  d vis_pre = da2 * lang * GELU'(vis_pre);  g = da2 * vis is the gradient of lang,
// whose InstanceNorm reductions sum(g), sum(g * lang) accumulate into sums [B,2,C].  grid (chunks of 256 rows, B).
__global__ void __launch_bounds__(256) pwam_mul_bwd_kernel(const __nv_bfloat16* __restrict__ da2, const __nv_bfloat16* __restrict__ vis,
                                                           const __nv_bfloat16* __restrict__ vispre, const float* __restrict__ langpre,
                                                           const float* __restrict__ stats, __nv_bfloat16* __restrict__ dvispre,
                                                           float* __restrict__ sums, int n, int C) {
  extern __shared__ float pmb_sm[];       // [rg][C][2]
  const int b = blockIdx.y;
  const int tpr = C / 4, rg = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2) * C) + tc);
  const float4 rs = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2 + 1) * C) + tc);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int r0 = blockIdx.x * 256, r1 = min(n, r0 + 256);
  if (tr < rg) {
    for (int r = r0 + tr; r < r1; r += rg) {
      const long long e = (static_cast<long long>(b) * n + r) * C + tc * 4;
      const uint2 ua = __ldg(reinterpret_cast<const uint2*>(da2 + e)), uv = __ldg(reinterpret_cast<const uint2*>(vis + e));
      const uint2 up = __ldg(reinterpret_cast<const uint2*>(vispre + e));
      const float4 lp = __ldg(reinterpret_cast<const float4*>(langpre + e));
      const float2 a0 = unpack_bf16x2(ua.x), a1 = unpack_bf16x2(ua.y), v0 = unpack_bf16x2(uv.x), v1 = unpack_bf16x2(uv.y);
      const float2 p0 = unpack_bf16x2(up.x), p1 = unpack_bf16x2(up.y);
      const float l0 = (lp.x - mu.x) * rs.x, l1 = (lp.y - mu.y) * rs.y, l2 = (lp.z - mu.z) * rs.z, l3 = (lp.w - mu.w) * rs.w;
      const float g0 = a0.x * v0.x, g1 = a0.y * v0.y, g2 = a1.x * v1.x, g3 = a1.y * v1.y;
      s1.x += g0; s1.y += g1; s1.z += g2; s1.w += g3;
      s2.x += g0 * l0; s2.y += g1 * l1; s2.z += g2 * l2; s2.w += g3 * l3;
      *reinterpret_cast<uint2*>(dvispre + e) = make_uint2(pack_bf16x2(a0.x * l0 * gelu_grad(p0.x), a0.y * l1 * gelu_grad(p0.y)),
                                                          pack_bf16x2(a1.x * l2 * gelu_grad(p1.x), a1.y * l3 * gelu_grad(p1.y)));
    }
    float* o = pmb_sm + (static_cast<long long>(tr) * C + tc * 4) * 2;
    o[0] = s1.x; o[1] = s2.x; o[2] = s1.y; o[3] = s2.y; o[4] = s1.z; o[5] = s2.z; o[6] = s1.w; o[7] = s2.w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int g = 0; g < rg; ++g) {
      a += pmb_sm[(g * C + c) * 2 + 0];
      q += pmb_sm[(g * C + c) * 2 + 1];
    }
    atomicAdd(sums + (static_cast<long long>(b) * 2) * C + c, a);
    atomicAdd(sums + (static_cast<long long>(b) * 2 + 1) * C + c, q);
  }
}

int pwam_mul_bwd_dispatch(const __nv_bfloat16* da2, const __nv_bfloat16* vis, const __nv_bfloat16* vispre, const float* langpre,
                          const float* stats, __nv_bfloat16* dvispre, float* sums, int B, long long n, int C, cudaStream_t st) {
  LAVT_REQUIRE(C % 4 == 0 && C <= 1024, "pwam mul backward: C=%d unsupported (need a multiple of 4 up to 1024)", C);
  LAVT_REQUIRE(B > 0 && n > 0 && n < (1LL << 30), "pwam mul backward: bad sizes");
  const int rg = 256 / (C / 4);
  pwam_mul_bwd_kernel<<<dim3(static_cast<unsigned>((n + 255) / 256), B), 256, static_cast<size_t>(rg) * C * 2 * sizeof(float), st>>>(
      da2, vis, vispre, langpre, stats, dvispre, sums, static_cast<int>(n), C);
  LAVT_LAUNCH_CHECK("pwam_mul_bwd_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// InstanceNorm reductions of an fp32 gradient g w.r.t. x^ = IN(x_pre): sums [B,2,C] += (sum_n g, sum_n g * x^).  SepTPWAM sums two
// InstanceNorm'd branches (lib/video_swin_transformer.py:1512-1524, 1556-1561): the same g flows into both, each needs its own
// reductions.  grid (chunks of 256 rows, B).
__global__ void __launch_bounds__(256) instnorm_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ xpre,
                                                                  const float* __restrict__ stats, float* __restrict__ sums, int n, int C) {
  extern __shared__ float ibr_sm[];       // [rg][C][2]
  const int b = blockIdx.y;
  const int tpr = C / 4, rg = blockDim.x / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int r0 = blockIdx.x * 256, r1 = min(n, r0 + 256);
  if (tr < rg) {
    const float4 mu = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2) * C) + tc);
    const float4 rs = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2 + 1) * C) + tc);
    for (int r = r0 + tr; r < r1; r += rg) {
      const long long e = (static_cast<long long>(b) * n + r) * C + tc * 4;
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + e)), x = __ldg(reinterpret_cast<const float4*>(xpre + e));
      s1.x += gv.x; s1.y += gv.y; s1.z += gv.z; s1.w += gv.w;
      s2.x += gv.x * (x.x - mu.x) * rs.x; s2.y += gv.y * (x.y - mu.y) * rs.y;
      s2.z += gv.z * (x.z - mu.z) * rs.z; s2.w += gv.w * (x.w - mu.w) * rs.w;
    }
    float* o = ibr_sm + (static_cast<long long>(tr) * C + tc * 4) * 2;
    o[0] = s1.x; o[1] = s2.x; o[2] = s1.y; o[3] = s2.y; o[4] = s1.z; o[5] = s2.z; o[6] = s1.w; o[7] = s2.w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int k = 0; k < rg; ++k) {
      a += ibr_sm[(k * C + c) * 2 + 0];
      q += ibr_sm[(k * C + c) * 2 + 1];
    }
    atomicAdd(sums + (static_cast<long long>(b) * 2) * C + c, a);
    atomicAdd(sums + (static_cast<long long>(b) * 2 + 1) * C + c, q);
  }
}

int instnorm_bwd_reduce_dispatch(const float* g, const float* xpre, const float* stats, float* sums, int B, long long n, int C,
                                 cudaStream_t st) {
  LAVT_REQUIRE(C % 4 == 0 && C <= 1024 && B > 0 && n > 0 && n < (1LL << 30), "instance-norm backward reduce: bad sizes");
  const int rg = 256 / (C / 4);
  instnorm_bwd_reduce_kernel<<<dim3(static_cast<unsigned>((n + 255) / 256), B), 256, static_cast<size_t>(rg) * C * 2 * sizeof(float), st>>>(
      g, xpre, stats, sums, static_cast<int>(n), C);
  LAVT_LAUNCH_CHECK("instnorm_bwd_reduce_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// InstanceNorm backward from the accumulated reductions: out = rstd * (g - S1/n - x^ * S2/n); g = g_f32, or ga * gb (bf16)
__global__ void __launch_bounds__(256) instnorm_bwd_kernel(const float* __restrict__ g32, const __nv_bfloat16* __restrict__ ga,
                                                           const __nv_bfloat16* __restrict__ gb, const float* __restrict__ xpre,
                                                           const float* __restrict__ stats, const float* __restrict__ sums,
                                                           __nv_bfloat16* __restrict__ out, long long n, int C, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % (C / 4));
  const long long row = i / (C / 4);
  const int b = static_cast<int>(row / n);
  const float inv_n = 1.0f / static_cast<float>(n);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2) * C) + c4);
  const float4 rs = __ldg(reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * 2 + 1) * C) + c4);
  const float4 S1 = __ldg(reinterpret_cast<const float4*>(sums + (static_cast<long long>(b) * 2) * C) + c4);
  const float4 S2 = __ldg(reinterpret_cast<const float4*>(sums + (static_cast<long long>(b) * 2 + 1) * C) + c4);
  const float4 x = __ldg(reinterpret_cast<const float4*>(xpre) + i);
  float4 g;
  if (g32) {
    g = __ldg(reinterpret_cast<const float4*>(g32) + i);
  } else {
    const uint2 ua = __ldg(reinterpret_cast<const uint2*>(ga) + i), ub = __ldg(reinterpret_cast<const uint2*>(gb) + i);
    const float2 a0 = unpack_bf16x2(ua.x), a1 = unpack_bf16x2(ua.y), b0 = unpack_bf16x2(ub.x), b1 = unpack_bf16x2(ub.y);
    g = make_float4(a0.x * b0.x, a0.y * b0.y, a1.x * b1.x, a1.y * b1.y);
  }
  const float h0 = (x.x - mu.x) * rs.x, h1 = (x.y - mu.y) * rs.y, h2 = (x.z - mu.z) * rs.z, h3 = (x.w - mu.w) * rs.w;
  const float o0 = rs.x * (g.x - S1.x * inv_n - h0 * S2.x * inv_n), o1 = rs.y * (g.y - S1.y * inv_n - h1 * S2.y * inv_n);
  const float o2 = rs.z * (g.z - S1.z * inv_n - h2 * S2.z * inv_n), o3 = rs.w * (g.w - S1.w * inv_n - h3 * S2.w * inv_n);
  reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
}

int instnorm_bwd_dispatch(const float* g32, const __nv_bfloat16* ga, const __nv_bfloat16* gb, const float* xpre, const float* stats,
                          const float* sums, __nv_bfloat16* out, int B, long long n, int C, cudaStream_t st) {
  LAVT_REQUIRE(C % 4 == 0 && B > 0 && n > 0, "instance-norm backward: bad sizes");
  LAVT_REQUIRE(g32 || (ga && gb), "instance-norm backward: no gradient input");
  const long long total4 = static_cast<long long>(B) * n * (C / 4);
  instnorm_bwd_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(g32, ga, gb, xpre, stats, sums, out, n, C, total4);
  LAVT_LAUNCH_CHECK("instnorm_bwd_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// k, v = (W l + b) * mask.  dkbuf / dvbuf fp32 [(b * heads + h) * NlPad + j, C]: the row of head h = channel's head is the valid one.
// Two small tiled contractions over G[(b, j), c] = mask[b, j] * d{k,v}[b, head(c), j, c]  (the first version walked all (clip, word)
// pairs / all channels in ONE thread per output element with strided loads: 330 us per launch for 60-250 MFLOP):
//   pwam_kv_wgrad_kernel   dW{k,v}[c, i] += sum_(b,j) G[(b,j), c] * l[b, i, j],  db{k,v}[c] += sum_(b,j) G[(b,j), c]
//                          block = 64 channels x 64 language channels, (b, j) staged 32 words at a time
//   pwam_kv_dl_kernel      dl[b, i, j] += mask[b, j] * sum_c (dk[.., c] * Wk[c, i] + dv[.., c] * Wv[c, i])
//                          block = 16 (b, j) rows x 64 language channels; the 2C-long contraction is split over gridDim.z
//                          (fp32 atomic adds into dl, like the LayerNorm gamma / beta gradients)
__global__ void __launch_bounds__(256) pwam_kv_wgrad_kernel(const float* __restrict__ dkbuf, const float* __restrict__ dvbuf,
                                                            const float* __restrict__ mask, const float* __restrict__ l,
                                                            float* __restrict__ dwk, float* __restrict__ dbk, float* __restrict__ dwv,
                                                            float* __restrict__ dbv, int B, int Nl, int NlPad, int Lin, int C, int heads) {
  __shared__ float sGk[32][64], sGv[32][64], sL[32][65];
  const int t = threadIdx.x;
  const int i0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tc = t & 15, ti = t >> 4;           // 4 channels x 4 language channels per thread
  const int ch = C / heads;
  float ak[4][4], av[4][4], bk[4], bv[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    bk[a] = 0.f;
    bv[a] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { ak[a][q] = 0.f; av[a][q] = 0.f; }
  }
  for (int b = 0; b < B; ++b) {
    for (int j0 = 0; j0 < Nl; j0 += 32) {
      const int jn = min(32, Nl - j0);
      __syncthreads();
      for (int e = t; e < 32 * 64; e += 256) {
        const int kk = e >> 6, c = c0 + (e & 63);
        float gk = 0.f, gv = 0.f;
        if (kk < jn && c < C) {
          const float m = __ldg(mask + b * Nl + j0 + kk);
          if (m != 0.f) {
            const long long r = (static_cast<long long>(b) * heads + c / ch) * NlPad + j0 + kk;
            gk = __ldg(dkbuf + r * C + c) * m;
            gv = __ldg(dvbuf + r * C + c) * m;
          }
        }
        sGk[kk][e & 63] = gk;
        sGv[kk][e & 63] = gv;
      }
      for (int e = t; e < 64 * jn; e += 256) {        // l[b, i0 + i, j0 + kk]: runs of jn consecutive floats
        const int i = e / jn, kk = e - i * jn;
        sL[kk][i] = (i0 + i < Lin) ? __ldg(l + (static_cast<long long>(b) * Lin + i0 + i) * Nl + j0 + kk) : 0.f;
      }
      __syncthreads();
      for (int kk = 0; kk < jn; ++kk) {
        const float4 gk = *reinterpret_cast<const float4*>(&sGk[kk][tc * 4]);
        const float4 gv = *reinterpret_cast<const float4*>(&sGv[kk][tc * 4]);
        const float g1[4] = {gk.x, gk.y, gk.z, gk.w}, g2[4] = {gv.x, gv.y, gv.z, gv.w};
        float lv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) lv[q] = sL[kk][ti * 4 + q];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          bk[a] += g1[a];
          bv[a] += g2[a];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ak[a][q] = fmaf(g1[a], lv[q], ak[a][q]);
            av[a][q] = fmaf(g2[a], lv[q], av[a][q]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int c = c0 + tc * 4 + a;
    if (c >= C) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ti * 4 + q;
      if (i >= Lin) continue;
      if (dwk) dwk[static_cast<long long>(c) * Lin + i] += ak[a][q];
      if (dwv) dwv[static_cast<long long>(c) * Lin + i] += av[a][q];
    }
    if (blockIdx.x == 0 && ti == 0) {
      if (dbk) dbk[c] += bk[a];
      if (dbv) dbv[c] += bv[a];
    }
  }
}

__global__ void __launch_bounds__(256) pwam_kv_dl_kernel(const float* __restrict__ dkbuf, const float* __restrict__ dvbuf,
                                                         const float* __restrict__ mask, const float* __restrict__ wk,
                                                         const float* __restrict__ wv, float* __restrict__ dl, int B, int Nl, int NlPad,
                                                         int Lin, int C, int heads) {
  __shared__ float sG[16][65];
  __shared__ __align__(16) float sW[64][64];
  const int t = threadIdx.x;
  const int i0 = blockIdx.x * 64, k0 = blockIdx.y * 16;
  const int ti = t & 15, tk = t >> 4;           // one (b, j) row x 4 language channels per thread
  const int ch = C / heads;
  const int KR = B * Nl;
  const int cchunks = (C + 63) / 64;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int cc = blockIdx.z; cc < 2 * cchunks; cc += gridDim.z) {
    const bool isv = cc >= cchunks;
    const int c0 = (isv ? cc - cchunks : cc) * 64;
    const float* g = isv ? dvbuf : dkbuf;
    const float* w = isv ? wv : wk;
    __syncthreads();
    for (int e = t; e < 16 * 64; e += 256) {
      const int kr = e >> 6, c = c0 + (e & 63), k = k0 + kr;
      float v = 0.f;
      if (k < KR && c < C) {
        const int b = k / Nl, j = k - b * Nl;
        v = __ldg(g + ((static_cast<long long>(b) * heads + c / ch) * NlPad + j) * C + c);
      }
      sG[kr][e & 63] = v;
    }
    for (int e = t; e < 64 * 64; e += 256) {
      const int cr = e >> 6, i = i0 + (e & 63), c = c0 + cr;
      sW[cr][e & 63] = (c < C && i < Lin) ? __ldg(w + static_cast<long long>(c) * Lin + i) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float gv = sG[tk][c];
      const float4 w4 = *reinterpret_cast<const float4*>(&sW[c][ti * 4]);
      acc[0] = fmaf(gv, w4.x, acc[0]);
      acc[1] = fmaf(gv, w4.y, acc[1]);
      acc[2] = fmaf(gv, w4.z, acc[2]);
      acc[3] = fmaf(gv, w4.w, acc[3]);
    }
  }
  const int k = k0 + tk;
  if (k >= KR) return;
  const int b = k / Nl, j = k - b * Nl;
  const float m = __ldg(mask + b * Nl + j);
  if (m == 0.f) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ti * 4 + q;
    if (i < Lin) atomicAdd(dl + (static_cast<long long>(b) * Lin + i) * Nl + j, acc[q] * m);
  }
}

int pwam_kv_bwd_dispatch(const float* dkbuf, const float* dvbuf, const float* mask, const float* l, const float* wk, const float* wv,
                         float* dwk, float* dbk, float* dwv, float* dbv, float* dl, int B, int Nl, int NlPad, int Lin, int C, int heads,
                         cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && Nl > 0 && Lin > 0 && C > 0 && heads > 0 && C % heads == 0, "pwam kv backward: bad sizes");
  if (dwk || dwv || dbk || dbv) {
    dim3 grid(static_cast<unsigned>((Lin + 63) / 64), static_cast<unsigned>((C + 63) / 64));
    pwam_kv_wgrad_kernel<<<grid, 256, 0, st>>>(dkbuf, dvbuf, mask, l, dwk, dbk, dwv, dbv, B, Nl, NlPad, Lin, C, heads);
    LAVT_LAUNCH_CHECK("pwam_kv_wgrad_kernel");
  }
  if (dl) {
    const int cchunks = (C + 63) / 64;
    dim3 grid(static_cast<unsigned>((Lin + 63) / 64), static_cast<unsigned>((B * Nl + 15) / 16), static_cast<unsigned>(min(4, 2 * cchunks)));
    pwam_kv_dl_kernel<<<grid, 256, 0, st>>>(dkbuf, dvbuf, mask, wk, wv, dl, B, Nl, NlPad, Lin, C, heads);
    LAVT_LAUNCH_CHECK("pwam_kv_dl_kernel");
  }
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// LangProject backward (--fuse simple, reference lib/video_swin_transformer.py:1012-1039): s = W2 relu(W0 mean + b0) + b2 with
// mean[b, :] = sum_j l[b, :, j] m[b, j] / sum_j m[b, j].  ds fp32 [B, C] arrives as the first row of pwam_mul_bwd's reductions
// (sum over the pixels of d a2 * vis).  Everything here is clip-sized (B x C, B x Lin): four small launches, the forward's mean / hidden
// vectors are recomputed.  workspace fp32: h [B, C] | dh [B, C] | mean [B, Lin] | dmean [B, Lin].
__global__ void __launch_bounds__(256) lang_project_bwd_hidden_kernel(const float* __restrict__ l, const float* __restrict__ mask,
                                                                      const float* __restrict__ w0, const float* __restrict__ b0,
                                                                      float* __restrict__ h, float* __restrict__ mean, int Nl, int Lin, int C) {
  extern __shared__ float lpb_sm[];      // [Lin] pooled sentence vector
  const int b = blockIdx.y;
  float cnt = 0.f;
  for (int j = 0; j < Nl; ++j) cnt += mask[b * Nl + j];
  for (int i = threadIdx.x; i < Lin; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Nl; ++j) s += l[(static_cast<long long>(b) * Lin + i) * Nl + j] * mask[b * Nl + j];
    lpb_sm[i] = s / cnt;
    if (blockIdx.x == 0) mean[static_cast<long long>(b) * Lin + i] = s / cnt;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  float acc = 0.f;
  for (int i = lane; i < Lin; i += 32) acc = fmaf(__ldg(w0 + static_cast<long long>(c) * Lin + i), lpb_sm[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) h[static_cast<long long>(b) * C + c] = fmaxf(acc + b0[c], 0.f);
}
// dW[r, q] += sum_b u[b, r] * v[b, q];  dbias[r] += sum_b u[b, r]        (u [B, R], v [B, Q])
__global__ void __launch_bounds__(256) outer_accumulate_kernel(const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ dW,
                                                               float* __restrict__ dbias, int B, int R, int Q) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(R) * Q) return;
  const int r = static_cast<int>(idx / Q), q = static_cast<int>(idx - static_cast<long long>(r) * Q);
  float acc = 0.f, bs = 0.f;
  for (int b = 0; b < B; ++b) {
    const float uv = __ldg(u + static_cast<long long>(b) * R + r);
    acc = fmaf(uv, __ldg(v + static_cast<long long>(b) * Q + q), acc);
    bs += uv;
  }
  if (dW) dW[idx] += acc;
  if (dbias && q == 0) dbias[r] += bs;
}
// out[b, q] = (sum_r W[r, q] * u[b, r]) * (gate == nullptr || gate[b, q] > 0)        (W [R, Q])
__global__ void __launch_bounds__(256) matvec_t_kernel(const float* __restrict__ W, const float* __restrict__ u, const float* __restrict__ gate,
                                                       float* __restrict__ out, int B, int R, int Q) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (q >= Q) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc = fmaf(__ldg(W + static_cast<long long>(r) * Q + q), __ldg(u + static_cast<long long>(b) * R + r), acc);
  if (gate && !(gate[static_cast<long long>(b) * Q + q] > 0.f)) acc = 0.f;
  out[static_cast<long long>(b) * Q + q] = acc;
}
// dl[b, i, j] += dmean[b, i] * m[b, j] / sum_j m[b, j]
__global__ void __launch_bounds__(256) lang_mean_bwd_kernel(const float* __restrict__ dmean, const float* __restrict__ mask, float* __restrict__ dl,
                                                            int Nl, int Lin, long long total) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int j = static_cast<int>(e % Nl);
  const long long bi = e / Nl;
  const int b = static_cast<int>(bi / Lin);
  float cnt = 0.f;
  for (int t = 0; t < Nl; ++t) cnt += mask[b * Nl + t];
  dl[e] += dmean[bi] * mask[b * Nl + j] / cnt;
}

int lang_project_bwd_dispatch(const float* l, const float* mask, const float* w0, const float* b0, const float* w2, const float* ds, float* dw0,
                              float* db0, float* dw2, float* db2, float* dl, float* workspace, int B, int Nl, int Lin, int C, cudaStream_t st) {
  LAVT_REQUIRE(B > 0 && B < 65536 && Nl > 0 && Lin > 0 && C > 0 && workspace, "lang_project backward: bad arguments");
  LAVT_REQUIRE(static_cast<size_t>(Lin) * sizeof(float) <= 48 * 1024, "lang_project backward: language width %d too large", Lin);
  float* h = workspace;
  float* dh = h + static_cast<long long>(B) * C;
  float* mean = dh + static_cast<long long>(B) * C;
  float* dmean = mean + static_cast<long long>(B) * Lin;
  lang_project_bwd_hidden_kernel<<<dim3((C + 7) / 8, B), 256, static_cast<size_t>(Lin) * sizeof(float), st>>>(l, mask, w0, b0, h, mean, Nl, Lin, C);
  LAVT_LAUNCH_CHECK("lang_project_bwd_hidden_kernel");
  // s = W2 h + b2
  outer_accumulate_kernel<<<static_cast<unsigned>((1LL * C * C + 255) / 256), 256, 0, st>>>(ds, h, dw2, db2, B, C, C);
  LAVT_LAUNCH_CHECK("outer_accumulate_kernel");
  matvec_t_kernel<<<dim3((C + 255) / 256, B), 256, 0, st>>>(w2, ds, h, dh, B, C, C);            // dh = (W2^T ds) * [h > 0]
  LAVT_LAUNCH_CHECK("matvec_t_kernel");
  // hpre = W0 mean + b0
  outer_accumulate_kernel<<<static_cast<unsigned>((1LL * C * Lin + 255) / 256), 256, 0, st>>>(dh, mean, dw0, db0, B, C, Lin);
  LAVT_LAUNCH_CHECK("outer_accumulate_kernel");
  if (dl) {
    matvec_t_kernel<<<dim3((Lin + 255) / 256, B), 256, 0, st>>>(w0, dh, nullptr, dmean, B, C, Lin);
    LAVT_LAUNCH_CHECK("matvec_t_kernel");
    const long long total = 1LL * B * Lin * Nl;
    lang_mean_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dmean, mask, dl, Nl, Lin, total);
    LAVT_LAUNCH_CHECK("lang_mean_bwd_kernel");
  }
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// LanguageGate elementwise pieces (x' = x + g2 * r, g2 = tanh(g1 G2^T), g1 = relu(r G0^T)); 8 elements per thread
// (the tanh PRE-activation is what is saved: 1 - tanh^2 recomputed from a bf16-rounded tanh output loses all precision near
//  saturation)
// mode 0: out_f32 = x + tanh(a) * r                                     (forward; a = gate pre-activation, b = r, f = x)
// mode 1: out_bf16 = f * r * (1 - tanh(a)^2);  out_f32 = f2 + f * tanh(a)  (f = dx', f2 = gradient of r so far or NULL)
// mode 2: out_bf16 = a * [b > 0]                                         (relu backward; a = dg1, b = g1)
// mode 3: out_bf16 = GELU(a), out_f32 = the same in fp32                 (forward with fp32 copy; a = pre-activation)
// mode 4: out_bf16 = f * GELU'(a)                                        (GELU backward with an fp32 gradient)
// mode 5: out_bf16 = out_f32 = GELU(a) + f                               (SepTPWAM: sum of the two GELU'd branches)
// mode 6: out_f32 = f + f2                                               (--version no_gate: x' = x + r, and its adjoint dr += dx')
// mode 7 / 8: modes 0 / 1 with a sigmoid gate (--lg_act_layer sigmoid, reference lib/backbone.py:552-554): g = sigmoid(a), g' = g (1 - g)
template <int MODE>
__global__ void __launch_bounds__(256) gate_elem_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const float4* __restrict__ f,
                                                        const float4* __restrict__ f2, uint4* __restrict__ out_bf16,
                                                        float4* __restrict__ out_f32, long long count8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count8) return;
  float av[8], bv[8], fv[8], gv[8], ob[8], of[8];
  {
    const uint4 u = __ldg(a + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(w[j]); av[2 * j] = t.x; av[2 * j + 1] = t.y; }
  }
  if (MODE <= 2 || MODE == 7 || MODE == 8) {
    const uint4 u = __ldg(b + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(w[j]); bv[2 * j] = t.x; bv[2 * j + 1] = t.y; }
  }
  if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 8) {
    const float4 x0 = __ldg(f + 2 * i), x1 = __ldg(f + 2 * i + 1);
    fv[0] = x0.x; fv[1] = x0.y; fv[2] = x0.z; fv[3] = x0.w; fv[4] = x1.x; fv[5] = x1.y; fv[6] = x1.z; fv[7] = x1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) gv[j] = 0.f;
  if ((MODE == 1 || MODE == 6 || MODE == 8) && f2) {
    const float4 x0 = __ldg(f2 + 2 * i), x1 = __ldg(f2 + 2 * i + 1);
    gv[0] = x0.x; gv[1] = x0.y; gv[2] = x0.z; gv[3] = x0.w; gv[4] = x1.x; gv[5] = x1.y; gv[6] = x1.z; gv[7] = x1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ob[j] = 0.f;
    of[j] = 0.f;
    if (MODE == 0) of[j] = fv[j] + tanhf(av[j]) * bv[j];
    if (MODE == 1) { const float t = tanhf(av[j]); ob[j] = fv[j] * bv[j] * (1.0f - t * t); of[j] = gv[j] + fv[j] * t; }
    if (MODE == 2) ob[j] = bv[j] > 0.f ? av[j] : 0.f;
    if (MODE == 3) { ob[j] = gelu_erf(av[j]); of[j] = ob[j]; }
    if (MODE == 4) ob[j] = fv[j] * gelu_grad(av[j]);
    if (MODE == 5) { ob[j] = gelu_erf(av[j]) + fv[j]; of[j] = ob[j]; }
    if (MODE == 6) of[j] = fv[j] + gv[j];
    if (MODE == 7) of[j] = fv[j] + bv[j] / (1.0f + __expf(-av[j]));
    if (MODE == 8) { const float t = 1.0f / (1.0f + __expf(-av[j])); ob[j] = fv[j] * bv[j] * t * (1.0f - t); of[j] = gv[j] + fv[j] * t; }
  }
  if (MODE != 0 && MODE != 6 && MODE != 7) out_bf16[i] = make_uint4(pack_bf16x2(ob[0], ob[1]), pack_bf16x2(ob[2], ob[3]), pack_bf16x2(ob[4], ob[5]), pack_bf16x2(ob[6], ob[7]));
  if (MODE == 0 || MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 8) {
    out_f32[2 * i] = make_float4(of[0], of[1], of[2], of[3]);
    out_f32[2 * i + 1] = make_float4(of[4], of[5], of[6], of[7]);
  }
}

int gate_elem_dispatch(int mode, const __nv_bfloat16* a, const __nv_bfloat16* b, const float* f, const float* f2, __nv_bfloat16* out_bf16,
                       float* out_f32, long long count, cudaStream_t st) {
  LAVT_REQUIRE(count > 0 && count % 8 == 0, "gate kernels: element count must be a multiple of 8");
  const long long c8 = count / 8;
  const unsigned grid = static_cast<unsigned>((c8 + 255) / 256);
  const uint4* a4 = reinterpret_cast<const uint4*>(a);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* g4 = reinterpret_cast<const float4*>(f2);
  uint4* ob = reinterpret_cast<uint4*>(out_bf16);
  float4* of = reinterpret_cast<float4*>(out_f32);
  switch (mode) {
    case 0: LAVT_REQUIRE(a && b && f && out_f32, "gate apply: missing tensor"); gate_elem_kernel<0><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 1: LAVT_REQUIRE(a && b && f && out_bf16 && out_f32, "gate backward: missing tensor"); gate_elem_kernel<1><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 2: LAVT_REQUIRE(a && b && out_bf16, "relu backward: missing tensor"); gate_elem_kernel<2><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 3: LAVT_REQUIRE(a && out_bf16 && out_f32, "gelu forward: missing tensor"); gate_elem_kernel<3><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 4: LAVT_REQUIRE(a && f && out_bf16, "gelu backward: missing tensor"); gate_elem_kernel<4><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 5: LAVT_REQUIRE(a && f && out_bf16 && out_f32, "gelu sum: missing tensor"); gate_elem_kernel<5><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 6: LAVT_REQUIRE(a && f && g4 && out_f32, "add: missing tensor"); gate_elem_kernel<6><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 7: LAVT_REQUIRE(a && b && f && out_f32, "sigmoid gate apply: missing tensor"); gate_elem_kernel<7><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    case 8: LAVT_REQUIRE(a && b && f && out_bf16 && out_f32, "sigmoid gate backward: missing tensor"); gate_elem_kernel<8><<<grid, 256, 0, st>>>(a4, b4, f4, g4, ob, of, c8); break;
    default: set_last_error("gate kernels: bad mode %d", mode); return LAVT_ERR_SHAPE;
  }
  LAVT_LAUNCH_CHECK("gate_elem_kernel");
  return LAVT_OK;
}

}  // namespace lavt
