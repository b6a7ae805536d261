// Bandwidth-bound row kernels: LayerNorm fused with the gathers that feed the GEMMs.
//
//   ln_rows<MODE>      one warp per OUTPUT row, fp32 in, two-pass statistics in registers, bf16 (and/or fp32) out
//     MODE_IDENTITY    LN2 before the MLP (reference lib/video_swin_transformer.py:250), patch_embed.norm (:630),
//                      per-stage output norm (:871)
//     MODE_WINDOW      LN1 + pad + cyclic shift + window_partition as ONE gather (:218-234); pad rows -> 0
//     MODE_MERGE       PatchMerging 2x2 gather + LN(4C) (:302-308); odd H/W padded with zeros BEFORE the norm
//   im2col_patch4      (B,3,T,H,W) fp32 -> (tokens, 64) bf16 rows for the patch-embed GEMM (K = 3*4*4 = 48, zero padded)
//   colstats_*         InstanceNorm1d statistics over all tokens of a clip (PWAM, :959-962, :970-973)
//   pwam_mul           vis * InstanceNorm(lang_pre)  -> bf16 A operand of project_mm (:929-930)
#include "kernels.cuh"

#include <cstdlib>

namespace lavt {

// NV = float4 slots per lane: Cn <= NV * 128 (a lane's slot i covers channels [(i*32+lane)*4, +4); slots past Cn are
// masked, which is what Swin-T/S widths such as 96 or 192 need).  Each warp handles RPW consecutive OUTPUT rows: their
// loads are all in flight together, and the closed-form window gather (a dozen integer divisions per row,
// geom.cuh::win_token) is evaluated by RPW lanes in parallel and broadcast by shuffle instead of being recomputed by all
// 32 lanes of a one-row warp (at C = 128 that index arithmetic, not memory, bounded the gather kernel).
template <int MODE, int NV, int RPW>
__global__ void __launch_bounds__(256) ln_rows_kernel(const LnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long m0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (m0 >= p.M) return;
  const int Cn = (MODE == MODE_MERGE) ? 4 * p.C : p.C;
  const float inv_cn = 1.0f / static_cast<float>(Cn);
  bool slot[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) slot[i] = (i * 32 + lane) * 4 < Cn;
  float4 v[RPW][NV];
  long long myrow = -1;
  if (MODE == MODE_WINDOW) {
    if (lane < RPW && m0 + lane < p.M) myrow = win_token(p.win, m0 + lane).row;
  }
  bool live[RPW];

#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long long m = m0 + r;
    live[r] = m < p.M;
    if (MODE == MODE_MERGE) {
      // out row m <-> (b, d, h2, w2); channel block q of 4 <-> source pixel (2*h2 + (q&1), 2*w2 + (q>>1))
      const int H2 = (p.mH + 1) >> 1, W2 = (p.mW + 1) >> 1;
      const long long mm = live[r] ? m : p.M - 1;
      const int w2 = static_cast<int>(mm % W2);
      const int h2 = static_cast<int>((mm / W2) % H2);
      const long long bd = mm / (static_cast<long long>(W2) * H2);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;           // column in [0, 4C)
        const int q = col / p.C, c = col - q * p.C;
        const int h = 2 * h2 + (q & 1), w = 2 * w2 + (q >> 1);
        if (slot[i] && h < p.mH && w < p.mW) {
          const float* src = p.x + ((bd * p.mH + h) * p.mW + w) * p.ldx + c;
          v[r][i] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      long long row = live[r] ? m : p.M - 1;
      if (MODE == MODE_WINDOW) {
        row = __shfl_sync(0xffffffffu, myrow, r);
        if (row < 0) live[r] = false;                  // pad row (zeros written below) or past the end
      }
      const float4* src = reinterpret_cast<const float4*>(p.x + (row < 0 ? 0 : row) * p.ldx);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        v[r][i] = (live[r] && slot[i]) ? __ldg(src + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  const float4* g4 = reinterpret_cast<const float4*>(p.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(p.beta);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long long m = m0 + r;
    if (m >= p.M) break;
    if (!live[r]) {   // window pad row: zeros AFTER the norm (F.pad follows norm1 in the reference)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (!slot[i]) continue;
        if (p.out_bf16) reinterpret_cast<uint2*>(p.out_bf16 + m * Cn)[i * 32 + lane] = make_uint2(0u, 0u);
        if (p.out_f32) reinterpret_cast<float4*>(p.out_f32 + m * Cn)[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);     // masked slots hold zeros
    const float mean = warp_sum(s) * inv_cn;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (!slot[i]) continue;
      const float a = v[r][i].x - mean, b = v[r][i].y - mean, c = v[r][i].z - mean, d = v[r][i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_cn + p.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (!slot[i]) continue;
      const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
      float4 y;
      y.x = (v[r][i].x - mean) * rstd * g.x + b.x;
      y.y = (v[r][i].y - mean) * rstd * g.y + b.y;
      y.z = (v[r][i].z - mean) * rstd * g.z + b.z;
      y.w = (v[r][i].w - mean) * rstd * g.w + b.w;
      if (p.out_bf16)
        reinterpret_cast<uint2*>(p.out_bf16 + m * Cn)[i * 32 + lane] = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      if (p.out_f32) reinterpret_cast<float4*>(p.out_f32 + m * Cn)[i * 32 + lane] = y;
    }
  }
}

// Narrow rows (C = 128 / 256, stages 0 / 1: 590 K / 147 K rows per 8-clip step): LPR = C / 16 lanes share a row (4 float4 per lane), a warp
// works on 32 / LPR rows at once and G such row sets per warp.  A row's two reductions cost log2(LPR) shuffle steps that serve 32 / LPR rows
// per instruction: 1.5 (C = 128) / 4 (C = 256) warp shuffles per row instead of 10.  ncu on the one-row-per-warp kernel at C = 128: issue
// slots 59 %, short-scoreboard (shuffle) stalls 33 %, DRAM 50 % -- as much shuffle- as bandwidth-bound.
template <int MODE, int LPR, int G>
__global__ void __launch_bounds__(256) ln_rows_narrow_kernel(const LnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int RPS = 32 / LPR;                     // rows per set
  const int lane = threadIdx.x & 31, sl = lane % LPR, sr = lane / LPR;
  const long long m0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (RPS * G);
  if (m0 >= p.M) return;
  const int Cn = p.C;                               // == 16 * LPR
  const float inv_cn = 1.0f / static_cast<float>(Cn);
  float4 v[G][4];
  bool live[G], inrange[G];
  long long myrow = -1;                              // closed-form window gather: lane l evaluates row m0 + l once, the sub-groups fetch theirs
  if (MODE == MODE_WINDOW) {
    if (lane < RPS * G && m0 + lane < p.M) myrow = win_token(p.win, m0 + lane).row;
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const long long m = m0 + g * RPS + sr;
    inrange[g] = m < p.M;
    long long row = inrange[g] ? m : p.M - 1;
    if (MODE == MODE_WINDOW) row = __shfl_sync(0xffffffffu, myrow, g * RPS + sr);
    live[g] = inrange[g] && row >= 0;
    const float4* src = reinterpret_cast<const float4*>(p.x + (row < 0 ? 0 : row) * p.ldx);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[g][i] = live[g] ? __ldg(src + i * LPR + sl) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 gm[4], bt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(p.gamma) + i * LPR + sl);
    bt[i] = __ldg(reinterpret_cast<const float4*>(p.beta) + i * LPR + sl);
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (v[g][i].x + v[g][i].y) + (v[g][i].z + v[g][i].w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_cn;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = v[g][i].x - mean, b = v[g][i].y - mean, c = v[g][i].z - mean, d = v[g][i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_cn + p.eps);
    const long long m = m0 + g * RPS + sr;
    if (!inrange[g]) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);             // window pad row: zeros AFTER the norm
      if (live[g]) {
        y.x = (v[g][i].x - mean) * rstd * gm[i].x + bt[i].x;
        y.y = (v[g][i].y - mean) * rstd * gm[i].y + bt[i].y;
        y.z = (v[g][i].z - mean) * rstd * gm[i].z + bt[i].z;
        y.w = (v[g][i].w - mean) * rstd * gm[i].w + bt[i].w;
      }
      if (p.out_bf16) reinterpret_cast<uint2*>(p.out_bf16 + m * Cn)[i * LPR + sl] = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      if (p.out_f32) reinterpret_cast<float4*>(p.out_f32 + m * Cn)[i * LPR + sl] = y;
    }
  }
}

template <int MODE, int LPR, int G>
static void launch_ln_narrow(const LnParams& p, cudaStream_t st) {
  const long long rows_per_block = 8LL * (32 / LPR) * G;
  const long long blocks = (p.M + rows_per_block - 1) / rows_per_block;
  launch_pdl(ln_rows_narrow_kernel<MODE, LPR, G>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, 1, p);
}

template <int MODE, int NV, int RPW>
static void launch_ln_nv(const LnParams& p, cudaStream_t st) {
  const int warps = 8;
  const long long rows_per_block = static_cast<long long>(warps) * RPW;
  const long long blocks = (p.M + rows_per_block - 1) / rows_per_block;
  launch_pdl(ln_rows_kernel<MODE, NV, RPW>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, 1, p);
}

template <int MODE>
static int launch_ln(const LnParams& p, int Cn, cudaStream_t st) {
  LAVT_REQUIRE((p.M + 7) / 8 < (1LL << 31), "layernorm: too many rows");
  if constexpr (MODE != MODE_MERGE) {
    static int narrow = -1;
    if (narrow < 0) {
      const char* e = getenv("LAVT_LN_NARROW");       // A/B runs: 0 = one row per warp everywhere
      narrow = (e && e[0] == '0') ? 0 : 1;
    }
    if (narrow && (Cn == 128 || Cn == 256)) {
      if (Cn == 128) launch_ln_narrow<MODE, 8, 2>(p, st); else launch_ln_narrow<MODE, 16, 4>(p, st);
      LAVT_LAUNCH_CHECK("ln_rows_narrow_kernel");
      return LAVT_OK;
    }
  }
  switch ((Cn + 127) / 128) {   // rows per warp chosen so that the row data of a warp stays in registers (<= 16 float4 per lane)
    case 1: launch_ln_nv<MODE, 1, 8>(p, st); break;
    case 2: launch_ln_nv<MODE, 2, 8>(p, st); break;
    case 3: launch_ln_nv<MODE, 3, 4>(p, st); break;
    case 4: launch_ln_nv<MODE, 4, 4>(p, st); break;
    case 6: launch_ln_nv<MODE, 6, 2>(p, st); break;
    case 8: launch_ln_nv<MODE, 8, 2>(p, st); break;
    case 12: launch_ln_nv<MODE, 12, 1>(p, st); break;
    case 16: launch_ln_nv<MODE, 16, 1>(p, st); break;
    case 24: launch_ln_nv<MODE, 24, 1>(p, st); break;
    default:
      set_last_error("layernorm: normalised width %d not supported (ceil(width / 128) must be in {1,2,3,4,6,8,12,16,24})", Cn);
      return LAVT_ERR_SHAPE;
  }
  LAVT_LAUNCH_CHECK("ln_rows_kernel");
  return LAVT_OK;
}

int ln_rows_dispatch(int mode, const LnParams& p, cudaStream_t st) {
  LAVT_REQUIRE(p.M > 0, "layernorm: empty input");
  LAVT_REQUIRE(p.out_bf16 || p.out_f32, "layernorm: no output");
  LAVT_REQUIRE(p.ldx % 4 == 0 && p.C % 4 == 0, "layernorm: pitch / channels must be multiples of 4");
  const int Cn = (mode == MODE_MERGE) ? 4 * p.C : p.C;
  LAVT_REQUIRE(Cn % 4 == 0, "layernorm: normalised width %d must be a multiple of 4", Cn);
  if (mode == MODE_IDENTITY) return launch_ln<MODE_IDENTITY>(p, Cn, st);
  if (mode == MODE_WINDOW) return launch_ln<MODE_WINDOW>(p, Cn, st);
  if (mode == MODE_MERGE) return launch_ln<MODE_MERGE>(p, Cn, st);
  set_last_error("layernorm: bad mode %d", mode);
  return LAVT_ERR_SHAPE;
}

// ---------------------------------------------------------------------------------------------
// patch-embed im2col: x (B, 3, T, H, W) fp32  ->  rows (B*T*Hp*Wp, 64) bf16, col = c*16 + ph*4 + pw (48..63 = 0)
// (Conv3d k = s = (1,4,4) == GEMM, reference lib/video_swin_transformer.py:616-628; right/bottom zero pad)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_patch4_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            long long sB, long long sC, long long sT,
                                                            int B, int T, int H, int W, int Hp, int Wp) {
  // one thread per (token, c, ph): 4 consecutive pixels -> 4 bf16
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * T * Hp * Wp * 16;
  if (idx >= total) return;
  const int sub = static_cast<int>(idx & 15);
  const long long tok = idx >> 4;
  uint2 o = make_uint2(0u, 0u);
  if (sub < 12) {
    const int c = sub >> 2, ph = sub & 3;
    const int wp = static_cast<int>(tok % Wp);
    const int hp = static_cast<int>((tok / Wp) % Hp);
    const int t = static_cast<int>((tok / (static_cast<long long>(Wp) * Hp)) % T);
    const int b = static_cast<int>(tok / (static_cast<long long>(Wp) * Hp * T));
    const int h = hp * 4 + ph;
    if (h < H) {
      const float* src = x + b * sB + c * sC + t * sT + static_cast<long long>(h) * W + wp * 4;
      float f[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = (wp * 4 + j < W) ? __ldg(src + j) : 0.f;
      o = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
    }
  }
  reinterpret_cast<uint2*>(out + tok * 64)[sub] = o;
}

int im2col_patch4_dispatch(const float* x, long long sB, long long sC, long long sT, __nv_bfloat16* out, int B, int T, int H,
                           int W, cudaStream_t st) {
  const int Hp = (H + 3) / 4, Wp = (W + 3) / 4;
  const long long total = static_cast<long long>(B) * T * Hp * Wp * 16;
  LAVT_REQUIRE(total > 0, "patch embed: empty input");
  im2col_patch4_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, out, sB, sC, sT, B, T, H, W, Hp, Wp);
  LAVT_LAUNCH_CHECK("im2col_patch4_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm statistics per (clip, channel) over n tokens of a bf16 (B, n, C) tensor.
// pass 1: per-chunk sums of (x - pivot) and (x - pivot)^2 (pivot = token 0 of the clip, kills the
//         E[x^2]-E[x]^2 cancellation);  pass 2: deterministic reduction over chunks -> mean, rstd.
// ---------------------------------------------------------------------------------------------
constexpr int CS_ROWS_PER_BLOCK = 256;
// rows per block: 256 while that still gives every SM several blocks, fewer for short clips (stage 2 / 3 of the bench have 4608 / 1152 tokens
// per clip: with 256-row chunks those launches ran 144 / 40 blocks and took 41 / 76 us for 75 / 38 MB)
static int cs_rows_per_block(int B, long long n) {
  int rows = CS_ROWS_PER_BLOCK;
  while (rows > 16 && static_cast<long long>(B) * ((n + rows - 1) / rows) < 592) rows >>= 1;
  return rows;
}

__global__ void __launch_bounds__(256) colstats_partial_kernel(const float* __restrict__ x, float* __restrict__ part,
                                                               int n, int C, int chunks, int rows_per_block) {
  // grid: (chunks, B).  thread -> 4 channels (one float4); row groups stride over the chunk's rows
  extern __shared__ float sm[];   // [rowgroups][C][2]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tpr = C / 4;                       // threads per row
  const int rg = blockDim.x / tpr;             // row groups
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const float* base = x + static_cast<long long>(b) * n * C;
  const float4 pv = __ldg(reinterpret_cast<const float4*>(base) + tc);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const int r0 = chunk * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  if (tr < rg) {
    int r = r0 + tr;
    for (; r + 3 * rg < r1; r += 4 * rg) {      // 4 independent 16-byte loads in flight per thread
      float4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const float4*>(base + static_cast<long long>(r + k * rg) * C) + tc);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = u[k].x - pv.x, bb = u[k].y - pv.y, c = u[k].z - pv.z, d = u[k].w - pv.w;
        s1.x += a; s1.y += bb; s1.z += c; s1.w += d;
        s2.x += a * a; s2.y += bb * bb; s2.z += c * c; s2.w += d * d;
      }
    }
    for (; r < r1; r += rg) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(base + static_cast<long long>(r) * C) + tc);
      const float a = u.x - pv.x, bb = u.y - pv.y, c = u.z - pv.z, d = u.w - pv.w;
      s1.x += a; s1.y += bb; s1.z += c; s1.w += d;
      s2.x += a * a; s2.y += bb * bb; s2.z += c * c; s2.w += d * d;
    }
    float* o = sm + (static_cast<long long>(tr) * C + tc * 4) * 2;
    o[0] = s1.x; o[1] = s2.x; o[2] = s1.y; o[3] = s2.y; o[4] = s1.z; o[5] = s2.z; o[6] = s1.w; o[7] = s2.w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int g = 0; g < rg; ++g) {
      a += sm[(g * C + c) * 2 + 0];
      q += sm[(g * C + c) * 2 + 1];
    }
    float* o = part + ((static_cast<long long>(b) * chunks + chunk) * 2) * C;
    o[c] = a;
    o[C + c] = q;
  }
}

// grid (C / 32, B), 256 threads = 8 chunk slices x 32 channels: slice s sums the partials of chunks s, s + 8, ... in double, the slices are
// combined in a fixed order (deterministic).  (One thread per channel walking all chunks serially took 43 us for 288 chunks per clip.)
__global__ void __launch_bounds__(256) colstats_final_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                                             float* __restrict__ stats, int n, int C, int chunks, float eps) {
  __shared__ double sa[8][32], sq[8][32];
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double a = 0.0, q = 0.0;
  if (c < C) {
    for (int k = sl; k < chunks; k += 8) {
      const float* o = part + ((static_cast<long long>(b) * chunks + k) * 2) * C;
      a += o[c];
      q += o[C + c];
    }
  }
  sa[sl][cl] = a;
  sq[sl][cl] = q;
  __syncthreads();
  if (sl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      a += sa[k][cl];
      q += sq[k][cl];
    }
    const float pivot = x[static_cast<long long>(b) * n * C + c];
    const double m1 = a / n;
    const double var = fmax(q / n - m1 * m1, 0.0);
    stats[(static_cast<long long>(b) * 2 + 0) * C + c] = pivot + static_cast<float>(m1);
    stats[(static_cast<long long>(b) * 2 + 1) * C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

long long colstats_workspace_floats(int B, long long n, int C) {
  const int rows = cs_rows_per_block(B, n);
  const long long chunks = (n + rows - 1) / rows;
  return static_cast<long long>(B) * chunks * 2 * C;
}

int colstats_dispatch(const float* x, float* stats, float* workspace, int B, long long n_ll, int C, float eps, cudaStream_t st) {
  LAVT_REQUIRE(n_ll < (1LL << 30), "instance-norm stats: too many tokens");
  const int n = static_cast<int>(n_ll);
  LAVT_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, "instance-norm stats: C=%d unsupported", C);
  LAVT_REQUIRE(B > 0 && n > 0, "instance-norm stats: empty input");
  const int rows = cs_rows_per_block(B, n);
  const int chunks = (n + rows - 1) / rows;
  const int tpr = C / 4;
  const int rg = 256 / tpr;
  const size_t smem = static_cast<size_t>(rg) * C * 2 * sizeof(float);   // <= 8 KB
  colstats_partial_kernel<<<dim3(chunks, B), 256, smem, st>>>(x, workspace, n, C, chunks, rows);
  LAVT_LAUNCH_CHECK("colstats_partial_kernel");
  colstats_final_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(x, workspace, stats, n, C, chunks, eps);
  LAVT_LAUNCH_CHECK("colstats_final_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// out = vis * (lang_pre - mean) * rstd   (bf16 in / out, stats fp32 (B,2,C))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pwam_mul_kernel(const __nv_bfloat16* __restrict__ vis, const float* __restrict__ lang,
                                                       const float* __restrict__ stats, __nv_bfloat16* __restrict__ out,
                                                       long long n, int C, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = static_cast<int>(i % (C / 8));
  const long long row = i / (C / 8);
  const int b = static_cast<int>(row / n);
  const float* mu = stats + (static_cast<long long>(b) * 2) * C + c8 * 8;
  const float* rs = mu + C;
  const uint4 uv = __ldg(reinterpret_cast<const uint4*>(vis) + i);
  const float4 l0 = __ldg(reinterpret_cast<const float4*>(lang) + 2 * i), l1 = __ldg(reinterpret_cast<const float4*>(lang) + 2 * i + 1);
  const uint32_t vv[4] = {uv.x, uv.y, uv.z, uv.w};
  const float ll[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
  uint32_t oo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = unpack_bf16x2(vv[j]);
    oo[j] = pack_bf16x2(a.x * (ll[2 * j] - mu[2 * j]) * rs[2 * j], a.y * (ll[2 * j + 1] - mu[2 * j + 1]) * rs[2 * j + 1]);
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
}

int pwam_mul_dispatch(const __nv_bfloat16* vis, const float* lang, const float* stats, __nv_bfloat16* out, int B,
                      long long n, int C, cudaStream_t st) {
  LAVT_REQUIRE(C % 8 == 0, "pwam_mul: C must be a multiple of 8");
  const long long total8 = static_cast<long long>(B) * n * (C / 8);
  LAVT_REQUIRE(total8 > 0, "pwam_mul: empty input");
  pwam_mul_kernel<<<static_cast<unsigned>((total8 + 255) / 256), 256, 0, st>>>(vis, lang, stats, out, n, C, total8);
  LAVT_LAUNCH_CHECK("pwam_mul_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// out = IN(a) + IN(b) = (a - mean_a) * rstd_a + (b - mean_b) * rstd_b   (fp32, stats (B,2,C) each; out may alias a)
// SepTPWAM sums a temporal (3x3x3) and a spatial (1x1x1) branch AFTER their InstanceNorm3d
// (reference lib/video_swin_transformer.py:1513-1524, 1556-1561)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) instnorm_sum2_kernel(const float* __restrict__ a, const float* __restrict__ sa,
                                                            const float* __restrict__ b, const float* __restrict__ sb,
                                                            float* __restrict__ out, long long n, int C, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % (C / 4));
  const long long row = i / (C / 4);
  const int bi = static_cast<int>(row / n);
  const float4 ma = __ldg(reinterpret_cast<const float4*>(sa + (static_cast<long long>(bi) * 2) * C) + c4);
  const float4 ra = __ldg(reinterpret_cast<const float4*>(sa + (static_cast<long long>(bi) * 2 + 1) * C) + c4);
  const float4 mb = __ldg(reinterpret_cast<const float4*>(sb + (static_cast<long long>(bi) * 2) * C) + c4);
  const float4 rb = __ldg(reinterpret_cast<const float4*>(sb + (static_cast<long long>(bi) * 2 + 1) * C) + c4);
  const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
  float4 o;
  o.x = (x.x - ma.x) * ra.x + (y.x - mb.x) * rb.x;
  o.y = (x.y - ma.y) * ra.y + (y.y - mb.y) * rb.y;
  o.z = (x.z - ma.z) * ra.z + (y.z - mb.z) * rb.z;
  o.w = (x.w - ma.w) * ra.w + (y.w - mb.w) * rb.w;
  reinterpret_cast<float4*>(out)[i] = o;
}

int instnorm_sum2_dispatch(const float* a, const float* sa, const float* b, const float* sb, float* out, int B, long long n,
                           int C, cudaStream_t st) {
  LAVT_REQUIRE(C % 4 == 0, "instnorm_sum2: C must be a multiple of 4");
  const long long total4 = static_cast<long long>(B) * n * (C / 4);
  LAVT_REQUIRE(total4 > 0 && (total4 + 255) / 256 < (1LL << 31), "instnorm_sum2: empty or too large input");
  instnorm_sum2_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(a, sa, b, sb, out, n, C, total4);
  LAVT_LAUNCH_CHECK("instnorm_sum2_kernel");
  return LAVT_OK;
}

}  // namespace lavt
