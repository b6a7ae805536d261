// fp32 VALIDATION build of the two contraction families of the Swin block (north_star tolerance: "1e-4 in fp32 with fp32 accumulate").
// The production kernels (gemm_tc.cu, attn_tc*.cu) take bf16 operands; these twins take fp32 operands, multiply-accumulate in fp32 on
// the CUDA cores and share the production epilogue / index math (same lavt_epilogue_t, same closed-form window geometry, same
// exp2-domain softmax), so that a whole block can be replayed at fp32 accuracy on the device (engine.set_precision("fp32")) and every
// production op can be bracketed: |bf16 path - fp32 path| is the operand-rounding error, |fp32 path - oracle| <= 1e-4 is the kernel logic.
// They are slow by design (no tensor cores) and never selected by default.
#include "../../include/lavt_b200.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"

#include <cstring>

namespace lavt {

constexpr int RG_TM = 64, RG_TN = 64, RG_TK = 16;

__device__ __forceinline__ float ref_act(float x, int act) {
  if (act == ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_TANH) return tanhf(x);
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  return x;
}

// C[M,N] = A[M,K] W[N,K]^T, 64 x 64 tile, 16 x 16 threads, 4 x 4 outputs per thread, epilogue of GemmParams
__global__ void __launch_bounds__(256) gemm_f32_ref_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                                                           const GemmParams p) {
  __shared__ float sa[RG_TK][RG_TM + 4];
  __shared__ float sb[RG_TK][RG_TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = static_cast<long long>(blockIdx.y) * RG_TM;
  const int n0 = blockIdx.x * RG_TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += RG_TK) {
    for (int i = threadIdx.x; i < RG_TM * RG_TK; i += 256) {
      const int r = i / RG_TK, k = i - r * RG_TK;
      const long long m = m0 + r;
      sa[k][r] = (m < p.M && k0 + k < p.K) ? A[m * lda + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < RG_TN * RG_TK; i += 256) {
      const int r = i / RG_TK, k = i - r * RG_TK;
      const int n = n0 + r;
      sb[k][r] = (n < p.N && k0 + k < p.K) ? W[static_cast<long long>(n) * ldw + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RG_TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    long long orow = m;
    if (p.rowmap == ROWMAP_WINDOW) {
      orow = win_token(p.win, m).row;
      if (orow < 0) continue;                       // pad row of a window: nothing to write back
    }
    const float rs = p.rscale ? p.rscale[orow / p.rs_rows] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.cscale) v *= p.cscale[n];
      if (p.bias) v += p.bias[n];
      v = ref_act(v, p.act);
      if (p.mul) v *= __bfloat162float(p.mul[m * p.ldm + n]);
      v *= rs;
      if (p.resid) v += p.resid[orow * p.ldo + n];
      if (p.out_f32) p.out_f32[orow * p.ldo + n] = v;
      if (p.out_bf16) p.out_bf16[orow * p.ldo + n] = __float2bfloat16(v);
    }
  }
}

// One warp per (window, head, query row); lane l scores the keys l, l + 32, ...; q is pre-scaled by head_dim^-0.5 * log2(e) like in the
// production path, the table is multiplied by log2(e) here, masked pairs get -100 * log2(e), softmax in base 2.
constexpr int RA_MAX_PER_LANE = 36;      // N <= 1152
__global__ void __launch_bounds__(128) window_attn_f32_ref_kernel(const float* __restrict__ qkv, const float* __restrict__ table_t,
                                                                  float* __restrict__ out, int C, int nH, int L, const WinGeom g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long rows = 1LL * g.B * g.nwd * g.nwh * g.nww * g.N;
  if (wid >= rows * nH) return;
  const int h = static_cast<int>(wid % nH);
  const long long m = wid / nH;
  const long long win0 = (m / g.N) * g.N;
  const WinTok ti = win_token(g, m);
  const bool shifted = (g.sd | g.sh | g.sw) != 0;
  const float* qr = qkv + m * 3 * C + h * 32;
  float q[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = qr[d];
  const float* tab = table_t + static_cast<long long>(h) * L;
  const int rc = rel_const(g);
  float sc[RA_MAX_PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < RA_MAX_PER_LANE; ++t) {
    const int j = lane + 32 * t;
    float s = -INFINITY;
    if (j < g.N) {
      const WinTok tj = win_token(g, win0 + j);
      const float* kr = qkv + (win0 + j) * 3 * C + C + h * 32;
      s = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) s = fmaf(q[d], kr[d], s);
      s += tab[ti.code - tj.code + rc] * 1.4426950408889634f;
      if (shifted && ti.rid != tj.rid) s += -100.0f * 1.4426950408889634f;
    }
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
#pragma unroll
  for (int t = 0; t < RA_MAX_PER_LANE; ++t) {
    const int j = lane + 32 * t;
    if (j < g.N) {
      const float pj = exp2f(sc[t] - mx);
      sum += pj;
      const float* vr = qkv + (win0 + j) * 3 * C + 2 * C + h * 32;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] = fmaf(pj, vr[d], acc[d]);
    }
  }
  sum = warp_sum(sum);
  float mine = 0.f;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const float tot = warp_sum(acc[d]);
    if (lane == d) mine = tot;
  }
  out[m * C + h * 32 + lane] = mine / sum;
}

}  // namespace lavt

using namespace lavt;

extern "C" int lavt_gemm_f32_ref(const float* A, int64_t lda, const float* Wt, int64_t ldw, int32_t M, int32_t N, int32_t K,
                                 const lavt_epilogue_t* e, void* stream) {
  LAVT_REQUIRE(A && Wt && e && M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K, "gemm_f32_ref: bad arguments");
  LAVT_REQUIRE(e->out_f32 || e->out_bf16, "gemm_f32_ref: no output");
  LAVT_REQUIRE(e->act >= 0 && e->act <= 4, "gemm_f32_ref: bad activation id %d", e->act);
  LAVT_REQUIRE(e->mul_act == 0 && !e->out_pre && e->pre_mode == 0, "gemm_f32_ref: mul_act / out_pre are training-path epilogues of the tcgen05 kernel only");
  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.cscale = e->cscale; p.bias = e->bias; p.act = e->act;
  p.mul = static_cast<const __nv_bfloat16*>(e->mul); p.ldm = e->ldm;
  p.resid = e->resid; p.out_f32 = e->out_f32; p.out_bf16 = static_cast<__nv_bfloat16*>(e->out_bf16); p.ldo = e->ldo;
  p.rscale = e->rscale; p.rs_rows = e->rscale_rows;
  LAVT_REQUIRE(!e->rscale || e->rscale_rows > 0, "gemm_f32_ref: rscale needs rscale_rows > 0");
  p.rowmap = ROWMAP_IDENTITY;
  if (e->win) {
    p.rowmap = ROWMAP_WINDOW;
    std::memcpy(&p.win, e->win, sizeof(WinGeom));
    const long long rows = 1LL * p.win.B * p.win.nwd * p.win.nwh * p.win.nww * p.win.N;
    LAVT_REQUIRE(rows == M, "gemm_f32_ref: window geometry rows %lld != M %d", rows, M);
  }
  const dim3 grid((N + RG_TN - 1) / RG_TN, (M + RG_TM - 1) / RG_TM);
  LAVT_REQUIRE(grid.y < 65536, "gemm_f32_ref: M too large for the validation kernel");
  gemm_f32_ref_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, Wt, ldw, p);
  LAVT_LAUNCH_CHECK("gemm_f32_ref_kernel");
  return LAVT_OK;
}

extern "C" int lavt_window_attention_f32_ref(const float* qkv, const float* table_t, int32_t L, int32_t nH, const lavt_win_geom_t* geom,
                                             float* out, void* stream) {
  LAVT_REQUIRE(qkv && table_t && geom && out && nH > 0, "attention_f32_ref: bad arguments");
  WinGeom g;
  std::memcpy(&g, geom, sizeof(WinGeom));
  LAVT_REQUIRE(g.N == g.wd * g.wh * g.ww && g.N > 0 && g.N <= 32 * RA_MAX_PER_LANE, "attention_f32_ref: window of %d tokens not supported", g.N);
  LAVT_REQUIRE(L == (2 * g.Wd - 1) * (2 * g.Wh - 1) * (2 * g.Ww - 1), "attention_f32_ref: bias table rows %d do not match the window", L);
  const long long warps = 1LL * g.B * g.nwd * g.nwh * g.nww * g.N * nH;
  LAVT_REQUIRE((warps + 3) / 4 < (1LL << 31), "attention_f32_ref: too many rows");
  window_attn_f32_ref_kernel<<<static_cast<unsigned>((warps + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(qkv, table_t, out, nH * 32,
                                                                                                              nH, L, g);
  LAVT_LAUNCH_CHECK("window_attn_f32_ref_kernel");
  return LAVT_OK;
}

extern "C" int lavt_layernorm_window_gather_f32(const float* x, int32_t C, const lavt_win_geom_t* geom, const float* gamma, const float* beta,
                                                float eps, float* out_f32, void* stream) {
  LAVT_REQUIRE(geom != nullptr && out_f32 != nullptr, "window gather (fp32): missing arguments");
  LnParams p;
  std::memset(&p, 0, sizeof(p));
  std::memcpy(&p.win, geom, sizeof(WinGeom));
  p.x = x; p.ldx = C; p.gamma = gamma; p.beta = beta; p.out_f32 = out_f32;
  p.M = 1LL * geom->B * geom->nwd * geom->nwh * geom->nww * geom->N; p.C = C; p.eps = eps;
  return ln_rows_dispatch(MODE_WINDOW, p, static_cast<cudaStream_t>(stream));
}
