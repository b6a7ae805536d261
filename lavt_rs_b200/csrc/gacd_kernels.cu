// GA-CD fusion (reference lib/bcam.py:78-127, the --gacd ablation of the 2-D image backbone, lib/backbone.py:578-582).
// With one query vector per image the "attention" collapses to vector algebra, so nothing of size n x n or n x C x C is formed:
//     ls   = LangProject(l)                                  (lavt_lang_project)
//     xm   = relu(Linear(ls * x))                            (lavt_pwam_mul_norm trick + tcgen05 GEMM with ReLU epilogue)
//     q    = Wq ls + bq ;  u_c = Wc^T q, k_c = bc . q ;  u_d = Wd^T q, k_d = bd . q                 gacd_query_kernel, gacd_vec_kernel
//     s_c[n] = (xm[n] . u_c + k_c) C^-0.5 ,  s_d[n] likewise; per-block softmax partials of s_c       gacd_scores_kernel
//     xbar = sum_n softmax(s_c)[n] xm[n] ;  f_col = Wv xbar + bv     (sum_n A_c = 1 folds value's bias)  gacd_finish_kernel
//     out[n] = xm[n] + sigmoid(s_d[n]) f_col                                                             gacd_apply_kernel
#include "../../include/lavt_b200.h"
#include "kernels.cuh"

namespace lavt {

// q[b, o] = Wq[o, :] . ls[b] + bq[o]: one warp per output row, 8 rows per CTA, grid (C / 8, B).  (These per-image mat-vecs ran in one CTA
// per image at first: 0.5 ms per launch at C = 1024, profiles/r1_ncu_launches_image_gacd.txt.)
__global__ void __launch_bounds__(256) gacd_query_kernel(const float* __restrict__ stats, const float* __restrict__ wq, const float* __restrict__ bq,
                                                         float* __restrict__ q, int C) {
  extern __shared__ float gv_sm[];          // ls[C]
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) gv_sm[c] = -stats[(static_cast<long long>(b) * 2) * C + c];   // lang_project stores -ls
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= C) return;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wq + static_cast<long long>(o) * C + c), gv_sm[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) q[static_cast<long long>(b) * C + o] = acc + bq[o];
}

// u_c = Wc^T q, u_d = Wd^T q (column sums: thread x = column, thread y = one eighth of the rows), k_c = bc . q, k_d = bd . q.
// grid (C / 32, B), block (32, 8)
__global__ void __launch_bounds__(256) gacd_vec_kernel(const float* __restrict__ q, const float* __restrict__ wc, const float* __restrict__ bc,
                                                       const float* __restrict__ wd, const float* __restrict__ bd, float* __restrict__ u,
                                                       float* __restrict__ k0, int C) {
  extern __shared__ float gv_sm[];          // q[C], part[2][8][32]
  float* qs = gv_sm;
  float* part = gv_sm + C;
  const int b = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  for (int c = tid; c < C; c += 256) qs[c] = q[static_cast<long long>(b) * C + c];
  __syncthreads();
  const int c = blockIdx.x * 32 + tx;
  float ac = 0.f, ad = 0.f;
  if (c < C) {
    for (int o = ty; o < C; o += 8) {
      ac = fmaf(__ldg(wc + static_cast<long long>(o) * C + c), qs[o], ac);
      ad = fmaf(__ldg(wd + static_cast<long long>(o) * C + c), qs[o], ad);
    }
  }
  part[ty * 32 + tx] = ac;
  part[256 + ty * 32 + tx] = ad;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sc = 0.f, sd = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) { sc += part[y * 32 + tx]; sd += part[256 + y * 32 + tx]; }
    u[(static_cast<long long>(b) * 2) * C + c] = sc;
    u[(static_cast<long long>(b) * 2 + 1) * C + c] = sd;
  }
  if (blockIdx.x == 0 && ty == 1) {
    float kc = 0.f, kd = 0.f;
    for (int o = tx; o < C; o += 32) { kc = fmaf(bc[o], qs[o], kc); kd = fmaf(bd[o], qs[o], kd); }
    kc = warp_sum(kc);
    kd = warp_sum(kd);
    if (tx == 0) { k0[b * 2] = kc; k0[b * 2 + 1] = kd; }
  }
}

constexpr int GACD_ROWS = 256;

// grid (chunks, B), block 256 (8 warps).  scores [B, n, 2]; partials pm / ps [B, chunks], px [B, chunks, C]
__global__ void __launch_bounds__(256) gacd_scores_kernel(const float* __restrict__ xm, const float* __restrict__ u, const float* __restrict__ k0,
                                                          float* __restrict__ scores, float* __restrict__ pm, float* __restrict__ ps,
                                                          float* __restrict__ px, int n, int C, float scale) {
  extern __shared__ float gs_sm[];           // sc[256], red[8], slab[8][C]
  float* sc = gs_sm;
  float* red = gs_sm + GACD_ROWS;
  float* slab = red + 8;
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = chunk * GACD_ROWS, r1 = min(n, r0 + GACD_ROWS);
  const float* uc = u + (static_cast<long long>(b) * 2) * C;
  const float* ud = uc + C;
  const float kc = k0[b * 2], kd = k0[b * 2 + 1];
  for (int r = r0 + warp; r < r1; r += 8) {
    const float* row = xm + (static_cast<long long>(b) * n + r) * C;
    float ac = 0.f, ad = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + c));
      const float4 a = __ldg(reinterpret_cast<const float4*>(uc + c)), d = __ldg(reinterpret_cast<const float4*>(ud + c));
      ac += v.x * a.x + v.y * a.y + v.z * a.z + v.w * a.w;
      ad += v.x * d.x + v.y * d.y + v.z * d.z + v.w * d.w;
    }
    ac = warp_sum(ac);
    ad = warp_sum(ad);
    if (lane == 0) {
      const float s_c = (ac + kc) * scale, s_d = (ad + kd) * scale;
      sc[r - r0] = s_c;
      *reinterpret_cast<float2*>(scores + (static_cast<long long>(b) * n + r) * 2) = make_float2(s_c, s_d);
    }
  }
  __syncthreads();
  float m = -INFINITY;
  for (int i = threadIdx.x; i < r1 - r0; i += blockDim.x) m = fmaxf(m, sc[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  // weighted sum of the block's rows with weights exp(s - m): lane owns channels lane*4 + 128*i
  float* mine = slab + warp * C;
  for (int c = lane; c < C; c += 32) mine[c] = 0.f;
  float wsum = 0.f;
  for (int r = r0 + warp; r < r1; r += 8) {
    const float w = __expf(sc[r - r0] - m);
    wsum += w;
    const float* row = xm + (static_cast<long long>(b) * n + r) * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + c));
      float4* d = reinterpret_cast<float4*>(mine + c);
      float4 o = *d;
      o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y); o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
      *d = o;
    }
  }
  __syncthreads();                              // red[] reads above are done
  if (lane == 0) red[warp] = wsum;
  __syncthreads();
  const long long pi = static_cast<long long>(b) * chunks + chunk;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += slab[w * C + c];
    px[pi * C + c] = a;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    pm[pi] = m;
    ps[pi] = s;
  }
}

// grid (C / 8, B): every CTA merges the block partials into xbar (chunks * C L2 reads), then one warp per row of f_col = Wv xbar + bv
__global__ void __launch_bounds__(256) gacd_finish_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ px,
                                                          const float* __restrict__ wv, const float* __restrict__ bv, float* __restrict__ fcol,
                                                          int chunks, int C) {
  extern __shared__ float gf_sm[];            // xbar[C]
  const int b = blockIdx.y;
  float M = -INFINITY;
  for (int k = 0; k < chunks; ++k) M = fmaxf(M, pm[b * chunks + k]);
  float S = 0.f;
  for (int k = 0; k < chunks; ++k) S += ps[b * chunks + k] * __expf(pm[b * chunks + k] - M);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < chunks; ++k) a = fmaf(px[(static_cast<long long>(b) * chunks + k) * C + c], __expf(pm[b * chunks + k] - M), a);
    gf_sm[c] = a / S;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= C) return;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wv + static_cast<long long>(o) * C + c), gf_sm[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) fcol[static_cast<long long>(b) * C + o] = acc + bv[o];
}

__global__ void __launch_bounds__(256) gacd_apply_kernel(const float4* __restrict__ xm, const float* __restrict__ scores, const float* __restrict__ fcol,
                                                         float4* __restrict__ out_f32, uint2* __restrict__ out_bf16, long long n, int C,
                                                         long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % (C / 4));
  const long long row = i / (C / 4);
  const int b = static_cast<int>(row / n);
  const float g = 1.0f / (1.0f + __expf(-__ldg(scores + row * 2 + 1)));
  const float4 f = __ldg(reinterpret_cast<const float4*>(fcol + static_cast<long long>(b) * C) + c4);
  float4 v = __ldg(xm + i);
  v.x = fmaf(g, f.x, v.x); v.y = fmaf(g, f.y, v.y); v.z = fmaf(g, f.z, v.z); v.w = fmaf(g, f.w, v.w);
  if (out_f32) out_f32[i] = v;
  if (out_bf16) out_bf16[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

}  // namespace lavt

using namespace lavt;

static inline long long al4(long long v) { return (v + 3) / 4 * 4; }

extern "C" int64_t lavt_gacd_workspace_floats(int32_t B, int64_t n, int32_t C) {
  const int64_t chunks = (n + GACD_ROWS - 1) / GACD_ROWS;
  // every segment starts on a 16-byte boundary (float4 reads of the partial sums / f_col): sizes rounded up to 4 floats
  return al4(1LL * B * 2 * C) + al4(2LL * B) + al4(2LL * B * n) + 2 * al4(1LL * B * chunks) + al4(1LL * B * chunks * C) + 2 * al4(1LL * B * C);
}

// Everything of GA-CD after mm_gen: xm fp32 [B,n,C] -> out (fp32 and / or bf16) [B,n,C]
extern "C" int lavt_gacd_fuse(const float* xm, const float* lang_stats, const float* wq, const float* bq, const float* wc, const float* bc,
                              const float* wd, const float* bd, const float* wv, const float* bv, float* workspace, float* out_f32,
                              void* out_bf16, int32_t B, int64_t n, int32_t C, void* stream) {
  LAVT_REQUIRE(B > 0 && n > 0 && n < (1LL << 30) && C % 4 == 0 && C <= 1024, "gacd: bad sizes (C=%d)", C);
  LAVT_REQUIRE(workspace && (out_f32 || out_bf16), "gacd: missing buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = static_cast<int>((n + GACD_ROWS - 1) / GACD_ROWS);
  float* u = workspace;
  float* k0 = u + al4(1LL * B * 2 * C);
  float* scores = k0 + al4(2LL * B);
  float* pm = scores + al4(2LL * B * n);
  float* ps = pm + al4(1LL * B * chunks);
  float* px = ps + al4(1LL * B * chunks);
  float* fcol = px + al4(1LL * B * chunks * C);
  float* q = fcol + al4(1LL * B * C);
  LAVT_REQUIRE(B < 65536, "gacd: batch %d too large", B);
  gacd_query_kernel<<<dim3((C + 7) / 8, B), 256, C * sizeof(float), st>>>(lang_stats, wq, bq, q, C);
  LAVT_LAUNCH_CHECK("gacd_query_kernel");
  gacd_vec_kernel<<<dim3((C + 31) / 32, B), dim3(32, 8), (C + 512) * sizeof(float), st>>>(q, wc, bc, wd, bd, u, k0, C);
  LAVT_LAUNCH_CHECK("gacd_vec_kernel");
  const size_t smem = (GACD_ROWS + 8 + 8 * static_cast<size_t>(C)) * sizeof(float);
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    LAVT_CUDA(cudaFuncSetAttribute(gacd_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>((GACD_ROWS + 8 + 8 * 1024) * sizeof(float))));
    configured = true;
  }
  gacd_scores_kernel<<<dim3(chunks, B), 256, smem, st>>>(xm, u, k0, scores, pm, ps, px, static_cast<int>(n), C, 1.0f / sqrtf(static_cast<float>(C)));
  LAVT_LAUNCH_CHECK("gacd_scores_kernel");
  gacd_finish_kernel<<<dim3((C + 7) / 8, B), 256, C * sizeof(float), st>>>(pm, ps, px, wv, bv, fcol, chunks, C);
  LAVT_LAUNCH_CHECK("gacd_finish_kernel");
  const long long total4 = 1LL * B * n * (C / 4);
  gacd_apply_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(xm), scores, fcol,
                                                                               reinterpret_cast<float4*>(out_f32), reinterpret_cast<uint2*>(out_bf16),
                                                                               n, C, total4);
  LAVT_LAUNCH_CHECK("gacd_apply_kernel");
  return LAVT_OK;
}
