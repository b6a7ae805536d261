// Bandwidth-bound kernels of the backward pass (training step, BASELINE config 4; the reference differentiates the forward
// with autograd, train.py:330-360 -- these are the hand-written adjoints of the forward kernels in norm_kernels.cu and of the
// GEMM epilogues):
//   splitk_reduce     sums the split-K partials of a weight-gradient GEMM into the (accumulating) fp32 gradient
//   transpose_bf16    [M, N] -> [N, M]: dY^T and X^T operands of  dW = dY^T X  (the tcgen05 GEMM takes K-major operands)
//   colsum            bias gradients: dst[n] += sum_m x[m, n]
//   cast_rows         fp32 rows -> bf16 rows, optionally gathered into window order (adjoint of the proj epilogue's
//                     window_reverse scatter, lib/video_swin_transformer.py:238-247)
//   gelu_fwd / bwd    exact-erf GELU on the saved pre-activation (Mlp, :30-36)
//   ln_bwd<MODE>      LayerNorm backward fused with the adjoint of the forward gather: MODE_IDENTITY (norm2, patch_embed.norm,
//                     norm{i}), MODE_WINDOW (norm1 + pad + roll + window_partition, :218-234), MODE_MERGE (PatchMerging
//                     2x2 gather + LN(4C), :298-308); adds the result to the gradient on the residual stream
#include "kernels.cuh"
#include "gemm_tc.cuh"

namespace lavt {

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float4* __restrict__ part, int splits, long long count4, int ncols4,
                                                            float* __restrict__ dst, long long ldd, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float4 a = __ldg(part + i);
  for (int s = 1; s < splits; ++s) {
    const float4 b = __ldg(part + s * count4 + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const long long m = i / ncols4;
  const int n = static_cast<int>(i - m * ncols4) * 4;
  float4* d = reinterpret_cast<float4*>(dst + m * ldd + n);
  if (accumulate) {
    const float4 o = *d;
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
  }
  *d = a;
}

// Many splits (single-tile weight gradients split 100-300 ways): one thread walking all partials of its element is a chain of `splits`
// L2 round trips on a handful of SMs (272 splits of a 128 x 128 output: ~40 us).  Here 8 thread groups of a block sum every 8th split
// of 32 consecutive float4 elements and combine through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) splitk_reduce_wide_kernel(const float4* __restrict__ part, int splits, long long count4, int ncols4,
                                                                 float* __restrict__ dst, long long ldd, int accumulate) {
  __shared__ float4 sm[8][32];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + e;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < count4) {
#pragma unroll 4
    for (int s = g; s < splits; s += 8) {
      const float4 b = __ldg(part + s * count4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
  }
  sm[g][e] = a;
  __syncthreads();
  if (g != 0 || i >= count4) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) {
    const float4 b = sm[q][e];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const long long m = i / ncols4;
  const int n = static_cast<int>(i - m * ncols4) * 4;
  float4* d = reinterpret_cast<float4*>(dst + m * ldd + n);
  if (accumulate) {
    const float4 o = *d;
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
  }
  *d = a;
}

// out = act((sum_s partial[s]) * cscale[col] + bias[col]) (+ resid): the epilogue of a split-K forward GEMM (small-M GEMMs of the text encoder:
// 160 rows give 2 x 3..12 output tiles, each walking up to 48 k-blocks alone -- 26-43 us per launch; split over K they fill the machine)
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const float4* __restrict__ part, int splits, long long count4, int ncols4,
                                                              const float4* __restrict__ cscale, const float4* __restrict__ bias, int act,
                                                              const float* __restrict__ resid, float* __restrict__ out_f32,
                                                              __nv_bfloat16* __restrict__ out_bf16, long long ldo) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float4 a = __ldg(part + i);
  for (int s2 = 1; s2 < splits; ++s2) {
    const float4 b = __ldg(part + s2 * count4 + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const long long row = i / ncols4;
  const int c4 = static_cast<int>(i - row * ncols4);
  if (cscale) { const float4 sc = __ldg(cscale + c4); a.x *= sc.x; a.y *= sc.y; a.z *= sc.z; a.w *= sc.w; }
  if (bias) { const float4 b = __ldg(bias + c4); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  if (act == 1) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
  const long long o = row * ldo + c4 * 4;
  if (resid) { const float4 r = *reinterpret_cast<const float4*>(resid + o); a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w; }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = a;
  if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + o) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
}

int splitk_epilogue_dispatch(const float* partials, int splits, long long M, int N, const float* cscale, const float* bias, int act,
                             const float* resid, float* out_f32, __nv_bfloat16* out_bf16, long long ldo, cudaStream_t st) {
  LAVT_REQUIRE(N % 4 == 0 && ldo % 4 == 0 && splits >= 1 && (act == 0 || act == 1), "split-K epilogue: bad arguments");
  const long long c4 = M * N / 4;
  splitk_epilogue_kernel<<<static_cast<unsigned>((c4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(partials), splits, c4, N / 4,
                                                                                 reinterpret_cast<const float4*>(cscale),
                                                                                 reinterpret_cast<const float4*>(bias), act, resid, out_f32,
                                                                                 out_bf16, ldo);
  LAVT_LAUNCH_CHECK("splitk_epilogue_kernel");
  return LAVT_OK;
}

int splitk_reduce_dispatch(const float* partials, int splits, long long count, int ncols, float* dst, long long ldd, int accumulate,
                           cudaStream_t st) {
  LAVT_REQUIRE(count % 4 == 0 && ncols % 4 == 0 && ldd % 4 == 0, "splitk reduce: sizes must be multiples of 4");
  const long long c4 = count / 4;
  if (splits >= 16 && (c4 + 255) / 256 < 4 * 148) {
    splitk_reduce_wide_kernel<<<static_cast<unsigned>((c4 + 31) / 32), 256, 0, st>>>(reinterpret_cast<const float4*>(partials), splits, c4,
                                                                                  ncols / 4, dst, ldd, accumulate);
    LAVT_LAUNCH_CHECK("splitk_reduce_wide_kernel");
    return LAVT_OK;
  }
  splitk_reduce_kernel<<<static_cast<unsigned>((c4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(partials), splits, c4,
                                                                              ncols / 4, dst, ldd, accumulate);
  LAVT_LAUNCH_CHECK("splitk_reduce_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// out[n, m] = in[m, n]   (64 x 64 tiles through shared memory; M, N even)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                             __nv_bfloat16* __restrict__ out, long long ldo, long long M, int N) {
  __shared__ unsigned short tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long m0 = static_cast<long long>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  for (int r = ty; r < 64; r += 8) {
    uint32_t v = 0;
    if (m0 + r < M && n0 + 2 * tx < N) v = *reinterpret_cast<const uint32_t*>(in + (m0 + r) * ldi + n0 + 2 * tx);
    tile[r][2 * tx] = static_cast<unsigned short>(v & 0xffffu);
    tile[r][2 * tx + 1] = static_cast<unsigned short>(v >> 16);
  }
  __syncthreads();
  for (int c = ty; c < 64; c += 8) {
    if (n0 + c < N && m0 + 2 * tx < M) {
      const uint32_t v = static_cast<uint32_t>(tile[2 * tx][c]) | (static_cast<uint32_t>(tile[2 * tx + 1][c]) << 16);
      *reinterpret_cast<uint32_t*>(out + static_cast<long long>(n0 + c) * ldo + m0 + 2 * tx) = v;
    }
  }
}

int transpose_bf16_dispatch(const __nv_bfloat16* in, long long ldi, __nv_bfloat16* out, long long ldo, long long M, int N,
                            cudaStream_t st) {
  LAVT_REQUIRE(M > 0 && N > 0 && N % 2 == 0 && ldi % 2 == 0 && ldo % 2 == 0, "transpose: N and the pitches must be even");
  // the kernel stores pairs along M: an odd M writes one extra (zero) element, which must fall inside the row pitch
  LAVT_REQUIRE(M % 2 == 0 || ldo > M, "transpose: odd M=%lld needs an output pitch > M (got %lld)", M, ldo);
  const long long mt = (M + 63) / 64, nt = (N + 63) / 64;
  LAVT_REQUIRE(mt < (1LL << 31) && nt < 65536, "transpose: grid too large (M=%lld, N=%d)", M, N);
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>(mt), static_cast<unsigned>(nt)), 256, 0, st>>>(in, ldi, out, ldo, M, N);
  LAVT_LAUNCH_CHECK("transpose_bf16_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// dst[n] += sum_m x[m, n]; block = 128 columns x one row chunk, 8 warps stride the rows
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ xv, long long ldx, long long M, int N, long long rows_per_block,
                                                     float* __restrict__ dst) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      if (BF16) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(xv) + r * ldx + col));
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(xv) + r * ldx + col));
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    }
  }
  *reinterpret_cast<float4*>(&red[warp][lane * 4]) = acc;
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dst + blockIdx.x * 128 + threadIdx.x, s);
  }
}

int colsum_dispatch(const void* x, int is_bf16, long long ldx, long long M, int N, float* dst, cudaStream_t st) {
  LAVT_REQUIRE(M > 0 && N > 0 && N % 4 == 0 && ldx % 4 == 0, "colsum: N and the pitch must be multiples of 4");
  const int cb = (N + 127) / 128;
  long long rb = (148 * 8 + cb - 1) / cb;
  if (rb > (M + 63) / 64) rb = (M + 63) / 64;
  if (rb < 1) rb = 1;
  if (rb > 65535) rb = 65535;
  const long long rpb = (M + rb - 1) / rb;
  dim3 grid(cb, static_cast<unsigned>((M + rpb - 1) / rpb));
  if (is_bf16) colsum_kernel<true><<<grid, 256, 0, st>>>(x, ldx, M, N, rpb, dst);
  else colsum_kernel<false><<<grid, 256, 0, st>>>(x, ldx, M, N, rpb, dst);
  LAVT_LAUNCH_CHECK("colsum_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// out[m, :] = bf16(x[src(m), :]);  src = identity, or the token of window row m (pad rows -> 0)
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out,
                                                        long long M, int C, const WinGeom win, const int use_win,
                                                        const float* __restrict__ rscale, const int rs_rows) {
  const int lane = threadIdx.x & 31;
  const long long m0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * 32;
  if (m0 >= M) return;
  long long myrow = m0 + lane;
  if (use_win) myrow = (m0 + lane < M) ? win_token(win, m0 + lane).row : -1;
  for (int r = 0; r < 32; ++r) {
    const long long m = m0 + r;
    if (m >= M) break;
    const long long row = __shfl_sync(0xffffffffu, myrow, r);
    const float sc = (rscale && row >= 0) ? __ldg(rscale + row / rs_rows) : 1.0f;
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
      uint2 o = make_uint2(0u, 0u);
      if (row >= 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * ldx) + c4);
        o = make_uint2(pack_bf16x2(v.x * sc, v.y * sc), pack_bf16x2(v.z * sc, v.w * sc));
      }
      reinterpret_cast<uint2*>(out + m * C)[c4] = o;
    }
  }
}

int cast_rows_dispatch(const float* x, long long ldx, __nv_bfloat16* out, long long M, int C, const WinGeom* win, const float* rscale,
                       int rs_rows, cudaStream_t st) {
  LAVT_REQUIRE(M > 0 && C % 4 == 0 && ldx % 4 == 0, "cast rows: channels / pitch must be multiples of 4");
  WinGeom g{};
  if (win) g = *win;
  const long long blocks = (M + 255) / 256;
  LAVT_REQUIRE(blocks < (1LL << 31), "cast rows: too many rows");
  cast_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, ldx, out, M, C, g, win ? 1 : 0, rscale, rs_rows);
  LAVT_LAUNCH_CHECK("cast_rows_kernel");
  return LAVT_OK;
}

// The same cast with the COLUMN SUMS of what it writes (fp32, before the bf16 rounding) added into colsum[C]: the bias gradient of the Linear
// layer whose output gradient this is (fc2 / proj of a Swin block) comes out of the pass that already reads every element, instead of a
// second pass over the bf16 copy.  A warp takes groups of 32 output rows (closed-form gather evaluated one row per lane), keeps its column
// sums in registers over all its groups, the block combines them in shared memory and issues one atomicAdd per column.
template <int NV>
__global__ void __launch_bounds__(256) cast_rows_colsum_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out,
                                                               long long M, int C, const WinGeom win, const int use_win,
                                                               const float* __restrict__ rscale, const int rs_rows, float* __restrict__ colsum) {
  extern __shared__ float crc_red[];      // [8][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long groups = (M + 31) / 32;
  for (long long g = static_cast<long long>(blockIdx.x) * 8 + warp; g < groups; g += static_cast<long long>(gridDim.x) * 8) {
    const long long m0 = g * 32;
    long long myrow = (m0 + lane < M) ? m0 + lane : -1;
    if (use_win && myrow >= 0) myrow = win_token(win, m0 + lane).row;
    const int rows = static_cast<int>((M - m0 < 32) ? M - m0 : 32);
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      const long long row = __shfl_sync(0xffffffffu, myrow, r);
      const float sc = (rscale && row >= 0) ? __ldg(rscale + row / rs_rows) : 1.0f;
      uint2* dst = reinterpret_cast<uint2*>(out + (m0 + r) * C);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 * 4 >= C) continue;
        uint2 o = make_uint2(0u, 0u);
        if (row >= 0) {
          float4 v = __ldg(reinterpret_cast<const float4*>(x + row * ldx) + c4);
          v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
          acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
          o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
        dst[c4] = o;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = i * 32 + lane;
    if (c4 * 4 < C) *reinterpret_cast<float4*>(&crc_red[warp * C + c4 * 4]) = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += crc_red[w * C + c];
    atomicAdd(colsum + c, s);
  }
}

int cast_rows_colsum_dispatch(const float* x, long long ldx, __nv_bfloat16* out, long long M, int C, const WinGeom* win, const float* rscale,
                              int rs_rows, float* colsum, cudaStream_t st) {
  LAVT_REQUIRE(M > 0 && C % 4 == 0 && ldx % 4 == 0 && colsum, "cast rows + column sums: channels / pitch must be multiples of 4");
  LAVT_REQUIRE(C <= 1024, "cast rows + column sums: at most 1024 channels (got %d)", C);
  WinGeom g{};
  if (win) g = *win;
  long long blocks = ((M + 31) / 32 + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;       // every block ends with C atomics onto the same C words: few, long-running blocks
  const size_t smem = static_cast<size_t>(8) * C * sizeof(float);
#define LAVT_CRC_CASE(nv)                                                                                                           \
  case nv: {                                                                                                                          \
    cast_rows_colsum_kernel<nv><<<static_cast<unsigned>(blocks), 256, smem, st>>>(x, ldx, out, M, C, g, win ? 1 : 0, rscale, rs_rows, colsum); \
    break;                                                                                                                            \
  }
  switch ((C + 127) / 128) {
    LAVT_CRC_CASE(1)
    LAVT_CRC_CASE(2)
    LAVT_CRC_CASE(3)
    LAVT_CRC_CASE(4)
    LAVT_CRC_CASE(5)
    LAVT_CRC_CASE(6)
    LAVT_CRC_CASE(7)
    LAVT_CRC_CASE(8)
    default:
      set_last_error("cast rows + column sums: %d channels unsupported", C);
      return LAVT_ERR_SHAPE;
  }
#undef LAVT_CRC_CASE
  LAVT_LAUNCH_CHECK("cast_rows_colsum_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(256) gelu_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ out,
                                                   long long count8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count8) return;
  const uint4 xv = __ldg(x + i);
  uint4 gv = make_uint4(0u, 0u, 0u, 0u);
  if (BWD) gv = __ldg(dy + i);
  const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = unpack_bf16x2(xs[j]);
    float r0, r1;
    if (BWD) {
      const float2 d = unpack_bf16x2(gs[j]);
      // d/dx [x Phi(x)] = Phi(x) + x phi(x)
      r0 = d.x * (0.5f * (1.0f + erff(a.x * 0.70710678118654752f)) + a.x * 0.3989422804014327f * __expf(-0.5f * a.x * a.x));
      r1 = d.y * (0.5f * (1.0f + erff(a.y * 0.70710678118654752f)) + a.y * 0.3989422804014327f * __expf(-0.5f * a.y * a.y));
    } else {
      r0 = gelu_erf(a.x);
      r1 = gelu_erf(a.y);
    }
    o[j] = pack_bf16x2(r0, r1);
  }
  out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

int gelu_fwd_dispatch(const __nv_bfloat16* x, __nv_bfloat16* y, long long count, cudaStream_t st) {
  LAVT_REQUIRE(count > 0 && count % 8 == 0, "gelu: element count must be a multiple of 8");
  const long long c8 = count / 8;
  gelu_kernel<false><<<static_cast<unsigned>((c8 + 255) / 256), 256, 0, st>>>(nullptr, reinterpret_cast<const uint4*>(x),
                                                                             reinterpret_cast<uint4*>(y), c8);
  LAVT_LAUNCH_CHECK("gelu_kernel");
  return LAVT_OK;
}
int gelu_bwd_dispatch(const __nv_bfloat16* dy, const __nv_bfloat16* x, __nv_bfloat16* dx, long long count, cudaStream_t st) {
  LAVT_REQUIRE(count > 0 && count % 8 == 0, "gelu backward: element count must be a multiple of 8");
  const long long c8 = count / 8;
  gelu_kernel<true><<<static_cast<unsigned>((c8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(x),
                                                                            reinterpret_cast<uint4*>(dx), c8);
  LAVT_LAUNCH_CHECK("gelu_kernel");
  return LAVT_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  A warp walks groups of LNB_G OUTPUT rows (the closed-form gathers are evaluated one row per lane and
// broadcast); d gamma / d beta accumulate in registers over all rows of the block, are combined in shared memory and leave
// the block as one atomicAdd per channel.  Each row is a chain of three dependent global round trips and four warp reductions,
// so throughput comes from resident warps: small row groups (many blocks) and, for widths up to 512, two blocks per SM
// (the first version -- 32-row groups, one 254-register block per SM -- ran at 1/7 of the HBM rate: 156 us for 18 432 x 512).
constexpr int LNB_G = 8;
template <int MODE, int NV>
__global__ void __launch_bounds__(256, (NV <= 4) ? 2 : 1) ln_bwd_kernel(const LnBwdParams p, const long long rows_per_block) {
  extern __shared__ float lnb_red[];      // [2 * Cn]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cn = (MODE == MODE_MERGE) ? 4 * p.C : p.C;
  const float inv_cn = 1.0f / static_cast<float>(Cn);
  for (int i = threadIdx.x; i < 2 * Cn; i += blockDim.x) lnb_red[i] = 0.f;
  __syncthreads();
  bool slot[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) slot[i] = (i * 32 + lane) * 4 < Cn;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long b0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long b1 = (b0 + rows_per_block < p.M) ? b0 + rows_per_block : p.M;
  const int H2 = (p.mH + 1) >> 1, W2 = (p.mW + 1) >> 1;
  const float4* g4 = reinterpret_cast<const float4*>(p.gamma);

  for (long long m0 = b0 + warp * LNB_G; m0 < b1; m0 += 8 * LNB_G) {
    long long myrow = m0 + lane;
    if (MODE == MODE_WINDOW) myrow = (lane < LNB_G && m0 + lane < b1) ? win_token(p.win, m0 + lane).row : -1;
#pragma unroll 1
    for (int r = 0; r < LNB_G; ++r) {
      const long long m = m0 + r;
      if (m >= b1) break;
      const long long row = __shfl_sync(0xffffffffu, myrow, r);
      if (row < 0) continue;                       // window pad row: produced as zeros, carries no gradient
      // every global operand of the row (x, dy, the gradient already in dx) is requested up front: one DRAM round trip per row
      // instead of three dependent ones (x -> statistics -> dy -> reductions -> dres; 87 us for 18 432 x 512 = 1.5 TB/s before)
      constexpr bool EARLY = NV <= 8;                // wider rows (PatchMerging at 1536 / 2048 channels) do not have the registers
      float4 v[NV], o[EARLY ? NV : 1];
      uint2 u[EARLY ? NV : 1];
      int moff[MODE == MODE_MERGE ? NV : 1];       // MODE_MERGE: element offset of the source token's channels, -1 = zero padding
      const long long rbase = row * static_cast<long long>(p.C);
      const uint2* dy2 = reinterpret_cast<const uint2*>(p.dy + m * p.lddy);
      if (MODE == MODE_MERGE) {
        const int w2 = static_cast<int>(m % W2);
        const int h2 = static_cast<int>((m / W2) % H2);
        const long long bd = m / (static_cast<long long>(W2) * H2);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int col = (i * 32 + lane) * 4;
          const int q = col / p.C, c = col - q * p.C;
          const int h = 2 * h2 + (q & 1), w = 2 * w2 + (q >> 1);
          const bool src_ok = slot[i] && h < p.mH && w < p.mW;          // zero padding of an odd grid: no source token
          const long long tok = (bd * p.mH + h) * p.mW + w;
          moff[i] = src_ok ? static_cast<int>(tok * p.C + c) : -1;
          v[i] = src_ok ? __ldg(reinterpret_cast<const float4*>(p.x + tok * p.ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>(p.x + row * p.ldx);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i] = slot[i] ? __ldg(src + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if constexpr (EARLY) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          u[i] = slot[i] ? __ldg(dy2 + i * 32 + lane) : make_uint2(0u, 0u);
          const bool live = (MODE == MODE_MERGE) ? moff[i] >= 0 : slot[i];
          const long long off = (MODE == MODE_MERGE) ? static_cast<long long>(moff[i]) : rbase + (i * 32 + lane) * 4;
          o[i] = (p.dres && live) ? *reinterpret_cast<const float4*>(p.dres + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) * inv_cn;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (!slot[i]) continue;
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = rsqrtf(warp_sum(ss) * inv_cn + p.eps);
      // xhat in place; gy = dy * gamma
      float4 gy[NV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (!slot[i]) { gy[i] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        uint2 uu;
        if constexpr (EARLY) uu = u[i];
        else uu = __ldg(dy2 + i * 32 + lane);
        const float2 d01 = unpack_bf16x2(uu.x), d23 = unpack_bf16x2(uu.y);
        const float4 gm = __ldg(g4 + i * 32 + lane);
        v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
        v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd;
        dg[i].x += d01.x * v[i].x; dg[i].y += d01.y * v[i].y; dg[i].z += d23.x * v[i].z; dg[i].w += d23.y * v[i].w;
        db[i].x += d01.x; db[i].y += d01.y; db[i].z += d23.x; db[i].w += d23.y;
        gy[i] = make_float4(d01.x * gm.x, d01.y * gm.y, d23.x * gm.z, d23.y * gm.w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * v[i].x + gy[i].y * v[i].y) + (gy[i].z * v[i].z + gy[i].w * v[i].w);
      }
      s1 = warp_sum(s1) * inv_cn;
      s2 = warp_sum(s2) * inv_cn;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool live = (MODE == MODE_MERGE) ? moff[i] >= 0 : slot[i];
        if (!live) continue;
        const long long off = (MODE == MODE_MERGE) ? static_cast<long long>(moff[i]) : rbase + (i * 32 + lane) * 4;
        float4 oo;
        if constexpr (EARLY) oo = o[i];
        else oo = p.dres ? *reinterpret_cast<const float4*>(p.dres + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 d;
        d.x = rstd * (gy[i].x - s1 - v[i].x * s2) + oo.x;
        d.y = rstd * (gy[i].y - s1 - v[i].y * s2) + oo.y;
        d.z = rstd * (gy[i].z - s1 - v[i].z * s2) + oo.z;
        d.w = rstd * (gy[i].w - s1 - v[i].w * s2) + oo.w;
        *reinterpret_cast<float4*>(p.dx + off) = d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (!slot[i]) continue;
    const int c = (i * 32 + lane) * 4;
    atomicAdd(&lnb_red[c + 0], dg[i].x); atomicAdd(&lnb_red[c + 1], dg[i].y);
    atomicAdd(&lnb_red[c + 2], dg[i].z); atomicAdd(&lnb_red[c + 3], dg[i].w);
    atomicAdd(&lnb_red[Cn + c + 0], db[i].x); atomicAdd(&lnb_red[Cn + c + 1], db[i].y);
    atomicAdd(&lnb_red[Cn + c + 2], db[i].z); atomicAdd(&lnb_red[Cn + c + 3], db[i].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cn; i += blockDim.x) {
    if (p.dgamma) atomicAdd(p.dgamma + i, lnb_red[i]);
    if (p.dbeta) atomicAdd(p.dbeta + i, lnb_red[Cn + i]);
  }
}

template <int MODE, int NV>
static void launch_ln_bwd_nv(const LnBwdParams& p, int Cn, cudaStream_t st) {
  constexpr int ROWS = 8 * LNB_G;                 // rows per block iteration
  long long blocks = (p.M + ROWS - 1) / ROWS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  long long rpb = (p.M + blocks - 1) / blocks;
  rpb = (rpb + ROWS - 1) / ROWS * ROWS;
  blocks = (p.M + rpb - 1) / rpb;
  ln_bwd_kernel<MODE, NV><<<static_cast<unsigned>(blocks), 256, 2 * Cn * sizeof(float), st>>>(p, rpb);
}

template <int MODE>
static int launch_ln_bwd(const LnBwdParams& p, int Cn, cudaStream_t st) {
  switch ((Cn + 127) / 128) {
    case 1: launch_ln_bwd_nv<MODE, 1>(p, Cn, st); break;
    case 2: launch_ln_bwd_nv<MODE, 2>(p, Cn, st); break;
    case 3: launch_ln_bwd_nv<MODE, 3>(p, Cn, st); break;
    case 4: launch_ln_bwd_nv<MODE, 4>(p, Cn, st); break;
    case 6: launch_ln_bwd_nv<MODE, 6>(p, Cn, st); break;
    case 8: launch_ln_bwd_nv<MODE, 8>(p, Cn, st); break;
    case 12: launch_ln_bwd_nv<MODE, 12>(p, Cn, st); break;
    case 16: launch_ln_bwd_nv<MODE, 16>(p, Cn, st); break;
    default:
      set_last_error("layernorm backward: normalised width %d not supported", Cn);
      return LAVT_ERR_SHAPE;
  }
  LAVT_LAUNCH_CHECK("ln_bwd_kernel");
  return LAVT_OK;
}

int ln_bwd_dispatch(int mode, const LnBwdParams& p, cudaStream_t st) {
  LAVT_REQUIRE(p.M > 0 && p.dy && p.dx && p.x && p.gamma, "layernorm backward: missing tensor");
  LAVT_REQUIRE(p.ldx % 4 == 0 && p.C % 4 == 0 && p.lddy % 4 == 0, "layernorm backward: pitch / channels must be multiples of 4");
  const int Cn = (mode == MODE_MERGE) ? 4 * p.C : p.C;
  LAVT_REQUIRE(mode != MODE_MERGE || 1LL * p.mH * p.mW * p.C * ((p.M + 1LL * ((p.mH + 1) / 2) * ((p.mW + 1) / 2) - 1) / (1LL * ((p.mH + 1) / 2) * ((p.mW + 1) / 2))) < (1LL << 31),
               "layernorm backward (merge): more than 2^31 source elements");
  if (mode == MODE_IDENTITY) return launch_ln_bwd<MODE_IDENTITY>(p, Cn, st);
  if (mode == MODE_WINDOW) return launch_ln_bwd<MODE_WINDOW>(p, Cn, st);
  if (mode == MODE_MERGE) return launch_ln_bwd<MODE_MERGE>(p, Cn, st);
  set_last_error("layernorm backward: bad mode %d", mode);
  return LAVT_ERR_SHAPE;
}

}  // namespace lavt
