// Shifted-window attention core (reference WindowAttention3D.forward, lib/video_swin_transformer.py:147-165;
// 2-D twin lib/backbone.py:127-138):  per (window, head)
//     S = q k^T  (+ relative-position bias, + shifted-window mask)  ->  softmax  ->  O = P v
// computed flash-style: the N x N score matrix never leaves registers.  q arrives pre-scaled by head_dim^-0.5
// (folded into the qkv GEMM epilogue).  The bias is gathered from the per-head table column held in shared
// memory through the closed form  idx(i,j) = code(i) - code(j) + const  and the -100 mask from per-token
// region ids (geom.cuh) -- neither the (N,N) index buffer nor the (nW,N,N) mask tensor exists on the device.
//
// Round-1 implementation: mma.sync.m16n8k16 (bf16 -> fp32) with cp.async double-buffered K/V tiles.  head_dim
// is 32 at every Swin stage, which makes this core exp/ALU bound rather than tensor bound (128 MMA flop per
// exp); a tcgen05/TMEM version is the next step.
#include "kernels.cuh"

namespace lavt {

constexpr int AT_HD = 32;        // head dim
constexpr int AT_KV = 64;        // keys per tile

// 16-byte chunk swizzle inside a 64-byte row so that ldmatrix (8 rows x 16 B) is bank-conflict free
__device__ __forceinline__ int kv_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) window_attn_kernel(const AttnParams p) {
  constexpr int BQ = WARPS * 16;
  constexpr int THREADS = WARPS * 32;
  extern __shared__ __align__(16) uint8_t smem[];
  const int N = p.win.N;
  uint8_t* ks = smem;                                   // [2][64][64 B]
  uint8_t* vs = ks + 2 * AT_KV * 64;                    // [2][64][64 B]
  int* info = reinterpret_cast<int*>(vs + 2 * AT_KV * 64);   // [N] code | rid << 16
  float* tab = reinterpret_cast<float*>(info + ((N + 3) & ~3));   // [L]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const long long wlin = blockIdx.z;                    // window index over B * nW
  const long long row0 = wlin * N;
  const int ld = 3 * p.C;
  const __nv_bfloat16* qbase = p.qkv + row0 * ld + head * AT_HD;
  const __nv_bfloat16* kbase = qbase + p.C;
  const __nv_bfloat16* vbase = qbase + 2 * p.C;
  const bool masked = (p.win.sd | p.win.sh | p.win.sw) != 0;
  const int ntiles = (N + AT_KV - 1) / AT_KV;

  auto load_tile = [&](int tile, int buf) {
    // 64 keys x 4 chunks for K and V = 512 16-byte copies
    for (int i = threadIdx.x; i < 2 * AT_KV * 4; i += THREADS) {
      const int isv = i >= AT_KV * 4;
      const int j = isv ? i - AT_KV * 4 : i;
      const int r = j >> 2, c = j & 3;
      const int key = tile * AT_KV + r;
      const bool ok = key < N;
      const __nv_bfloat16* src = (isv ? vbase : kbase) + static_cast<long long>(ok ? key : 0) * ld + c * 8;
      uint8_t* dst = (isv ? vs : ks) + buf * AT_KV * 64 + kv_off(r, c);
      cp_async_16(dst, src, ok);
    }
  };

  load_tile(0, 0);
  cp_async_commit();

  // per-token relative-position code and mask region (same for every head / q-tile of this window)
  for (int i = threadIdx.x; i < N; i += THREADS) {
    const WinTok tk = win_token(p.win, row0 + i);
    info[i] = tk.code | (tk.rid << 16);
  }
  for (int i = threadIdx.x; i < p.L; i += THREADS) tab[i] = __ldg(p.table + static_cast<long long>(i) * p.nH + head);
  const int rc = rel_const(p.win);

  // Q fragments (A operand, 16 rows x 32 d = 2 k-steps), straight from global
  uint32_t qf[2][4];
  const int qr0 = q0 + warp * 16 + g, qr1 = qr0 + 8;
  {
    const __nv_bfloat16* r0p = qbase + static_cast<long long>(min(qr0, N - 1)) * ld;
    const __nv_bfloat16* r1p = qbase + static_cast<long long>(min(qr1, N - 1)) * ld;
#pragma unroll
    for (int ksb = 0; ksb < 2; ++ksb) {
      qf[ksb][0] = __ldg(reinterpret_cast<const uint32_t*>(r0p + ksb * 16 + 2 * t));
      qf[ksb][1] = __ldg(reinterpret_cast<const uint32_t*>(r1p + ksb * 16 + 2 * t));
      qf[ksb][2] = __ldg(reinterpret_cast<const uint32_t*>(r0p + ksb * 16 + 8 + 2 * t));
      qf[ksb][3] = __ldg(reinterpret_cast<const uint32_t*>(r1p + ksb * 16 + 8 + 2 * t));
    }
  }
  __syncthreads();   // info / tab visible
  const int iq0 = info[min(qr0, N - 1)], iq1 = info[min(qr1, N - 1)];
  const int cq0 = (iq0 & 0xffff) + rc, cq1 = (iq1 & 0xffff) + rc;
  const int rq0 = iq0 >> 16, rq1 = iq1 >> 16;

  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float LOG2E = 1.4426950408889634f;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const uint8_t* kt = ks + buf * AT_KV * 64;
    const uint8_t* vt = vs + buf * AT_KV * 64;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kf[4];
      ldmatrix_x4(kf, kt + kv_off(nt * 8 + (lane & 7), lane >> 3));
      mma_bf16_16816(s[nt], qf[0], kf[0], kf[1]);
      mma_bf16_16816(s[nt], qf[1], kf[2], kf[3]);
    }
    // bias + mask + key padding
    const int kv0 = tile * AT_KV;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = kv0 + nt * 8 + 2 * t + e;
        if (key < N) {
          const int ik = info[key];
          const int ck = ik & 0xffff, rk = ik >> 16;
          float a = s[nt][e] + tab[cq0 - ck];
          float b = s[nt][2 + e] + tab[cq1 - ck];
          if (masked) {
            if (rk != rq0) a -= 100.0f;
            if (rk != rq1) b -= 100.0f;
          }
          s[nt][e] = a;
          s[nt][2 + e] = b;
          mx0 = fmaxf(mx0, a);
          mx1 = fmaxf(mx1, b);
        } else {
          s[nt][e] = -INFINITY;
          s[nt][2 + e] = -INFINITY;
        }
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float nm0 = fmaxf(m0, mx0), nm1 = fmaxf(m1, mx1);
    const float c0 = exp2f((m0 - nm0) * LOG2E), c1 = exp2f((m1 - nm1) * LOG2E);
    m0 = nm0;
    m1 = nm1;
    const float ms0 = nm0 * LOG2E, ms1 = nm1 * LOG2E;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * LOG2E - ms0), p1 = exp2f(s[nt][1] * LOG2E - ms0);
      const float p2 = exp2f(s[nt][2] * LOG2E - ms1), p3 = exp2f(s[nt][3] * LOG2E - ms1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {        // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {      // two d-chunks (8 each) per ldmatrix.x4
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, vt + kv_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        mma_bf16_16816(o[dp * 2 + 0], pf[kk], vf[0], vf[1]);
        mma_bf16_16816(o[dp * 2 + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();   // all warps done with buf before it is refilled
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = p.out + row0 * p.C + head * AT_HD;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    if (qr0 < N)
      *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qr0) * p.C + dt * 8 + 2 * t) = pack_bf16x2(o[dt][0] * i0, o[dt][1] * i0);
    if (qr1 < N)
      *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qr1) * p.C + dt * 8 + 2 * t) = pack_bf16x2(o[dt][2] * i1, o[dt][3] * i1);
  }
}

template <int WARPS>
static int launch_attn(const AttnParams& p, long long nwin, cudaStream_t st) {
  const int N = p.win.N;
  const size_t smem = 4 * AT_KV * 64 + static_cast<size_t>((N + 3) & ~3) * 4 + static_cast<size_t>(p.L) * 4;
  LAVT_REQUIRE(smem <= 200 * 1024, "attention: window too large for shared memory (N=%d, L=%d)", N, p.L);
  auto kfn = window_attn_kernel<WARPS>;
  static size_t configured = 0;
  if (smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  constexpr int BQ = WARPS * 16;
  dim3 grid((N + BQ - 1) / BQ, p.nH, static_cast<unsigned>(nwin));
  kfn<<<grid, WARPS * 32, smem, st>>>(p);
  LAVT_LAUNCH_CHECK("window_attn_kernel");
  return LAVT_OK;
}

int window_attn_dispatch(const AttnParams& p, cudaStream_t st) {
  const WinGeom& g = p.win;
  LAVT_REQUIRE(p.C == p.nH * AT_HD, "attention: head_dim must be 32 (C=%d, heads=%d)", p.C, p.nH);
  LAVT_REQUIRE(g.N == g.wd * g.wh * g.ww && g.N > 0, "attention: inconsistent window geometry");
  LAVT_REQUIRE(p.L == (2 * g.Wd - 1) * (2 * g.Wh - 1) * (2 * g.Ww - 1), "attention: bias table rows %d do not match window", p.L);
  LAVT_REQUIRE(p.L < 65536, "attention: bias table too large");
  LAVT_REQUIRE(g.N <= g.Wd * g.Wh * g.Ww, "attention: effective window larger than configured window");
  const long long nwin = 1LL * g.B * g.nwd * g.nwh * g.nww;
  LAVT_REQUIRE(nwin > 0 && nwin < 65536, "attention: window count %lld out of range", nwin);
  if (g.N <= 64) return launch_attn<4>(p, nwin, st);
  return launch_attn<8>(p, nwin, st);
}

}  // namespace lavt
