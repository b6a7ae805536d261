// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate)
// with a fused epilogue (column scale, bias, GELU/ReLU/tanh, elementwise multiply, residual add,
// window-reverse / un-shift scatter) and an implicit-GEMM 3x3 convolution mode whose A tiles are
// fetched tap-by-tap with 4-D TMA boxes (zero fill outside the image == conv padding).
//
// This one kernel carries every dense contraction of the hot path except the N x N attention
// core: qkv / proj / fc1 / fc2 (reference lib/video_swin_transformer.py:144,166,30-36),
// PatchMerging.reduction (:309), PatchEmbed3D.proj (:627), the PWAM 1x1 convs (:900-973),
// LanguageGate (:519-525) and the SimpleDecoding conv3x3+BN+ReLU stack (lib/mask_predictor.py:56-87).
//
// Structure: PERSISTENT, one CTA per SM, 384 threads, tiles of 128 x 128 handed out round-robin
// (N fastest, so concurrently running CTAs share the same A rows through L2):
//   warp 0    : TMA producer (one lane) -- smem ring of STAGES x {A 128x64, B 128x64} bf16, 128-byte swizzle,
//               mbarrier full/empty pairs; the ring streams straight across tile boundaries
//   warp 1    : MMA issuer (one lane)   -- tcgen05.mma.cta_group::1.kind::f16 into one of TWO TMEM accumulators
//   warp 2    : TMEM allocator (256 columns)
//   warps 4-7 : epilogue warpgroup 0 (even local tiles, accumulator 0)
//   warps 8-11: epilogue warpgroup 1 (odd local tiles, accumulator 1)
// so the epilogue of tile i (tcgen05.ld -> fused math -> global stores) overlaps the MMAs of tiles i+1, i+2.
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lavt {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;            // 64 bf16 = one 128-byte swizzle row

// GEMM_BN columns per tile, GEMM_STAGES smem ring slots, EPI_WGS epilogue warpgroups (1 -> 256 threads, 2 CTAs/SM;
// 2 -> 384 threads, 1 CTA/SM).  Two TMEM accumulators of GEMM_BN columns in every configuration.
// CTA2: CTA-pair variant (tcgen05 cta_group::2): a pair of CTAs computes a 256 x GEMM_BN tile; each CTA holds its 128 rows of A and HALF
// of the B rows, the leader issues the MMAs for both, every CTA drains its own 128 accumulator lanes.
template <int GEMM_BN, int GEMM_STAGES, int EPI_WGS, int CTA2 = 0>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (CTA2 ? GEMM_BN / 2 : GEMM_BN) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = GEMM_STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = RING_BYTES;                  // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem ptr
  static constexpr int VEC_OFF = BAR_OFF + 256;               // per epilogue warpgroup: scale[BN], bias[BN]
  static constexpr int VEC_TILE_BYTES = EPI_WGS * 2 * GEMM_BN * 4;   // per-tile scale / bias staging (fallback)
  static constexpr int TOTAL = VEC_OFF + VEC_TILE_BYTES + 1024 /*alignment slack*/;
  static constexpr int THREADS = 128 + 128 * EPI_WGS;
  static constexpr int MIN_CTAS = (EPI_WGS == 1) ? 2 : 1;
  // TMEM accumulator stages: all 512 columns when the CTA owns the SM (4 x 128 or 2 x 256), half of them with 2 CTAs/SM.
  // The epilogue is bound by TMEM -> register bandwidth while the mainloop runs (~30 B/clk/SM, ncu + clock64 traces), so
  // with short K it takes longer than the MMAs of a tile; more stages let the MMA warp run ahead instead of idling.
  static constexpr int ACC = (MIN_CTAS == 2) ? 2 : 512 / GEMM_BN;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100): one issue slot for two elements
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// exact-erf GELU (nn.GELU default) for two values at once, no MUFU:  erf(z) = z * P(z^2) on |z| <= 3 (degree-8 minimax fit,
// |err| <= 4e-5 including the tail beyond 3 where erf is taken as +-1 exactly; the coefficients are normalised so that
// 3 * P(9) == 1).  Horner in packed fp32x2: 8.5 issue slots per element instead of ~18 + 2 MUFU.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t x = pk2(x0, x1);
  float z0, z1;
  upk2(mul2(x, pk2(0.70710678118654752f, 0.70710678118654752f)), z0, z1);
  z0 = fminf(fmaxf(z0, -3.0f), 3.0f);
  z1 = fminf(fmaxf(z1, -3.0f), 3.0f);
  const uint64_t z = pk2(z0, z1);
  const uint64_t t = mul2(z, z);
  constexpr float K = 1.0000220904f;      // 1 / (3 * P(9)) of the raw fit
  uint64_t p = pk2(4.0742095563e-08f * K, 4.0742095563e-08f * K);
  p = fma2(p, t, pk2(-1.9448222676e-06f * K, -1.9448222676e-06f * K));
  p = fma2(p, t, pk2(4.1060515983e-05f * K, 4.1060515983e-05f * K));
  p = fma2(p, t, pk2(-5.1103678896e-04f * K, -5.1103678896e-04f * K));
  p = fma2(p, t, pk2(4.2354270142e-03f * K, 4.2354270142e-03f * K));
  p = fma2(p, t, pk2(-2.5102860078e-02f * K, -2.5102860078e-02f * K));
  p = fma2(p, t, pk2(1.1107933337e-01f * K, 1.1107933337e-01f * K));
  p = fma2(p, t, pk2(-3.7531487405e-01f * K, -3.7531487405e-01f * K));
  p = fma2(p, t, pk2(1.1282684285e+00f * K, 1.1282684285e+00f * K));
  const uint64_t e = mul2(z, p);
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  upk2(fma2(hx, e, hx), x0, x1);
}
// GELU and GELU'(x) = Phi(x) + x phi(x) for two values: Phi from the same erf polynomial as gelu_erf2, phi from one ex2.
// x0 / x1 become GELU(x), d0 / d1 the derivative.
__device__ __forceinline__ void gelu_both2(float& x0, float& x1, float& d0, float& d1) {
  const uint64_t x = pk2(x0, x1);
  float z0, z1;
  upk2(mul2(x, pk2(0.70710678118654752f, 0.70710678118654752f)), z0, z1);
  const uint64_t zz = pk2(z0, z1);
  float q0, q1;
  upk2(mul2(zz, zz), q0, q1);                  // x^2 / 2
  const float e0 = ex2_approx(-1.4426950408889634f * q0), e1 = ex2_approx(-1.4426950408889634f * q1);
  z0 = fminf(fmaxf(z0, -3.0f), 3.0f);
  z1 = fminf(fmaxf(z1, -3.0f), 3.0f);
  const uint64_t z = pk2(z0, z1);
  const uint64_t t = mul2(z, z);
  constexpr float K = 1.0000220904f;
  uint64_t p = pk2(4.0742095563e-08f * K, 4.0742095563e-08f * K);
  p = fma2(p, t, pk2(-1.9448222676e-06f * K, -1.9448222676e-06f * K));
  p = fma2(p, t, pk2(4.1060515983e-05f * K, 4.1060515983e-05f * K));
  p = fma2(p, t, pk2(-5.1103678896e-04f * K, -5.1103678896e-04f * K));
  p = fma2(p, t, pk2(4.2354270142e-03f * K, 4.2354270142e-03f * K));
  p = fma2(p, t, pk2(-2.5102860078e-02f * K, -2.5102860078e-02f * K));
  p = fma2(p, t, pk2(1.1107933337e-01f * K, 1.1107933337e-01f * K));
  p = fma2(p, t, pk2(-3.7531487405e-01f * K, -3.7531487405e-01f * K));
  p = fma2(p, t, pk2(1.1282684285e+00f * K, 1.1282684285e+00f * K));
  const uint64_t Phi = fma2(mul2(z, p), pk2(0.5f, 0.5f), pk2(0.5f, 0.5f));                    // 0.5 + 0.5 erf(x / sqrt 2)
  const uint64_t xs = mul2(x, pk2(0.3989422804014327f, 0.3989422804014327f));
  upk2(fma2(xs, pk2(e0, e1), Phi), d0, d1);
  upk2(mul2(x, Phi), x0, x1);
}
__device__ __forceinline__ void gelu_grad2(float& x0, float& x1) {
  float d0, d1;
  gelu_both2(x0, x1, d0, d1);
  x0 = d0;
  x1 = d1;
}
// 256-bit global accesses (sm_100): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void ldg_v8(uint32_t* r, const void* p) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// tanh(x) = 1 - 2 / (exp(2x) + 1), abs error ~1e-7
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return fmaf(-2.0f, rcp_approx(e + 1.0f), 1.0f);
}

// SIMPLE = 1: the epilogue is compiled for bias / column scale + activation + bf16 output in GEMM-row order only (no mul, residual, DropPath
// scale, fp32 or pre-activation output, no row map), which leaves room for FOUR epilogue warpgroups (640 threads, <= 102 registers): the
// epilogue of the short-K launches is a flat chain of dependent-issue latencies at 2 warps per scheduler (ncu: issue slots 28 %, no dominant
// stall), so it needs warps, not bandwidth.
template <int GEMM_BN, int GEMM_STAGES, int EPI_WGS, int CTA2 = 0, int SIMPLE = 0>
__global__ void __launch_bounds__(GemmSmem<GEMM_BN, GEMM_STAGES, EPI_WGS, CTA2>::THREADS, GemmSmem<GEMM_BN, GEMM_STAGES, EPI_WGS, CTA2>::MIN_CTAS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p, const int m_tiles, const int vec_all, long long* const trace_buf) {
  using L = GemmSmem<GEMM_BN, GEMM_STAGES, EPI_WGS, CTA2>;
  // CTA pair: rank 0 (leader) issues the MMAs; work items are handed to PAIRS (cluster index), the pair's CTA r owns m-tile 2 * mp + r
  const int crank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int wid0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int wstep = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  // aligned as an OFFSET so the compiler keeps the shared address space (LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  constexpr int ACC = L::ACC;
  uint64_t* tmem_full_bar = empty_bar + GEMM_STAGES;      // [ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;         // [ACC]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* const trace = (blockIdx.x == 0) ? trace_buf : nullptr;     // debug (LAVT_GEMM_TRACE): clock64 stamps of CTA 0
#define GEMM_TRACE(ev, lt) do { if (trace && (lt) < 32) trace[(ev) * 32 + (lt)] = clock64(); } while (0)
  // N and K need not be multiples of the tile: TMA zero-fills operand rows / columns past the tensor edge (so the extra
  // accumulator columns are exact zeros and the extra K contributes nothing) and the epilogue skips columns >= N.
  // conv: each tap spans ceil(Cin / 64) k-blocks; a block's surplus channels are zero-filled on the A side, which makes the
  // (finite) next-tap weights the B box picks up there irrelevant.
  const int n_tiles = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int total_tiles = (CTA2 ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;      // CTA pair: tiles of 256 rows
  const int conv_cpb = (p.cCin + GEMM_BK - 1) / GEMM_BK;
  const int num_kb = (p.rowmap == ROWMAP_CONV) ? p.taps * conv_cpb
                     : (p.rowmap == ROWMAP_WGCONV) ? (p.K / GEMM_BK) : (p.K + GEMM_BK - 1) / GEMM_BK;
  // split-K: work item = (tile, split); every split owns at least one k-block (host guarantees (ksplit - 1) * kbs < num_kb)
  const int nsplit = p.ksplit > 1 ? p.ksplit : 1;
  const int kbs = p.ksplit > 1 ? p.kbs : num_kb;
  const int total_work = total_tiles * nsplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], CTA2 ? 8 : 4);      // one arrive per epilogue warp of the owning warpgroup (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_2cta(tmem_ptr_smem, ACC * GEMM_BN);
    else tmem_alloc(tmem_ptr_smem, ACC * GEMM_BN);  // ACC fp32 accumulators of GEMM_BN columns
  }
  pdl_wait();                      // everything above touched no global data: it overlaps the previous kernel's tail
  pdl_launch_dependents();
  if (vec_all) {
    // column scale / bias of ALL N columns staged once per CTA: the epilogue never waits on global memory for them
    float* vec = reinterpret_cast<float*>(smem + L::VEC_OFF);
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      vec[i] = p.cscale ? __ldg(p.cscale + i) : 1.0f;
      vec[p.N + i] = p.bias ? __ldg(p.bias + i) : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();        // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const bool conv = (p.rowmap == ROWMAP_CONV);
      const int cpb = conv ? conv_cpb : 1;
      int it = 0;
      for (int work = wid0; work < total_work; work += wstep) {
        const int ks = work / total_tiles, tile = work - ks * total_tiles;
        const int kb_lo = ks * kbs, kb_hi = min(num_kb, kb_lo + kbs);
        const int mt = CTA2 ? 2 * (tile / n_tiles) + crank : tile / n_tiles;       // (a pair's odd m-tile may lie past the end: TMA zero-fills)
        const int n0 = (tile - (tile / n_tiles) * n_tiles) * GEMM_BN;
        const int nb0 = CTA2 ? n0 + crank * (GEMM_BN / 2) : n0;                    // this CTA's B rows
        const uint32_t lbar_base = CTA2 ? leader_bar_addr(&full_bar[0]) : 0u;
        int img = 0, h0 = 0, w0 = 0, d0 = 0;
        if (conv) {
          const int tw = mt % p.cTilesW;
          const int th = (mt / p.cTilesW) % p.cTilesH;
          img = mt / (p.cTilesW * p.cTilesH);          // 2-D: image; 3-D: clip * D + frame
          if (p.taps == 27) {
            d0 = img % p.cD;
            img /= p.cD;
          }
          h0 = th * p.cTH;
          w0 = tw * p.cTW;
        }
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % GEMM_STAGES;
          const uint32_t ph = (it / GEMM_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (kb == kb_lo) GEMM_TRACE(5, it / num_kb);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if constexpr (CTA2) {
            // the leader's barrier counts the bytes of BOTH CTAs; each CTA's loads complete on it
            if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          } else {
            mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          }
          const uint32_t lbar = lbar_base + s * 8;
          int bk0 = kb * GEMM_BK;                      // first K index of this block in the weight matrix
          if (conv) {
            const int tap = kb / cpb, cc = kb - tap * cpb;
            bk0 = tap * p.cCin + cc * GEMM_BK;
            if (p.taps == 27) {
              // 3x3x3: tap = (kz * 3 + ky) * 3 + kx; frames outside the clip are zero-filled by TMA like the spatial border
              const int dz = tap / 9 - 1, r9 = tap % 9;
              if constexpr (CTA2) tma_load_5d_2sm(sa, &tmA, lbar, cc * GEMM_BK, w0 + r9 % 3 - 1, h0 + r9 / 3 - 1, d0 + dz, img);
              else tma_load_5d(sa, &tmA, &full_bar[s], cc * GEMM_BK, w0 + r9 % 3 - 1, h0 + r9 / 3 - 1, d0 + dz, img);
            } else {
              const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
              const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
              if constexpr (CTA2) tma_load_4d_2sm(sa, &tmA, lbar, cc * GEMM_BK, w0 + dx, h0 + dy, img);
              else tma_load_4d(sa, &tmA, &full_bar[s], cc * GEMM_BK, w0 + dx, h0 + dy, img);
            }
          } else if (p.rowmap == ROWMAP_WGCONV) {
            // conv weight gradient: k-block = one 64-pixel tile of one image / frame; A = dz channels, B = x channels shifted by the tap
            const int tw = kb % p.cTilesW;
            const int th = (kb / p.cTilesW) % p.cTilesH;
            int im = kb / (p.cTilesW * p.cTilesH);
            const int ph0 = th * p.cTH, pw0 = tw * p.cTW;
            int fr = 0;
            if (p.taps == 27) { fr = im % p.cD; im /= p.cD; }     // 3-D: im = clip, fr = frame
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) {
              if (p.taps == 27) tma_load_5d(sa + j * 8192, &tmA, &full_bar[s], mt * GEMM_BM + j * 64, pw0, ph0, fr, im);
              else tma_load_4d(sa + j * 8192, &tmA, &full_bar[s], mt * GEMM_BM + j * 64, pw0, ph0, im);
            }
#pragma unroll
            for (int j = 0; j < GEMM_BN / 64; ++j) {
              const int col = n0 + j * 64;                       // column of C = tap * Cin + ci (Cin % 64 == 0: an atom never straddles taps)
              int tap = col / p.cCin, ci0 = col - tap * p.cCin;
              if (tap >= p.taps) { tap = p.taps - 1; ci0 = p.cCin; }   // past N: channel coordinate out of bounds -> zero fill
              if (p.taps == 27) {
                const int r9 = tap % 9;
                tma_load_5d(sb + j * 8192, &tmB, &full_bar[s], ci0, pw0 + r9 % 3 - 1, ph0 + r9 / 3 - 1, fr + tap / 9 - 1, im);
              } else {
                tma_load_4d(sb + j * 8192, &tmB, &full_bar[s], ci0, pw0 + tap % 3 - 1, ph0 + tap / 3 - 1, im);
              }
            }
            continue;
          } else if (p.mnmajor) {
            // operands stored [K, MN] row-major: boxes of 64 k-rows x 64 MN elements, one 8 KB box per 64-wide MN atom
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], mt * GEMM_BM + j * 64, kb * GEMM_BK);
#pragma unroll
            for (int j = 0; j < GEMM_BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, kb * GEMM_BK + p.b_koff);
            continue;
          } else {
            if constexpr (CTA2) tma_load_2d_2sm(sa, &tmA, lbar, kb * GEMM_BK, mt * GEMM_BM);
            else tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM);
          }
          if constexpr (CTA2) tma_load_2d_2sm(sb, &tmB, lbar, bk0 + p.b_koff, nb0);
          else tma_load_2d(sb, &tmB, &full_bar[s], bk0 + p.b_koff, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc_k = make_idesc_bf16_f32(CTA2 ? 2 * GEMM_BM : GEMM_BM, GEMM_BN);
      const uint32_t idesc = p.mnmajor ? (idesc_k | (1u << 15) | (1u << 16)) : idesc_k;      // bits 15 / 16: A / B are MN-major
      int it = 0, lt = 0;
      for (int work = wid0; work < total_work; work += wstep, ++lt) {
        const int ks = work / total_tiles;
        const int kb_lo = ks * kbs, kb_hi = min(num_kb, kb_lo + kbs);
        const int acc = lt % ACC;
        const uint32_t use = static_cast<uint32_t>(lt / ACC);
        mbar_wait(&tmem_empty_bar[acc], (use & 1) ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        GEMM_TRACE(0, lt);
        const uint32_t tmem_d = tmem_base + acc * GEMM_BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % GEMM_STAGES;
          const uint32_t ph = (it / GEMM_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == kb_lo) GEMM_TRACE(1, lt);
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          if (p.mnmajor) {
            const uint64_t da = make_mnmajor_sw128_desc(sa, 8192);
            const uint64_t db = make_mnmajor_sw128_desc(sb, 8192);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // 16 k-rows of 128 B = 2048 bytes per k-step: +128 in (addr >> 4) units
              umma_bf16_ss(tmem_d, da + 128 * k, db + 128 * k, idesc, ((kb - kb_lo) | k) != 0);
            }
          } else {
            const uint64_t da = make_kmajor_sw128_desc(sa);
            const uint64_t db = make_kmajor_sw128_desc(sb);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4) units
              if constexpr (CTA2) umma_bf16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb_lo) | k) != 0);
              else umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb_lo) | k) != 0);
            }
          }
          if constexpr (CTA2) umma_commit_2cta(&empty_bar[s]);     // frees the slot in BOTH CTAs
          else umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs retire
        }
        if constexpr (CTA2) umma_commit_2cta(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);      // accumulator complete
        GEMM_TRACE(2, lt);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups =====================
    const int wg = (warp - 4) >> 2;            // epilogue warpgroup: local tiles with lt % EPI_WGS == wg
    const int ew = warp & 3;                   // TMEM lanes [32*ew, 32*ew+32)
    const int et = threadIdx.x - 128 - wg * 128;
    float* const vec = reinterpret_cast<float*>(smem + L::VEC_OFF);
    const float* s_scale = vec + wg * 2 * GEMM_BN;
    const float* s_bias = s_scale + GEMM_BN;
    const int r = ew * 32 + lane;              // row inside the tile
    int lt = 0;
    for (int work = wid0; work < total_work; work += wstep, ++lt) {
      if ((lt % EPI_WGS) != wg) continue;
      const int ks = work / total_tiles, tile = work - ks * total_tiles;
      const int acc = lt % ACC;
      const uint32_t use = static_cast<uint32_t>(lt / ACC);
      const int mt = CTA2 ? 2 * (tile / n_tiles) + crank : tile / n_tiles;
      const int n0 = (tile - (tile / n_tiles) * n_tiles) * GEMM_BN;

      if (vec_all) {
        s_scale = vec + n0;
        s_bias = vec + p.N + n0;
      } else {
        // previous tile's readers of the staging vectors are done -> refill for this tile's columns
        float* w_scale = vec + wg * 2 * GEMM_BN;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int i = et; i < GEMM_BN; i += 128) {
          const bool in = n0 + i < p.N;
          w_scale[i] = (p.cscale && in) ? __ldg(p.cscale + n0 + i) : 1.0f;
          w_scale[GEMM_BN + i] = (p.bias && in) ? __ldg(p.bias + n0 + i) : 0.0f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
      }

      long long m = -1, orow = -1;
      if (!SIMPLE && p.rowmap == ROWMAP_CONV) {
        const int tw = mt % p.cTilesW;
        const int th = (mt / p.cTilesW) % p.cTilesH;
        const int img = mt / (p.cTilesW * p.cTilesH);
        const int h = th * p.cTH + r / p.cTW, w = tw * p.cTW + r % p.cTW;
        if (h < p.cH && w < p.cW && mt < m_tiles) {
          m = (static_cast<long long>(img) * p.cH + h) * p.cW + w;
          orow = m;
        }
      } else {
        const long long mm = static_cast<long long>(mt) * GEMM_BM + r;
        if (mm < p.M) {
          m = mm;
          orow = (!SIMPLE && p.rowmap == ROWMAP_WINDOW) ? win_token(p.win, mm).row : mm;
        }
      }
      const bool live = (orow >= 0);
      const float row_scale = (!SIMPLE && p.rscale && live) ? __ldg(p.rscale + orow / p.rs_rows) : 1.0f;
      const __nv_bfloat16* mul_row = (!SIMPLE && p.mul && live) ? p.mul + m * p.ldm + n0 : nullptr;
      const float* res_row = (!SIMPLE && p.resid && live) ? p.resid + orow * p.ldo + n0 : nullptr;
      float* of_row = (!SIMPLE && p.out_f32 && live) ? p.out_f32 + ks * p.split_stride + orow * p.ldo + n0 : nullptr;
      __nv_bfloat16* ob_row = (p.out_bf16 && live) ? p.out_bf16 + orow * p.ldo + n0 : nullptr;
      __nv_bfloat16* op_row = (!SIMPLE && p.out_pre && live) ? p.out_pre + orow * p.ldo + n0 : nullptr;

      const uint32_t tbase = tmem_base + acc * GEMM_BN + (static_cast<uint32_t>(ew * 32) << 16);
      bool acc_ready = false;
      const int nch = min(GEMM_BN / 32, (p.N - n0 + 31) / 32);      // 32-column chunks of this tile that exist (N % 32 == 0)

#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 32; ++c) {
        // operands that do not depend on the accumulator are requested first (for c == 0: before the MMAs finish)
        uint32_t rr[32];
        if (res_row && c < nch) {
#pragma unroll
          for (int q = 0; q < 4; ++q) ldg_v8(&rr[q * 8], res_row + c * 32 + q * 8);
        }
        if (!acc_ready) {
          mbar_wait(&tmem_full_bar[acc], use & 1);
          tc_fence_after();
          acc_ready = true;
          if (ew == 0 && lane == 0) GEMM_TRACE(3, lt);
        }
        uint32_t v[32];
        tmem_ld_32x32b_x32(tbase + c * 32, v);          // warp-collective: executed by all lanes
        tmem_ld_wait();
        if (c == GEMM_BN / 32 - 1) {
          // last read of this accumulator: hand it back to the MMA warp before finishing the math / stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_leader(&tmem_empty_bar[acc]);      // the leader's MMA thread waits for both CTAs' drains
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        if (!live || c >= nch) continue;
        float f[32];
        if (p.cscale) {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c * 32);
          const float4* bi4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 sc = sc4[q], bi = bi4[q];
            upk2(fma2(pk2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), pk2(sc.x, sc.y), pk2(bi.x, bi.y)), f[4 * q], f[4 * q + 1]);
            upk2(fma2(pk2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), pk2(sc.z, sc.w), pk2(bi.z, bi.w)), f[4 * q + 2],
                 f[4 * q + 3]);
          }
        } else if (p.bias) {       // no column scale (most launches): half the shared-memory loads of the staged vectors
          const float4* bi4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bi = bi4[q];
            f[4 * q] = __uint_as_float(v[4 * q]) + bi.x;
            f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bi.y;
            f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bi.z;
            f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bi.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        }
        if (op_row && p.pre_mode == 1) {     // training forward of fc1: GELU(pre) and GELU'(pre) from one polynomial
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float d0, d1;
              gelu_both2(f[q * 16 + 2 * j], f[q * 16 + 2 * j + 1], d0, d1);
              u[j] = pack_bf16x2(d0, d1);
            }
            stg_v8(op_row + c * 32 + q * 16, u);
          }
        } else if (op_row) {                 // training forward: the pre-activation is saved next to the activated output
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(f[q * 16 + 2 * j], f[q * 16 + 2 * j + 1]);
            stg_v8(op_row + c * 32 + q * 16, u);
          }
        }
        if (op_row && p.pre_mode == 1) {
          // activation already applied above
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1]);
        } else if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
        } else if (!SIMPLE && p.act == ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
        } else if (!SIMPLE && p.act == ACT_SIGMOID) {       // sigmoid(x) = 0.5 tanh(x / 2) + 0.5
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaf(0.5f, tanh_fast(0.5f * f[j]), 0.5f);
        }
        if (mul_row) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t u[8];
            ldg_v8(u, mul_row + c * 32 + q * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float2 a = unpack_bf16x2(u[j]);
              if (p.mul_act == ACT_GELU) gelu_grad2(a.x, a.y);
              f[q * 16 + 2 * j] *= a.x;
              f[q * 16 + 2 * j + 1] *= a.y;
            }
          }
        }
        if (!SIMPLE && p.rscale) {       // DropPath: the whole branch of a dropped sample is scaled (0 or 1 / keep_prob) before the shortcut add
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= row_scale;
        }
        if (res_row) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(rr[j]);
        }
        if (of_row) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = __float_as_uint(f[q * 8 + j]);
            stg_v8(of_row + c * 32 + q * 8, u);
          }
        }
        if (ob_row) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(f[q * 16 + 2 * j], f[q * 16 + 2 * j + 1]);
            stg_v8(ob_row + c * 32 + q * 16, u);
          }
        }
      }
      if (ew == 0 && lane == 0) GEMM_TRACE(4, lt);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();        // the peer may still be read by the leader's MMAs / signalled by its commits
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, ACC * GEMM_BN);
    else tmem_dealloc(tmem_base, ACC * GEMM_BN);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int BN, int STAGES, int EPI_WGS, int CTA2 = 0, int SIMPLE = 0>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int m_tiles, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, EPI_WGS, CTA2>;
  auto kfn = gemm_bf16_tc_kernel<BN, STAGES, EPI_WGS, CTA2, SIMPLE>;
  // stage scale / bias of all N columns when that fits next to the operand ring (per-CTA budget: the whole SM, or half of
  // it for the 2-CTAs/SM variant); otherwise the kernel refills a per-tile staging area
  const int budget = (L::MIN_CTAS == 1 ? 227 * 1024 : 113 * 1024) - (L::VEC_OFF + 1024);
  const int vec_all = (2 * p.N * 4 <= budget) ? 1 : 0;
  const int smem = L::VEC_OFF + 1024 + (vec_all ? (2 * p.N * 4 > L::VEC_TILE_BYTES ? 2 * p.N * 4 : L::VEC_TILE_BYTES) : L::VEC_TILE_BYTES);
  static int configured = 0;
  if (smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const long long total = 1LL * (CTA2 ? (m_tiles + 1) / 2 : m_tiles) * ((p.N + BN - 1) / BN) * (p.ksplit > 1 ? p.ksplit : 1);
  LAVT_REQUIRE(total < (1LL << 30), "gemm: too many tiles (%lld)", total);
  const long long slots = CTA2 ? sm_count() / 2 : 1LL * sm_count() * L::MIN_CTAS;      // CTA pairs: one pair per TPC
  const int grid = static_cast<int>(total < slots ? total : slots) * (CTA2 ? 2 : 1);
  long long* trace = nullptr;
  const char* trace_path = getenv("LAVT_GEMM_TRACE");
  if (trace_path) {
    LAVT_CUDA(cudaMalloc(&trace, 6 * 32 * sizeof(long long)));
    LAVT_CUDA(cudaMemsetAsync(trace, 0, 6 * 32 * sizeof(long long), stream));
  }
  LAVT_CUDA(launch_pdl(kfn, dim3(grid), dim3(L::THREADS), smem, stream, CTA2 ? 2 : 1, tmA, tmB, p, m_tiles, vec_all, trace));
  LAVT_LAUNCH_CHECK("gemm_bf16_tc_kernel");
  if (trace) {
    // debug only: synchronous dump of CTA 0's event clocks (rows: acc free, first operands, MMAs issued, acc ready, epilogue end, first TMA)
    long long host[6 * 32];
    LAVT_CUDA(cudaStreamSynchronize(stream));
    LAVT_CUDA(cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int e = 0; e < 6; ++e) {
        for (int t = 0; t < 32; ++t) fprintf(f, "%lld ", host[e * 32 + t]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return LAVT_OK;
}

// 0: 128-wide tiles, 6 stages, two epilogue warpgroups, 1 CTA/SM
// 1: 128-wide tiles, 3 stages, one epilogue warpgroup, 2 CTAs/SM
// 2: 256-wide tiles, 4 stages, two epilogue warpgroups, 1 CTA/SM (N % 256 == 0)
// 4: CTA PAIRS (cta_group::2): 256 x 256 tiles over two SMs, 6 stages of (A 128 x 64 | half of B 128 x 64) per CTA (N % 256 == 0,
//    K-major operands, no split-K)
static int gemm_variant(const GemmParams& p) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("LAVT_GEMM_VARIANT");
    forced = e ? atoi(e) : -1;
  }
  int v;
  if (forced >= 0) {
    v = forced;
  } else {
    // measured on B200 (tools/bench_gemm.py): 256-wide tiles halve the B-operand smem traffic per FLOP and win
    // whenever they still fill the machine; 2 CTAs/SM wins for long K at N = 128; otherwise the 1-CTA/SM pipeline
    const long long m_tiles = (p.rowmap == ROWMAP_CONV) ? (1LL * (p.M / (p.cH * p.cW)) * p.cTilesH * p.cTilesW)
                                                        : ((p.M + GEMM_BM - 1) / GEMM_BM);
    const long long splits = p.ksplit > 1 ? p.ksplit : 1;                         // split-K work items fill the machine too
    if ((p.N % 256) == 0 && m_tiles * splits * (p.N / 256) >= sm_count() / 2) v = 2;      // (partial 256-wide tiles never pay off)
    else v = (p.K >= 1024) ? 1 : 0;
    static int pair = -1;
    if (pair < 0) {
      const char* e = getenv("LAVT_GEMM_PAIR");
      pair = e ? atoi(e) : 1;
    }
    // CTA pairs (measured, tools/bench_gemm.py B = 8 shapes): qkv 942 -> 1004, fc1 932 -> 990, fc2 1129 -> 1207, 8192^3 1237 -> 1362 TFLOP/s;
    // the implicit-GEMM convs do not gain (their A tiles are re-fetched tap by tap from L2, 1.20-1.37 PFLOP/s either way) and stay single-CTA
    if (v == 2 && pair && !p.mnmajor && p.ksplit <= 1 && (p.rowmap == ROWMAP_IDENTITY || p.rowmap == ROWMAP_WINDOW || pair == 2) &&
        p.rowmap != ROWMAP_WGCONV && (m_tiles + 1) / 2 * (p.N / 256) >= sm_count() / 4)
      v = 4;
  }
  if (v == 4 && ((p.N % 256) != 0 || p.mnmajor || p.ksplit > 1 || p.rowmap == ROWMAP_WGCONV)) v = 2;
  if (v == 2 && (p.N % 256) != 0) v = 0;
  // 6: four epilogue warpgroups with the plain bf16-output epilogue for the short-K, output-bound launches of stages 0 / 1 (measured:
  //    qkv M 614656 x N 384 x K 128: 185 -> 138 us, fc1 + GELU M 589824 x N 512 x K 128: 294 -> 236 us; at K = 512 the 128-wide tiles lose:
  //    fc1 104 -> 115 us).  A TMA-store epilogue (64-column halves staged in 128-byte-swizzled shared memory, parity-green) did NOT help
  //    these launches (185 -> 185 us): their L1 store wavefronts (ncu: LSU data pipe 68 %) were never what the epilogue warps waited for.
  static int simple_kmax = -1;
  if (simple_kmax < 0) {
    const char* e = getenv("LAVT_GEMM_SIMPLE_KMAX");
    simple_kmax = e ? atoi(e) : 256;
  }
  if (forced < 0 && p.rowmap == ROWMAP_IDENTITY && !p.mnmajor && p.ksplit <= 1 && p.K <= simple_kmax && p.out_bf16 && !p.out_f32 && !p.out_pre &&
      !p.mul && !p.resid && !p.rscale && p.act <= ACT_RELU && p.M >= 32768)
    v = 6;
  return v;
}

int gemm_dispatch(const void* A, long long lda, const void* Bw, long long ldb, const GemmParams& p,
                  cudaStream_t stream) {
  LAVT_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  LAVT_REQUIRE(p.N % 32 == 0, "gemm: N=%d must be a multiple of 32", p.N);
  LAVT_REQUIRE(p.mnmajor || p.K % 8 == 0, "gemm: K=%d must be a multiple of 8", p.K);
  LAVT_REQUIRE(!p.mnmajor || (p.M % 8 == 0 && (p.rowmap == ROWMAP_IDENTITY || p.rowmap == ROWMAP_WGCONV)),
               "gemm (MN-major operands): M=%d must be a multiple of 8", p.M);
  LAVT_REQUIRE(p.ldo % 16 == 0 && p.ldo >= p.N, "gemm: ldo=%d invalid for N=%d (need a multiple of 16)", p.ldo, p.N);
  LAVT_REQUIRE((reinterpret_cast<uintptr_t>(p.out_f32) | reinterpret_cast<uintptr_t>(p.out_bf16) |
                reinterpret_cast<uintptr_t>(p.resid) | reinterpret_cast<uintptr_t>(p.mul)) % 32 == 0,
               "gemm: epilogue tensors must be 32-byte aligned");
  LAVT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
  LAVT_REQUIRE(p.out_f32 || p.out_bf16, "gemm: no output buffer");
  LAVT_REQUIRE(!p.mul || p.ldm % 16 == 0, "gemm: ldm must be a multiple of 16");

  CUtensorMap tmA, tmB;
  int m_tiles;
  if (p.rowmap == ROWMAP_CONV) {
    LAVT_REQUIRE(p.cCin % 8 == 0, "conv: Cin=%d must be a multiple of 8", p.cCin);
    LAVT_REQUIRE(p.cTH * p.cTW == GEMM_BM, "conv: tile %dx%d != 128 pixels", p.cTH, p.cTW);
    LAVT_REQUIRE(p.K == p.taps * p.cCin, "conv: K mismatch");
    const int n_img = p.M / (p.cH * p.cW);             // 2-D: images; 3-D: clips * frames
    int rc;
    if (p.taps == 27) {
      LAVT_REQUIRE(p.cD >= 1 && n_img % p.cD == 0, "conv3d: %d frames do not split into clips of %d", n_img, p.cD);
      uint64_t dims[5] = {(uint64_t)p.cCin, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)p.cD, (uint64_t)(n_img / p.cD)};
      uint64_t strides[4] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * p.cW, (uint64_t)lda * 2 * p.cW * p.cH,
                             (uint64_t)lda * 2 * p.cW * p.cH * p.cD};
      uint32_t box[5] = {GEMM_BK, (uint32_t)p.cTW, (uint32_t)p.cTH, 1, 1};
      rc = make_tmap_bf16(&tmA, A, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      uint64_t dims[4] = {(uint64_t)p.cCin, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)n_img};
      uint64_t strides[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * p.cW, (uint64_t)lda * 2 * p.cW * p.cH};
      uint32_t box[4] = {GEMM_BK, (uint32_t)p.cTW, (uint32_t)p.cTH, 1};
      rc = make_tmap_bf16(&tmA, A, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc) return rc;
    m_tiles = n_img * p.cTilesH * p.cTilesW;
  } else if (p.rowmap == ROWMAP_WGCONV) {
    LAVT_REQUIRE(p.mnmajor == 1 && p.cCin % 64 == 0 && p.cTH * p.cTW == 64 && (p.taps == 9 || p.taps == 27) && p.N == p.taps * p.cCin,
                 "conv wgrad: bad configuration");
    const int n_img = p.K / (p.cTilesH * p.cTilesW * 64);          // 2-D: images; 3-D: clips * frames
    uint32_t box[5] = {64, (uint32_t)p.cTW, (uint32_t)p.cTH, 1, 1};
    int rc;
    if (p.taps == 27) {
      LAVT_REQUIRE(p.cD >= 1 && n_img % p.cD == 0, "conv3d wgrad: %d frames do not split into clips of %d", n_img, p.cD);
      uint64_t dimsA[5] = {(uint64_t)p.M, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)p.cD, (uint64_t)(n_img / p.cD)};
      uint64_t sA[4] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * p.cW, (uint64_t)lda * 2 * p.cW * p.cH, (uint64_t)lda * 2 * p.cW * p.cH * p.cD};
      uint64_t dimsB[5] = {(uint64_t)p.cCin, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)p.cD, (uint64_t)(n_img / p.cD)};
      uint64_t sB[4] = {(uint64_t)ldb * 2, (uint64_t)ldb * 2 * p.cW, (uint64_t)ldb * 2 * p.cW * p.cH, (uint64_t)ldb * 2 * p.cW * p.cH * p.cD};
      rc = make_tmap_bf16(&tmA, A, 5, dimsA, sA, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tmap_bf16(&tmB, Bw, 5, dimsB, sB, box, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      uint64_t dimsA[4] = {(uint64_t)p.M, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)n_img};
      uint64_t sA[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * p.cW, (uint64_t)lda * 2 * p.cW * p.cH};
      uint64_t dimsB[4] = {(uint64_t)p.cCin, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)n_img};
      uint64_t sB[3] = {(uint64_t)ldb * 2, (uint64_t)ldb * 2 * p.cW, (uint64_t)ldb * 2 * p.cW * p.cH};
      rc = make_tmap_bf16(&tmA, A, 4, dimsA, sA, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tmap_bf16(&tmB, Bw, 4, dimsB, sB, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc) return rc;
    m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  } else if (p.mnmajor) {
    // A stored [K, M] row-major, B stored [K, N] row-major: 64 x 64 boxes (inner = 64 MN elements = one 128-byte swizzle row)
    uint64_t dimsA[2] = {(uint64_t)p.M, (uint64_t)p.K}, dimsB[2] = {(uint64_t)p.N, (uint64_t)p.K};
    uint64_t sA[1] = {(uint64_t)lda * 2}, sB[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {64, GEMM_BK};
    int rc = make_tmap_bf16(&tmA, A, 2, dimsA, sA, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB, Bw, 2, dimsB, sB, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  } else {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  }
  if (!p.mnmajor) {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t strides[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {GEMM_BK, gemm_variant(p) == 2 ? 256u : 128u};      // (CTA pairs: each CTA loads 128 of the tile's 256 B rows)
    int rc = make_tmap_bf16(&tmB, Bw, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int variant = gemm_variant(p);
  if (variant == 6) return launch_gemm<128, 6, 4, 0, 1>(tmA, tmB, p, m_tiles, stream);
  if (variant == 1) return launch_gemm<128, 3, 1>(tmA, tmB, p, m_tiles, stream);
  if (variant == 2) return launch_gemm<256, 4, 2>(tmA, tmB, p, m_tiles, stream);
  if (variant == 4) return launch_gemm<256, 6, 2, 1>(tmA, tmB, p, m_tiles, stream);
  return launch_gemm<128, 6, 2>(tmA, tmB, p, m_tiles, stream);
}

}  // namespace lavt
