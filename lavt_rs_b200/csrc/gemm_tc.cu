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
// Structure (one 128 x BLOCK_N output tile per CTA, 256 threads):
//   warp 0   : TMA producer  (one elected lane)   smem ring of STAGES x {A 128x64, B BLOCK_Nx64} bf16,
//                                                  128-byte swizzle, mbarrier full/empty pairs
//   warp 1   : MMA issuer    (one elected lane)   tcgen05.mma.cta_group::1.kind::f16, D in TMEM
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue       tcgen05.ld 32x32b -> registers -> fused math -> vectorised global stores
#include "common.cuh"
#include "gemm_tc.cuh"

namespace lavt {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 256;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // after the ring: barriers, tmem pointer, epilogue vectors
  static constexpr int BAR_OFF = RING_BYTES;
  static constexpr int VEC_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = VEC_OFF + 2 * BN * 4 + 1024 /*alignment slack*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_scale = reinterpret_cast<float*>(smem + L::VEC_OFF);
  float* s_bias = s_scale + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;

  // conv tile decomposition
  int img = 0, h0 = 0, w0 = 0;
  if (p.rowmap == ROWMAP_CONV) {
    int tw = mt % p.cTilesW;
    int th = (mt / p.cTilesW) % p.cTilesH;
    img = mt / (p.cTilesW * p.cTilesH);
    h0 = th * p.cTH;
    w0 = tw * p.cTW;
  }
  const int num_kb = p.K / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, BN);   // BN fp32 accumulator columns (power of two >= 32)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int cpb = (p.rowmap == ROWMAP_CONV) ? (p.cCin / GEMM_BK) : 1;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if (p.rowmap == ROWMAP_CONV) {
          const int tap = kb / cpb, cc = kb - tap * cpb;
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          tma_load_4d(sa, &tmA, &full_bar[s], cc * GEMM_BK, w0 + dx, h0 + dy, img);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], kb * GEMM_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
        const uint64_t da = make_kmajor_sw128_desc(sa);
        const uint64_t db = make_kmajor_sw128_desc(sb);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4) units
          umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);            // accumulator complete
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;                 // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int et = threadIdx.x - 128;
    for (int i = et; i < BN; i += 128) {
      s_scale[i] = p.cscale ? p.cscale[n0 + i] : 1.0f;
      s_bias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only

    const int r = ew * 32 + lane;            // row inside the tile
    long long m = -1, orow = -1;
    if (p.rowmap == ROWMAP_CONV) {
      const int h = h0 + r / p.cTW, w = w0 + r % p.cTW;
      if (h < p.cH && w < p.cW) {
        m = (static_cast<long long>(img) * p.cH + h) * p.cW + w;
        orow = m;
      }
    } else {
      const long long mm = static_cast<long long>(mt) * GEMM_BM + r;
      if (mm < p.M) {
        m = mm;
        orow = (p.rowmap == ROWMAP_WINDOW) ? win_token(p.win, mm).row : mm;
      }
    }
    const bool live = (orow >= 0);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const __nv_bfloat16* mul_row = (p.mul && live) ? p.mul + m * p.ldm + n0 : nullptr;
    const float* res_row = (p.resid && live) ? p.resid + orow * p.ldo + n0 : nullptr;
    float* of_row = (p.out_f32 && live) ? p.out_f32 + orow * p.ldo + n0 : nullptr;
    __nv_bfloat16* ob_row = (p.out_bf16 && live) ? p.out_bf16 + orow * p.ldo + n0 : nullptr;

#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + c * 32;
      tmem_ld_32x32b_x32(taddr, v);          // warp-collective: executed by all lanes
      tmem_ld_wait();
      if (!live) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]) * s_scale[c * 32 + j] + s_bias[c * 32 + j];
        if (p.act == ACT_GELU) x = gelu_erf(x);
        else if (p.act == ACT_RELU) x = fmaxf(x, 0.0f);
        else if (p.act == ACT_TANH) x = tanhf(x);
        f[j] = x;
      }
      if (mul_row) {
        const uint4* mp = reinterpret_cast<const uint4*>(mul_row + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u = __ldg(mp + q);
          float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc2 = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          f[q * 8 + 0] *= a.x; f[q * 8 + 1] *= a.y; f[q * 8 + 2] *= b.x; f[q * 8 + 3] *= b.y;
          f[q * 8 + 4] *= cc2.x; f[q * 8 + 5] *= cc2.y; f[q * 8 + 6] *= d.x; f[q * 8 + 7] *= d.y;
        }
      }
      if (res_row) {
        const float4* rp = reinterpret_cast<const float4*>(res_row + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 u = __ldg(rp + q);
          f[q * 4 + 0] += u.x; f[q * 4 + 1] += u.y; f[q * 4 + 2] += u.z; f[q * 4 + 3] += u.w;
        }
      }
      if (of_row) {
        float4* op = reinterpret_cast<float4*>(of_row + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
      }
      if (ob_row) {
        uint4* op = reinterpret_cast<uint4*>(ob_row + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          op[q] = make_uint4(pack_bf16x2(f[q * 8], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                             pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int m_tiles,
                       cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  static bool configured = false;
  auto kfn = gemm_bf16_tc_kernel<BN, STAGES>;
  if (!configured) {
    LAVT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid(p.N / BN, m_tiles, 1);
  kfn<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, p);
  LAVT_LAUNCH_CHECK("gemm_bf16_tc_kernel");
  return LAVT_OK;
}

int gemm_dispatch(const void* A, long long lda, const void* Bw, long long ldb, const GemmParams& p,
                  cudaStream_t stream) {
  LAVT_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  LAVT_REQUIRE(p.N % 128 == 0, "gemm: N=%d must be a multiple of 128", p.N);
  LAVT_REQUIRE(p.K % GEMM_BK == 0, "gemm: K=%d must be a multiple of 64", p.K);
  LAVT_REQUIRE(p.ldo % 8 == 0 && p.ldo >= p.N, "gemm: ldo=%d invalid for N=%d", p.ldo, p.N);
  LAVT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
  LAVT_REQUIRE(p.out_f32 || p.out_bf16, "gemm: no output buffer");
  LAVT_REQUIRE(!p.mul || p.ldm % 8 == 0, "gemm: ldm must be a multiple of 8");

  CUtensorMap tmA, tmB;
  int m_tiles;
  if (p.rowmap == ROWMAP_CONV) {
    LAVT_REQUIRE(p.cCin % GEMM_BK == 0, "conv: Cin=%d must be a multiple of 64", p.cCin);
    LAVT_REQUIRE(p.cTH * p.cTW == GEMM_BM, "conv: tile %dx%d != 128 pixels", p.cTH, p.cTW);
    LAVT_REQUIRE(p.K == p.taps * p.cCin, "conv: K mismatch");
    const int n_img = p.M / (p.cH * p.cW);
    uint64_t dims[4] = {(uint64_t)p.cCin, (uint64_t)p.cW, (uint64_t)p.cH, (uint64_t)n_img};
    uint64_t strides[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * p.cW, (uint64_t)lda * 2 * p.cW * p.cH};
    uint32_t box[4] = {GEMM_BK, (uint32_t)p.cTW, (uint32_t)p.cTH, 1};
    int rc = make_tmap_bf16(&tmA, A, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    m_tiles = n_img * p.cTilesH * p.cTilesW;
  } else {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t strides[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {GEMM_BK, 128};
    int rc = make_tmap_bf16(&tmB, Bw, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  LAVT_REQUIRE(m_tiles <= 65535 * 16, "gemm: too many M tiles (%d)", m_tiles);
  if (p.K <= 2 * GEMM_BK) return launch_gemm<128, 2>(tmA, tmB, p, m_tiles, stream);
  return launch_gemm<128, 3>(tmA, tmB, p, m_tiles, stream);
}

}  // namespace lavt
