// tcgen05 / TMEM shifted-window attention core for windows of up to 400 tokens
// (reference WindowAttention3D.forward, lib/video_swin_transformer.py:147-165; 2-D twin lib/backbone.py:127-138):
//     S = q k^T + relative-position bias (+ shifted-window mask)  ->  softmax  ->  O = P v       per (window, head)
//
// Persistent, one CTA per SM, 544 threads:
//   warps 0-15 : softmax warps.  Warp w owns TMEM lanes [32*(w%4), +32) (= 32 query rows of the 128-row tile) and the
//                key-column range of warpgroup w/4, so a 128 x NP score tile is processed by all 16 warps at once:
//                  pass 1  tcgen05.ld S -> + bias (gathered from the per-head table in smem through the closed form
//                          idx = code(i) - code(j) + const) (+ mask) -> running max -> tcgen05.st back
//                  (row max exchanged between the four column owners of a row through smem + a 128-thread barrier)
//                  pass 2  tcgen05.ld -> exp2(s - max) -> row sum -> bf16 pack -> tcgen05.st P over the S columns
//                one warpgroup per tile then runs the epilogue (tcgen05.ld O, 1/sum, bf16, 64-byte row stores)
//   warp 16    : one elected thread issues the TMA loads (Q, K, V of one (window, head) = three [N x 32] bf16 boxes,
//                64-byte swizzle, double-buffered across units) and all tcgen05.mma:
//                  S[128 x NP]  = Q_tile (smem, K-major) x K^T (smem, K-major)        2 k-steps of 16
//                  O[128 x 32] += P (TMEM, bf16 pairs)   x V   (smem, MN-major)       NP/16 k-steps
// q arrives pre-scaled by head_dim^-0.5 * log2(e) (qkv GEMM epilogue), the table is pre-multiplied by log2(e), so the
// softmax is one FADD + one MUFU.EX2 per score.  Neither the (N,N) index buffer nor the (nW,N,N) mask exists on the device.
#include "kernels.cuh"

#include <cstdlib>

namespace lavt {

constexpr int TC_HD = 32;
constexpr int TC_SOFTMAX_THREADS = 512;
constexpr int TC_THREADS = TC_SOFTMAX_THREADS + 32;
constexpr int TC_MAX_NP = 400;            // S columns (fp32) in TMEM; O accumulators live at columns 448 / 480
constexpr int TC_O_COL = 448;
constexpr int TC_MSTRIDE = 401;           // mask-table row stride (floats): distinct banks for distinct classes
constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr float TC_MASKV = -100.0f * TC_LOG2E;

struct AttnTcArgs {
  int N, NP, ntiles;        // tokens per window, padded to 16, 128-row query tiles
  int nwin, units;          // windows in the launch, units = nwin * heads
  int BR, nb;               // TMA box rows, boxes per operand
  int nA, nB;               // S column chunks of the QK^T MMA (nB == 0 -> single chunk)
  int stage_bytes;          // Q | K | V, each NP x 64 B
  int off_tab, off_mtab, off_codes, off_cls, off_negoff, off_pm, off_ps, off_bar;
  int rc;                   // rel_const
  int shifted;
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers local to this kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int W>
__device__ __forceinline__ void tmem_ld_w(uint32_t taddr, uint32_t* r) {
  if constexpr (W == 32) tmem_ld_x32(taddr, r); else tmem_ld_x16(taddr, r);
}
template <int W>
__device__ __forceinline__ void tmem_st_w(uint32_t taddr, const uint32_t* r) {
  if constexpr (W == 32) tmem_st_x32(taddr, r); else if constexpr (W == 16) tmem_st_x16(taddr, r); else tmem_st_x8(taddr, r);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 add (sm_100): one issue slot for two lanes of work
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory matrix descriptors, 64-byte swizzle (rows of 32 bf16 = 64 B, 8-row groups of 512 B).
//   K-major  (Q as A, K as B):  SBO = 512 B between 8-row groups, LBO unused
//   MN-major (V as B, N = 32 = one swizzle atom): SBO = 512 B between 8-key groups, LBO (next N atom) unused
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;      // SWIZZLE_64B
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// softmax passes over one piece of W score columns starting at column c
// ---------------------------------------------------------------------------------------------------------------
// pass 1: s += bias (+ mask); columns >= N become -inf; returns the running max; writes s back to TMEM
template <int W>
__device__ __forceinline__ void pass1_piece(uint32_t ts, int c, int N, uint32_t tq, const int* negoff, bool need_mask,
                                            uint32_t mrow, float& m0, float& m1) {
  uint32_t v[W];
  tmem_ld_w<W>(ts + c, v);
  int no[W];
#pragma unroll
  for (int j = 0; j < W; j += 4) *reinterpret_cast<int4*>(&no[j]) = *reinterpret_cast<const int4*>(negoff + c + j);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
    add2(a0, a1, lds_f32(tq + no[j]), lds_f32(tq + no[j + 1]));
    v[j] = __float_as_uint(a0);
    v[j + 1] = __float_as_uint(a1);
  }
  if (need_mask) {
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
      add2(a0, a1, lds_f32(mrow + 4 * (c + j)), lds_f32(mrow + 4 * (c + j + 1)));
      v[j] = __float_as_uint(a0);
      v[j + 1] = __float_as_uint(a1);
    }
  }
  if (c + W > N) {
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (c + j >= N) v[j] = __float_as_uint(-INFINITY);
  }
#pragma unroll
  for (int j = 0; j < W; j += 4) {
    m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
    m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
  }
  tmem_st_w<W>(ts + c, v);
}

// pass 2: p = exp2(s - m); accumulates the row sum; writes bf16 pairs to TMEM columns tp ..
template <int W>
__device__ __forceinline__ void pass2_piece(uint32_t ts, int c, uint32_t tp, float nm, float& l0, float& l1) {
  uint32_t v[W];
  tmem_ld_w<W>(ts + c, v);
  tmem_ld_wait();
  uint32_t pk[W / 2];
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
    add2(a0, a1, nm, nm);
    a0 = ex2_ftz(a0);
    a1 = ex2_ftz(a1);
    add2(l0, l1, a0, a1);
    pk[j >> 1] = pack_bf16x2(a0, a1);
  }
  tmem_st_w<W / 2>(tp, pk);
}

// column range of warpgroup g: the NP/16 sixteen-column groups are dealt out as evenly as possible
__host__ __device__ __forceinline__ void wg_range(int NP, int g, int& c0, int& c1) {
  const int n16 = NP >> 4, base = n16 >> 2, rem = n16 & 3;
  const int start = g * base + (g < rem ? g : rem);
  c0 = start << 4;
  c1 = c0 + ((base + (g < rem ? 1 : 0)) << 4);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tab = reinterpret_cast<float*>(smem + a.off_tab);
  float* mtab = reinterpret_cast<float*>(smem + a.off_mtab);
  uint16_t* codes = reinterpret_cast<uint16_t*>(smem + a.off_codes);
  uint8_t* cls = smem + a.off_cls;
  int* negoff = reinterpret_cast<int*>(smem + a.off_negoff);
  float* pm = reinterpret_cast<float*>(smem + a.off_pm);       // [4][128] partial row max
  float* ps = reinterpret_cast<float*>(smem + a.off_ps);       // [4][128] partial row sum
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* kv_full = bars;            // [2]
  uint64_t* stage_free = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;         // [2]
  uint64_t* o_free = bars + 8;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, NP = a.NP;

  // contiguous unit range of this CTA; unit u = head * nwin + window (head-major: the bias table is reloaded rarely)
  const int u_begin = static_cast<int>(1LL * a.units * blockIdx.x / gridDim.x);
  const int u_end = static_cast<int>(1LL * a.units * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&stage_free[0], 1);
    mbar_init(&stage_free[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, TC_SOFTMAX_THREADS / 32);
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_init(&o_free[0], 4);
    mbar_init(&o_free[1], 4);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_ptr_smem, 512);
  // static per-launch tables: relative-position codes of the window tokens; zero the K / V pad rows of both stages
  for (int j = threadIdx.x; j < NP; j += blockDim.x) {
    const int code = (j < N) ? win_token(p.win, j).code : 0;
    codes[j] = static_cast<uint16_t>(code);
    negoff[j] = -4 * code;
    cls[j] = 0;
  }
  for (int i = threadIdx.x; i < 2 * 2 * (NP - N) * 16; i += blockDim.x) {
    const int w = i & 15, rest = i >> 4;
    const int row = N + rest % (NP - N), which = rest / (NP - N);       // which: stage * 2 + {K, V}
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + (which >> 1) * a.stage_bytes + (1 + (which & 1)) * NP * 64 + row * 64);
    dst[w] = 0;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 16) {
    // =============================== TMA + MMA issue (one thread) ===============================
    if (lane == 0 && u_begin < u_end) {
      const uint32_t idesc_a = make_idesc_bf16_f32(128, a.nA);
      const uint32_t idesc_b = make_idesc_bf16_f32(128, a.nB > 0 ? a.nB : 16);
      const uint32_t idesc_pv = make_idesc_bf16_f32(128, TC_HD) | (1u << 16);     // B (= V) is MN-major
      const uint32_t tx_bytes = 3u * N * 64u;
      auto issue_loads = [&](int u, int s) {
        const int head = u / a.nwin, win = u - head * a.nwin;
        uint8_t* st = smem + s * a.stage_bytes;
        mbar_expect_tx(&kv_full[s], tx_bytes);
        for (int op = 0; op < 3; ++op)
          for (int b = 0; b < a.nb; ++b)
            tma_load_2d(st + op * NP * 64 + b * a.BR * 64, &tmQKV, &kv_full[s], op * p.C + head * TC_HD, win * N + b * a.BR);
      };
      issue_loads(u_begin, 0);
      int gt = 0;
      for (int u = u_begin, lu = 0; u < u_end; ++u, ++lu) {
        const int s = lu & 1;
        if (u + 1 < u_end) {
          const int n = (lu + 1) >> 1;                      // n-th fill of stage s^1
          if (n >= 1) mbar_wait(&stage_free[s ^ 1], (n - 1) & 1);
          issue_loads(u + 1, s ^ 1);
        }
        mbar_wait(&kv_full[s], (lu >> 1) & 1);
        const uint32_t sq = smem_u32(smem + s * a.stage_bytes);
        const uint32_t sk = sq + NP * 64, sv = sk + NP * 64;
        for (int qt = 0; qt < a.ntiles; ++qt, ++gt) {
          const int ob = gt & 1;
          tc_fence_after();
          // S = Q_tile K^T
          const uint64_t dq = make_sw64_desc(sq + qt * 128 * 64);
          const uint64_t dk0 = make_sw64_desc(sk);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem_base, dq + 2 * k, dk0 + 2 * k, idesc_a, k);
          if (a.nB > 0) {
            const uint64_t dk1 = make_sw64_desc(sk + a.nA * 64);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16_ss(tmem_base + a.nA, dq + 2 * k, dk1 + 2 * k, idesc_b, k);
          }
          umma_commit(s_full);
          if (gt >= 2) mbar_wait(&o_free[ob], ((gt >> 1) - 1) & 1);     // epilogue of tile gt-2 drained this accumulator
          mbar_wait(p_ready, gt & 1);
          tc_fence_after();
          // O = P V : one k-step per 16 keys; P sits packed inside the S columns of its owner warpgroup
          const uint32_t tmem_o = tmem_base + TC_O_COL + ob * TC_HD;
          int ks = 0;
          for (int g = 0; g < 4; ++g) {
            int c0, c1;
            wg_range(NP, g, c0, c1);
            for (int c = c0; c < c1; c += 16, ++ks)
              umma_bf16_ts(tmem_o, tmem_base + c0 + ((c - c0) >> 1), make_sw64_desc(sv + ks * 16 * 64), idesc_pv, ks);
          }
          umma_commit(&o_full[ob]);
        }
        umma_commit(&stage_free[s]);
      }
    }
  } else {
    // =============================== softmax / epilogue warps ===============================
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const int st = threadIdx.x;                       // 0..511
    int c0, c1;
    wg_range(NP, g, c0, c1);
    const uint32_t ts = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tab_addr = smem_u32(tab), mtab_addr = smem_u32(mtab);
    const WinGeom& wg = p.win;
    const int nW = wg.nwd * wg.nwh * wg.nww;
    int cur_head = -1;
    int gt = 0;
    for (int u = u_begin, lu = 0; u < u_end; ++u, ++lu) {
      const int head = u / a.nwin, win = u - head * a.nwin;
      // ---- per-unit tables (all softmax warps are past the previous unit's passes after this barrier) ----
      named_bar(5, TC_SOFTMAX_THREADS);
      if (head != cur_head) {
        const float* src = p.table_t + static_cast<long long>(head) * p.L;
        for (int i = st; i < p.L; i += TC_SOFTMAX_THREADS) tab[i] = __ldg(src + i) * TC_LOG2E;
        cur_head = head;
      }
      bool need_mask = false;
      if (a.shifted) {
        const int wi = win % nW;
        const int wc = wi % wg.nww, wb = (wi / wg.nww) % wg.nwh, wa = wi / (wg.nww * wg.nwh);
        need_mask = (wg.sd && wa == wg.nwd - 1) || (wg.sh && wb == wg.nwh - 1) || (wg.sw && wc == wg.nww - 1);
        if (need_mask) {
          // class = per-axis (region - region of the window's first token): at most two regions per axis in a window
          const int Dp = wg.nwd * wg.wd, Hp = wg.nwh * wg.wh, Wp = wg.nww * wg.ww;
          const int rd0 = shift_region(wa * wg.wd, Dp, wg.wd, wg.sd);
          const int rh0 = shift_region(wb * wg.wh, Hp, wg.wh, wg.sh);
          const int rw0 = shift_region(wc * wg.ww, Wp, wg.ww, wg.sw);
          for (int j = st; j < NP; j += TC_SOFTMAX_THREADS) {
            int cj = 0;
            if (j < N) {
              const int tw = j % wg.ww, th = (j / wg.ww) % wg.wh, td = j / (wg.ww * wg.wh);
              cj = 4 * (shift_region(wa * wg.wd + td, Dp, wg.wd, wg.sd) - rd0) +
                   2 * (shift_region(wb * wg.wh + th, Hp, wg.wh, wg.sh) - rh0) +
                   (shift_region(wc * wg.ww + tw, Wp, wg.ww, wg.sw) - rw0);
            }
            cls[j] = static_cast<uint8_t>(cj);
#pragma unroll
            for (int k = 0; k < 8; ++k) mtab[k * TC_MSTRIDE + j] = (k != cj) ? TC_MASKV : 0.0f;
          }
        }
      }
      named_bar(5, TC_SOFTMAX_THREADS);

      for (int qt = 0; qt < a.ntiles; ++qt, ++gt) {
        const int i = qt * 128 + r;
        const bool wvalid = (qt * 128 + q * 32) < N;          // warp-uniform: any live query row in this warp?
        mbar_wait(s_full, gt & 1);
        tc_fence_after();
        uint32_t tq = 0, mrow = 0;
        if (wvalid) {
          const int ic = i < N ? i : N - 1;
          tq = tab_addr + 4 * (static_cast<int>(codes[ic]) + a.rc);
          mrow = mtab_addr + 4 * TC_MSTRIDE * (need_mask ? cls[ic] : 0);
          float m0 = -INFINITY, m1 = -INFINITY;
          int c = c0;
          for (; c + 32 <= c1; c += 32) pass1_piece<32>(ts, c, N, tq, negoff, need_mask, mrow, m0, m1);
          if (c < c1) pass1_piece<16>(ts, c, N, tq, negoff, need_mask, mrow, m0, m1);
          pm[g * 128 + r] = fmaxf(m0, m1);
          tmem_st_wait();
        }
        named_bar(1 + q, 128);                                 // the four column owners of these 32 rows
        if (wvalid) {
          const float nm = -fmaxf(fmaxf(pm[r], pm[128 + r]), fmaxf(pm[256 + r], pm[384 + r]));
          float l0 = 0.f, l1 = 0.f;
          int c = c0;
          for (; c + 32 <= c1; c += 32) pass2_piece<32>(ts, c, ts + c0 + ((c - c0) >> 1), nm, l0, l1);
          if (c < c1) pass2_piece<16>(ts, c, ts + c0 + ((c - c0) >> 1), nm, l0, l1);
          ps[g * 128 + r] = l0 + l1;
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);

        if (g == (gt & 3)) {
          // ---- epilogue of this tile: O / l -> bf16 rows ----
          const int ob = gt & 1;
          mbar_wait(&o_full[ob], (gt >> 1) & 1);
          tc_fence_after();
          if (wvalid) {
            uint32_t o[32];
            tmem_ld_x32(ts + TC_O_COL + ob * TC_HD, o);
            tmem_ld_wait();
            const float l = (ps[r] + ps[128 + r]) + (ps[256 + r] + ps[384 + r]);
            const float inv = 1.0f / l;
            if (i < N) {
              __nv_bfloat16* dst = p.out + (static_cast<long long>(win) * N + i) * p.C + head * TC_HD;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t w8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  w8[j] = pack_bf16x2(__uint_as_float(o[h * 16 + 2 * j]) * inv, __uint_as_float(o[h * 16 + 2 * j + 1]) * inv);
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + h * 16), "r"(w8[0]),
                             "r"(w8[1]), "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                             : "memory");
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[ob]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool window_attn_tc_supported(const AttnParams& p) {
  const WinGeom& g = p.win;
  if (g.N < 16 || g.N > TC_MAX_NP) return false;
  const int nb = (g.N + 255) / 256;
  if (g.N % nb != 0) return false;
  if (p.L > 8192) return false;
  return true;
}

int window_attn_tc_dispatch(const AttnParams& p, cudaStream_t st) {
  const WinGeom& g = p.win;
  LAVT_REQUIRE(window_attn_tc_supported(p), "attention(tc): unsupported window (N=%d, L=%d)", g.N, p.L);
  const long long nwin = 1LL * g.B * g.nwd * g.nwh * g.nww;
  LAVT_REQUIRE(nwin * p.nH < (1LL << 30), "attention(tc): too many units");
  AttnTcArgs a;
  a.N = g.N;
  a.NP = (g.N + 15) & ~15;
  a.ntiles = (g.N + 127) / 128;
  a.nwin = static_cast<int>(nwin);
  a.units = static_cast<int>(nwin * p.nH);
  a.nb = (g.N + 255) / 256;
  a.BR = g.N / a.nb;
  if (a.NP <= 256) {
    a.nA = a.NP;
    a.nB = 0;
  } else {
    a.nA = ((a.NP / 2) + 15) & ~15;
    a.nB = a.NP - a.nA;
  }
  a.stage_bytes = 3 * a.NP * 64;
  int off = 2 * a.stage_bytes;
  // the last query tile reads up to 128 rows past N from the Q region: keep that inside the stage (it runs into K / V)
  a.off_tab = off;        off += ((p.L * 4 + 127) / 128) * 128;
  a.off_mtab = off;       off += ((8 * TC_MSTRIDE * 4 + 127) / 128) * 128;
  a.off_codes = off;      off += ((a.NP * 2 + 127) / 128) * 128;
  a.off_cls = off;        off += ((a.NP + 127) / 128) * 128;
  a.off_negoff = off;     off += ((a.NP * 4 + 127) / 128) * 128;
  a.off_pm = off;         off += 4 * 128 * 4;
  a.off_ps = off;         off += 4 * 128 * 4;
  a.off_bar = off;        off += 128;
  const int smem = off + 1024;
  LAVT_REQUIRE(smem <= 227 * 1024, "attention(tc): shared memory %d B exceeds the SM", smem);
  LAVT_REQUIRE(a.ntiles * 128 * 64 <= a.stage_bytes + a.stage_bytes, "attention(tc): query tile overrun");
  a.rc = rel_const(g);
  a.shifted = (g.sd | g.sh | g.sw) != 0;

  CUtensorMap tm;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3 * p.C), static_cast<uint64_t>(nwin * g.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(3 * p.C) * 2};
    uint32_t box[2] = {TC_HD, static_cast<uint32_t>(a.BR)};
    int rc = make_tmap_bf16(&tm, p.qkv, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  static int configured = 0;
  if (smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  const int grid = a.units < sms ? a.units : sms;
  window_attn_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tm, p, a);
  LAVT_LAUNCH_CHECK("window_attn_tc_kernel");
  return LAVT_OK;
}

}  // namespace lavt
