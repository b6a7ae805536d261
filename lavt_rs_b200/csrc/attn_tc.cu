// tcgen05 / TMEM shifted-window attention core for windows of up to 400 tokens
// (reference WindowAttention3D.forward, lib/video_swin_transformer.py:147-165; 2-D twin lib/backbone.py:127-138):
//     S = q k^T + relative-position bias (+ shifted-window mask)  ->  softmax  ->  O = P v       per (window, head)
//
// Persistent, one CTA per SM, 16 warps.  A unit of work is one (window, head); its 128-row query tiles are processed
// one after the other, and the KEY axis of a tile is split into three column groups that run as independent pipelines:
//   warps 0-11 : softmax warps.  Warp w owns TMEM lanes [32*(w%4), +32) (32 query rows) and column group w/4.
//                  pass 1  tcgen05.ld S -> + bias (per-head table in smem, closed-form index) (+ mask) -> max -> tcgen05.st
//                  pass 2  tcgen05.ld -> exp2(s - group max) -> row sum -> bf16 pack -> tcgen05.st P over the S columns
//                each group keeps its OWN row max / row sum / O accumulator (flash-style partials), so the groups never
//                synchronise with each other: while one group is in the shared-memory bound pass 1 another is in the
//                MUFU bound pass 2 and a third waits for its MMAs.  The last group owns fewer columns and also runs the
//                epilogue of the previous tile:  O = sum_g 2^(m_g - m) O_g / sum_g 2^(m_g - m) l_g -> bf16 rows.
//   warp 12    : one thread issues the TMA loads (Q, K, V of a unit = three [N x 32] bf16 boxes, 64-byte swizzle,
//                double-buffered across units)
//   warps 13-15: one thread each issues the tcgen05.mma of one column group:
//                  S_g[128 x n_g] = Q_tile (smem, K-major) x K_g^T (smem, K-major)          2 k-steps of 16
//                  O_g[128 x 32]  = P_g (TMEM, bf16 pairs) x V_g (smem, MN-major)           n_g/16 k-steps
// q arrives pre-scaled by head_dim^-0.5 * log2(e) (qkv GEMM epilogue), the table is pre-multiplied by log2(e), so the
// softmax is one FADD + one MUFU.EX2 per score.  Neither the (N,N) index buffer nor the (nW,N,N) mask exists on the device.
//
// Bias table layout in shared memory: idx(i,j) = code(i) - code(j) + const with code(t) = cd*SD + ch*SH + cw and strides
// chosen so that code(t) == t (mod 32): the 32 lanes of a warp (32 consecutive tokens) then always hit 32 distinct banks.
#include "kernels.cuh"
#include "attn_tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace lavt {

constexpr int TC_HD = 32;
constexpr int TC_GROUPS = 3;
constexpr int TC_SOFTMAX_WARPS = 4 * TC_GROUPS;
constexpr int TC_SOFTMAX_THREADS = 32 * TC_SOFTMAX_WARPS;
constexpr int TC_TMA_WARP = TC_SOFTMAX_WARPS;          // warp 12: TMA producer
constexpr int TC_MMA_WARP0 = TC_SOFTMAX_WARPS + 1;     // warps 13-15: MMA issuer of group 0 / 1 / 2
constexpr int TC_THREADS = 32 * (TC_MMA_WARP0 + TC_GROUPS);
constexpr int TC_EPI_GROUP = TC_GROUPS - 1;             // the group that runs the epilogue (last in phase order)
constexpr int TC_MAX_NP = 400;            // S columns (fp32) in TMEM; the three O accumulators live at columns 416..511
constexpr int TC_O_COL = 416;
constexpr int TC_MSTRIDE = 401;           // mask-table row stride (floats): distinct banks for distinct classes
constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr float TC_MASKV = -100.0f * TC_LOG2E;
constexpr int TC_TRACE_TILES = 64;

// smallest s >= lo with s == r (mod 32)
__host__ __device__ constexpr int tc_stride(int lo, int r) { return lo + ((r - lo) % 32 + 32) % 32; }
// column range of group g over the NP/16 sixteen-column blocks; the LAST group (which also runs the epilogue) gets
// about 20 % fewer columns than the other two
__host__ __device__ constexpr int tc_group_begin(int NP, int g) {
  const int n16 = NP / 16;
  int nl = (n16 * 4 + 7) / 14;                    // ~ 0.29 * n16
  if (nl < 1) nl = 1;
  const int n0 = (n16 - nl + 1) / 2;
  return 16 * (g == 0 ? 0 : g == 1 ? n0 : g == 2 ? n16 - nl : n16);
}
__host__ __device__ constexpr int tc_group_end(int NP, int g) { return tc_group_begin(NP, g + 1); }

struct AttnTcArgs {
  int N, NP, ntiles;        // tokens per window, padded to 16, 128-row query tiles
  int nwin, units;          // windows in the launch, units = nwin * heads
  int BR, nb;               // TMA box rows, boxes per operand
  int SH, SD, L2, rc;       // expanded bias-table strides, size (floats) and rel_const in that layout
  int stage_bytes;          // Q | K | V, each NP x 64 B
  int off_tab, off_mtab, off_negoff, off_pm, off_ps, off_bar;
  int shifted;
  int r4, tail_rows;        // tail tile replicated x4 across the TMEM lane quadrants (see softmax warps)
  int off_xq;               // [4][16][34] cross-quadrant partials of the replicated tail tile
  int stagger;              // phase offset (clocks) between consecutive group pipelines
  long long* trace;        // debug (LAVT_ATTN_TRACE): clock64 stamps of CTA 0, [17 events][TC_TRACE_TILES]
};

// ---------------------------------------------------------------------------------------------------------------
// softmax passes over one piece of W score columns starting at column c.
//   tabq = table + code(i) + rc (this thread's row): bias(i, j) = tabq[negoff[j]],  negoff[j] = -code(j)
//   mrow = mask-table row of this thread's region class
// ---------------------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void pass1_piece(uint32_t ts, int c, int N, const float* tabq, const int* negoff, bool need_mask,
                                            const float* mrow, float& m0, float& m1) {
  uint32_t v[W];
  tmem_ld_w<W>(ts + c, v);
  // plain shared-memory loads: the compiler batches the index loads ahead of the dependent gathers
  int no[W];
#pragma unroll
  for (int j = 0; j < W; j += 4) *reinterpret_cast<int4*>(&no[j]) = *reinterpret_cast<const int4*>(negoff + c + j);
  float b[W];
#pragma unroll
  for (int j = 0; j < W; ++j) b[j] = tabq[no[j]];
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
    add2(a0, a1, b[j], b[j + 1]);
    v[j] = __float_as_uint(a0);
    v[j + 1] = __float_as_uint(a1);
  }
  if (need_mask) {
    const float* mp = mrow + c;
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
      add2(a0, a1, mp[j], mp[j + 1]);
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
__device__ __forceinline__ void pass2_piece(uint32_t ts_c, uint32_t tp, float nm, float& l0, float& l1) {
  uint32_t v[W];
  tmem_ld_w<W>(ts_c, v);
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

__global__ void __launch_bounds__(TC_THREADS, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQT, const AttnParams p,
                      const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment as an OFFSET (keeps the shared address space visible to the compiler: LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* tab = reinterpret_cast<float*>(smem + a.off_tab);
  float* mtab = reinterpret_cast<float*>(smem + a.off_mtab);
  int* negoff = reinterpret_cast<int*>(smem + a.off_negoff);
  float* pm = reinterpret_cast<float*>(smem + a.off_pm);       // [2][3][128] group row max (tile parity)
  float* ps = reinterpret_cast<float*>(smem + a.off_ps);       // [2][3][128] group row sum
  float* xq = reinterpret_cast<float*>(smem + a.off_xq);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* kv_full = bars;            // [2]
  uint64_t* stage_free = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;         // [3]
  uint64_t* p_ready = bars + 7;        // [3]
  uint64_t* o_full = bars + 10;
  uint64_t* o_free = bars + 11;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, NP = a.NP, ntiles = a.ntiles;
  long long* const trace = (blockIdx.x == 0) ? a.trace : nullptr;
#define TC_TRACE(ev, t) do { if (trace && (t) < TC_TRACE_TILES) trace[(ev) * TC_TRACE_TILES + (t)] = clock64(); } while (0)
  const int SH = a.SH, SD = a.SD;
  const WinGeom& wg = p.win;

  // contiguous unit range of this CTA; unit u = head * nwin + window (head-major: the bias table is reloaded rarely)
  const int u_begin = static_cast<int>(1LL * a.units * blockIdx.x / gridDim.x);
  const int u_end = static_cast<int>(1LL * a.units * (blockIdx.x + 1) / gridDim.x);
  const int nunits = u_end - u_begin;
  const int T = nunits * ntiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&stage_free[0], TC_GROUPS);
    mbar_init(&stage_free[1], TC_GROUPS);
    for (int g = 0; g < TC_GROUPS; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], 4);
    }
    mbar_init(o_full, TC_GROUPS);
    mbar_init(o_free, 4);
    fence_mbar_init();
  }
  if (warp == TC_TMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  // per-launch table: -code(j) of the key tokens; zero the K / V pad rows of both stages
  for (int j = threadIdx.x; j < NP; j += blockDim.x) {
    int code = 0;
    if (j < N) code = (j / (wg.Wh * wg.Ww)) * SD + ((j / wg.Ww) % wg.Wh) * SH + j % wg.Ww;
    negoff[j] = -code;
  }
  const int npad = (NP - N) > 0 ? (NP - N) : 1;
  for (int i = threadIdx.x; i < 2 * 2 * (NP - N) * 16; i += blockDim.x) {
    const int w = i & 15, rest = i >> 4;
    const int row = N + rest % npad, which = rest / npad;               // which: stage * 2 + {K, V}
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + (which >> 1) * a.stage_bytes + (1 + (which & 1)) * NP * 64 + row * 64);
    dst[w] = 0;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == TC_TMA_WARP) {
    // =============================== TMA producer (one thread) ===============================
    if (lane == 0) {
      const uint32_t tx_bytes = 3u * N * 64u + (a.r4 ? 4u * 32u * 64u : 0u);
      for (int lu = 0; lu < nunits; ++lu) {
        const int u = u_begin + lu, s = lu & 1;
        if (lu >= 2) mbar_wait(&stage_free[s], ((lu - 2) >> 1) & 1);     // all MMAs reading unit lu-2 have retired
        const int head = u / a.nwin, win = u - head * a.nwin;
        uint8_t* st = smem + s * a.stage_bytes;
        mbar_expect_tx(&kv_full[s], tx_bytes);
        for (int op = 0; op < 3; ++op)
          for (int b = 0; b < a.nb; ++b)
            tma_load_2d(st + op * NP * 64 + b * a.BR * 64, &tmQKV, &kv_full[s], op * p.C + head * TC_HD, win * N + b * a.BR);
        if (a.r4) {
          // tail query rows, one copy per TMEM lane quadrant (rows past the tensor end are zero-filled by TMA)
          for (int qd = 0; qd < 4; ++qd)
            tma_load_2d(st + 3 * NP * 64 + qd * 32 * 64, &tmQT, &kv_full[s], head * TC_HD, win * N + (ntiles - 1) * 128);
        }
      }
    }
  } else if (warp >= TC_MMA_WARP0) {
    // =============================== MMA issuer of one column group (one thread) ===============================
    // QK(t) -> [softmax warps] -> PV(t) -> QK(t+1) ...  tcgen05.mma of one thread execute in issue order, so the
    // S_g / P_g columns are safely overwritten by QK(t+1) after PV(t) has read them.
    // The whole warp runs this loop with warp-uniform values (so descriptors live in uniform registers and each
    // tcgen05.mma is a single predicated instruction); one elected lane issues.
    if (nunits > 0) {
      const int g = __shfl_sync(0xffffffffu, warp - TC_MMA_WARP0, 0);
      const int c0 = tc_group_begin(NP, g), len = tc_group_end(NP, g) - c0;
      const uint32_t idesc_qk = make_idesc_bf16_f32(128, len);
      const uint32_t idesc_pv = make_idesc_bf16_f32(128, TC_HD) | (1u << 16);     // B (= V) is MN-major
      const uint32_t tmem_s = tmem_base + c0;
      const uint32_t tmem_o = tmem_base + TC_O_COL + g * TC_HD;
      const int nks = len >> 4;
      int t = 0;
      for (int lu = 0; lu < nunits; ++lu) {
        const int s = lu & 1;
        const uint32_t sq = smem_u32(smem + s * a.stage_bytes);
        const uint64_t dk = make_sw64_desc(sq + NP * 64 + c0 * 64);
        const uint64_t dv = make_sw64_desc(sq + 2 * NP * 64 + c0 * 64);
        mbar_wait(&kv_full[s], (lu >> 1) & 1);
        for (int qt = 0; qt < ntiles; ++qt, ++t) {
          const uint64_t dq = make_sw64_desc((a.r4 && qt == ntiles - 1) ? sq + 3 * NP * 64 : sq + qt * 128 * 64);
          tc_fence_after();
          if (elect_one_sync()) {
            umma_bf16_ss(tmem_s, dq, dk, idesc_qk, 0);
            umma_bf16_ss(tmem_s, dq + 2, dk + 2, idesc_qk, 1);
            umma_commit(&s_full[g]);
          }
          __syncwarp();
          if (lane == 0) TC_TRACE(0 + g, t);
          mbar_wait(&p_ready[g], t & 1);
          if (t >= 1) mbar_wait(o_free, (t - 1) & 1);      // the epilogue of the previous tile no longer reads O_g
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll 1
            for (int ks = 0; ks < nks; ++ks)               // 16 keys per step: 8 packed P columns, 16 V rows (1 KB)
              umma_bf16_ts(tmem_o, tmem_s + 8 * ks, dv + 64 * ks, idesc_pv, ks);
            umma_commit(o_full);
          }
          __syncwarp();
          if (lane == 0) TC_TRACE(3 + g, t);
        }
        if (elect_one_sync()) umma_commit(&stage_free[s]);
        __syncwarp();
      }
    }
  } else {
    // =============================== softmax warps (last group: + epilogue) ===============================
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const int c0 = tc_group_begin(NP, g), c1 = tc_group_end(NP, g);
    const uint32_t ts = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nW = wg.nwd * wg.nwh * wg.nww;
    const int Dp = wg.nwd * wg.wd, Hp = wg.nwh * wg.wh, Wp = wg.nww * wg.ww;

    // epilogue of tile te (group TC_EPI_GROUP only): O = sum_g 2^(m_g - m) O_g / sum_g 2^(m_g - m) l_g
    auto epilogue = [&](int te) {
      const int lu = te / ntiles, qt = te - lu * ntiles;
      const int u = u_begin + lu;
      const int head = u / a.nwin, win = u - head * a.nwin;
      const bool rep = a.r4 && qt == ntiles - 1;            // replicated tail tile: every quadrant holds the same rows
      const int i = rep ? qt * 128 + lane : qt * 128 + r;
      const bool wvalid = rep || (qt * 128 + q * 32) < N;
      if (q == 0 && lane == 0) TC_TRACE(15, te);
      mbar_wait(o_full, te & 1);
      tc_fence_after();
      if (wvalid) {
        const float* pmb = pm + (te & 1) * TC_GROUPS * 128 + r;
        const float* psb = ps + (te & 1) * TC_GROUPS * 128 + r;
        const float m0 = pmb[0], m1 = pmb[128], m2 = pmb[256];
        const float m = fmaxf(m0, fmaxf(m1, m2));
        const float w0 = ex2_ftz(m0 - m), w1 = ex2_ftz(m1 - m), w2 = ex2_ftz(m2 - m);
        const float l = w0 * psb[0] + w1 * psb[128] + w2 * psb[256];
        const float inv = rep ? 1.0f : 1.0f / l;
        const float wgt[3] = {w0 * inv, w1 * inv, w2 * inv};
        float lse_val = 0.f;                                // row statistic for the backward pass (base 2, like the scores)
        if (p.lse) lse_val = m + __log2f(l);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
        for (int gg = 0; gg < TC_GROUPS; ++gg) {
          uint32_t o[32];
          tmem_ld_x32(ts + TC_O_COL + gg * TC_HD, o);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(wgt[gg], __uint_as_float(o[j]), acc[j]);
        }
        if (rep) {
          // this quadrant only saw a quarter of each group's keys: merge the four partials through shared memory
          if (lane < a.tail_rows) {
            float* dstq = xq + (q * 16 + lane) * 34;
            dstq[0] = m;
            dstq[1] = l;
#pragma unroll
            for (int j = 0; j < 32; ++j) dstq[2 + j] = acc[j];
          }
          named_bar(1 + TC_EPI_GROUP, 128);
          if (q == 0 && lane < a.tail_rows) {
            const float* x0 = xq + lane * 34;
            const float mm = fmaxf(fmaxf(x0[0], x0[16 * 34]), fmaxf(x0[2 * 16 * 34], x0[3 * 16 * 34]));
            float wq[4], ll = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              wq[k] = ex2_ftz(x0[k * 16 * 34] - mm);
              ll = fmaf(wq[k], x0[k * 16 * 34 + 1], ll);
            }
            const float iv = 1.0f / ll;
            if (p.lse) lse_val = mm + __log2f(ll);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = 0.f;
#pragma unroll
              for (int k = 0; k < 4; ++k) v = fmaf(wq[k], x0[k * 16 * 34 + 2 + j], v);
              acc[j] = v * iv;
            }
          }
          named_bar(1 + TC_EPI_GROUP, 128);
        }
        if (i < N && (!rep || q == 0)) {
          if (p.lse) p.lse[(static_cast<long long>(win) * N + i) * p.nH + head] = lse_val;
          __nv_bfloat16* dst = p.out + (static_cast<long long>(win) * N + i) * p.C + head * TC_HD;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = pack_bf16x2(acc[h * 16 + 2 * j], acc[h * 16 + 2 * j + 1]);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + h * 16), "r"(w8[0]),
                         "r"(w8[1]), "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                         : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (q == 0 && lane == 0) TC_TRACE(16, te);
    };

    int cur_head = -1;
    int t = 0;
    for (int lu = 0; lu < nunits; ++lu) {
      const int u = u_begin + lu;
      const int head = u / a.nwin, win = u - head * a.nwin;
      if (head != cur_head) {
        // all softmax warps are past the previous head's tiles -> rebuild the expanded (bank-conflict-free) table
        named_bar(5, TC_SOFTMAX_THREADS);
        const float* src = p.table_t + static_cast<long long>(head) * p.L;
        const int e2 = 2 * wg.Ww - 1, e1 = 2 * wg.Wh - 1;
        for (int i = threadIdx.x; i < p.L; i += TC_SOFTMAX_THREADS) {
          const int cc = i % e2, bb = (i / e2) % e1, aa = i / (e2 * e1);
          tab[aa * SD + bb * SH + cc] = __ldg(src + i) * TC_LOG2E;
        }
        cur_head = head;
        named_bar(5, TC_SOFTMAX_THREADS);
        if (g > 0) {
          // phase-shift the group pipelines (the barrier aligned them): pass 1 (LDS bound) of one group then overlaps
          // pass 2 (MUFU bound) of another instead of all groups fighting for the same pipe at the same time
          const long long t0 = clock64();
          while (clock64() - t0 < static_cast<long long>(g) * a.stagger) {
          }
        }
      }
      bool need_mask = false;
      int wa = 0, wb = 0, wc = 0, rd0 = 0, rh0 = 0, rw0 = 0;
      if (a.shifted) {
        const int wi = win % nW;
        wc = wi % wg.nww; wb = (wi / wg.nww) % wg.nwh; wa = wi / (wg.nww * wg.nwh);
        need_mask = (wg.sd && wa == wg.nwd - 1) || (wg.sh && wb == wg.nwh - 1) || (wg.sw && wc == wg.nww - 1);
        if (need_mask) {
          // region class = per-axis (region - region of the window's first token): at most two regions per axis.
          // each group fills (and later reads) only its own key columns of the mask table
          rd0 = shift_region(wa * wg.wd, Dp, wg.wd, wg.sd);
          rh0 = shift_region(wb * wg.wh, Hp, wg.wh, wg.sh);
          rw0 = shift_region(wc * wg.ww, Wp, wg.ww, wg.sw);
          named_bar(1 + g, 128);
          for (int j = c0 + (threadIdx.x & 127); j < c1; j += 128) {
            int cj = 0;
            if (j < N) {
              const int tw = j % wg.ww, th = (j / wg.ww) % wg.wh, td = j / (wg.ww * wg.wh);
              cj = 4 * (shift_region(wa * wg.wd + td, Dp, wg.wd, wg.sd) - rd0) +
                   2 * (shift_region(wb * wg.wh + th, Hp, wg.wh, wg.sh) - rh0) +
                   (shift_region(wc * wg.ww + tw, Wp, wg.ww, wg.sw) - rw0);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mtab[k * TC_MSTRIDE + j] = (k != cj) ? TC_MASKV : 0.0f;
          }
          named_bar(1 + g, 128);
        }
      }

      for (int qt = 0; qt < ntiles; ++qt, ++t) {
        // Tail tile with <= 16 live rows: the rows are replicated into all four lane quadrants (see the TMA producer) and
        // quadrant q processes only a quarter of the group's keys, so the tile costs a quarter of a full one instead of
        // leaving three of the four SM sub-partitions idle.
        const bool rep = a.r4 && qt == ntiles - 1;
        const int i = rep ? qt * 128 + lane : qt * 128 + r;
        const bool wvalid = rep || (qt * 128 + q * 32) < N;   // warp-uniform: any live query row in this warp?
        int s0 = c0, s1 = c1;
        if (rep) {
          const int nblk = (c1 - c0) >> 4;
          s0 = c0 + 16 * ((q * nblk) >> 2);
          s1 = c0 + 16 * (((q + 1) * nblk) >> 2);
        }
        // the last group (in phase order) drains the PREVIOUS tile's accumulators before its own softmax: by then the
        // other groups' P.V of that tile were issued long ago, and no P.V of this tile ever waits for the epilogue
        if (g == TC_EPI_GROUP && t >= 1) epilogue(t - 1);
        mbar_wait(&s_full[g], t & 1);
        tc_fence_after();
        if (q == 0 && lane == 0) TC_TRACE(6 + g, t);
        if (wvalid) {
          const int ic = i < N ? i : N - 1;
          const int code_i = (ic / (wg.Wh * wg.Ww)) * SD + ((ic / wg.Ww) % wg.Wh) * SH + ic % wg.Ww;
          const float* tabq = tab + code_i + a.rc;
          const float* mrow = mtab;
          if (need_mask) {
            const int tw = ic % wg.ww, th = (ic / wg.ww) % wg.wh, td = ic / (wg.ww * wg.wh);
            const int ci = 4 * (shift_region(wa * wg.wd + td, Dp, wg.wd, wg.sd) - rd0) +
                           2 * (shift_region(wb * wg.wh + th, Hp, wg.wh, wg.sh) - rh0) +
                           (shift_region(wc * wg.ww + tw, Wp, wg.ww, wg.sw) - rw0);
            mrow = mtab + ci * TC_MSTRIDE;
          }
          float m0 = -INFINITY, m1 = -INFINITY;
          int c = s0;
          for (; c + 32 <= s1; c += 32) pass1_piece<32>(ts, c, N, tabq, negoff, need_mask, mrow, m0, m1);
          if (c < s1) pass1_piece<16>(ts, c, N, tabq, negoff, need_mask, mrow, m0, m1);
          tmem_st_wait();
          if (q == 0 && lane == 0) TC_TRACE(9 + g, t);
          const float m = fmaxf(fmaxf(m0, m1), -1e30f);       // finite even for an empty key range (replicated tail)
          float l0 = 0.f, l1 = 0.f;
          for (c = s0; c + 32 <= s1; c += 32) pass2_piece<32>(ts + c, ts + c0 + ((c - c0) >> 1), -m, l0, l1);
          if (c < s1) pass2_piece<16>(ts + c, ts + c0 + ((c - c0) >> 1), -m, l0, l1);
          if (rep) {
            // keys of this group owned by the other quadrants contribute nothing to these lanes: P = 0 there
            uint32_t z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = 0u;
            for (c = c0; c < c1; c += 16)
              if (c < s0 || c >= s1) tmem_st_x8(ts + c0 + ((c - c0) >> 1), z);
          }
          tmem_st_wait();
          pm[((t & 1) * TC_GROUPS + g) * 128 + r] = m;
          ps[((t & 1) * TC_GROUPS + g) * 128 + r] = l0 + l1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        if (q == 0 && lane == 0) TC_TRACE(12 + g, t);
      }
    }
    if (g == TC_EPI_GROUP && T > 0) epilogue(T - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool window_attn_tc_supported(const AttnParams& p) {
  const WinGeom& g = p.win;
  if (g.N < 48 || g.N > TC_MAX_NP) return false;          // at least one 16-column block per group
  const int nb = (g.N + 255) / 256;
  if (g.N % nb != 0) return false;
  const int SH = tc_stride(2 * g.Ww - 1, g.Ww % 32);
  const int SD = tc_stride((2 * g.Wh - 1) * SH, (g.Wh * g.Ww) % 32);
  if ((2 * g.Wd - 1) * SD > 12288) return false;           // expanded table <= 48 KB
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
  a.SH = tc_stride(2 * g.Ww - 1, g.Ww % 32);
  a.SD = tc_stride((2 * g.Wh - 1) * a.SH, (g.Wh * g.Ww) % 32);
  a.L2 = (2 * g.Wd - 1) * a.SD;
  a.rc = (g.Wd - 1) * a.SD + (g.Wh - 1) * a.SH + (g.Ww - 1);
  a.tail_rows = g.N - (a.ntiles - 1) * 128;
  a.r4 = (a.ntiles >= 2 && a.tail_rows <= 16) ? 1 : 0;
  {
    static int no_r4 = -1;
    if (no_r4 < 0) {
      const char* e = getenv("LAVT_ATTN_NO_R4");
      no_r4 = (e && e[0] == '1') ? 1 : 0;
    }
    if (no_r4) a.r4 = 0;
  }
  a.stage_bytes = 3 * a.NP * 64 + (a.r4 ? 128 * 64 : 0);
  // the last query tile reads up to 128 rows past N from the Q region: that runs into the K / V regions of the stage
  LAVT_REQUIRE(a.ntiles * 128 <= 3 * a.NP, "attention(tc): query tile overrun");
  int off = 2 * a.stage_bytes;
  a.off_tab = off;        off += ((a.L2 * 4 + 127) / 128) * 128;
  a.off_mtab = off;       off += ((8 * TC_MSTRIDE * 4 + 127) / 128) * 128;
  a.off_negoff = off;     off += ((a.NP * 4 + 127) / 128) * 128;
  a.off_pm = off;         off += 2 * TC_GROUPS * 128 * 4;
  a.off_ps = off;         off += 2 * TC_GROUPS * 128 * 4;
  a.off_xq = off;         off += a.r4 ? 4 * 16 * 34 * 4 : 0;
  a.off_bar = off;        off += 128;
  const int smem = off + 1024;
  LAVT_REQUIRE(smem <= 227 * 1024, "attention(tc): shared memory %d B exceeds the SM", smem);
  a.shifted = (g.sd | g.sh | g.sw) != 0;

  CUtensorMap tm, tm_tail;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3 * p.C), static_cast<uint64_t>(nwin * g.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(3 * p.C) * 2};
    uint32_t box[2] = {TC_HD, static_cast<uint32_t>(a.BR)};
    int rc = make_tmap_bf16_l2_64b(&tm, p.qkv, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    uint32_t box_t[2] = {TC_HD, 32};
    rc = make_tmap_bf16_l2_64b(&tm_tail, p.qkv, 2, dims, strides, box_t, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  const int grid = a.units < sms ? a.units : sms;
  static int configured = 0;
  if (smem > configured) {
    LAVT_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  static int stagger = -1;
  if (stagger < 0) {
    const char* e = getenv("LAVT_ATTN_STAGGER");
    stagger = e ? atoi(e) : 2000;
  }
  a.stagger = stagger;
  a.trace = nullptr;
  const char* trace_path = getenv("LAVT_ATTN_TRACE");
  if (trace_path) LAVT_CUDA(cudaMalloc(&a.trace, 17 * TC_TRACE_TILES * sizeof(long long)));
  if (a.trace) LAVT_CUDA(cudaMemsetAsync(a.trace, 0, 17 * TC_TRACE_TILES * sizeof(long long), st));
  window_attn_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tm, tm_tail, p, a);
  LAVT_LAUNCH_CHECK("window_attn_tc_kernel");
  if (a.trace) {
    // debug only: synchronous dump of CTA 0's event clocks
    static long long host[17 * TC_TRACE_TILES];
    LAVT_CUDA(cudaStreamSynchronize(st));
    LAVT_CUDA(cudaMemcpy(host, a.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int e = 0; e < 17; ++e) {
        for (int t = 0; t < TC_TRACE_TILES; ++t) fprintf(f, "%lld ", host[e * TC_TRACE_TILES + t]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return LAVT_OK;
}

}  // namespace lavt
