// tcgen05 / TMEM shifted-window attention core, third generation: windows whose (h, w) extent is the configured 7 x 7 or 12 x 12
// (N = 49 * wd or 144 * wd tokens, wd <= 8: every unclamped window of the 8 x 7 x 7 / 8 x 12 x 12 video models and of the 2-D image models).
// The description below is written for 7 x 7; the 12 x 12 instantiation (T3G<12>) differs in the numbers only: runs of 12 keys need no
// padding, a chunk is 4 runs = 48 columns (three chunks per frame), and the 33 KB table exists once (scalar bias loads at immediate
// offsets: four shifted copies do not fit next to 147 KB of K / V, which is also why that shape has a single K / V stage).
// (reference WindowAttention3D.forward, lib/video_swin_transformer.py:147-165; 2-D twin lib/backbone.py:127-138)
//     S = q k^T + relative-position bias (+ shifted-window mask)  ->  softmax  ->  O = P v       per (window, head)
//
// What the profiles of the first two generations said (profiles/r2_ncu_attn_tc2_*): 12.6 issued instructions per score against 1
// MUFU.EX2; one shared-memory gather + index arithmetic per score for the bias; ~15 % of the issue slots in mbarrier polling of
// dedicated MMA-issuer warps; the key-split groups merge partial (m, l, O) per tile.  This kernel is built around the exponential:
//   * KEY LAYOUT.  K and V of a unit arrive through ONE 4-D TMA box each, (32 ch, 8 of 7 w, 8 of 7 h, wd frames): the out-of-bounds
//     w = 7 column and h = 7 row are zero-filled by TMA, so in shared memory (and therefore along the S columns) every run of 7 keys
//     that share (frame, h) starts at a multiple of 8 and every frame at a multiple of 64.  A frame = one 64-column key chunk.
//   * BIAS.  Along such a run the table index is contiguous, so with the w axis of the table flipped a run's 7 bias values are 7
//     consecutive floats.  Four copies of the table, shifted by 0..3 floats, make the run start 16-byte aligned in one of them: the
//     bias of a run is TWO LDS.128 at an immediate offset from one per-row base register -- 0.29 loads per score, no index math.
//   * ROW-PARALLEL warpgroups.  Each of the three softmax warpgroups owns whole 128-row query tiles (thread = query row, all chunks of
//     the tile, online softmax with a lazily re-based maximum like attn_tc2.cu) -- no cross-group merge of partial results -- with two
//     64-column S buffers and one O accumulator in TMEM (3 x 160 columns).  One issuer warp per warpgroup: P.V of chunk n, then
//     Q.K^T of chunk n + 2 into the buffer that P.V just read (tcgen05.mma of one thread execute in order).
//   * the tile epilogue (O / l -> bf16) is deferred until the first chunk of the warpgroup's next tile has been exponentiated, when
//     the last P.V of the old tile has long retired.
//   * a tail tile of <= 32 rows lands in a different TMEM lane quadrant for every unit, so its single active warp loads the four SM
//     sub-partitions evenly.
// Warp roles: warps 0-11 softmax (warpgroup g = warp / 4, lane quadrant q = warp % 4), warps 12-14 MMA issuers of warpgroup 0 / 1 / 2
// (the softmax warps only arrive on an mbarrier when their P rows are written and never wait for the issue: measured with the
// in-kernel timeline, issuing from a softmax warp put ~900 cycles per chunk on that warp's critical path), warp 15 TMA producer + TMEM
// allocation.
// q arrives pre-scaled by head_dim^-0.5 * log2(e) (qkv GEMM epilogue); the table is multiplied by log2(e) when it is staged.
#include "kernels.cuh"
#include "attn_tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace lavt {

constexpr int T3_HD = 32;
constexpr int T3_WGS = 3;
constexpr int T3_SM_THREADS = 128 * T3_WGS;
constexpr int T3_TMA_WARP = 4 * T3_WGS + 3;   // warp 15: the SM sub-partition (warp % 4 = 3) that hosts no MMA-issuing softmax warp
constexpr int T3_THREADS = 512;               // warps 12-14 only run to the final barrier (registers are granted per 4 warps anyway)
// Window-shape traits.  RW keys per run (= window width), RP S columns per run, WH runs per frame, RC runs per chunk (incl. the all-padding
// run of a 7 x 7 frame), CPF chunks per frame, CW S columns per chunk, SH / EH / EW bias-table geometry in shared memory (w' in [0, EW),
// h in [0, EH), row stride SH, frame stride SD), NCOPY shifted table copies, NR0 + NR1 runs in the two pieces of a chunk.
template <int RW_> struct T3G;
template <> struct T3G<7> {
  static constexpr int RW = 7, RP = 8, WH = 7, RC = 8, CPF = 1, CW = 64, SH = 16, EH = 13, EW = 13, NCOPY = 4, NR0 = 4, NR1 = 3;
};
template <> struct T3G<12> {
  static constexpr int RW = 12, RP = 12, WH = 12, RC = 4, CPF = 3, CW = 48, SH = 24, EH = 23, EW = 23, NCOPY = 1, NR0 = 2, NR1 = 2;
};
constexpr int T3_NQ = 4;                    // Q-tile ring slots
constexpr float T3_LOG2E = 1.4426950408889634f;
constexpr float T3_MASKV = -100.0f * T3_LOG2E;
constexpr float T3_PSUM_LIMIT = 1048576.0f;  // 2^20

#ifdef T3_WATCHDOG
// debug build (-DT3_WATCHDOG): a wait that does not complete within ~0.5 s reports where it is stuck and traps
__device__ __noinline__ void t3_stuck(int tag, uint32_t parity) {
  printf("[tc3 stuck] block %d warp %d lane %d tag %d parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, tag, parity);
  __trap();
}
__device__ __forceinline__ void t3_wait(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 1000000000LL) t3_stuck(tag, parity);
}
#else
__device__ __forceinline__ void t3_wait(uint64_t* bar, uint32_t parity, int) { mbar_wait(bar, parity); }
#endif

struct AttnTc3Args {
  int N, nch, ntiles;       // tokens per window, key chunks per window, 128-row query tiles
  int nkv;                  // K | V stages (2, or 1 when a unit's K / V fill most of the shared memory)
  int BR, nb;               // 12 x 12 windows: K / V arrive as nb plain 2-D boxes of BR rows
  int nwin, units;
  int tail_rows, rot;       // rot: the last tile has <= 32 rows and rotates through the lane quadrants
  int CS;                   // floats between the four shifted table copies
  int kv_bytes;             // K | V of one stage
  int off_q, off_tab, off_bar;
  int shifted;
  long long* trace;         // -DT3_TRACE builds: clock64 stamps of CTA 0, [16 warps][8 events][T3_TRACE_ITEMS]
};
constexpr int T3_TRACE_ITEMS = 96;

struct T3Row {
  const float* tb;          // this row's bias address of (frame 0, run h_j = 6); run h_j: + (6 - h_j) * SH, frame t_j: - t_j * SD
  float m, l;
  float mw[12];             // masked windows: w-axis mask of the keys of a run (0 or -100 log2 e)
  uint32_t dm, hm;          // masked windows: bit t_j / h_j set = that frame / run lies in another region than this row
};

// running maximum / sum helpers over the RW live values of a run
template <int RW>
__device__ __forceinline__ void t3_run_max(const float* x, float& m0, float& m1) {
#pragma unroll
  for (int e = 0; e + 1 < RW; e += 2) {
    if ((e >> 1) & 1) m1 = max3(m1, x[e], x[e + 1]); else m0 = max3(m0, x[e], x[e + 1]);
  }
  if (RW & 1) m1 = fmaxf(m1, x[RW - 1]);
}

// One piece = NR runs of RP columns (RW live) of the chunk in TMEM at ts (fp32 scores), written back as bf16 pairs at tp.
//   fb  : bias address of run h_j = WH - 1 of this frame for this row, moved back by the chunk's first run;   HJ0: first run of the piece
//   pdone : P columns of this chunk already written (slow path rescales them), o_acc: the O accumulator holds earlier chunks
template <typename G, int NR, int HJ0, bool MASKED, bool FIRST>
__device__ __forceinline__ void t3_piece(uint32_t ts, uint32_t tp, const float* fb, T3Row& r, uint32_t rmask, uint32_t tp_chunk, int pdone,
                                         bool o_acc, uint64_t* pv_done, uint32_t pvp, uint32_t tmem_o) {
  constexpr int RW = G::RW, RP = G::RP;
  constexpr int W = RP * NR;
  static_assert(W == 32 || W == 24, "pieces are 32 or 24 columns");
  uint32_t v[W];
#ifdef T3_X_NOLD
#pragma unroll
  for (int j = 0; j < W; ++j) v[j] = ts + j;
#else
  if constexpr (W == 32) {
    tmem_ld_x32(ts, v);
  } else {
    tmem_ld_x16(ts, v);
    tmem_ld_x8(ts + 16, v + 16);
  }
#endif
  float x[W];
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const float* rb = fb + (G::WH - 1 - HJ0 - k) * G::SH;
#ifdef T3_X_NOBIAS
#pragma unroll
    for (int e = 0; e < RP; ++e) x[RP * k + e] = 0.f;
#else
    if constexpr (G::NCOPY == 4) {          // 16-byte aligned run start in this row's table copy: two LDS.128 (the 8th value is unused)
      const float4 b0 = *reinterpret_cast<const float4*>(rb);
      const float4 b1 = *reinterpret_cast<const float4*>(rb + 4);
      x[8 * k + 0] = b0.x; x[8 * k + 1] = b0.y; x[8 * k + 2] = b0.z; x[8 * k + 3] = b0.w;
      x[8 * k + 4] = b1.x; x[8 * k + 5] = b1.y; x[8 * k + 6] = b1.z; x[8 * k + 7] = b1.w;
    } else {
#pragma unroll
      for (int e = 0; e < RP; ++e) x[RP * k + e] = rb[e];
    }
#endif
  }
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < W; j += 2) add2(x[j], x[j + 1], __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
  if constexpr (MASKED) {
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const bool rm = (rmask >> (HJ0 + k)) & 1u;
#pragma unroll
      for (int e = 0; e < RP; e += 2) add2(x[RP * k + e], x[RP * k + e + 1], rm ? T3_MASKV : r.mw[e], rm ? T3_MASKV : r.mw[e + 1]);
    }
  }
  if constexpr (FIRST) {
    float m0 = -1e30f, m1 = -1e30f;
#pragma unroll
    for (int k = 0; k < NR; ++k) t3_run_max<RW>(x + RP * k, m0, m1);
    r.m = fmaxf(m0, m1);
  }
  float p[W];
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  {
    const float nm = -r.m;
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      p[j] = x[j];
      p[j + 1] = x[j + 1];
      add2(p[j], p[j + 1], nm, nm);
    }
#ifndef T3_X_NOMUFU
#pragma unroll
    for (int k = 0; k < NR; ++k) {
#pragma unroll
      for (int e = 0; e < RW; ++e) p[RP * k + e] = ex2_ftz(p[RP * k + e]);
    }
#endif
#pragma unroll
    for (int k = 0; k < NR; ++k) {
#pragma unroll
      for (int e = 0; e + 1 < RW; e += 2) {
        if ((e >> 1) & 1) add2(l2, l3, p[RP * k + e], p[RP * k + e + 1]); else add2(l0, l1, p[RP * k + e], p[RP * k + e + 1]);
      }
      if (RW & 1) l2 += p[RP * k + RW - 1];
    }
  }
  float ps = (l0 + l1) + (l2 + l3);
  if constexpr (!FIRST) {
    if (__any_sync(0xffffffffu, !(ps <= T3_PSUM_LIMIT))) {
      // ---- slow path (rare): some score of this piece exceeds the running maximum by ~2^14: re-base the row ----
      float m0 = r.m, m1 = r.m;
#pragma unroll
      for (int k = 0; k < NR; ++k) t3_run_max<RW>(x + RP * k, m0, m1);
      const float m2 = fmaxf(m0, m1);
      const float f = ex2_ftz(r.m - m2);          // 1 for the lanes whose maximum did not move
      r.l *= f;
      tmem_st_wait();                              // the P columns stored by the previous piece are read back below
      for (int pc = 0; pc < pdone; pc += 2) {      // P columns of this chunk that were already written (bf16 pairs)
        uint32_t w2[2];
        tmem_ld_x2(tp_chunk + pc, w2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float2 t = unpack_bf16x2(w2[j]);
          w2[j] = pack_bf16x2(t.x * f, t.y * f);
        }
        tmem_st_x2(tp_chunk + pc, w2);
      }
      if (o_acc) {
        t3_wait(pv_done, pvp, 1);                  // every P.V issued so far into this accumulator has retired
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_x32(tmem_o, o);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
        tmem_st_x32(tmem_o, o);
      }
      tmem_st_wait();
      r.m = m2;
      ps = 0.f;
      const float nm = -m2;
#pragma unroll
      for (int k = 0; k < NR; ++k) {
#pragma unroll
        for (int e = 0; e < RW; ++e) {
          p[RP * k + e] = ex2_ftz(x[RP * k + e] + nm);
          ps += p[RP * k + e];
        }
      }
    }
  }
  r.l += ps;
  uint32_t pk[16];
#pragma unroll
  for (int k = 0; k < NR; ++k) {
#pragma unroll
    for (int e = 0; e < RP; e += 2) pk[(RP / 2) * k + (e >> 1)] = pack_bf16x2(p[RP * k + e], (e + 1 < RW) ? p[RP * k + e + 1] : 0.f);
  }
#pragma unroll
  for (int j = NR * RP / 2; j < 16; ++j) pk[j] = 0u;          // 7 x 7: the all-padding run that ends the frame
#ifdef T3_X_NOST
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc ^= pk[j];
  if (acc == 0x12345678u) tmem_st_x16(tp, pk);
#else
  if constexpr (G::RP == 8) {
    tmem_st_x16(tp, pk);                         // 4 runs (or 3 + the padding run) x 4 words
  } else {
    tmem_st_x8(tp, pk);                          // 2 runs x 6 words
    tmem_st_x4(tp + 8, pk + 8);
  }
#endif
}

// one chunk (chunk index c of the window) for this thread's row
template <typename G, bool MASKED, bool FIRST>
__device__ __forceinline__ void t3_chunk(uint32_t ts_buf, int c, T3Row& r, bool o_acc, uint64_t* pv_done, uint32_t pvp, uint32_t tmem_o) {
  const int tj = G::CPF == 1 ? c : c / G::CPF;
  const int run0 = G::CPF == 1 ? 0 : (c - tj * G::CPF) * G::RC;
  const float* fb = r.tb - tj * (G::EH * G::SH) - run0 * G::SH;
  uint32_t rmask = 0;
  if constexpr (MASKED) rmask = (((r.dm >> tj) & 1u) ? 0xfffu : r.hm) >> run0;
  constexpr int P0 = G::NR0 * G::RP;             // S columns of the first piece (its P takes half as many)
  t3_piece<G, G::NR0, 0, MASKED, FIRST>(ts_buf, ts_buf, fb, r, rmask, ts_buf, 0, o_acc, pv_done, pvp, tmem_o);
  t3_piece<G, G::NR1, G::NR0, MASKED, false>(ts_buf + P0, ts_buf + P0 / 2, fb, r, rmask, ts_buf, P0 / 2, o_acc, pv_done, pvp, tmem_o);
}

template <int RW>
__global__ void __launch_bounds__(T3_THREADS, 1)
window_attn_tc3_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ,
                       const __grid_constant__ CUtensorMap tmQT, const AttnParams p, const AttnTc3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* qring = smem + a.off_q;
  float* tab = reinterpret_cast<float*>(smem + a.off_tab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* kv_full = bars;                       // [2]
  uint64_t* kv_free = bars + 2;                   // [2]   one arrival per tile of the unit (its last P.V retired)
  uint64_t* q_full = bars + 4;                    // [T3_NQ]
  uint64_t* q_free = bars + 4 + T3_NQ;            // [T3_NQ]
  uint64_t* s_full = bars + 4 + 2 * T3_NQ;        // [T3_WGS][2]
  uint64_t* pv_done = s_full + 2 * T3_WGS;        // [T3_WGS]
  uint64_t* o_full = pv_done + T3_WGS;            // [T3_WGS]  the last P.V of a tile of that warpgroup retired
  uint64_t* p_ready = o_full + T3_WGS;            // [T3_WGS][2] the four warps of the warpgroup wrote their P rows of that S buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(p_ready + 2 * T3_WGS);

  using G = T3G<RW>;
  constexpr int T3_CW = G::CW, T3_WG_COLS = 2 * G::CW + T3_HD, T3_CHUNK_BYTES = G::CW * 64, T3_SH = G::SH, T3_SD = G::EH * G::SH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, nch = a.nch, ntiles = a.ntiles, nkv = a.nkv;
  const WinGeom& wg = p.win;
#ifdef T3_WATCHDOG
  volatile int* prog = reinterpret_cast<volatile int*>(smem + a.off_bar + 256);     // progress code per warp
  if (threadIdx.x < 16) prog[threadIdx.x] = 0;
#define T3_PROG(code) do { if (lane == 0) prog[warp] = (code); } while (0)
#else
#define T3_PROG(code) do { } while (0)
#endif
#ifdef T3_TRACE
#define T3_STAMP(ev, item) do { if (blockIdx.x == 0 && lane == 0 && a.trace && (item) < T3_TRACE_ITEMS) a.trace[(warp * 8 + (ev)) * T3_TRACE_ITEMS + (item)] = clock64(); } while (0)
#else
#define T3_STAMP(ev, item) do { } while (0)
#endif

  // contiguous unit range of this CTA; unit u = head * nwin + window (head-major: the bias table is restaged rarely)
  const int u_begin = static_cast<int>(1LL * a.units * blockIdx.x / gridDim.x);
  const int u_end = static_cast<int>(1LL * a.units * (blockIdx.x + 1) / gridDim.x);
  const int nunits = u_end - u_begin;
  const int T = nunits * ntiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmQ);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], ntiles);
    }
    for (int i = 0; i < T3_NQ; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_free[i], 1);
    }
    for (int i = 0; i < 2 * T3_WGS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
    }
    for (int i = 0; i < T3_WGS; ++i) {
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == T3_TMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  for (int i = threadIdx.x; i < G::NCOPY * a.CS; i += blockDim.x) tab[i] = 0.f;      // padding entries of the table copies: any finite value
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();                      // barriers, TMEM and the zeroed table are set up under the previous kernel's tail
  pdl_launch_dependents();

  // Register split (the launch grants 128 per thread): the control warpgroup (issuers + producer) keeps 56, each softmax warpgroup
  // grows to 152:  3 x 128 x 152 + 128 x 56 = 65 536.
  if (warp >= 4 * T3_WGS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
  }
  if (warp == T3_TMA_WARP) {
#ifdef T3_WATCHDOG
    if (lane == 1) {
      const long long t0 = clock64();
      for (;;) {
        bool all = true;
        for (int w = 0; w < 4 * T3_WGS; ++w) all = all && prog[w] == 1000;
        if (all) break;
        if (clock64() - t0 > 4000000000LL) {
          printf("[tc3 watchdog] block %d progress:", blockIdx.x);
          for (int w = 0; w < 16; ++w) printf(" %d", prog[w]);
          printf("\n bars:");
          for (int w = 0; w < 30; ++w) printf(" %d:%llx", w, *reinterpret_cast<volatile unsigned long long*>(&bars[w]));
          printf("\n");
          __trap();
        }
      }
    }
#endif
    // =============================== TMA producer (one thread) ===============================
    if (lane == 0) {
      for (int lu = 0; lu < nunits; ++lu) {
        const int u = u_begin + lu;
        const int s = nkv == 2 ? (lu & 1) : 0, use = nkv == 2 ? (lu >> 1) : lu;      // stage and how often it was used before
        const int head = u / a.nwin, win = u - head * a.nwin;
        if (use >= 1) t3_wait(&kv_free[s], (use - 1) & 1, 2);                        // every P.V of the unit that used this stage retired
        uint8_t* st = smem + s * a.kv_bytes;
        mbar_expect_tx(&kv_full[s], 2u * nch * T3_CHUNK_BYTES);
        if constexpr (RW == 7) {
          tma_load_4d(st, &tmKV, &kv_full[s], p.C + head * T3_HD, 0, 0, win * nch);
          tma_load_4d(st + nch * T3_CHUNK_BYTES, &tmKV, &kv_full[s], 2 * p.C + head * T3_HD, 0, 0, win * nch);
        } else {
          for (int op = 0; op < 2; ++op)
            for (int b = 0; b < a.nb; ++b)
              tma_load_2d(st + op * nch * T3_CHUNK_BYTES + b * a.BR * 64, &tmKV, &kv_full[s], (1 + op) * p.C + head * T3_HD, win * N + b * a.BR);
        }
        T3_PROG(5 + lu * 100);
        for (int qt = 0; qt < ntiles; ++qt) {
          const int tq = lu * ntiles + qt, slot = tq & (T3_NQ - 1);
          if (tq >= T3_NQ) t3_wait(&q_free[slot], ((tq >> 2) - 1) & 1, 3);
          uint8_t* qs = qring + slot * (128 * 64);
          mbar_expect_tx(&q_full[slot], 128u * 64u);
          if (a.rot && qt == ntiles - 1) {
            // tail query rows, one copy per TMEM lane quadrant (rows past the tensor end are zero-filled by TMA)
            for (int qd = 0; qd < 4; ++qd) tma_load_2d(qs + qd * 32 * 64, &tmQT, &q_full[slot], head * T3_HD, win * N + qt * 128);
          } else {
            tma_load_2d(qs, &tmQ, &q_full[slot], head * T3_HD, win * N + qt * 128);
          }
        }
      }
    }
  } else if (warp >= 4 * T3_WGS) {
    // =============================== MMA issuer of warpgroup g (warp-uniform loop, one elected lane issues) ===============================
    // item n = (tile, chunk) of this warpgroup:  ... P(n) announced -> P.V(n) -> Q K^T(n + 2) into the S buffer that P.V(n) reads
    // (tcgen05.mma of one thread execute in issue order) ...; Q K^T(n + 1) already sits in the other buffer.
    const int g = warp - 4 * T3_WGS;
    const uint32_t tmem_wg = tmem_base + g * T3_WG_COLS;
    const uint32_t idesc_qk = make_idesc_bf16_f32(128, T3_CW);
    const uint32_t idesc_pv = make_idesc_bf16_f32(128, T3_HD) | (1u << 16);     // B (= V) is MN-major
    const uint32_t kv_base = smem_u32(smem), q_base = smem_u32(qring);
    int qk_tq = g, qk_lu = g / ntiles, qk_c = 0, qk_n = 0;
    int qk_qt = g - qk_lu * ntiles;
    // Look-ahead issues never block: the K/V stage of a unit two ahead is only released by P.V's this warp has not issued yet
    // (units of one or two tiles); the issue for the item that is due next may block -- it only depends on earlier tiles.
    auto issue_qk = [&](bool blocking) -> bool {
      if (qk_tq >= T) return false;
      const int s = nkv == 2 ? (qk_lu & 1) : 0, slot = qk_tq & (T3_NQ - 1);
      if (qk_c == 0) {
        const uint32_t pk = (nkv == 2 ? (qk_lu >> 1) : qk_lu) & 1, pq = (qk_tq >> 2) & 1;
        if (blocking) {
          t3_wait(&kv_full[s], pk, 4);
          t3_wait(&q_full[slot], pq, 5);
        } else if (!(mbar_test(&kv_full[s], pk) && mbar_test(&q_full[slot], pq))) {
          return false;
        }
        tc_fence_after();
      }
      const uint64_t dq = make_sw64_desc(q_base + slot * (128 * 64));
      const uint64_t dk = make_sw64_desc(kv_base + s * a.kv_bytes + qk_c * T3_CHUNK_BYTES);
      const uint32_t ts = tmem_wg + (qk_n & 1) * T3_CW;
      if (elect_one_sync()) {
        umma_bf16_ss(ts, dq, dk, idesc_qk, 0);
        umma_bf16_ss(ts, dq + 2, dk + 2, idesc_qk, 1);
        umma_commit(&s_full[g * 2 + (qk_n & 1)]);
        if (qk_c == nch - 1) umma_commit(&q_free[slot]);
      }
      __syncwarp();
      ++qk_n;
      if (++qk_c == nch) {
        qk_c = 0;
        qk_tq += T3_WGS;
        qk_qt += T3_WGS;
        while (qk_qt >= ntiles) {
          qk_qt -= ntiles;
          ++qk_lu;
        }
      }
      return true;
    };
    if (issue_qk(false)) issue_qk(false);
    int n = 0, lu = g / ntiles, qt = g - lu * ntiles;
    for (int tq = g; tq < T; tq += T3_WGS) {
      const int s = nkv == 2 ? (lu & 1) : 0;
      for (int c = 0; c < nch; ++c, ++n) {
        const int buf = n & 1;
        while (qk_n <= n) issue_qk(true);                         // normally issued long ago
        T3_STAMP(3, n);
        t3_wait(&p_ready[g * 2 + buf], (n >> 1) & 1, 8);           // the four softmax warps wrote their P rows of this chunk
        T3_STAMP(4, n);
        tc_fence_after();
        const uint64_t dv = make_sw64_desc(kv_base + s * a.kv_bytes + (nch + c) * T3_CHUNK_BYTES);
        const uint32_t tp = tmem_wg + buf * T3_CW;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < T3_CW / 16; ++ks)               // 16 keys per step: 8 packed P columns, 16 V rows (1 KB)
            umma_bf16_ts(tmem_wg + 2 * T3_CW, tp + 8 * ks, dv + 64 * ks, idesc_pv, (c > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&pv_done[g]);
          if (c == nch - 1) {
            umma_commit(&o_full[g]);
            umma_commit(&kv_free[s]);
          }
        }
        __syncwarp();
        T3_STAMP(5, n);
        while (qk_n < n + 3 && issue_qk(false)) {
        }
        T3_STAMP(6, n);
      }
      qt += T3_WGS;
      while (qt >= ntiles) {
        qt -= ntiles;
        ++lu;
      }
    }
  } else {
    // =============================== softmax warpgroups ===============================
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * T3_WG_COLS;
    const uint32_t tmem_o = tlane + 2 * T3_CW;
    uint64_t* const my_pv = &pv_done[g];
    const int nW = wg.nwd * wg.nwh * wg.nww;
    const int head_first = u_begin / a.nwin;
    const int nevents = nunits > 0 ? (u_end - 1) / a.nwin - head_first + 1 : 0;

    // ---- bias table of one head: four copies shifted by 0..3 floats, w axis flipped, strides (SD, SH, 1) ----
    auto stage_table = [&](int head) {
      // the compact table is read with independent coalesced loads (a dependent load per expanded entry cost 20 000 cycles per head: ncu
      // showed the softmax warps parked on it) and every entry is scattered into the shifted copies; w axis flipped, strides (SD, SH, 1)
      const float* src = p.table_t + static_cast<long long>(head) * p.L;
      named_bar(4, T3_SM_THREADS);                 // every warpgroup is done with the previous head's tiles
      for (int base = 0; base < p.L; base += 8 * T3_SM_THREADS) {
        float vals[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = base + threadIdx.x + j * T3_SM_THREADS;
          vals[j] = i < p.L ? __ldg(src + i) * T3_LOG2E : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = base + threadIdx.x + j * T3_SM_THREADS;
          if (i < p.L) {
            const int aa = i / (G::EH * G::EW), rem = i - aa * (G::EH * G::EW);
            const int bb = rem / G::EW, co = rem - bb * G::EW;
            const int nat = aa * T3_SD + bb * T3_SH + (G::EW - 1 - co);
#pragma unroll
            for (int k = 0; k < G::NCOPY; ++k)
              if (nat >= k) tab[k * a.CS + nat - k] = vals[j];
          }
        }
      }
      named_bar(4, T3_SM_THREADS);
    };

    int ev_done = 0;
    int n = 0;                                      // chunk items processed by this warpgroup
    int ntile_done = 0;                             // tiles whose chunks were all processed
    // deferred epilogue state (previous tile of this warpgroup)
    bool have_prev = false, prev_store = false, prev_wvalid = false;
    float prev_m = 0.f, prev_l = 1.f;
    long long prev_row = 0;
    int prev_head = 0;

    auto epilogue = [&]() {
      // the last P.V of that tile retired (its own barrier: pv_done's parity only tells the current phase from the one before, and at
      // the end of the warpgroup's work nothing guarantees that P.V(n - 2) is already complete when this wait starts)
      t3_wait(&o_full[g], static_cast<uint32_t>((ntile_done - 1) & 1), 6);
      tc_fence_after();
      if (prev_wvalid) {                        // warp-uniform: tcgen05.ld is .sync.aligned
        uint32_t o[32];
        tmem_ld_x32(tmem_o, o);
        tmem_ld_wait();
        if (prev_store) {
          const float inv = 1.0f / prev_l;
          if (p.lse) p.lse[prev_row * p.nH + prev_head] = prev_m + __log2f(prev_l);
          __nv_bfloat16* dst = p.out + prev_row * p.C + prev_head * T3_HD;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              w8[j] = pack_bf16x2(__uint_as_float(o[h * 16 + 2 * j]) * inv, __uint_as_float(o[h * 16 + 2 * j + 1]) * inv);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + h * 16), "r"(w8[0]), "r"(w8[1]),
                         "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                         : "memory");
          }
        }
        __syncwarp();
      }
    };

    T3_PROG(1);
    T3_PROG(2);
    // (unit, tile, head, window) of this warpgroup's tiles advance incrementally: no runtime divisions per tile
    int lu = g / ntiles, qt = g - lu * ntiles;
    int head = (u_begin + lu) / a.nwin, win = (u_begin + lu) - head * a.nwin;
    for (int tq = g; tq < T; tq += T3_WGS) {
      T3_PROG(10 + tq * 100);
      while (ev_done <= head - head_first) {
        stage_table(head_first + ev_done);
        ++ev_done;
      }
      T3_PROG(11 + tq * 100);
      // ---- this thread's query row ----
      const bool rot = a.rot && qt == ntiles - 1;
      // a rotated tail tile is replicated into all four lane quadrants (see the producer); the warp of quadrant lu % 4 processes it
      const int i = qt * 128 + (rot ? lane : r);
      const bool wvalid = rot ? (q == (lu & 3)) : (qt * 128 + q * 32) < N;  // warp-uniform: does this warp own live query rows?
      const int ic = i < N ? i : N - 1;
      const int ti = ic / (G::WH * RW), hi = (ic / RW) % G::WH, wi = ic % RW;
      T3Row row;
      {
        const int k = G::NCOPY == 4 ? ((RW - 1 - wi) & 3) : 0;
        const int A = (ti + wg.Wd - 1) * T3_SD + (hi + G::WH - 1) * T3_SH + (RW - 1 - wi);
        row.tb = tab + k * a.CS + (A - k) - (G::WH - 1) * T3_SH;
        row.m = -1e30f;
        row.l = 0.f;
        row.dm = 0u;
        row.hm = 0u;
#pragma unroll
        for (int e = 0; e < 12; ++e) row.mw[e] = 0.f;
      }
      bool need_mask = false;
      if (a.shifted) {
        const int wi_ = win % nW;
        const int wc = wi_ % wg.nww, wb = (wi_ / wg.nww) % wg.nwh, wa = wi_ / (wg.nww * wg.nwh);
        // class boundary along an axis (local position): only the last window of a shifted axis holds two regions
        const int bd = (wg.sd && wa == wg.nwd - 1) ? wg.wd - wg.sd : 64;
        const int bh = (wg.sh && wb == wg.nwh - 1) ? wg.wh - wg.sh : 64;
        const int bw = (wg.sw && wc == wg.nww - 1) ? wg.ww - wg.sw : 64;
        need_mask = (bd < 64) || (bh < 64) || (bw < 64);
        if (need_mask) {
          const bool cd = ti >= bd, chh = hi >= bh, cw = wi >= bw;
#pragma unroll
          for (int e = 0; e < 12; ++e) {
            if (e < 8 && ((e >= bd) != cd)) row.dm |= 1u << e;
            if (e < G::WH && ((e >= bh) != chh)) row.hm |= 1u << e;
            row.mw[e] = (e < RW && ((e >= bw) != cw)) ? T3_MASKV : 0.f;
          }
        }
      }

      for (int c = 0; c < nch; ++c, ++n) {
        const int buf = n & 1;
        T3_PROG(20 + c + tq * 100);
        T3_STAMP(0, n);
        t3_wait(&s_full[g * 2 + buf], (n >> 1) & 1, 7);
        tc_fence_after();
        T3_STAMP(1, n);
        T3_PROG(30 + c + tq * 100);
#ifdef T3_X_NOSOFTMAX
        if (wvalid && n < 0) {
#else
        if (wvalid) {
#endif
          const uint32_t ts_buf = tlane + buf * T3_CW;
          const uint32_t pvp = static_cast<uint32_t>((n - 1) & 1);
          if (need_mask) {
            if (c == 0) t3_chunk<G, true, true>(ts_buf, c, row, false, my_pv, pvp, tmem_o);
            else t3_chunk<G, true, false>(ts_buf, c, row, true, my_pv, pvp, tmem_o);
          } else {
            if (c == 0) t3_chunk<G, false, true>(ts_buf, c, row, false, my_pv, pvp, tmem_o);
            else t3_chunk<G, false, false>(ts_buf, c, row, true, my_pv, pvp, tmem_o);
          }
        }
        T3_PROG(40 + c + tq * 100);
        T3_STAMP(2, n);
        if (c == 0 && have_prev) epilogue();                       // previous tile of this warpgroup: its P.V retired long ago
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        T3_STAMP(3, n);
        T3_PROG(50 + c + tq * 100);
        // announce this warp's P rows to the warpgroup's MMA-issuing warp and move on to the next chunk
        if (lane == 0) mbar_arrive(&p_ready[g * 2 + buf]);
      }
      T3_STAMP(7, n - 1);
      have_prev = true;
      ++ntile_done;
      prev_store = wvalid && i < N;
      prev_wvalid = wvalid;
      prev_m = row.m;
      prev_l = row.l;
      prev_row = static_cast<long long>(win) * N + ic;
      prev_head = head;
      qt += T3_WGS;
      while (qt >= ntiles) {
        qt -= ntiles;
        ++lu;
        if (++win == a.nwin) {
          win = 0;
          ++head;
        }
      }
    }
    T3_PROG(900);
    if (have_prev) epilogue();
    T3_PROG(901);
    while (ev_done < nevents) {
      stage_table(head_first + ev_done);
      ++ev_done;
    }
    T3_PROG(1000);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == T3_TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
template <int RW>
static bool t3_plan_rw(const AttnParams& p, AttnTc3Args& a, int& smem_out) {
  using G = T3G<RW>;
  const WinGeom& g = p.win;
  if (g.Wh != RW || g.Ww != RW || g.wh != RW || g.ww != RW) return false;
  if (g.wd < 1 || g.wd > 8 || g.Wd < g.wd || g.Wd > 8 || g.N != RW * RW * g.wd) return false;
  if ((3 * p.C * 2) % 16 != 0 || p.L != (2 * g.Wd - 1) * G::EH * G::EW) return false;
  a.N = g.N;
  a.nch = g.wd * G::CPF;
  a.ntiles = (g.N + 127) / 128;
  a.tail_rows = g.N - (a.ntiles - 1) * 128;
  a.rot = (a.ntiles >= 2 && a.tail_rows <= 32) ? 1 : 0;
  a.shifted = (g.sd | g.sh | g.sw) != 0;
  a.nb = a.BR = 0;
  if (RW != 7) {          // plain 2-D K / V boxes: the fewest boxes of <= 256 rows that tile N exactly
    for (int nb = (g.N + 255) / 256; nb <= 16; ++nb)
      if (g.N % nb == 0) { a.nb = nb; break; }
    if (a.nb == 0) return false;
    a.BR = g.N / a.nb;
  }
  const int L2 = (2 * g.Wd - 1) * G::EH * G::SH + 32;
  a.CS = L2 + ((8 - L2 % 32) + 32) % 32;          // CS = 8 (mod 32): the four copies start 2 sixteen-byte bank groups apart
  a.kv_bytes = 2 * a.nch * G::CW * 64;
  for (a.nkv = 2; a.nkv >= 1; --a.nkv) {
    int off = a.nkv * a.kv_bytes;
    a.off_q = off;          off += T3_NQ * 128 * 64;
    a.off_tab = off;        off += ((G::NCOPY * a.CS * 4 + 127) / 128) * 128;
    a.off_bar = off;        off += 384;
    smem_out = off + 1024;
    if (smem_out <= 227 * 1024) return true;
  }
  return false;
}

static bool t3_plan(const AttnParams& p, AttnTc3Args& a, int& smem_out) {
  return p.win.Ww == 12 ? t3_plan_rw<12>(p, a, smem_out) : t3_plan_rw<7>(p, a, smem_out);
}

bool window_attn_tc3_supported(const AttnParams& p) {
  AttnTc3Args a;
  int smem = 0;
  return t3_plan(p, a, smem);
}

int window_attn_tc3_dispatch(const AttnParams& p, cudaStream_t st) {
  const WinGeom& g = p.win;
  AttnTc3Args a;
  int smem = 0;
  LAVT_REQUIRE(t3_plan(p, a, smem), "attention(tc3): unsupported window (N=%d, L=%d)", g.N, p.L);
  const long long nwin = 1LL * g.B * g.nwd * g.nwh * g.nww;
  LAVT_REQUIRE(nwin * p.nH < (1LL << 30), "attention(tc3): too many units");
  a.nwin = static_cast<int>(nwin);
  a.units = static_cast<int>(nwin * p.nH);

  CUtensorMap tm_kv, tm_q, tm_tail;
  {
    const uint64_t rowb = static_cast<uint64_t>(3 * p.C) * 2;
    uint64_t dims[2] = {static_cast<uint64_t>(3 * p.C), static_cast<uint64_t>(nwin * g.N)};
    uint64_t strides[1] = {rowb};
    int rc;
    if (g.Ww == 7) {
      // K / V: (channel, w, h, frame) with a box one larger than the 7 x 7 window: TMA zero-fills w = 7 and h = 7
      uint64_t dims4[4] = {static_cast<uint64_t>(3 * p.C), 7, 7, static_cast<uint64_t>(nwin * a.nch)};
      uint64_t str4[3] = {rowb, 7 * rowb, 49 * rowb};
      uint32_t box4[4] = {T3_HD, 8, 8, static_cast<uint32_t>(a.nch)};
      rc = make_tmap_bf16_l2_64b(&tm_kv, p.qkv, 4, dims4, str4, box4, CU_TENSOR_MAP_SWIZZLE_64B);
    } else {
      uint32_t box_kv[2] = {T3_HD, static_cast<uint32_t>(a.BR)};
      rc = make_tmap_bf16_l2_64b(&tm_kv, p.qkv, 2, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_64B);
    }
    if (rc) return rc;
    uint32_t box_q[2] = {T3_HD, 128};
    rc = make_tmap_bf16_l2_64b(&tm_q, p.qkv, 2, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    uint32_t box_t[2] = {T3_HD, 32};
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
  const int ki = g.Ww == 12 ? 1 : 0;
  auto kfn = ki ? window_attn_tc3_kernel<12> : window_attn_tc3_kernel<7>;
  static int configured[2] = {0, 0};
  if (smem > configured[ki]) {
    LAVT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[ki] = smem;
  }
  a.trace = nullptr;
#ifdef T3_TRACE
  const char* trace_path = getenv("LAVT_ATTN_TRACE");
  const size_t trace_n = 16 * 8 * T3_TRACE_ITEMS;
  if (trace_path) {
    LAVT_CUDA(cudaMalloc(&a.trace, trace_n * sizeof(long long)));
    LAVT_CUDA(cudaMemsetAsync(a.trace, 0, trace_n * sizeof(long long), st));
  }
#endif
  LAVT_CUDA(launch_pdl(kfn, dim3(grid), dim3(T3_THREADS), smem, st, 1, tm_kv, tm_q, tm_tail, p, a));
  LAVT_LAUNCH_CHECK("window_attn_tc3_kernel");
#ifdef T3_TRACE
  if (a.trace) {
    // debug only: synchronous dump of CTA 0's event clocks, one line per (warp, event)
    static long long host[16 * 8 * T3_TRACE_ITEMS];
    LAVT_CUDA(cudaStreamSynchronize(st));
    LAVT_CUDA(cudaMemcpy(host, a.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int e = 0; e < 16 * 8; ++e) {
        for (int t = 0; t < T3_TRACE_ITEMS; ++t) fprintf(f, "%lld ", host[e * T3_TRACE_ITEMS + t]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#endif
  return LAVT_OK;
}

}  // namespace lavt
