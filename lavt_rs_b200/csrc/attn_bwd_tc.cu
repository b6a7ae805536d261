// tcgen05 / TMEM backward of the shifted-window attention core for 7 x 7 windows (N = 49 * wd tokens, wd in {2, 4, 6, 8}: every window of
// the 8 x 7 x 7 video models).  Same mathematics and ABI as attn_bwd.cu (reference WindowAttention3D.forward,
// lib/video_swin_transformer.py:147-165, differentiated by autograd in the reference's training loop, train.py:330-360):
//     Z = q' k^T + table[idx(i,j)] log2 e + mask(i,j)    P = 2^(Z - lse_i)    (lse saved by the forward kernel)
//     dV = P^T dO     dP = dO V^T     dZ = P o (dP - delta_i)     dQ = dZ K     dK = dZ^T Q     dtable[idx(i,j)] += dZ[i,j]
//
// ONE pass over the N x N matrices, in the TRANSPOSED orientation: TMEM lane = key row, column = query.
//   * S^T = K_tile Q^T and dP^T = V_tile dO^T (SS MMAs, M = 128 keys x N = 64 queries = one frame of queries) land in TMEM;
//   * four column groups of four warps (warp = lane quadrant, 16 query columns = two runs of the frame each) turn them into P^T and
//     dZ^T (bf16, written back over the fp32 columns they came from) -- purely pointwise, no row reduction, so the groups never merge;
//   * dV += P^T dO and dK += dZ^T Q take their A operand straight from TMEM (TS MMAs, B = dO / Q rows as MN-major operands);
//   * dQ += dZ K needs dZ with queries along M: each thread also stores its 16 dZ values as one 32-byte piece of a 128-byte-swizzled
//     MN-major shared-memory tile (row = key = K dimension, 64 queries contiguous); every second chunk one SS MMA (M = 128 queries,
//     K = the tile's keys) accumulates the window's dQ, which stays in TMEM (4 query tiles x 32 columns) across the key tiles.
//   * Q and dO arrive through 4-D TMA boxes (32 ch, 8 of 7 w, 8 of 7 h, wd frames) like K / V of attn_tc3.cu: zero-filled pad columns make
//     every run of 7 queries start at a multiple of 8 and every frame a 64-column chunk, so the bias of a run is 7 consecutive floats at
//     an immediate offset from a per-thread base, and so is its gradient entry.
//   * the table gradient accumulates per unit in fixed point with native integer shared atomics (fp32 shared atomics are CAS loops,
//     attn_bwd.cu); the scale comes from a rigorous bound |dZ| <= 2 max||dO_i|| max||V_j||.
// TMEM (512 columns): dQ 4 x 32 | S^T 2 x 64 | dP^T 2 x 64 | (dK | dV) 2 x 64.
// Warps 0-15: pointwise groups (g = warp / 4, lane quadrant q = warp % 4); warp 16: MMA issue (one thread); warp 17: TMA producer + TMEM.
#include "kernels.cuh"
#include "attn_tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace lavt {

constexpr int BT_HD = 32;
constexpr int BT_NG = 4;
constexpr int BT_SM_THREADS = 128 * BT_NG;
constexpr int BT_MMA_WARP = 4 * BT_NG;
constexpr int BT_TMA_WARP = 4 * BT_NG + 1;
constexpr int BT_THREADS = BT_SM_THREADS + 128;
constexpr int BT_SH = 16, BT_SD = 13 * BT_SH;          // table strides in shared memory: (frame offset, h offset, w offset)
constexpr float BT_LOG2E = 1.4426950408889634f;
constexpr float BT_MASKV = -100.0f * BT_LOG2E;
constexpr int BT_COL_DQ = 0, BT_COL_ST = 128, BT_COL_DP = 256, BT_COL_DKV = 384;

struct AttnBwdTcArgs {
  int N, nch, ntk, nqt;     // tokens per window, query chunks (= frames), 128-row key tiles, 128-row query tiles (= nch / 2)
  int nwin, units;
  int shifted;
  int tab_floats;           // (2 Wd - 1) * BT_SD
  int off_q, off_do, off_k, off_v, off_dz, off_tab, off_itab, off_lse, off_del, off_bar;
  unsigned ld_bytes;        // TMA bytes per unit
  long long* trace;         // -DBT_DEBUG builds: clock64 stamps of CTA 0 ([2 roles][BT_TRACE_N])
  int dbg;                  // -DBT_DEBUG builds: elimination switches (LAVT_BT_DBG bit mask), timing experiments only
};

constexpr int BT_TRACE_N = 512;
#ifdef BT_DEBUG
#define BT_STAMP(role, idx) do { if (blockIdx.x == 0 && a.trace && (idx) < BT_TRACE_N) a.trace[(role) * BT_TRACE_N + (idx)] = clock64(); } while (0)
#else
#define BT_STAMP(role, idx) do { } while (0)
#endif
#ifdef BT_WATCHDOG
__device__ __noinline__ void bt_stuck(int tag, uint32_t parity) {
  printf("[bwd_tc stuck] block %d warp %d lane %d tag %d parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, tag, parity);
  __trap();
}
__device__ __forceinline__ void bt_wait(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000LL) bt_stuck(tag, parity);
}
#else
__device__ __forceinline__ void bt_wait(uint64_t* bar, uint32_t parity, int) { mbar_wait(bar, parity); }
__device__ __forceinline__ void bt_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) {
  }
}
#endif

__device__ __forceinline__ void mul2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}

#ifdef BT_DEBUG
#define BT_DBG(bit) ((dbg & (bit)) != 0)
#else
#define BT_DBG(bit) false
#endif

struct BtKey {              // per-thread state of this thread's key row in the current key tile
  const float* tb;          // bias address of query (frame 0, run 2 g, w 0) for this key
  int* ib;                  // the same entry of the fixed-point gradient table
  float kpen;               // 0, or -1e30 for a lane past the last key of the window
  float mw[8];              // masked windows: w-axis mask of the queries of a run (0 or -100 log2 e)
  uint32_t dm, hm;          // masked windows: bit t_i / h_i set = that query frame / run lies in another region than this key
};

// One chunk (query frame c) for this thread's key row: 16 query columns = runs 2 g and 2 g + 1 of the frame.
//   ts / td : TMEM address of this group's 16 fp32 columns of S^T / dP^T (P^T / dZ^T go back to the first 8 of them as bf16 pairs)
//   nl / nd : -lse / -delta of the 16 queries (shared memory, warp-uniform addresses)
//   dzrow   : this key's 128-byte row in the MN-major dZ tile of the chunk;  r7 = row & 7 (swizzle phase)
template <bool MASKED, bool TAIL>
__device__ __forceinline__ void bt_chunk(uint32_t ts, uint32_t td, int c, int g, const BtKey& k, const float* nl, const float* nd,
                                         uint8_t* dzrow, int r7, float fix, bool do_tab, bool kvalid, int dbg) {
  uint32_t sv[16], dv[16];
  tmem_ld_x16(ts, sv);
  tmem_ld_x16(td, dv);
  const float* fb = k.tb + c * BT_SD;
  int* ib = k.ib + c * BT_SD;
  uint32_t rmask = 0;
  if constexpr (MASKED) rmask = (((k.dm >> c) & 1u) ? 0xffu : k.hm) >> (2 * g);
  uint32_t pw[8], zw[8];
  bool waited = false;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    if (g == BT_NG - 1 && kk == 1) {           // run 7 of the frame is all padding (warp-uniform)
      if (!waited) tmem_ld_wait();
      waited = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) { pw[4 + j] = 0u; zw[4 + j] = 0u; }
      continue;
    }
    float t[8], d[8];
    {
      const float4 l0 = *reinterpret_cast<const float4*>(nl + 8 * kk), l1 = *reinterpret_cast<const float4*>(nl + 8 * kk + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(nd + 8 * kk), d1 = *reinterpret_cast<const float4*>(nd + 8 * kk + 4);
      t[0] = l0.x; t[1] = l0.y; t[2] = l0.z; t[3] = l0.w; t[4] = l1.x; t[5] = l1.y; t[6] = l1.z; t[7] = l1.w;
      d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
    }
    float b[8];
#pragma unroll
    for (int e = 0; e < 7; ++e) b[e] = BT_DBG(32) ? 0.f : fb[kk * BT_SH + e];
    b[7] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; e += 2) add2(t[e], t[e + 1], b[e], b[e + 1]);
    if constexpr (MASKED) {
      const bool rm = (rmask >> kk) & 1u;
#pragma unroll
      for (int e = 0; e < 8; e += 2) add2(t[e], t[e + 1], rm ? BT_MASKV : k.mw[e], rm ? BT_MASKV : k.mw[e + 1]);
    }
    if constexpr (TAIL) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) add2(t[e], t[e + 1], k.kpen, k.kpen);
    }
    if (!waited) tmem_ld_wait();
    waited = true;
    float pr[8], dz[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      pr[e] = __uint_as_float(sv[8 * kk + e]);
      pr[e + 1] = __uint_as_float(sv[8 * kk + e + 1]);
      add2(pr[e], pr[e + 1], t[e], t[e + 1]);
      dz[e] = __uint_as_float(dv[8 * kk + e]);
      dz[e + 1] = __uint_as_float(dv[8 * kk + e + 1]);
      add2(dz[e], dz[e + 1], d[e], d[e + 1]);
    }
#pragma unroll
    for (int e = 0; e < 7; ++e) pr[e] = BT_DBG(4) ? pr[e] : ex2_ftz(pr[e]);
    pr[7] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; e += 2) mul2(dz[e], dz[e + 1], pr[e], pr[e + 1]);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      pw[4 * kk + (e >> 1)] = pack_bf16x2(pr[e], pr[e + 1]);
      zw[4 * kk + (e >> 1)] = pack_bf16x2(dz[e], dz[e + 1]);
    }
    if (do_tab && (!TAIL || kvalid)) {
#pragma unroll
      for (int e = 0; e < 7; ++e) atomicAdd(ib + kk * BT_SH + e, __float2int_rn(dz[e] * fix));
    }
  }
  if (!BT_DBG(8)) {
    tmem_st_x8(ts, pw);
    tmem_st_x8(td, zw);
  }
  if (!BT_DBG(16)) {
    *reinterpret_cast<uint4*>(dzrow + (((2 * g) ^ r7) << 4)) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
    *reinterpret_cast<uint4*>(dzrow + (((2 * g + 1) ^ r7) << 4)) = make_uint4(zw[4], zw[5], zw[6], zw[7]);
  }
}

__global__ void __launch_bounds__(BT_THREADS, 1)
window_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                          const __grid_constant__ CUtensorMap tmKV, const AttnBwdParams p, const AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* tab = reinterpret_cast<float*>(smem + a.off_tab);
  int* itab = reinterpret_cast<int*>(smem + a.off_itab);
  float* nlse_s = reinterpret_cast<float*>(smem + a.off_lse);
  float* ndel_s = reinterpret_cast<float*>(smem + a.off_del);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* ld_full = bars;            // Q, dO, K, V of the unit landed
  uint64_t* ops_free = bars + 1;       // every MMA of the unit retired: the operands may be overwritten
  uint64_t* s_full = bars + 2;         // [2] S^T and dP^T of an item are in TMEM
  uint64_t* p_ready = bars + 4;        // [2] the 16 pointwise warps wrote P^T / dZ^T (TMEM) and dZ (shared memory) of an item
  uint64_t* dz_free = bars + 6;        // [2] the dQ MMA that read this pair of dZ chunks retired
  uint64_t* dkv_full = bars + 8;       // [2] dK | dV of a key tile are complete
  uint64_t* dkv_free = bars + 10;      // [2] ... and have been read out
  uint64_t* dq_full = bars + 12;       // dQ of the unit is complete
  uint64_t* dq_free = bars + 13;       // ... and has been read out
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);
  int* umax = reinterpret_cast<int*>(bars + 15);          // [2][2] bit patterns of max ||dO_i||^2, max ||V_j||^2 per unit parity

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, nch = a.nch, ntk = a.ntk;
  const int NT = ntk * nch;                                // (key tile, chunk) items per unit
  const WinGeom& wg = p.win;
  const int u_begin = static_cast<int>(1LL * a.units * blockIdx.x / gridDim.x);
  const int u_end = static_cast<int>(1LL * a.units * (blockIdx.x + 1) / gridDim.x);
  const int nunits = u_end - u_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmKV);
    mbar_init(ld_full, 1);
    mbar_init(ops_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4 * BT_NG);
      mbar_init(&dz_free[i], 1);
      mbar_init(&dkv_full[i], 1);
      mbar_init(&dkv_free[i], 4 * BT_NG);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4 * BT_NG);
    umax[0] = umax[1] = umax[2] = umax[3] = 0;
    fence_mbar_init();
  }
  if (warp == BT_TMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  for (int i = threadIdx.x; i < a.tab_floats + 32; i += blockDim.x) {
    tab[i] = 0.f;
    itab[i] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Register split (the launch grants 96 per thread): the control warpgroup keeps 56, the pointwise warpgroups grow to 104 (only registers the CTA itself released can be re-granted: 128 x 40 >= 512 x 8); each
  // setmaxnreg sits inside its role's branch so that it dominates the role's code (ptxas budgets a region by the setmaxnreg that dominates it)
  if (warp >= 4 * BT_NG) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == BT_TMA_WARP) {
    // =============================== TMA producer (one thread) ===============================
    if (lane == 0) {
      for (int lu = 0; lu < nunits; ++lu) {
        const int u = u_begin + lu;
        const int head = u / a.nwin, win = u - head * a.nwin;
        if (lu >= 1) bt_wait(ops_free, (lu - 1) & 1, 1);
        mbar_expect_tx(ld_full, a.ld_bytes);
        tma_load_4d(smem + a.off_q, &tmQ, ld_full, head * BT_HD, 0, 0, win * nch);
        tma_load_4d(smem + a.off_do, &tmDO, ld_full, head * BT_HD, 0, 0, win * nch);
        for (int j = 0; j < ntk; ++j) {
          tma_load_2d(smem + a.off_k + j * 8192, &tmKV, ld_full, p.C + head * BT_HD, win * N + j * 128);
          tma_load_2d(smem + a.off_v + j * 8192, &tmKV, ld_full, 2 * p.C + head * BT_HD, win * N + j * 128);
        }
      }
    }
  } else if (warp == BT_MMA_WARP) {
    // =============================== MMA issue (one thread; its tcgen05.mma execute in issue order) ===============================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16_f32(128, 64);
      const uint32_t idesc_kv = make_idesc_bf16_f32(128, BT_HD) | (1u << 16);                 // B (dO / Q rows) MN-major
      const uint32_t idesc_dq = make_idesc_bf16_f32(128, BT_HD) | (1u << 15) | (1u << 16);    // A (dZ tile) and B (K rows) MN-major
      const uint32_t sbase = smem_u32(smem);
      const uint32_t q_base = sbase + a.off_q, do_base = sbase + a.off_do, k_base = sbase + a.off_k, v_base = sbase + a.off_v;
      const uint32_t dz_base = sbase + a.off_dz;
      int n = 0, kt = 0;
      for (int lu = 0; lu < nunits; ++lu) {
        int tr = lu * 40;
        BT_STAMP(0, tr); ++tr;
        bt_wait(ld_full, lu & 1, 2);
        tc_fence_after();
        BT_STAMP(0, tr); ++tr;
        auto issue_sdp = [&](int nl, int ng) {
          const int j = nl / nch, c = nl - j * nch, buf = ng & 1;
          const uint64_t dk = make_sw64_desc(k_base + j * 8192), dv = make_sw64_desc(v_base + j * 8192);
          const uint64_t dq = make_sw64_desc(q_base + c * 4096), dd = make_sw64_desc(do_base + c * 4096);
          const uint32_t ts = tmem_base + BT_COL_ST + buf * 64, td = tmem_base + BT_COL_DP + buf * 64;
#ifdef BT_DEBUG
          if (!(a.dbg & 256)) {
#endif
          umma_bf16_ss(ts, dk, dq, idesc_s, 0);
          umma_bf16_ss(ts, dk + 2, dq + 2, idesc_s, 1);
          umma_bf16_ss(td, dv, dd, idesc_s, 0);
          umma_bf16_ss(td, dv + 2, dd + 2, idesc_s, 1);
#ifdef BT_DEBUG
          }
#endif
          umma_commit(&s_full[buf]);
        };
        issue_sdp(0, n);
        if (NT > 1) issue_sdp(1, n + 1);
        for (int nl = 0; nl < NT; ++nl, ++n) {
          const int j = nl / nch, c = nl - j * nch, buf = n & 1;
#ifdef BT_DEBUG
          if (a.dbg & 512) bt_spin(&p_ready[buf], (n >> 1) & 1); else
#endif
          bt_wait(&p_ready[buf], (n >> 1) & 1, 3);
          BT_STAMP(0, tr); ++tr;
          if (c == 0 && kt >= 2) bt_wait(&dkv_free[kt & 1], ((kt >> 1) - 1) & 1, 4);
          tc_fence_after();
          {
            const uint32_t tdk = tmem_base + BT_COL_DKV + (kt & 1) * 64, tdv = tdk + BT_HD;
            const uint32_t tp = tmem_base + BT_COL_ST + buf * 64, tz = tmem_base + BT_COL_DP + buf * 64;
            const uint64_t bq = make_sw64_desc(q_base + c * 4096), bd = make_sw64_desc(do_base + c * 4096);
#ifdef BT_DEBUG
            if (!(a.dbg & 128)) {
#endif
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(tdv, tp + 16 * ks, bd + 64 * ks, idesc_kv, (c > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16_ts(tdk, tz + 16 * ks, bq + 64 * ks, idesc_kv, (c > 0 || ks > 0) ? 1u : 0u);
#ifdef BT_DEBUG
            }
#endif
          }
          if (c & 1) {
            if (nl == 1 && lu >= 1) {
              bt_wait(dq_free, (lu - 1) & 1, 5);
              tc_fence_after();
            }
            const int slot = (n >> 1) & 1;
            const int nk = min(128, N - j * 128), nks = (nk + 15) >> 4;
            const uint64_t da = make_mnmajor_sw128_desc(dz_base + slot * 32768, 16384);
            const uint64_t db = make_sw64_desc(k_base + j * 8192);
            const uint32_t tq = tmem_base + BT_COL_DQ + (c >> 1) * BT_HD;
#ifdef BT_DEBUG
            if (!(a.dbg & 64))
#endif
            for (int ks = 0; ks < nks; ++ks) umma_bf16_ss(tq, da + 128 * ks, db + 64 * ks, idesc_dq, (j > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&dz_free[slot]);
          }
          if (c == nch - 1) {
            umma_commit(&dkv_full[kt & 1]);
            ++kt;
          }
          if (nl == NT - 1) {
            umma_commit(dq_full);
            umma_commit(ops_free);
          }
          if (nl + 2 < NT) issue_sdp(nl + 2, n + 2);
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // =============================== pointwise warps ===============================
    const int g = warp >> 2, q = warp & 3, r = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nW = wg.nwd * wg.nwh * wg.nww;
    const bool do_tab = p.dtable_t != nullptr;
    const int tid = threadIdx.x;                        // 0 .. 511
    int cur_head = -1;
    int n = 0, kt = 0;
    float prev_inv_fix = 0.f;
    int prev_head = 0;

    auto drain_dkv = [&](int ktile, int j, long long row0, int head) {
      bt_wait(&dkv_full[ktile & 1], (ktile >> 1) & 1, 6);
      tc_fence_after();
      if (j * 128 + q * 32 < N) {                        // warp-uniform: this warp holds live key rows
        uint32_t v[16];
        tmem_ld_x16(tlane + BT_COL_DKV + (ktile & 1) * 64 + 16 * g, v);
        tmem_ld_wait();
        const int kr = j * 128 + r;
        if (kr < N) {
          // dy_k = dZ^T q' / log2 e  (q' = y_q hd^-0.5 log2 e),  dy_v = P^T dO
          const float sc = g < 2 ? (1.0f / BT_LOG2E) : 1.0f;
          __nv_bfloat16* dst = p.dqkv + (row0 + kr) * (3 * p.C) + (g < 2 ? p.C : 2 * p.C) + head * BT_HD + (g & 1) * 16;
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = pack_bf16x2(__uint_as_float(v[2 * e]) * sc, __uint_as_float(v[2 * e + 1]) * sc);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w8[0]), "r"(w8[1]), "r"(w8[2]),
                       "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dkv_free[ktile & 1]);
    };

    for (int lu = 0; lu < nunits; ++lu) {
      const int u = u_begin + lu;
      const int head = u / a.nwin, win = u - head * a.nwin;
      const long long row0 = static_cast<long long>(win) * N;
      int tr = lu * 80;
      if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
      // ---- unit prologue: fold the previous unit's table gradient, (re)stage the bias table, -lse / -delta of every query, scale ----
      if (lu >= 1 && do_tab) {
        for (int pos = tid; pos < a.tab_floats; pos += BT_SM_THREADS) {
          const int co = pos & (BT_SH - 1);
          const int v = itab[pos];
          if (co < 13 && v != 0) {
            atomicAdd(p.dtable_t + static_cast<long long>(prev_head) * p.L + (pos >> 4) * 13 + co, static_cast<float>(v) * prev_inv_fix);
            itab[pos] = 0;
          }
        }
      }
      if (head != cur_head) {
        const float* src = p.table_t + static_cast<long long>(head) * p.L;
        for (int pos = tid; pos < a.tab_floats; pos += BT_SM_THREADS) {
          const int co = pos & (BT_SH - 1);
          tab[pos] = co < 13 ? __ldg(src + (pos >> 4) * 13 + co) * BT_LOG2E : 0.f;
        }
        cur_head = head;
      }
      {
        const int c = tid >> 6, hi = (tid >> 3) & 7, wi = tid & 7;
        const bool qvalid = c < nch && hi < 7 && wi < 7;
        float nl = 0.f, ndl = 0.f, ndo = 0.f, nv = 0.f;
        if (qvalid) {
          const long long row = row0 + c * 49 + hi * 7 + wi;
          const uint4* o4 = reinterpret_cast<const uint4*>(p.out + row * p.C + head * BT_HD);
          const uint4* d4 = reinterpret_cast<const uint4*>(p.dout + row * p.C + head * BT_HD);
          float acc = 0.f;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 x = __ldg(o4 + ch), y = __ldg(d4 + ch);
            const uint32_t xx[4] = {x.x, x.y, x.z, x.w}, yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 xo = unpack_bf16x2(xx[e]), yd = unpack_bf16x2(yy[e]);
              acc += xo.x * yd.x + xo.y * yd.y;
              ndo += yd.x * yd.x + yd.y * yd.y;
            }
          }
          nl = -__ldg(p.lse + row * p.nH + head);
          ndl = -acc;
        }
        if (tid < N) {
          const uint4* v4 = reinterpret_cast<const uint4*>(p.qkv + (row0 + tid) * (3 * p.C) + 2 * p.C + head * BT_HD);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 x = __ldg(v4 + ch);
            const uint32_t xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 xv = unpack_bf16x2(xx[e]);
              nv += xv.x * xv.x + xv.y * xv.y;
            }
          }
        }
        nlse_s[tid] = nl;
        ndel_s[tid] = ndl;
        ndo = warp_max(ndo);
        nv = warp_max(nv);
        if (lane == 0) {                       // non-negative floats order like their bit patterns
          atomicMax(&umax[(lu & 1) * 2], __float_as_int(ndo));
          atomicMax(&umax[(lu & 1) * 2 + 1], __float_as_int(nv));
        }
        if (tid == 0) umax[((lu + 1) & 1) * 2] = umax[((lu + 1) & 1) * 2 + 1] = 0;
      }
      named_bar(2, BT_SM_THREADS);
      if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
      // |dZ| <= |dP| + |delta| <= 2 max||dO_i|| max||V_j|| (Cauchy-Schwarz; O is a convex combination of V rows), <= N terms per entry:
      // scale = 2^30 / (2.5 N bound) cannot overflow int32
      const float bound = 2.5f * static_cast<float>(N) * sqrtf(__int_as_float(umax[(lu & 1) * 2]) * __int_as_float(umax[(lu & 1) * 2 + 1]));
      const float fix = (bound > 0.f && isfinite(bound)) ? 1073741824.0f / bound : 1.0f;

      // ---- shifted windows: class boundaries of this window (only the last window of a shifted axis holds two regions) ----
      bool need_mask = false;
      int bd = 64, bh = 64, bw = 64;
      if (a.shifted) {
        const int wi_ = win % nW;
        const int wc = wi_ % wg.nww, wb = (wi_ / wg.nww) % wg.nwh, wa = wi_ / (wg.nww * wg.nwh);
        bd = (wg.sd && wa == wg.nwd - 1) ? wg.wd - wg.sd : 64;
        bh = (wg.sh && wb == wg.nwh - 1) ? wg.wh - wg.sh : 64;
        bw = (wg.sw && wc == wg.nww - 1) ? wg.ww - wg.sw : 64;
        need_mask = (bd < 64) || (bh < 64) || (bw < 64);
      }

      for (int j = 0; j < ntk; ++j, ++kt) {
        const int kr = j * 128 + r;
        const bool kvalid = kr < N;
        const bool wvalid = j * 128 + q * 32 < N;           // warp-uniform
        const bool tail = (j + 1) * 128 > N;
        const int krc = kvalid ? kr : N - 1;
        const int tj = krc / 49, hj = (krc - tj * 49) / 7, wj = krc - tj * 49 - hj * 7;
        BtKey key;
        {
          const int off = (wg.Wd - 1 - tj) * BT_SD + (6 - hj + 2 * g) * BT_SH + (6 - wj);
          key.tb = tab + off;
          key.ib = itab + off;
          key.kpen = kvalid ? 0.f : -1e30f;
          key.dm = 0u;
          key.hm = 0u;
#pragma unroll
          for (int e = 0; e < 8; ++e) key.mw[e] = 0.f;
          if (need_mask) {
            const bool cd = tj >= bd, chh = hj >= bh, cw = wj >= bw;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if ((e >= bd) != cd) key.dm |= 1u << e;
              if (e < 7 && ((e >= bh) != chh)) key.hm |= 1u << e;
              key.mw[e] = (e < 7 && ((e >= bw) != cw)) ? BT_MASKV : 0.f;
            }
          }
        }
        for (int c = 0; c < nch; ++c, ++n) {
          const int buf = n & 1;
#ifdef BT_DEBUG
          if (a.dbg & 1024) {
            if (lane == 0) bt_spin(&s_full[buf], (n >> 1) & 1);
            __syncwarp();
          } else if (a.dbg & 2048) {
            if (lane == 0) bt_wait(&s_full[buf], (n >> 1) & 1, 7);
            __syncwarp();
          } else
#endif
          bt_wait(&s_full[buf], (n >> 1) & 1, 7);
          if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
          if ((n & 1) == 0 && (n >> 1) >= 2) bt_wait(&dz_free[(n >> 1) & 1], (((n >> 1) >> 1) - 1) & 1, 8);
          tc_fence_after();
#ifdef BT_DEBUG
          if (wvalid && !(a.dbg & 1)) {
#else
          if (wvalid) {
#endif
            const uint32_t ts = tlane + BT_COL_ST + buf * 64 + 16 * g, td = tlane + BT_COL_DP + buf * 64 + 16 * g;
            uint8_t* dzrow = smem + a.off_dz + ((n >> 1) & 1) * 32768 + (c & 1) * 16384 + r * 128;
            const float* nl = nlse_s + c * 64 + 16 * g;
            const float* nd = ndel_s + c * 64 + 16 * g;
            if (need_mask) {
              if (tail) bt_chunk<true, true>(ts, td, c, g, key, nl, nd, dzrow, r & 7, fix, do_tab, kvalid, a.dbg);
              else bt_chunk<true, false>(ts, td, c, g, key, nl, nd, dzrow, r & 7, fix, do_tab, kvalid, a.dbg);
            } else {
              if (tail) bt_chunk<false, true>(ts, td, c, g, key, nl, nd, dzrow, r & 7, fix, do_tab, kvalid, a.dbg);
              else bt_chunk<false, false>(ts, td, c, g, key, nl, nd, dzrow, r & 7, fix, do_tab, kvalid, a.dbg);
            }
          }
          tmem_st_wait();
          tc_fence_before();
#ifdef BT_DEBUG
          if (!(a.dbg & 2))
#endif
          fence_proxy_async_smem();                          // the dZ rows are read by the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[buf]);
          if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
          if (c == 0 && j > 0) drain_dkv(kt - 1, j - 1, row0, head);    // previous key tile: its last MMAs retired long ago
        }
      }
      // ---- unit epilogue: last key tile, dQ, and everybody's atomics before the table is folded ----
      if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
      drain_dkv(kt - 1, ntk - 1, row0, head);
      bt_wait(dq_full, lu & 1, 9);
      if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
      tc_fence_after();
      if (g < a.nqt) {
        uint32_t v[32];
        tmem_ld_x32(tlane + BT_COL_DQ + g * BT_HD, v);
        tmem_ld_wait();
        const int c = 2 * g + (r >> 6), hi = (r >> 3) & 7, wi = r & 7;
        if (hi < 7 && wi < 7) {
          __nv_bfloat16* dst = p.dqkv + (row0 + c * 49 + hi * 7 + wi) * (3 * p.C) + head * BT_HD;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              w8[e] = pack_bf16x2(__uint_as_float(v[h * 16 + 2 * e]) * p.qscale, __uint_as_float(v[h * 16 + 2 * e + 1]) * p.qscale);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + h * 16), "r"(w8[0]), "r"(w8[1]),
                         "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                         : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      prev_inv_fix = 1.0f / fix;
      prev_head = head;
      named_bar(1, BT_SM_THREADS);
      if (threadIdx.x == 0) { BT_STAMP(1, tr); } ++tr;
    }
    if (nunits > 0 && do_tab) {
      for (int pos = tid; pos < a.tab_floats; pos += BT_SM_THREADS) {
        const int co = pos & (BT_SH - 1);
        const int v = itab[pos];
        if (co < 13 && v != 0)
          atomicAdd(p.dtable_t + static_cast<long long>(prev_head) * p.L + (pos >> 4) * 13 + co, static_cast<float>(v) * prev_inv_fix);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BT_TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool bt_plan(const AttnBwdParams& p, AttnBwdTcArgs& a, int& smem_out) {
  const WinGeom& g = p.win;
  if (g.Wh != 7 || g.Ww != 7 || g.wh != 7 || g.ww != 7) return false;
  if (g.wd < 2 || g.wd > 8 || (g.wd & 1) || g.Wd < g.wd || g.Wd > 8 || g.N != 49 * g.wd) return false;
  if (p.C != p.nH * BT_HD || p.L != (2 * g.Wd - 1) * 169 || p.lse == nullptr) return false;
  a.N = g.N;
  a.nch = g.wd;
  a.ntk = (g.N + 127) / 128;
  a.nqt = g.wd / 2;
  a.shifted = (g.sd | g.sh | g.sw) != 0;
  a.tab_floats = (2 * g.Wd - 1) * BT_SD;
  int off = 0;
  a.off_q = off;     off += a.nch * 4096;
  a.off_do = off;    off += a.nch * 4096;
  a.off_k = off;     off += a.ntk * 8192;
  a.off_v = off;     off += a.ntk * 8192;
  off = (off + 1023) / 1024 * 1024;
  a.off_dz = off;    off += 65536;
  const int tab_bytes = ((a.tab_floats + 32) * 4 + 127) / 128 * 128;
  a.off_tab = off;   off += tab_bytes;
  a.off_itab = off;  off += tab_bytes;
  a.off_lse = off;   off += 512 * 4;
  a.off_del = off;   off += 512 * 4;
  a.off_bar = off;   off += 256;
  a.ld_bytes = static_cast<unsigned>(2 * a.nch * 4096 + 2 * a.ntk * 8192);
  smem_out = off + 1024;
  return smem_out <= 227 * 1024;
}

bool window_attn_bwd_tc_supported(const AttnBwdParams& p) {
  AttnBwdTcArgs a;
  int smem = 0;
  return bt_plan(p, a, smem);
}

int window_attn_bwd_tc_dispatch(const AttnBwdParams& p, cudaStream_t st) {
  const WinGeom& g = p.win;
  AttnBwdTcArgs a;
  int smem = 0;
  LAVT_REQUIRE(bt_plan(p, a, smem), "attention backward (tcgen05): unsupported window (N=%d, L=%d)", g.N, p.L);
  const long long nwin = 1LL * g.B * g.nwd * g.nwh * g.nww;
  LAVT_REQUIRE(nwin * p.nH < (1LL << 30), "attention backward (tcgen05): too many units");
  a.nwin = static_cast<int>(nwin);
  a.units = static_cast<int>(nwin * p.nH);
  a.dbg = 0;
  a.trace = nullptr;
#ifdef BT_DEBUG
  if (const char* e = getenv("LAVT_BT_DBG")) a.dbg = atoi(e);
  const char* trace_path = getenv("LAVT_BT_TRACE");
  if (trace_path) {
    LAVT_CUDA(cudaMalloc(&a.trace, 2 * BT_TRACE_N * sizeof(long long)));
    LAVT_CUDA(cudaMemsetAsync(a.trace, 0, 2 * BT_TRACE_N * sizeof(long long), st));
  }
#endif

  CUtensorMap tm_q, tm_do, tm_kv;
  {
    // Q and dO: (channel, w, h, frame) with a box one larger than the 7 x 7 window: TMA zero-fills w = 7 and h = 7
    const uint64_t rowb = static_cast<uint64_t>(3 * p.C) * 2, rowd = static_cast<uint64_t>(p.C) * 2;
    uint64_t dq[4] = {static_cast<uint64_t>(3 * p.C), 7, 7, static_cast<uint64_t>(nwin * a.nch)};
    uint64_t sq[3] = {rowb, 7 * rowb, 49 * rowb};
    uint64_t dd[4] = {static_cast<uint64_t>(p.C), 7, 7, static_cast<uint64_t>(nwin * a.nch)};
    uint64_t sd[3] = {rowd, 7 * rowd, 49 * rowd};
    uint32_t box4[4] = {BT_HD, 8, 8, static_cast<uint32_t>(a.nch)};
    int rc = make_tmap_bf16_l2_64b(&tm_q, p.qkv, 4, dq, sq, box4, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_tmap_bf16_l2_64b(&tm_do, p.dout, 4, dd, sd, box4, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    uint64_t dims[2] = {static_cast<uint64_t>(3 * p.C), static_cast<uint64_t>(nwin * g.N)};
    uint64_t strides[1] = {rowb};
    uint32_t box_kv[2] = {BT_HD, 128};
    rc = make_tmap_bf16_l2_64b(&tm_kv, p.qkv, 2, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_64B);
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
    LAVT_CUDA(cudaFuncSetAttribute(window_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  window_attn_bwd_tc_kernel<<<grid, BT_THREADS, smem, st>>>(tm_q, tm_do, tm_kv, p, a);
  LAVT_LAUNCH_CHECK("window_attn_bwd_tc_kernel");
#ifdef BT_DEBUG
  if (a.trace) {            // debug only: synchronous dump of CTA 0's event clocks, one line per role
    static long long host[2 * BT_TRACE_N];
    LAVT_CUDA(cudaStreamSynchronize(st));
    LAVT_CUDA(cudaMemcpy(host, a.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 2; ++r) {
        for (int t = 0; t < BT_TRACE_N; ++t) fprintf(f, "%lld ", host[r * BT_TRACE_N + t]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#endif
  return LAVT_OK;
}

}  // namespace lavt
