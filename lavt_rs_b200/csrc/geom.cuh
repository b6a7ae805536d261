// Index math shared by the gather (LN -> windows) and scatter (proj epilogue) sides of a
// Swin block, so that cyclic shift + window partition/reverse never materialise a copy.
//
// Restates reference lib/video_swin_transformer.py:214-248 (forward_part1: pad, torch.roll,
// window_partition / window_reverse, un-roll, crop) and :70-83 (get_window_size) as a
// closed-form map  window-row m  ->  token row of the (B,D,H,W,C) tensor  (or "pad").
#pragma once
#include <stdint.h>

namespace lavt {

struct WinGeom {
  int B, D, H, W;        // token grid (channels-last tensor is (B,D,H,W,C))
  int wd, wh, ww;        // effective (clamped) window
  int sd, sh, sw;        // effective shift (0 where clamped / even blocks)
  int nwd, nwh, nww;     // windows per axis on the padded grid
  int N;                 // tokens per window = wd*wh*ww
  int Wd, Wh, Ww;        // CONFIGURED window (relative-position table geometry)
};

#if defined(__CUDACC__)
#define LAVT_HD __host__ __device__ __forceinline__
#else
#define LAVT_HD inline
#endif

// region id along one axis on the SHIFTED padded coordinate p in [0,P)
// (reference compute_mask, lib/video_swin_transformer.py:315-328)
LAVT_HD int shift_region(int p, int P, int w, int s) {
  if (s == 0) return 0;
  return (p >= P - w ? 1 : 0) + (p >= P - s ? 1 : 0);
}

struct WinTok {
  long long row;   // token row in the (B*D*H*W) tensor, or -1 for a pad row
  int code;        // linearised (d,h,w) inside the CONFIGURED window for the rel-pos index
  int rid;         // shifted-window region id (0..26)
};

// m = ((b*nW + window) * N + t)
LAVT_HD WinTok win_token(const WinGeom& g, long long m) {
  WinTok o;
  int t = static_cast<int>(m % g.N);
  long long wlin = m / g.N;
  int nW = g.nwd * g.nwh * g.nww;
  int wi = static_cast<int>(wlin % nW);
  int b = static_cast<int>(wlin / nW);
  int c = wi % g.nww;
  int bb = (wi / g.nww) % g.nwh;
  int a = wi / (g.nww * g.nwh);
  int tw = t % g.ww;
  int th = (t / g.ww) % g.wh;
  int td = t / (g.ww * g.wh);
  int Dp = g.nwd * g.wd, Hp = g.nwh * g.wh, Wp = g.nww * g.ww;
  int pd = a * g.wd + td, ph = bb * g.wh + th, pw = c * g.ww + tw;
  int d = pd + g.sd; if (d >= Dp) d -= Dp;
  int h = ph + g.sh; if (h >= Hp) h -= Hp;
  int w = pw + g.sw; if (w >= Wp) w -= Wp;
  o.row = (d < g.D && h < g.H && w < g.W)
              ? ((static_cast<long long>(b) * g.D + d) * g.H + h) * g.W + w
              : -1;
  // relative_position_index[:N,:N] slices the index built for the CONFIGURED window, so token
  // id t is unravelled on (Wd,Wh,Ww) -- reference :150 (differs from (td,th,tw) only when the
  // window was clamped along H or W).
  int cw = t % g.Ww;
  int ch = (t / g.Ww) % g.Wh;
  int cd = t / (g.Ww * g.Wh);
  o.code = (cd * (2 * g.Wh - 1) + ch) * (2 * g.Ww - 1) + cw;
  o.rid = 9 * shift_region(pd, Dp, g.wd, g.sd) + 3 * shift_region(ph, Hp, g.wh, g.sh) +
          shift_region(pw, Wp, g.ww, g.sw);
  return o;
}

// idx(i,j) = code(i) - code(j) + rel_const  (reference :113-126)
LAVT_HD int rel_const(const WinGeom& g) {
  return ((g.Wd - 1) * (2 * g.Wh - 1) + (g.Wh - 1)) * (2 * g.Ww - 1) + (g.Ww - 1);
}

}  // namespace lavt
