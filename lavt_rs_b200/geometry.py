"""Host-side window geometry (the integer part of a Swin block).

Mirrors ``csrc/geom.cuh`` so the host can size buffers and tests can check the device index math.
Reference behaviour restated: ``get_window_size`` (lib/video_swin_transformer.py:70-83), the padded
grid of ``forward_part1`` (:219-225), ``compute_mask`` (:315-328) and the ``relative_position_index``
construction (:107-127).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from ._cabi import WinGeom


def effective_window(size: Sequence[int], window: Sequence[int], shift: Sequence[int]):
    """Clamp the window to the extent; zero the shift on clamped axes."""
    ws, ss = list(window), list(shift)
    for i in range(3):
        if size[i] <= window[i]:
            ws[i] = size[i]
            ss[i] = 0
    return tuple(ws), tuple(ss)


def window_geometry(B: int, D: int, H: int, W: int, window: Sequence[int], shifted: bool,
                    clamp: bool = True) -> WinGeom:
    """Geometry of one block. ``shifted`` selects the odd-block shift (window // 2).

    ``clamp=False`` reproduces the 2-D backbone (lib/backbone.py:205-208), which never clamps the
    window and always pads up to a window multiple.
    """
    shift = tuple(w // 2 for w in window) if shifted else (0, 0, 0)
    if clamp:
        ws, ss = effective_window((D, H, W), window, shift)
    else:
        ws, ss = tuple(window), tuple(shift)
        if window[0] == 1:
            ss = (0, ss[1], ss[2])
    g = WinGeom()
    g.B, g.D, g.H, g.W = B, D, H, W
    g.wd, g.wh, g.ww = ws
    g.sd, g.sh, g.sw = ss
    g.nwd, g.nwh, g.nww = -(-D // ws[0]), -(-H // ws[1]), -(-W // ws[2])
    g.N = ws[0] * ws[1] * ws[2]
    g.Wd, g.Wh, g.Ww = window
    return g


def geom_tuple(g: WinGeom) -> Tuple[int, ...]:
    return tuple(getattr(g, n) for n, _ in WinGeom._fields_)


def window_row_map(g: WinGeom):
    """(rows, code, rid) int64 tensors of length B*nW*N, same contract as ``win_token`` in geom.cuh.

    rows[m] = token row of window-row m in the (B*D*H*W) tensor, or -1 for a pad row.
    """
    nW = g.nwd * g.nwh * g.nww
    m = torch.arange(g.B * nW * g.N, dtype=torch.int64)
    t = m % g.N
    wlin = m // g.N
    wi = wlin % nW
    b = wlin // nW
    c = wi % g.nww
    bb = (wi // g.nww) % g.nwh
    a = wi // (g.nww * g.nwh)
    tw = t % g.ww
    th = (t // g.ww) % g.wh
    td = t // (g.ww * g.wh)
    Dp, Hp, Wp = g.nwd * g.wd, g.nwh * g.wh, g.nww * g.ww
    pd, ph, pw = a * g.wd + td, bb * g.wh + th, c * g.ww + tw
    d, h, w = (pd + g.sd) % Dp, (ph + g.sh) % Hp, (pw + g.sw) % Wp
    ok = (d < g.D) & (h < g.H) & (w < g.W)
    rows = torch.where(ok, ((b * g.D + d) * g.H + h) * g.W + w, torch.full_like(m, -1))
    cw = t % g.Ww
    ch = (t // g.Ww) % g.Wh
    cd = t // (g.Ww * g.Wh)
    code = (cd * (2 * g.Wh - 1) + ch) * (2 * g.Ww - 1) + cw

    def region(p, P, w_, s_):
        if s_ == 0:
            return torch.zeros_like(p)
        return (p >= P - w_).long() + (p >= P - s_).long()

    rid = 9 * region(pd, Dp, g.wd, g.sd) + 3 * region(ph, Hp, g.wh, g.sh) + region(pw, Wp, g.ww, g.sw)
    return rows, code, rid


def rel_const(g: WinGeom) -> int:
    return ((g.Wd - 1) * (2 * g.Wh - 1) + (g.Wh - 1)) * (2 * g.Ww - 1) + (g.Ww - 1)
