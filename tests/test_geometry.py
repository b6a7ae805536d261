"""Host index math (lavt_rs_b200/geometry.py, mirrored on the device in csrc/geom.cuh) vs the oracle's closed forms and,
where the reference tree is present, vs the reference's own roll / window_partition / compute_mask /
relative_position_index (bit-exact integer work)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavt_rs_b200.geometry import rel_const, window_geometry, window_row_map  # noqa: E402
from oracle import lavt_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

CASES = [((2, 8, 14, 14), (8, 7, 7), True), ((1, 8, 10, 13), (8, 7, 7), True), ((1, 4, 24, 24), (8, 12, 12), True),
         ((1, 16, 14, 14), (8, 7, 7), True), ((2, 8, 12, 12), (8, 12, 12), False), ((1, 8, 10, 10), (8, 12, 12), True),
         ((1, 16, 30, 17), (8, 12, 12), True), ((3, 1, 15, 30), (1, 12, 12), True)]


@pytest.mark.parametrize("dims,window,shifted", CASES)
def test_row_map_matches_oracle(dims, window, shifted):
    B, D, H, W = dims
    g = window_geometry(B, D, H, W, window, shifted)
    rows, code, rid = window_row_map(g)
    shift = tuple(w // 2 for w in window) if shifted else (0, 0, 0)
    ws, ss = O.effective_window((D, H, W), window, shift)
    assert (g.wd, g.wh, g.ww) == ws and (g.sd, g.sh, g.sw) == ss
    src, valid, ocode, orid = O.window_tokens(D, H, W, ws, ss, window)
    nW, N = src.shape
    assert g.N == N and g.rows() == B * nW * N
    exp = torch.where(valid, src, torch.full_like(src, -1)).reshape(-1)
    for b in range(B):
        got = rows[b * nW * N:(b + 1) * nW * N]
        assert torch.equal(torch.where(got >= 0, got - b * D * H * W, got), exp)
    assert torch.equal(code[: nW * N], ocode.reshape(-1)) and torch.equal(rid[: nW * N], orid.reshape(-1))
    assert rel_const(g) == O.rel_const(window)
    live = rows[rows >= 0]
    assert live.numel() == B * D * H * W and torch.equal(live.sort().values, torch.arange(B * D * H * W))   # bijection


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("dims,window,shifted", CASES[:7])
def test_row_map_matches_reference_ops(dims, window, shifted):
    ref_shims.install_shims()
    import torch.nn.functional as F
    from lib import video_swin_transformer as R     # reference module
    B, D, H, W = dims
    g = window_geometry(B, D, H, W, window, shifted)
    rows, code, rid = window_row_map(g)
    shift = tuple(w // 2 for w in window) if shifted else (0, 0, 0)
    ws, ss = R.get_window_size((D, H, W), window, shift)
    ids = torch.arange(1, B * D * H * W + 1, dtype=torch.float32).reshape(B, D, H, W, 1)     # 0 = pad
    pd, pb, pr = (ws[0] - D % ws[0]) % ws[0], (ws[1] - H % ws[1]) % ws[1], (ws[2] - W % ws[2]) % ws[2]
    xp = F.pad(ids, (0, 0, 0, pr, 0, pb, 0, pd))
    if any(ss):
        xp = torch.roll(xp, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    win = R.window_partition(xp, ws).reshape(-1).long() - 1
    assert torch.equal(win, rows)
    Dp, Hp, Wp = xp.shape[1:4]
    if any(ss):
        mask = R.compute_mask(Dp, Hp, Wp, ws, ss, "cpu")          # (nW, N, N) 0 / -100
        nW, N = mask.shape[:2]
        mine = (rid[: nW * N].reshape(nW, N)[:, :, None] != rid[: nW * N].reshape(nW, N)[:, None, :]).float() * -100.0
        assert torch.equal(mine, mask)
    attn = R.WindowAttention3D(32, window, 1)
    N = g.N
    idx = attn.relative_position_index[:N, :N]
    c = code[:N]
    assert torch.equal(c[:, None] - c[None, :] + rel_const(g), idx)
