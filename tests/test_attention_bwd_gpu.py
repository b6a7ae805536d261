"""Window-attention backward (C-ABI lavt_window_attention_bwd) in isolation: the tcgen05 / TMEM kernel (attn_bwd_tc.cu: 7 x 7 windows with
an even number of frames) and the mma.sync kernel (attn_bwd.cu) against float64 autograd through a plain PyTorch evaluation of the same op
on the same bf16 q / k / v / dO (reference WindowAttention3D.forward, lib/video_swin_transformer.py:147-165, differentiated by autograd in
train.py:330-360).  Tolerance: rel-L2 <= 1e-2 per gradient (bf16 operands and outputs, fp32 accumulation; measured 2e-3 .. 3e-3)."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("run_attn_bwd", os.path.join(ROOT, "tools", "run_attn_bwd.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


CASES = [  # (B, D, H, W), shifted, heads: N = 392 / 196 / 98 / 294, masked and unmasked windows, one unit, many units and heads per CTA
    ((1, 8, 14, 14), True, 4), ((1, 4, 14, 14), True, 4), ((1, 2, 14, 14), True, 4), ((1, 8, 7, 7), False, 1), ((3, 8, 14, 21), False, 4),
    ((4, 8, 24, 24), True, 16), ((1, 6, 21, 14), True, 8), ((2, 16, 14, 14), True, 4),
]


@pytest.mark.parametrize("impl", ["tc", "mma"])
@pytest.mark.parametrize("dims,shifted,nH", CASES)
def test_window_attention_bwd_matches_autograd(dims, shifted, nH, impl):
    t = _tool()
    geom, qkv, table, dout, out, lse = t.setup(dims, shifted, nH)
    C = nH * 32
    ref_q, ref_t = t.reference(geom, qkv, table, dout)
    dqkv, dtab = t.run(impl, geom, qkv, table, dout, out, lse)
    errs = {"dq": t.rel(dqkv[:, :C], ref_q[:, :C]), "dk": t.rel(dqkv[:, C:2 * C], ref_q[:, C:2 * C]),
            "dv": t.rel(dqkv[:, 2 * C:], ref_q[:, 2 * C:]), "dtable": t.rel(dtab, ref_t)}
    assert all(e < 1e-2 for e in errs.values()), f"N={geom.N} {impl}: {errs}"


def test_window_attention_bwd_tc_accumulates_table_gradient():
    """dtable_t accumulates (+=) across calls and the two kernels agree on it."""
    t = _tool()
    geom, qkv, table, dout, out, lse = t.setup((2, 8, 14, 14), True, 4)
    _, d1 = t.run("tc", geom, qkv, table, dout, out, lse)
    _, d2 = t.run("mma", geom, qkv, table, dout, out, lse)
    assert t.rel(d1, d2) < 5e-3
    from lavt_rs_b200 import _cabi as K
    dqkv = torch.zeros_like(qkv)
    acc = d1.clone()
    K.window_attention_bwd(qkv, out, dout, table.t().contiguous(), geom, dqkv, acc, lse=lse)
    torch.cuda.synchronize()
    assert t.rel(acc, 2 * d1) < 1e-3
