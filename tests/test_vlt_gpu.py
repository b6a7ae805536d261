"""VLT fuse-and-classify head (reference lib/vlt.py) on the CUDA path: the glue kernels of csrc/vlt_kernels.cu against plain fp32 PyTorch,
and the whole head against the golden outputs of the UNMODIFIED reference module (oracle/make_golden_vlt.py) and the CPU oracle.
Tolerance: bf16 operands with fp32 accumulation through ~40 layers (two post-norm transformer stacks): logits rel-L2 <= 3e-2."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), torch.as_tensor(np.asarray(b)).float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_vlt_glue_kernels_vs_torch():
    from lavt_rs_b200 import _cabi as K
    g = torch.Generator(device="cuda").manual_seed(5)
    # rows_affine_act with every operand
    B, n, C = 3, 25, 64
    x = torch.randn(B * n, C, device="cuda", generator=g).bfloat16()
    add = torch.randn(B * n, C, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, C, device="cuda", generator=g)
    s = torch.rand(C, device="cuda", generator=g) + 0.5
    t = torch.randn(C, device="cuda", generator=g)
    out = torch.empty(B * n, C, device="cuda", dtype=torch.bfloat16)
    K.rows_affine_act(x, add=add, v=v, rows_per_image=n, s=s, t=t, act=K.ACT_RELU, out_bf16=out)
    ref = torch.relu((x.float() + add.float()).view(B, n, C) * v[:, None] * s + t).view(B * n, C)
    assert rel_l2(out, ref.cpu()) < 5e-3
    o32 = torch.empty(B * n, C, device="cuda")
    K.rows_affine_act(ref, act=K.ACT_SIGMOID, out_f32=o32)
    assert torch.allclose(o32, torch.sigmoid(ref), atol=1e-5)
    # average pool, coordinates, positional table
    img = torch.randn(2, 6, 10, 32, device="cuda", generator=g).bfloat16()
    pooled = torch.empty(2, 3, 5, 32, device="cuda", dtype=torch.bfloat16)
    K.avgpool2_nhwc(img, pooled)
    refp = torch.nn.functional.avg_pool2d(img.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel_l2(pooled, refp.cpu()) < 5e-3
    cc = torch.empty(2, 6, 10, 40, device="cuda", dtype=torch.bfloat16)
    K.append_coords(img, cc)
    assert torch.equal(cc[..., :32], img) and torch.all(cc[..., 38:] == 0)
    xs = (2.0 * torch.arange(10, device="cuda") / 9.0 - 1.0).bfloat16()
    ys = (2.0 * torch.arange(6, device="cuda") / 5.0 - 1.0).bfloat16()
    assert torch.equal(cc[0, 2, :, 32], xs) and torch.equal(cc[1, :, 3, 35], ys) and torch.equal(cc[0, 4, :, 34], xs)
    tab = torch.randn(7, 64, device="cuda", generator=g)
    rows = torch.randn(21, 64, device="cuda", generator=g)
    o32 = torch.empty(21, 64, device="cuda")
    K.rows_add_table(rows, tab, out_f32=o32)
    assert torch.allclose(o32, rows + tab.repeat(3, 1), atol=1e-6)
    # gate * x, transposed to NHWC
    xq = torch.randn(2 * 16, 40, device="cuda", generator=g)
    gate = torch.rand(2 * 16, 32, device="cuda", generator=g)
    om = torch.empty(2, 36, 16, device="cuda", dtype=torch.bfloat16)
    K.gate_transpose(xq[:, :36], gate, om, 2)
    refm = (gate[:, :1] * xq[:, :36]).view(2, 16, 36).permute(0, 2, 1)
    assert rel_l2(om, refm.cpu()) < 5e-3
    # upsample only (align_corners=True)
    up = torch.empty(2, 12, 20, 32, device="cuda", dtype=torch.bfloat16)
    K.upsample_nhwc(img, up)
    refu = torch.nn.functional.interpolate(img.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert rel_l2(up, refu.cpu()) < 5e-3


@pytest.mark.parametrize("B,Lq,S,heads,masked", [(2, 16, 11, 8, True), (1, 900, 900, 8, False), (3, 16, 16, 8, False), (2, 16, 676, 4, False),
                                                 (1, 5, 1024, 2, True)])
def test_mha_small_vs_torch(B, Lq, S, heads, masked):
    from lavt_rs_b200 import _cabi as K
    E = heads * 32
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + S)
    q = torch.randn(B * Lq, E, device="cuda", generator=g).bfloat16()
    kv = torch.randn(B * S, 2 * E, device="cuda", generator=g).bfloat16()
    mask = None
    if masked:
        mask = (torch.rand(B, S, device="cuda", generator=g) > 0.3).float()
        mask[:, 0] = 1
    out = torch.empty(B * Lq, E, device="cuda", dtype=torch.bfloat16)
    K.mha_small(q, kv[:, :E], kv[:, E:], out, B, heads, key_mask=mask)
    qh = q.float().view(B, Lq, heads, 32).permute(0, 2, 1, 3)
    kh = kv[:, :E].float().reshape(B, S, heads, 32).permute(0, 2, 1, 3)
    vh = kv[:, E:].float().reshape(B, S, heads, 32).permute(0, 2, 1, 3)
    sc = qh @ kh.transpose(-1, -2) / math.sqrt(32)
    if mask is not None:
        sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = (sc.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B * Lq, E)
    assert rel_l2(out, ref.cpu()) < 5e-3


@pytest.mark.parametrize("name", ["vlt_head_160", "vlt_head_480"])
def test_vlt_head_matches_reference_golden(name):
    from lavt_rs_b200.lib.vlt import VLTFuseAndClassify
    from oracle import vlt_oracle as VO     # checker
    from oracle.make_golden_vlt import OUT, VLT_CASES, vlt_case
    c = VLT_CASES[name]
    gold = np.load(os.path.join(OUT, name + ".npz"))["logits"]
    args, sd, (c4, c3, c2, l, mask) = vlt_case(c)
    head = VLTFuseAndClassify(d_model=256, nhead=8, d_hid=256, nlayers=2, args=args)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    with torch.no_grad():
        got = head(c4.cuda(), c3.cuda(), c2.cuda(), l.cuda(), mask.cuda())
        ora = VO.vlt_fuse_and_classify(sd, c4, c3, c2, l, mask)
    assert got.shape == gold.shape
    r_gold, r_ora = rel_l2(got, gold), rel_l2(got, ora)
    print(name, "rel-L2 vs reference golden", r_gold, "vs oracle", r_ora)
    assert r_gold < 3e-2 and r_ora < 3e-2, (r_gold, r_ora)
    agree = ((got[:, 1] > got[:, 0]).cpu().numpy() == (gold[:, 1] > gold[:, 0])).mean()
    margin = np.abs(gold[:, 1] - gold[:, 0])
    clear = margin > 4 * np.abs(got.cpu().numpy() - gold).max()
    if clear.any():
        assert ((got[:, 1] > got[:, 0]).cpu().numpy() == (gold[:, 1] > gold[:, 0]))[clear].mean() >= 0.999
    assert agree > 0.97


@pytest.mark.parametrize("model_name", ["lavt_vlt", "vlt"])
def test_vlt_models_end_to_end_480(model_name):
    """``segmentation.lavt_vlt`` / ``segmentation.vlt`` built through the reference's builder API at --img_size 480 (Swin-B, random init):
    text encoder + encoder (stages 1-3) + VLT head + upsample on the CUDA path vs the CPU oracles on the same state dict."""
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib import segmentation
    from oracle import bert_oracle as BO
    from oracle import lavt_oracle as O
    from oracle import vlt_oracle as VO
    torch.manual_seed(3)
    model = segmentation.__dict__[model_name](pretrained="", args=default_args(["--model", model_name, "--swin_type", "base", "--img_size", "480"]))
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in model.classifier.modules():                        # non-trivial eval statistics in the head
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(1.0 + 0.3 * torch.rand(m.running_var.shape, generator=g))
    model = model.eval()
    sd = {k: v.detach().float().clone() for k, v in model.state_dict().items()}
    x = torch.randn(1, 3, 480, 480, generator=g)
    ids = torch.randint(1000, 5000, (1, 20), generator=g)
    mask = torch.zeros(1, 20, dtype=torch.int64)
    mask[:, :13] = 1
    cfg = O.OracleConfig.swin("base", video=False)
    cfg.clamp_window = False
    if model_name == "vlt":
        cfg.version = "swin"
    with torch.no_grad():
        l = BO.bert_forward(sd, ids, mask).permute(0, 2, 1)
        feats = O.backbone_forward(sd, cfg, x, l, mask.float().unsqueeze(-1))
        c2, c3, c4 = feats[-3:]
        ref = torch.nn.functional.interpolate(VO.vlt_fuse_and_classify(sd, c4, c3, c2, l, mask.float().unsqueeze(-1), pre="classifier."),
                                              size=(480, 480), mode="bilinear", align_corners=True)
        model = model.cuda()
        got = model(x.cuda(), ids.cuda(), mask.cuda())
    assert got.shape == ref.shape == (1, 2, 480, 480)
    r = rel_l2(got, ref)
    print(model_name, "logits rel-L2 vs oracle", r)
    assert r < 3e-2, r
