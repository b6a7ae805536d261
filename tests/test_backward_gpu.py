"""Backward-pass parity: gradients from the hand-written sm_100a adjoint kernels (lavt_rs_b200.train_engine -> C ABI) against
autograd through the CPU oracle (oracle/lavt_oracle.py, pinned against the reference's forward; the reference itself trains with
autograd over the same op sequence, train.py:330-360) on identical seeded weights / inputs / output gradients.

Tolerance: operands and saved activations are bf16 with fp32 accumulation, so gradients are held to rel-L2 <= 3e-2 per tensor
(north_star: rtol 2e-2 per layer in bf16; gradients pass through roughly twice as many bf16 roundings as the forward)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lavt_oracle as O  # noqa: E402  (test infrastructure)

pytestmark = pytest.mark.gpu

GRAD_L2 = 3e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# The ReLU inside the LanguageGate sits on pre-activations r G0^T that straddle zero: rounding r and G0 to bf16 flips the mask of the
# elements closest to the kink, and d G0 inherits that (an fp32 emulation with ONLY r and the two weights rounded to bf16 -- no kernel
# involved -- already differs from the fp32 gradient by 4-6 % rel-L2 on these random gates; every other tensor stays below 1 %).
LOOSE = {"res_gate.0.weight": 1e-1}


def check_grads(got: dict, ref: dict, what: str, tol: float = GRAD_L2):
    """rel-L2 per tensor; gradients that are analytically zero (a bias in front of an InstanceNorm, the key bias under the softmax's
    shift invariance: reference norm < 1e-3 of the largest same-sized gradient) are compared against that peer instead."""
    bad = []
    for k, r in ref.items():
        assert k in got, f"{what}: no gradient produced for {k}"
        a, b = got[k].reshape(r.shape).float().cpu(), r.float()
        big = max(q.float().norm().item() for q in ref.values() if q.numel() == r.numel())
        if b.norm().item() < 1e-3 * big:     # analytically zero: what we produce must be negligible next to its same-sized peers
            e = (a.norm() / big).item()
        else:
            e = ((a - b).norm() / b.norm()).item()
        lim = max([tol] + [v for kk, v in LOOSE.items() if k.endswith(kk)])
        if not e < lim:
            bad.append((k, e))
    if os.environ.get("LAVT_TEST_VERBOSE"):
        print(f"[{what}] worst:", sorted(((rel_l2(got[k].reshape(r.shape), r), k) for k, r in ref.items()), reverse=True)[:12])
    assert not bad, f"{what}: gradients out of tolerance (rel-L2 > {tol}): {bad}"


def cos_and_ratio(a, b):
    a, b = a.float().cpu().reshape(-1), b.float().cpu().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item(), (a.norm() / (b.norm() + 1e-30)).item()


def check_direction(got: dict, ref: dict, what: str, min_cos: float = 0.95):
    """End-to-end criterion.  Six ReLUs (decoder) and one per LanguageGate sit between the loss and most parameters; the 0.3-0.5 %
    bf16 forward error flips that fraction of their masks, which moves an fp32 gradient by ~sqrt(eps) ~ 10-25 % in rel-L2 no matter how
    the backward is implemented (the per-unit tests above, where both sides see the same masks, hold 3e-2).  What an implementation
    error would change -- direction and scale of every gradient -- is what is asserted here."""
    bad = []
    for k, r in ref.items():
        assert k in got, f"{what}: no gradient produced for {k}"
        big = max(q.float().norm().item() for q in ref.values() if q.numel() == r.numel())
        if r.float().norm().item() < 1e-3 * big:
            if got[k].float().norm().item() > 0.05 * big:
                bad.append((k, "nonzero", got[k].float().norm().item() / big))
            continue
        c, ratio = cos_and_ratio(got[k], r)
        if not (c > min_cos and 0.85 < ratio < 1.18):
            bad.append((k, round(c, 4), round(ratio, 4)))
    assert not bad, f"{what}: gradient direction / scale off for {len(bad)} tensors, e.g. {bad[:8]}"


def test_transpose_colsum_splitk():
    from lavt_rs_b200 import _cabi as K
    g = torch.Generator().manual_seed(0)
    for M, N, Kin in ((1000, 96, 64), (4104, 384, 128), (392 * 9, 512, 2048), (3 * 148, 128, 128)):
        dy = torch.randn(M, N, generator=g).cuda().to(torch.bfloat16)
        x = torch.randn(M, Kin, generator=g).cuda().to(torch.bfloat16)
        M8 = (M + 7) // 8 * 8
        dyt = torch.zeros(N, M8, device="cuda", dtype=torch.bfloat16)
        xt = torch.zeros(Kin, M8, device="cuda", dtype=torch.bfloat16)
        K.transpose_bf16(dy, dyt[:, :M])
        K.transpose_bf16(x, xt[:, :M])
        assert torch.equal(dyt[:, :M], dy.t()) and torch.equal(xt[:, :M], x.t())
        dst = torch.full((N, Kin), 0.5, device="cuda")
        wsf = K.splitk_workspace_floats(N, Kin, M8)
        part = torch.empty(wsf, device="cuda")
        K.gemm_bf16_splitk(dyt, xt, dst, part, accumulate=True)
        ref = dy.float().t() @ x.float() + 0.5
        assert rel_l2(dst, ref) < 2e-3, (M, N, Kin, rel_l2(dst, ref))
        # the same product straight from the row-major operands (MN-major tcgen05 descriptors, no transposed copies)
        dst2 = torch.full((N, Kin), 0.5, device="cuda")
        K.gemm_bf16_wgrad(dy, x, dst2, torch.empty(K.splitk_workspace_floats(N, Kin, M), device="cuda"), accumulate=True)
        assert rel_l2(dst2, ref) < 2e-3, ("mn-major", M, N, Kin, rel_l2(dst2, ref))
        cs = torch.ones(N, device="cuda")
        K.colsum_accumulate(dy, cs)
        assert rel_l2(cs, dy.float().sum(0) + 1) < 1e-3
        cs32 = torch.zeros(N, device="cuda")
        K.colsum_accumulate(dy.float(), cs32)
        assert rel_l2(cs32, dy.float().sum(0)) < 1e-4


def test_gelu_and_layernorm_bwd():
    from lavt_rs_b200 import _cabi as K
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(1024, 512, generator=g) * 2).to(torch.bfloat16)
    dy = torch.randn(1024, 512, generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_()
    torch.nn.functional.gelu(xr).backward(dy.float())
    y = torch.empty_like(x, device="cuda")
    K.gelu_fwd(x.cuda(), y)
    assert rel_l2(y, torch.nn.functional.gelu(x.float())) < 4e-3
    dx = torch.empty_like(y)
    K.gelu_bwd(dy.cuda(), x.cuda(), dx)
    assert rel_l2(dx, xr.grad) < 4e-3
    for Cn in (96, 128, 512, 1024):
        M = 777
        xi = torch.randn(M, Cn, generator=g) * 1.5 + 0.3
        gam, bet = torch.rand(Cn, generator=g) + 0.5, torch.randn(Cn, generator=g)
        dyl = torch.randn(M, Cn, generator=g).to(torch.bfloat16)
        res = torch.randn(M, Cn, generator=g)
        xa, ga, ba = xi.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
        torch.nn.functional.layer_norm(xa, (Cn,), ga, ba, 1e-5).backward(dyl.float())
        dxo = res.clone().cuda()
        dga, dbe = torch.zeros(Cn, device="cuda"), torch.zeros(Cn, device="cuda")
        K.layernorm_rows_bwd(xi.cuda(), dyl.cuda(), gam.cuda(), dxo, dga, dbe, dres=dxo)
        assert rel_l2(dxo, xa.grad + res) < 1e-4, Cn
        assert rel_l2(dga, ga.grad) < 1e-4 and rel_l2(dbe, ba.grad) < 1e-4, Cn


def _block_setup(window, mha=(1, 1, 1, 1)):
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=window, fusion_heads=mha)
    sd = O.random_state_dict(cfg, seed=0)
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=window, drop_path_rate=0.0, patch_norm=True, num_heads_fusion=list(mha), args=None)
    load_reference_state_dict(bb, sd, "backbone.")
    return cfg, sd, bb.cuda().train()


def _oracle_grads(fn, sd, pre, inputs, gout):
    """autograd through an oracle function: returns (input grads, {param name without prefix: grad})."""
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items() if k.startswith(pre)}
    sd2 = dict(sd)
    sd2.update(leaf)
    ins = [t.clone().requires_grad_() for t in inputs]
    out = fn(sd2, *ins)
    out.backward(gout)
    return [t.grad for t in ins], {k[len(pre):]: v.grad for k, v in leaf.items() if v.grad is not None}


@pytest.mark.parametrize("stage,shifted,dims", [(0, False, (2, 8, 14, 14)), (0, True, (1, 8, 16, 12)), (1, True, (1, 4, 14, 21)),
                                                 (2, True, (2, 8, 7, 7))])
def test_swin_block_backward(stage, shifted, dims):
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    window = (8, 7, 7)
    cfg, sd, bb = _block_setup(window)
    B, D, H, W = dims
    C, nH = 128 * 2 ** stage, 4 * 2 ** stage
    bi = 1 if shifted else 0
    blk = bb.layers[stage].blocks[bi]
    pre = f"backbone.layers.{stage}.blocks.{bi}."
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, D, H, W, C, generator=g)
    gout = torch.randn(B, D, H, W, C, generator=g)

    def fn(sd2, xx):
        return O.swin_mlp_half(O.swin_attention_half(xx, sd2, pre, nH, window, shifted), sd2, pre)
    (dx_ref,), pg_ref = _oracle_grads(fn, sd, pre, [x], gout)

    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    out, saved = T.swin_block_fwd(xf, blk, B, D, H, W, window, shifted, True, ws)
    assert rel_l2(out, fn(sd, x).reshape(-1, C)) < 1e-2
    dx = gout.cuda().reshape(-1, C).contiguous()
    dx = T.swin_block_bwd(blk, saved, dx, grads, ws)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    check_grads(grads.named(blk), pg_ref, f"swin block stage {stage} shifted={shifted}")
    # the forward kernel saved the row log-sum-exp for the attention backward; without it the backward recomputes it in a first pass
    assert saved[-1] is not None and saved[-1].shape == (saved[8].rows(), nH)
    # (without the statistics the mma.sync kernel runs; with them the tcgen05 kernel attn_bwd_tc.cu where it applies: two implementations,
    # each held to the oracle above -- their mutual distance is bounded by twice their own bf16 error)
    grads2 = T.GradStore()
    dx2 = T.swin_block_bwd(blk, saved[:-1] + (None,), gout.cuda().reshape(-1, C).contiguous(), grads2, ws)
    assert rel_l2(dx2, dx) < 1e-2
    n1, n2 = grads.named(blk), grads2.named(blk)
    errs = {k: rel_l2(n2[k], n1[k]) for k in ("attn.qkv.weight", "attn.relative_position_bias_table")}
    assert all(e < 1e-2 for e in errs.values()), errs
    # the mma.sync kernel with and without the saved statistics
    from lavt_rs_b200 import _cabi as K
    prev = K.set_attention_bwd_impl("mma")
    try:
        grads3 = T.GradStore()
        dx3 = T.swin_block_bwd(blk, saved, gout.cuda().reshape(-1, C).contiguous(), grads3, ws)
        torch.cuda.synchronize()
    finally:
        K.set_attention_bwd_impl(prev)
    n3 = grads3.named(blk)
    assert rel_l2(dx3, dx2) < 5e-3 and all(rel_l2(n3[k], n2[k]) < 5e-3 for k in ("attn.qkv.weight", "attn.relative_position_bias_table"))


def test_swin_block_drop_path():
    """Stochastic depth (timm DropPath, reference :266/:271): per-sample branch scale inside the GEMM epilogues and its adjoint."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    window = (8, 7, 7)
    cfg, sd, bb = _block_setup(window)
    B, D, H, W, C, nH = 3, 4, 14, 10, 128, 4
    blk = bb.layers[0].blocks[1]
    pre = "backbone.layers.0.blocks.1."
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, D, H, W, C, generator=g)
    gout = torch.randn(B, D, H, W, C, generator=g)
    s_attn = torch.tensor([1 / 0.7, 0.0, 1 / 0.7])
    s_mlp = torch.tensor([0.0, 1 / 0.7, 1 / 0.7])

    def fn(sd2, xx):
        return O.swin_mlp_half(O.swin_attention_half(xx, sd2, pre, nH, window, True, branch_scale=s_attn), sd2, pre, branch_scale=s_mlp)
    (dx_ref,), pg_ref = _oracle_grads(fn, sd, pre, [x], gout)
    ws = E.workspace("cuda")
    grads = T.GradStore()
    out, saved = T.swin_block_fwd(x.cuda().reshape(-1, C).contiguous(), blk, B, D, H, W, window, True, True, ws,
                                  drop_scales=(s_attn.cuda(), s_mlp.cuda()))
    assert rel_l2(out, fn(sd, x).reshape(-1, C)) < 1e-2
    dx = T.swin_block_bwd(blk, saved, gout.cuda().reshape(-1, C).contiguous(), grads, ws)
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2
    check_grads(grads.named(blk), pg_ref, "swin block with DropPath")
    # the drawn scales take the two values of timm's DropPath
    blk.drop_path_rate = 0.25
    s = T.draw_drop_path(0.25, 4096, "cuda")
    vals = torch.unique(s).tolist()
    assert all(min(abs(v), abs(v - 1 / 0.75)) < 1e-6 for v in vals) and abs((s > 0).float().mean().item() - 0.75) < 0.05
    blk.drop_path_rate = 0.0


def test_patch_merging_backward():
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    cfg, sd, bb = _block_setup((8, 7, 7))
    ds = bb.layers[0].downsample
    pre = "backbone.layers.0.downsample."
    B, D, H, W, C = 2, 2, 9, 14, 128
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, D, H, W, C, generator=g)
    ref_out = O.patch_merging(x, sd, pre)
    gout = torch.randn(ref_out.shape, generator=g)
    (dx_ref,), pg_ref = _oracle_grads(lambda sd2, xx: O.patch_merging(xx, sd2, pre), sd, pre, [x], gout)
    ws = E.workspace("cuda")
    grads = T.GradStore()
    out, saved = T.patch_merging_fwd(x.cuda().reshape(-1, C).contiguous(), ds, B, D, H, W, ws)
    assert rel_l2(out, ref_out.reshape(-1, 2 * C)) < 1e-2
    dx = T.patch_merging_bwd(ds, saved, gout.cuda().reshape(-1, 2 * C).contiguous(), grads, ws)
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2
    check_grads(grads.named(ds), pg_ref, "patch merging")


@pytest.mark.parametrize("stage,heads,Nl,gate", [(0, 1, 20, True), (1, 4, 13, True), (3, 1, 22, False), (1, 1, 17, "sigmoid"), (0, 1, 20, "no_norm"),
                                                 (2, 2, 9, "no_norm")])
def test_pwam_gate_backward(stage, heads, Nl, gate):
    gate_act = "sigmoid" if gate == "sigmoid" else "tanh"          # --lg_act_layer sigmoid (reference lib/backbone.py:552-554)
    att_norm = "none" if gate == "no_norm" else "IN"               # --att_norm_layer_type none (reference lib/backbone.py:1297-1316): Identity
    gate = bool(gate)
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    mha = tuple(heads if i == stage else 1 for i in range(4))
    cfg, sd, bb = _block_setup((8, 7, 7), mha)
    layer = bb.layers[stage]
    C = 128 * 2 ** stage
    B, n = 2, 520
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, n, C, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl, 1)
    m[0, Nl - 5:] = 0
    m[1, Nl - 2:] = 0
    gr = torch.randn(B, n, C, generator=g)
    gx = torch.randn(B, n, C, generator=g)
    pre = f"backbone.layers.{stage}."
    # a zero-initialised gate (reference :525-526) has no gradient signal through tanh'; test with random gate weights
    sd = dict(sd)
    for k in ("res_gate.0.weight", "res_gate.2.weight"):
        sd[pre + k] = torch.randn(C, C, generator=g) * C ** -0.5
    with torch.no_grad():
        layer.res_gate[0].weight.copy_(sd[pre + "res_gate.0.weight"])
        layer.res_gate[2].weight.copy_(sd[pre + "res_gate.2.weight"])

    layer.fusion.image_lang_att.att_norm_layer_type = att_norm      # InstanceNorm1d has no parameters: the same module, other statistics

    def fn(sd2, xx, ll):
        r = O.pwam(xx, ll, m, sd2, pre + "fusion.", heads, att_norm=att_norm)
        if not gate:
            return r
        return torch.cat([r, O.language_gate(xx, r, sd2, pre + "res_gate.", act=gate_act)], 0)
    gout = torch.cat([gr, gx], 0) if gate else gr
    (dx_ref, dl_ref), pg_ref = _oracle_grads(fn, sd, pre, [x, l], gout)
    pg_ref = {k: v for k, v in pg_ref.items() if k.startswith("fusion.") or (gate and k.startswith("res_gate."))}

    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    r32, xg, saved = T.pwam_gate_fwd(xf, xf.to(torch.bfloat16), layer.fusion, layer.res_gate if gate else None, l.cuda(),
                                     m.squeeze(-1).cuda(), B, ws, gate_act=gate_act)
    ref = fn(sd, x, l)
    assert rel_l2(r32, ref[:B].reshape(-1, C)) < 1.5e-2
    if gate:
        assert rel_l2(xg, ref[B:].reshape(-1, C)) < 1.5e-2
    dl = torch.zeros(B, 768, Nl, device="cuda")
    dx = T.pwam_gate_bwd(layer.fusion, layer.res_gate if gate else None, saved, gr.cuda().reshape(-1, C).contiguous(),
                         gx.cuda().reshape(-1, C).contiguous() if gate else None, grads, ws, dl)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    assert rel_l2(dl, dl_ref) < GRAD_L2, ("dl", rel_l2(dl, dl_ref))
    check_grads(grads.named(layer), pg_ref, f"pwam stage {stage} heads {heads}")


def _decoder_setup():
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.weights import load_reference_state_dict
    cfg = O.OracleConfig(depths=(2, 2, 2, 2))
    sd = O.random_state_dict(cfg, seed=0)
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(dec, sd, "classifier.")
    return sd, dec.cuda().train()


def test_decoder_units_backward():
    """conv3x3 + BatchNorm(batch statistics) + ReLU and the upsample adjoint, each against autograd through the oracle on IDENTICAL
    bf16-representable inputs (one ReLU layer: both sides see the same mask, so the 3e-2 bar applies; a chain of them does not --
    see check_direction)."""
    from lavt_rs_b200 import _cabi as K
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    sd, dec = _decoder_setup()
    g = torch.Generator().manual_seed(22)
    ws = E.workspace("cuda")
    n, H, W = 3, 12, 10
    for conv_name, bn_name in (("conv2_3", "bn2_3"), ("conv1_3", "bn1_3"), ("conv1_2", "bn1_2")):
        Cin = sd[f"classifier.{conv_name}.weight"].shape[1]
        x = torch.randn(n, Cin, H, W, generator=g).to(torch.bfloat16).float()
        gout = torch.randn(n, 512, H, W, generator=g).to(torch.bfloat16).float()
        keys = [f"classifier.{conv_name}.weight", f"classifier.{bn_name}.weight", f"classifier.{bn_name}.bias"]
        leaf = dict(sd)
        leaf.update({k: sd[k].clone().requires_grad_() for k in keys})
        xr = x.clone().requires_grad_()
        y = O._cbr(xr, leaf, conv_name, bn_name, train_bn=True, emulate_bf16=True)
        y.backward(gout)
        grads = T.GradStore()
        t, saved = T._cbr_fwd(x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), dec, conv_name, bn_name, ws, False)
        assert rel_l2(t.permute(0, 3, 1, 2), y) < 5e-3
        dx = torch.empty(n, H, W, Cin, device="cuda", dtype=torch.bfloat16)
        T._cbr_bwd(dec, saved, gout.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16).view(-1, 512), grads, ws, False, dx)
        assert rel_l2(dx.permute(0, 3, 1, 2), xr.grad) < 1e-2, (conv_name, rel_l2(dx.permute(0, 3, 1, 2), xr.grad))
        check_grads({"classifier." + k: v for k, v in grads.named(dec).items()}, {k: leaf[k].grad for k in keys}, conv_name, tol=1e-2)
    # bilinear (align_corners) upsample half of upsample_concat
    for (ph, pw), (Ho, Wo) in (((6, 5), (12, 10)), ((12, 12), (24, 20)), ((7, 7), (7, 7))):
        prev = torch.randn(n, 512, ph, pw, generator=g).requires_grad_()
        go = torch.randn(n, 512 + 256, Ho, Wo, generator=g).to(torch.bfloat16).float()
        O._up_to(prev, go).backward(go[:, :512])
        dprev = torch.empty(n, ph, pw, 512, device="cuda", dtype=torch.bfloat16)
        K.upsample_concat_bwd(go.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), dprev)
        assert rel_l2(dprev.permute(0, 3, 1, 2), prev.grad) < 5e-3


def test_decoder_backward():
    """Whole SimpleDecoding in training mode: logits, BatchNorm running buffers, and direction / scale of every gradient."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    sd, dec = _decoder_setup()
    g = torch.Generator().manual_seed(21)
    n = 3
    shapes = [(n, 1024, 3, 3), (n, 512, 6, 6), (n, 256, 12, 12), (n, 128, 24, 20)]
    cs = [torch.randn(s, generator=g).to(torch.bfloat16).float() for s in shapes]        # c4, c3, c2, c1 (NCHW)
    gout = torch.randn(n, 2, 24, 20, generator=g)
    pre = "classifier."
    fn = lambda sd2, a, b, c, d: O.decoder_forward(sd2, a, b, c, d, train_bn=True, emulate_bf16=True)      # noqa: E731
    (d4, d3, d2, d1), pg_ref = _oracle_grads(fn, sd, pre, cs, gout)
    ref = O.decoder_forward(sd, *cs, train_bn=True)

    ws = E.workspace("cuda")
    grads = T.GradStore()
    nhwc = [c.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16) for c in cs]
    lg, saved = T.decoder_fwd(dec, *nhwc, ws)
    assert rel_l2(lg.permute(0, 3, 1, 2), ref) < 2e-2, rel_l2(lg.permute(0, 3, 1, 2), ref)
    # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    y = torch.cat([torch.nn.functional.interpolate(cs[0], size=(6, 6), mode="bilinear", align_corners=True), cs[1]], 1)
    z = torch.nn.functional.conv2d(y, sd["classifier.conv1_4.weight"], padding=1)
    assert rel_l2(dec.bn1_4.running_mean, 0.9 * sd["classifier.bn1_4.running_mean"] + 0.1 * z.mean((0, 2, 3))) < 2e-2
    assert rel_l2(dec.bn1_4.running_var, 0.9 * sd["classifier.bn1_4.running_var"] + 0.1 * z.var((0, 2, 3), unbiased=True)) < 2e-2
    dcs = T.decoder_bwd(dec, saved, gout.permute(0, 2, 3, 1).contiguous().cuda(), grads, ws)
    torch.cuda.synchronize()
    got = {nm: t for nm, t in zip(("dc4", "dc3", "dc2", "dc1"), dcs)}
    want = {nm: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]) for nm, t in zip(("dc4", "dc3", "dc2", "dc1"), (d4, d3, d2, d1))}
    check_direction(got, want, "decoder input gradients", min_cos=0.98)
    check_direction(grads.named(dec), pg_ref, "decoder parameters", min_cos=0.98)


def test_cross_entropy_and_upsample_bwd():
    from lavt_rs_b200 import _cabi as K
    g = torch.Generator().manual_seed(4)
    n, h, w, H, W = 3, 12, 10, 48, 40
    low = torch.randn(n, h, w, 2, generator=g).requires_grad_()
    target = torch.randint(0, 2, (n, H, W), generator=g)
    up = torch.nn.functional.interpolate(low.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True)
    loss = O.weighted_cross_entropy(up, target)
    loss.backward()
    logits = torch.empty(n, 2, H, W, device="cuda")
    K.upsample_logits(low.detach().cuda(), logits)
    acc = torch.zeros(2, device="cuda")
    K.cross_entropy(logits, target.cuda(), acc, phase=0)
    dlog = torch.empty_like(logits)
    K.cross_entropy(logits, target.cuda(), acc, dlog, phase=1)
    assert abs((acc[0] / acc[1]).item() - loss.item()) < 1e-4 * abs(loss.item()) + 1e-6
    dlow = torch.empty(n, h, w, 2, device="cuda")
    K.upsample_logits_bwd(dlog, dlow)
    assert rel_l2(dlow, low.grad) < 1e-4, rel_l2(dlow, low.grad)


def test_model_training_step():
    """Full video model: loss and every parameter gradient of one training step vs autograd through the oracle (train-mode BN,
    weighted CE), through the public autograd bridge (model.train(); loss.backward())."""
    from lavt_rs_b200.lib import segmentation
    from lavt_rs_b200.args import get_parser
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg = O.OracleConfig(depths=(2, 2, 2, 2))
    sd = O.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(31)
    for s in range(4):      # the reference zero-initialises the gates (:525-526): randomise them so that their path carries gradient
        C = 128 * 2 ** s
        for k in ("0", "2"):
            sd[f"backbone.layers.{s}.res_gate.{k}.weight"] = torch.randn(C, C, generator=g) * C ** -0.5
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib._utils import LAVT
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, num_heads_fusion=[1, 1, 1, 1], args=None)
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    model.video = True
    B, Tn, H, W, Nl = 2, 4, 96, 96, 12
    x = torch.randn(B, Tn, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    m[0, 8:] = 0
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)

    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    lr = l.clone().requires_grad_()
    out_ref = O.model_forward(leaf, cfg, x, lr, m, train_bn=True)
    loss_ref = O.weighted_cross_entropy(out_ref, target)
    loss_ref.backward()

    # (1) fused step
    grads = T.GradStore()
    loss, dl = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    got = {}
    got.update({"backbone." + k: v for k, v in grads.named(bb).items()})
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None and not k.startswith("backbone.layers.3.res_gate")}
    check_direction(got, ref, "training step")
    c, ratio = cos_and_ratio(dl, lr.grad)
    assert c > 0.95 and 0.85 < ratio < 1.18, ("dl", c, ratio)
    assert not any(k.startswith("backbone.layers.3.res_gate") for k in got), "the last stage's gate is dead (no gradient)"

    # (2) the autograd bridge gives the same gradients through loss.backward()
    for p in model.parameters():
        p.grad = None
    for bn in (mm for mm in dec.modules() if isinstance(mm, torch.nn.BatchNorm2d)):
        bn.reset_running_stats()
    lt = l.cuda().requires_grad_()
    out = TR.SegmentFunction.apply(x.cuda(), lt, m.cuda(), model, False)
    assert rel_l2(out, out_ref) < 3e-2
    loss2 = torch.nn.functional.cross_entropy(out, target.cuda(), weight=torch.tensor([0.9, 1.1], device="cuda"))
    loss2.backward()
    assert cos_and_ratio(lt.grad, lr.grad)[0] > 0.95
    for name, prm in (("backbone.layers.2.blocks.1.mlp.fc1.weight", bb.layers[2].blocks[1].mlp.fc1.weight),
                      ("classifier.conv2_3.weight", dec.conv2_3.weight), ("backbone.patch_embed.proj.weight", bb.patch_embed.proj.weight)):
        assert rel_l2(prm.grad, got[name].reshape(prm.shape)) < 3e-2, name      # same kernels, same inputs: both routes agree

    # (3) the TIGHT end-to-end criterion: the oracle differentiates the same linear branch of every ReLU as the kernels did (the six
    # decoder ReLUs and the LanguageGate ReLUs take their 0/1 masks from the kernels' saved activations, oracle.RELU_BRANCH), so what is
    # left between the two gradients is bf16 operand rounding: every parameter gradient and d l_feats within REL_TIGHT in rel-L2.
    for bn in (mm for mm in dec.modules() if isinstance(mm, torch.nn.BatchNorm2d)):
        bn.reset_running_stats()
    with torch.no_grad():
        _, tape = TR.segment_forward(model, x.cuda(), l.cuda(), m.cuda())
    branch = {}
    for _, s1, s2 in tape["dec"][0]:
        for sv in (s1, s2):
            branch["classifier." + sv[5]] = (sv[3] > 0).permute(0, 3, 1, 2).float().cpu()
    for s, st in enumerate(tape["stages"]):
        if st[1] is not None and st[1].get("g1") is not None:
            branch[f"backbone.layers.{s}.res_gate."] = (st[1]["g1"] > 0).float().cpu()
    assert len(branch) == 6 + 3
    leaf2 = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    lr2 = l.clone().requires_grad_()
    O.RELU_BRANCH = branch
    try:
        loss_b = O.weighted_cross_entropy(O.model_forward(leaf2, cfg, x, lr2, m, train_bn=True), target)
        loss_b.backward()
    finally:
        O.RELU_BRANCH = None
    ref2 = {k: v.grad for k, v in leaf2.items() if v.grad is not None and not k.startswith("backbone.layers.3.res_gate")}
    errs = {}
    for k, r in ref2.items():
        big = max(q.float().norm().item() for q in ref2.values() if q.numel() == r.numel())
        if r.float().norm().item() < 1e-3 * big:
            continue                                    # analytically zero (shift-invariant biases): covered by check_direction above
        errs[k] = rel_l2(got[k].reshape(r.shape), r)
    errs["d l_feats"] = rel_l2(dl, lr2.grad)
    worst = sorted(((e, k) for k, e in errs.items()), reverse=True)
    if os.environ.get("LAVT_TEST_VERBOSE"):
        print("[same-branch gradients] worst:", worst[:10], "median", sorted(errs.values())[len(errs) // 2])
    assert worst[0][0] < REL_TIGHT, f"same-branch end-to-end gradients: {worst[:6]}"


REL_TIGHT = 6e-2      # measured on B200: worst tensor 4.7e-2, median 3.7e-2 (bf16 operands and bf16 activation gradients end to end)


def test_fused_adamw_matches_torch():
    """lavt_adamw_step (one launch per parameter group) vs torch.optim.AdamW over three steps, with the reference's group structure
    (a no-decay group, default groups, train.py:615-692), odd sizes / unaligned views, AMSGrad on and off, and a LambdaLR schedule."""
    from lavt_rs_b200.optim import FusedAdamW, poly_lr_lambda
    g = torch.Generator().manual_seed(9)
    for amsgrad in (False, True):
        shapes = [(128, 64), (77,), (3, 5, 7), (4096 * 3 + 5,), (1,), (512, 1, 3, 3)]
        big = torch.randn(1000, generator=g).cuda()
        ours = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes] + [torch.nn.Parameter(big[3:3 + 500].clone())]
        ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]

        def groups(ps):
            return [{"params": ps[:2], "weight_decay": 0.0}, {"params": ps[2:5]}, {"params": ps[5:]}]
        o1 = FusedAdamW(groups(ours), lr=5e-3, weight_decay=1e-2, amsgrad=amsgrad)
        o2 = torch.optim.AdamW(groups(ref), lr=5e-3, weight_decay=1e-2, amsgrad=amsgrad)
        s1 = torch.optim.lr_scheduler.LambdaLR(o1, poly_lr_lambda(10))
        s2 = torch.optim.lr_scheduler.LambdaLR(o2, poly_lr_lambda(10))
        for it in range(3):
            for a, b in zip(ours, ref):
                gr = torch.randn(a.shape, generator=g).cuda()
                a.grad, b.grad = gr.clone(), gr.clone()
            if it == 1:
                ours[1].grad = None
                ref[1].grad = None           # a parameter without gradient is skipped
            o1.step(); o2.step(); s1.step(); s2.step()
        torch.cuda.synchronize()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (amsgrad, a.shape, (a - b).abs().max().item())
        st1, st2 = o1.state_dict()["state"], o2.state_dict()["state"]
        assert set(st1[0].keys()) == set(st2[0].keys())
        assert torch.allclose(st1[3]["exp_avg_sq"], st2[3]["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_fused_adamw_resumes_from_torch_adamw_state():
    """Optimizer state written by torch.optim.AdamW -- with the step counter as a Python int (torch 1.x checkpoints, the reference's
    era), a CPU tensor, or a CUDA tensor (what older load_state_dict produced) -- resumes in FusedAdamW and continues identically."""
    from lavt_rs_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(11)
    for kind in ("int", "cpu", "cuda"):
        ref = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in [(64, 32), (19,)]]
        o2 = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2)
        grads = [[torch.randn(p.shape, generator=g).cuda() for p in ref] for _ in range(4)]
        for it in range(2):
            for p, gr in zip(ref, grads[it]):
                p.grad = gr.clone()
            o2.step()
        ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
        import copy
        sd = copy.deepcopy(o2.state_dict())          # state_dict() hands out the optimizer's own tensors: edit a copy
        for st in sd["state"].values():
            stp = int(st["step"].item()) if torch.is_tensor(st["step"]) else int(st["step"])
            st["step"] = stp if kind == "int" else torch.tensor(float(stp), device=kind)
        o1 = FusedAdamW(ours, lr=1e-3, weight_decay=1e-2)
        o1.load_state_dict(sd)
        for it in range(2, 4):
            for p, q, gr in zip(ref, ours, grads[it]):
                p.grad, q.grad = gr.clone(), gr.clone()
            o2.step(); o1.step()
        for st in o1.state.values():
            assert torch.is_tensor(st["step"]) and not st["step"].is_cuda and int(st["step"].item()) == 4
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (kind, (a - b).abs().max().item())


def test_image_model_training_through_public_api(tmp_path):
    """2-D LAVT (windows (1,7,7), never clamped) in ``model.train()``: ``model(x, l_feats, l_mask)`` -> criterion -> ``backward()`` ->
    FusedAdamW step -> checkpoint round trip in the reference's format; gradients vs autograd through the oracle (direction / scale)."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200.optim import FusedAdamW, poly_lr_lambda, reference_param_groups
    from lavt_rs_b200.checkpoint import resume_checkpoint, save_checkpoint
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False)
    sd = O.random_state_dict(cfg, seed=0)
    bb = MultiModalSwinTransformer(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0,
                                   patch_norm=True, num_heads_fusion=[1, 1, 1, 1], args=None)
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    g = torch.Generator().manual_seed(41)
    B, H, W, Nl = 3, 160, 128, 15
    x = torch.randn(B, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    m[1, 9:] = 0
    target = torch.randint(0, 2, (B, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, l, m, train_bn=True), target)
    loss_ref.backward()

    opt = FusedAdamW(reference_param_groups(model), lr=1e-4, weight_decay=1e-2, amsgrad=True)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, poly_lr_lambda(100))
    out = model(x.cuda(), l.cuda(), m.cuda())                       # training mode: autograd-connected logits
    assert out.requires_grad and out.shape == (B, 2, H, W)
    loss = torch.nn.functional.cross_entropy(out, target.cuda(), weight=torch.tensor([0.9, 1.1], device="cuda"))
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    opt.zero_grad()
    loss.backward()
    got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    check_direction(got, ref, "2-D training step")
    before = bb.layers[1].blocks[0].mlp.fc1.weight.detach().clone()
    with torch.no_grad():
        model.eval()
        out_before = model(x.cuda(), l.cuda(), m.cuda()).clone()
        model.train()
    opt.step()
    sched.step()
    assert not torch.equal(before, bb.layers[1].blocks[0].mlp.fc1.weight)
    # eval mode still runs the inference path, without autograd
    model.eval()
    with torch.no_grad():
        out_after = model(x.cuda(), l.cuda(), m.cuda()).clone()
        assert not out_after.requires_grad
        # the optimizer kernel writes the fp32 masters through raw pointers: the cached bf16 / folded operand copies must follow
        # (version counters bumped in FusedAdamW.step) -- the forward after the step uses the NEW weights ...
        assert (out_after - out_before).abs().max().item() > 1e-4, "forward after optimizer.step() still uses the step-0 weights"
        # ... and equals a forward with every prepared-operand cache rebuilt from scratch
        for mod in model.modules():
            if hasattr(mod, "prepared"):
                mod.prepared.clear()
        out_fresh = model(x.cuda(), l.cuda(), m.cuda())
        assert torch.equal(out_after, out_fresh), (out_after - out_fresh).abs().max().item()
    # checkpoint round trip (train.py:752-762 format)
    path = os.path.join(tmp_path, "models", "checkpoint_00.pth")
    save_checkpoint(path, model, opt, sched, epoch=0, args={"lr": 1e-4})
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "optimizer", "epoch", "args", "lr_scheduler"} and set(ck["model"]) == set(model.state_dict())
    with torch.no_grad():
        bb.layers[1].blocks[0].mlp.fc1.weight.zero_()
    assert resume_checkpoint(path, model, opt, sched) == 0
    assert bb.layers[1].blocks[0].mlp.fc1.weight.abs().sum().item() > 0


def test_swin_tiny_width_training_step():
    """Swin-T / Swin-S widths (96, 192, 384, 768; 3-24 heads; decoder width 384) through the whole training step with stochastic
    depth off: tile remainders in the dgrad / wgrad GEMMs, masked LayerNorm-backward lanes, scalar PWAM-backward lanes."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg = O.OracleConfig(embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window=(8, 7, 7))
    sd = O.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(51)
    for s in range(3):
        C = 96 * 2 ** s
        for k in ("0", "2"):
            sd[f"backbone.layers.{s}.res_gate.{k}.weight"] = torch.randn(C, C, generator=g) * C ** -0.5
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=96, depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, args=None)
    dec = SimpleDecoding(768, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    B, Tn, H, W, Nl = 2, 4, 64, 96, 20
    x = torch.randn(B, Tn, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    m[1, 14:] = 0
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    lr = l.clone().requires_grad_()
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, lr, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, dl = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    check_direction(got, ref, "Swin-T width training step")
    c, ratio = cos_and_ratio(dl, lr.grad)
    assert c > 0.95 and 0.85 < ratio < 1.18, ("dl", c, ratio)


SEP_FLAGS = ["--sep_t_pwam", "--conv3d_kernel_size_t", "3-3-3", "--conv3d_kernel_size_s", "1-1-1", "--w_t3x3_s1x1", "--mm_t3x3_s1x1"]


def _sep_setup(embed=128, heads=(4, 8, 16, 32)):
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200.args import default_args
    cfg = O.OracleConfig(embed_dim=embed, depths=(2, 2, 2, 2), num_heads=heads, window=(8, 7, 7), sep_t_pwam=True)
    sd = O.random_state_dict(cfg, seed=0)
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=embed, depths=[2, 2, 2, 2], num_heads=list(heads), window_size=(8, 7, 7),
                                     drop_path_rate=0.0, patch_norm=True, args=default_args(SEP_FLAGS))
    load_reference_state_dict(bb, sd, "backbone.")
    return cfg, sd, bb.cuda().train()


@pytest.mark.parametrize("stage,gate", [(0, True), (2, False), (1, "sigmoid")])
def test_sep_t_pwam_gate_backward(stage, gate):
    """SepTPWAM under the README video flags (Conv3d(3,3,3) + Conv3d(1,1,1) branches, summed InstanceNorms) + LanguageGate (tanh, or the
    sigmoid of --lg_act_layer sigmoid): every gradient vs autograd through the oracle."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    gate_act = "sigmoid" if gate == "sigmoid" else "tanh"
    gate = bool(gate)
    cfg, sd, bb = _sep_setup()
    layer = bb.layers[stage]
    C = 128 * 2 ** stage
    B, D, H, W, Nl = 2, 4, 9, 10, 17
    n = D * H * W
    g = torch.Generator().manual_seed(61)
    x = torch.randn(B, D, H, W, C, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl, 1)
    m[0, Nl - 4:] = 0
    gr = torch.randn(B, n, C, generator=g)
    gx = torch.randn(B, n, C, generator=g)
    pre = f"backbone.layers.{stage}."
    sd = dict(sd)
    for k in ("res_gate.0.weight", "res_gate.2.weight"):
        sd[pre + k] = torch.randn(C, C, generator=g) * C ** -0.5
    with torch.no_grad():
        layer.res_gate[0].weight.copy_(sd[pre + "res_gate.0.weight"])
        layer.res_gate[2].weight.copy_(sd[pre + "res_gate.2.weight"])

    def fn(sd2, xx, ll):
        r = O.sep_t_pwam(xx, ll, m, sd2, pre + "fusion.", 1)
        if not gate:
            return r
        return torch.cat([r, O.language_gate(xx.reshape(B, n, C), r, sd2, pre + "res_gate.", act=gate_act)], 0)
    gout = torch.cat([gr, gx], 0) if gate else gr
    (dx_ref, dl_ref), pg_ref = _oracle_grads(fn, sd, pre, [x, l], gout)
    pg_ref = {k: v for k, v in pg_ref.items() if k.startswith("fusion.") or (gate and k.startswith("res_gate."))}
    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    r32, xg, saved = T.sep_t_pwam_gate_fwd(xf, xf.to(torch.bfloat16), layer.fusion, layer.res_gate if gate else None, l.cuda(),
                                           m.squeeze(-1).cuda(), B, D, H, W, ws, gate_act=gate_act)
    if gate:
        assert rel_l2(xg, fn(sd, x, l)[B:].reshape(-1, C)) < 1.5e-2
    ref = fn(sd, x, l)
    assert rel_l2(r32, ref[:B].reshape(-1, C)) < 1.5e-2
    dl = torch.zeros(B, 768, Nl, device="cuda")
    dx = T.sep_t_pwam_gate_bwd(layer.fusion, layer.res_gate if gate else None, saved, gr.cuda().reshape(-1, C).contiguous(),
                               gx.cuda().reshape(-1, C).contiguous() if gate else None, grads, ws, dl)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    assert rel_l2(dl, dl_ref) < GRAD_L2, ("dl", rel_l2(dl, dl_ref))
    check_grads(grads.named(layer), pg_ref, f"SepTPWAM stage {stage}")


def test_sep_t_pwam_tiny_training_step():
    """The README's video configuration (--swin_type tiny widths + SepTPWAM flags) through one whole training step."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg, sd, bb = _sep_setup(96, (3, 6, 12, 24))
    dec = SimpleDecoding(768, None)
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    g = torch.Generator().manual_seed(71)
    B, Tn, H, W, Nl = 2, 4, 64, 96, 20
    x = torch.randn(B, Tn, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    m[0, 15:] = 0
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    lr = l.clone().requires_grad_()
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, lr, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, dl = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    # the reference zero-initialises the gates: their gradients are exactly zero on both sides and carry no direction
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None and "res_gate" not in k}
    check_direction(got, ref, "SepTPWAM tiny training step")


def test_hs_training_step():
    """--hs (stage output = gated features E_i instead of the PWAM residual, reference :579-587): the decoder gradient enters the
    residual stream through the gate, including at the last stage."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), hs=True)
    sd = O.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(81)
    for s in range(4):
        C = 128 * 2 ** s
        for k in ("0", "2"):
            sd[f"backbone.layers.{s}.res_gate.{k}.weight"] = torch.randn(C, C, generator=g) * C ** -0.5
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, args=default_args(["--hs"]))
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    B, Tn, H, W, Nl = 2, 4, 96, 64, 12
    x = torch.randn(B, Tn, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, l, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, _ = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    assert any(k.startswith("backbone.layers.3.res_gate") for k in ref), "with --hs the last stage's gate is live"
    check_direction(got, ref, "--hs training step")


@pytest.mark.parametrize("version", ["no_gate", "none"])
def test_version_ablation_training_step(version):
    """--version no_gate (x' = x + r) / none (x' = x): reference lib/video_swin_transformer.py:561-575."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), version=version)
    sd = {k: v for k, v in O.random_state_dict(cfg, seed=0).items() if "res_gate" not in k}
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, args=default_args(["--version", version]))
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    g = torch.Generator().manual_seed(91)
    B, Tn, H, W, Nl = 2, 4, 64, 96, 10
    x = torch.randn(B, Tn, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, l, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, _ = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    check_direction(got, {k: v.grad for k, v in leaf.items() if v.grad is not None}, f"--version {version} training step")


@pytest.mark.parametrize("video", [True, False])
def test_lazy_pred_training_step(video):
    """--lazy_pred in training mode (reference lib/video_swin_transformer.py:556-558, lib/mask_predictor.py:77): the decoder reads
    norm_i(V_i) of stages 1-3 (features BEFORE fusion), the fusion residual only feeds the gate, the last stage's fusion is dead code
    and the logits come from the 1/8-scale level.  Loss, every parameter gradient and d l_feats vs autograd through the oracle."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    args = default_args(["--lazy_pred"])
    cfg = (O.OracleConfig(depths=(2, 2, 2, 2), lazy_pred=True) if video else
           O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False, lazy_pred=True))
    sd = O.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(5)
    for k in list(sd):      # a zero-initialised gate passes no gradient to the fusion: random gate weights (as in the unit test)
        if "res_gate" in k and k.endswith("weight"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * sd[k].shape[1] ** -0.5
    kw = dict(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], drop_path_rate=0.0, patch_norm=True, out_indices=(1, 2, 3), args=args)
    bb = (MultiModalSwinTransformer3D(patch_size=(1, 4, 4), window_size=(8, 7, 7), **kw) if video
          else MultiModalSwinTransformer(window_size=7, num_heads_fusion=[1, 1, 1, 1], **kw))
    dec = SimpleDecoding(1024, args)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    B, Tn, H, W, Nl = (2, 4, 64, 96, 10) if video else (3, 1, 160, 128, 10)
    x = torch.randn(B, Tn, 3, H, W, generator=g) if video else torch.randn(B, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    m[1, Nl - 3:] = 0
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    lq = l.clone().requires_grad_(True)
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, lq, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, dl = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    dead = [k for k in ref if k.startswith("backbone.layers.3.fusion") and ref[k].abs().max() > 0]
    assert not dead, f"the last stage's fusion is unused under --lazy_pred, yet the oracle has gradients for {dead[:3]}"
    check_direction(got, ref, "--lazy_pred training step")
    c, ratio = cos_and_ratio(dl, lq.grad)
    assert c > 0.95 and 0.85 < ratio < 1.18, ("d l_feats", c, ratio)


def test_text_encoder_on_side_stream():
    """``training.SideStreamText``: the text encoder's forward runs on a side stream under patch embedding + stage 0 (the hot path waits for
    its event at the first fusion) and its backward under the backward of stage 0's Swin blocks (``on_dl_ready``).  Same loss, d l_feats,
    hot-path gradients and text-encoder gradients as the serial order (a small autograd module stands in for BERT)."""
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    cfg = O.OracleConfig(depths=(2, 2, 2, 2))
    sd = O.random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(3)
    for k in list(sd):
        if "res_gate" in k and k.endswith("weight"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * sd[k].shape[1] ** -0.5
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, args=None)
    dec = SimpleDecoding(1024, None)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    B, Tn, H, W, Nl = 2, 4, 64, 96, 12
    x = torch.randn(B, Tn, 3, H, W, generator=g).cuda()
    m = torch.ones(B, Nl).cuda()
    target = torch.randint(0, 2, (B * Tn, H, W), generator=g).cuda()
    ids = torch.randint(0, 100, (B, Nl), generator=g).cuda()
    torch.manual_seed(0)
    emb = torch.nn.Embedding(100, 768).cuda()
    lin = torch.nn.Linear(768, 768).cuda()
    text_params = list(emb.parameters()) + list(lin.parameters())

    def text_fn(ids_, mask_):
        h = torch.tanh(lin(emb(ids_)))
        for _ in range(20):                       # a few hundred small launches, like the real encoder
            h = torch.tanh(lin(h)) + h
        return h.permute(0, 2, 1)

    def run(side_stream: bool):
        for prm in text_params:
            prm.grad = None
        bn = {k: v.clone() for k, v in model.named_buffers()}
        grads = T.GradStore()
        if side_stream:
            side = TR.SideStreamText(x.device)
            l_feats, ready = side.forward(text_fn, ids, m)
            loss, dl = TR.segment_forward_backward(model, x, l_feats.detach(), m, target, grads, lang_ready=ready, on_dl_ready=side.backward_hook())
            side.join()
        else:
            l_feats = text_fn(ids, m)
            loss, dl = TR.segment_forward_backward(model, x, l_feats.detach(), m, target, grads)
            l_feats.backward(dl)
        torch.cuda.synchronize()
        with torch.no_grad():
            for k, v in model.named_buffers():     # BatchNorm running statistics moved: both runs start from the same model
                v.copy_(bn[k])
        got = {k: v.clone() for k, v in grads.named(bb).items()}
        return loss.item(), dl.clone(), got, [prm.grad.clone() for prm in text_params]

    # Two serial runs differ too: fp32 atomics (InstanceNorm / LayerNorm reductions) sum in a different order and bf16 roundings downstream
    # of them flip.  That run-to-run noise is the yardstick: the side-stream runs must stay within 3x of it.
    loss0, dl0, g0, t0 = run(False)
    loss0b, dl0b, g0b, t0b = run(False)
    big = max(v.float().norm().item() for v in g0.values())

    def dist(ga, gb):
        return {k: (ga[k].float() - gb[k].float()).norm().item() for k in ga}
    noise_dl = rel_l2(dl0b, dl0)
    noise_t = max(rel_l2(a_, b_) for a_, b_ in zip(t0b, t0))
    noise_g = dist(g0b, g0)
    assert noise_dl < 2e-2 and noise_t < 2e-2, (noise_dl, noise_t)
    for _ in range(3):                             # repeated: a missing stream dependency shows up as a flaky mismatch
        loss1, dl1, g1, t1 = run(True)
        assert abs(loss0 - loss1) < 1e-5 * abs(loss0)
        assert rel_l2(dl1, dl0) < 3 * noise_dl + 1e-4, (rel_l2(dl1, dl0), noise_dl)
        for a_, b_ in zip(t1, t0):
            assert rel_l2(a_, b_) < 3 * noise_t + 1e-4, (rel_l2(a_, b_), noise_t)
        d1 = dist(g1, g0)
        for k in g0:
            assert d1[k] <= 3 * noise_g[k] + 1e-3 * g0[k].float().norm().item() + 1e-5 * big, (k, d1[k], noise_g[k])


@pytest.mark.parametrize("stage,gate", [(0, True), (2, False)])
def test_simple_fusion_backward(stage, gate):
    """--fuse simple in training mode (reference lib/video_swin_transformer.py:916-917, 929-930, LangProject :1012-1039): r = GELU(project_mm(
    GELU(vis_project(x)) * LangProject(l))) + LanguageGate; every parameter gradient, dx and d l_feats vs autograd through the oracle."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    from lavt_rs_b200.weights import load_reference_state_dict
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(8, 7, 7), fuse_simple=True)
    sd = dict(O.random_state_dict(cfg, seed=0))
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=(8, 7, 7), drop_path_rate=0.0, patch_norm=True, args=default_args(["--fuse", "simple"]))
    C = 128 * 2 ** stage
    pre = f"backbone.layers.{stage}."
    g = torch.Generator().manual_seed(17)
    for k in ("res_gate.0.weight", "res_gate.2.weight"):       # a zero-initialised gate has no gradient signal
        sd[pre + k] = torch.randn(C, C, generator=g) * C ** -0.5
    load_reference_state_dict(bb, sd, "backbone.")
    bb = bb.cuda().train()
    layer = bb.layers[stage]
    B, n, Nl = 3, 520, 11
    x = torch.randn(B, n, C, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl, 1)
    m[0, Nl - 4:] = 0
    m[2, Nl - 1:] = 0
    gr = torch.randn(B, n, C, generator=g)
    gx = torch.randn(B, n, C, generator=g)

    def fn(sd2, xx, ll):
        r = O.pwam(xx, ll, m, sd2, pre + "fusion.", 1)
        if not gate:
            return r
        return torch.cat([r, O.language_gate(xx, r, sd2, pre + "res_gate.")], 0)
    gout = torch.cat([gr, gx], 0) if gate else gr
    (dx_ref, dl_ref), pg_ref = _oracle_grads(fn, sd, pre, [x, l], gout)
    pg_ref = {k: v for k, v in pg_ref.items() if k.startswith("fusion.") or (gate and k.startswith("res_gate."))}
    assert any(k.startswith("fusion.image_lang_att.project.") for k in pg_ref)
    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    r32, xg, saved = T.pwam_gate_fwd(xf, xf.to(torch.bfloat16), layer.fusion, layer.res_gate if gate else None, l.cuda(), m.squeeze(-1).cuda(), B, ws)
    ref = fn(sd, x, l)
    assert rel_l2(r32, ref[:B].reshape(-1, C)) < 1.5e-2
    dl = torch.zeros(B, 768, Nl, device="cuda")
    dx = T.pwam_gate_bwd(layer.fusion, layer.res_gate if gate else None, saved, gr.cuda().reshape(-1, C).contiguous(),
                         gx.cuda().reshape(-1, C).contiguous() if gate else None, grads, ws, dl)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    assert rel_l2(dl, dl_ref) < GRAD_L2, ("dl", rel_l2(dl, dl_ref))
    check_grads(grads.named(layer), pg_ref, f"simple fusion stage {stage}")


@pytest.mark.parametrize("stage,heads", [(0, 1), (1, 2)])
def test_pwam_layernorm_att_norm_backward(stage, heads):
    """--att_norm_layer_type LN in training mode (2-D backbone, reference lib/backbone.py:1297-1316): a row LayerNorm behind f_query and W
    instead of InstanceNorm1d; PWAM + gate gradients (incl. the two LayerNorms' weight / bias) vs autograd through the oracle."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    from lavt_rs_b200.weights import load_reference_state_dict
    fh = [heads if i == stage else 1 for i in range(4)]
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False, att_norm="LN", fusion_heads=tuple(fh))
    sd = dict(O.random_state_dict(cfg, seed=0))
    bb = MultiModalSwinTransformer(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0,
                                   patch_norm=True, num_heads_fusion=fh, args=default_args(["--att_norm_layer_type", "LN"]))
    C = 128 * 2 ** stage
    pre = f"backbone.layers.{stage}."
    g = torch.Generator().manual_seed(23)
    for k in ("res_gate.0.weight", "res_gate.2.weight"):
        sd[pre + k] = torch.randn(C, C, generator=g) * C ** -0.5
    load_reference_state_dict(bb, sd, "backbone.")
    bb = bb.cuda().train()
    layer = bb.layers[stage]
    B, n, Nl = 2, 520, 14
    x = torch.randn(B, n, C, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl, 1)
    m[1, Nl - 3:] = 0
    gr = torch.randn(B, n, C, generator=g)
    gx = torch.randn(B, n, C, generator=g)

    def fn(sd2, xx, ll):
        r = O.pwam(xx, ll, m, sd2, pre + "fusion.", heads, att_norm="LN")
        return torch.cat([r, O.language_gate(xx, r, sd2, pre + "res_gate.")], 0)
    (dx_ref, dl_ref), pg_ref = _oracle_grads(fn, sd, pre, [x, l], torch.cat([gr, gx], 0))
    pg_ref = {k: v for k, v in pg_ref.items() if k.startswith("fusion.") or k.startswith("res_gate.")}
    assert "fusion.image_lang_att.f_query.1.weight" in pg_ref and "fusion.image_lang_att.W.1.bias" in pg_ref
    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    r32, xg, saved = T.pwam_gate_fwd(xf, xf.to(torch.bfloat16), layer.fusion, layer.res_gate, l.cuda(), m.squeeze(-1).cuda(), B, ws)
    ref = fn(sd, x, l)
    assert rel_l2(r32, ref[:B].reshape(-1, C)) < 1.5e-2
    assert rel_l2(xg, ref[B:].reshape(-1, C)) < 1.5e-2
    dl = torch.zeros(B, 768, Nl, device="cuda")
    dx = T.pwam_gate_bwd(layer.fusion, layer.res_gate, saved, gr.cuda().reshape(-1, C).contiguous(), gx.cuda().reshape(-1, C).contiguous(), grads, ws, dl)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    assert rel_l2(dl, dl_ref) < GRAD_L2, ("dl", rel_l2(dl, dl_ref))
    check_grads(grads.named(layer), pg_ref, f"PWAM with LayerNorm attention norms, stage {stage}")


@pytest.mark.parametrize("flags,cfgkw", [(["--interpolate_before_seg"], dict(interpolate_before_seg=True)),
                                         (["--interpolate_before_seg", "--seg_last"], dict(interpolate_before_seg=True, seg_last=True))])
def test_decoder_tails_training_step(flags, cfgkw):
    """--interpolate_before_seg / --seg_last in training mode (reference lib/mask_predictor.py:40-48, 88-97): bilinear x2 + conv3x3 + BN + ReLU
    levels between the top-down decoder and the 1x1 classifier.  Loss and every parameter gradient of a 2-D model step vs autograd through
    the oracle (direction / scale criterion of the end-to-end tests)."""
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib._utils import LAVT
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.weights import load_reference_state_dict
    from lavt_rs_b200 import training as TR
    from lavt_rs_b200 import train_engine as T
    args = default_args(flags)
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False, **cfgkw)
    sd = O.random_state_dict(cfg, seed=0)
    bb = MultiModalSwinTransformer(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0,
                                   patch_norm=True, num_heads_fusion=[1, 1, 1, 1], args=args)
    dec = SimpleDecoding(1024, args)
    load_reference_state_dict(bb, sd, "backbone.")
    load_reference_state_dict(dec, sd, "classifier.")
    model = LAVT(bb, dec).cuda().train()
    g = torch.Generator().manual_seed(29)
    B, H, W, Nl = 2, 96, 64, 9
    x = torch.randn(B, 3, H, W, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl)
    target = torch.randint(0, 2, (B, H, W), generator=g)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_ref = O.weighted_cross_entropy(O.model_forward(leaf, cfg, x, l, m, train_bn=True), target)
    loss_ref.backward()
    grads = T.GradStore()
    loss, _ = TR.segment_forward_backward(model, x.cuda(), l.cuda(), m.cuda(), target.cuda(), grads)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    got = {"backbone." + k: v for k, v in grads.named(bb).items()}
    got.update({"classifier." + k: v for k, v in grads.named(dec).items()})
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    assert "classifier.conv2_1.weight" in ref and (("classifier.conv1_0.weight" in ref) == bool(cfgkw.get("seg_last")))
    # --seg_last puts EIGHT conv + BN + ReLU layers between the loss and the backbone (six in every other configuration): the mask-flip noise
    # of check_direction's argument grows with them (measured: five cancellation-heavy bias / gate tensors at cosine 0.943-0.948, norm
    # ratios 0.96-1.02, every decoder tensor incl. the new levels above 0.95), so that variant is held to 0.93
    check_direction(got, ref, f"{flags} training step", min_cos=0.93 if cfgkw.get("seg_last") else 0.95)


def test_cast_rows_with_column_sums():
    """lavt_cast_rows_colsum_bf16: the fp32 -> bf16 cast of a gradient (identity rows, or gathered into window order with zero pad rows,
    optionally times the DropPath scale of each row's sample) that also adds the column sums of what it writes into the bias gradient."""
    from lavt_rs_b200 import _cabi as K
    from lavt_rs_b200.geometry import window_geometry, window_row_map
    g = torch.Generator().manual_seed(31)
    for C, dims, shifted in ((128, (2, 8, 14, 14), True), (512, (1, 4, 10, 13), True), (96, (2, 8, 7, 7), False), (1024, (3, 2, 7, 9), True)):
        B, D, H, W = dims
        geom = window_geometry(B, D, H, W, (8, 7, 7), shifted)
        n = geom.tokens()
        tok = n // B
        x = torch.randn(n, C, generator=g).cuda()
        sc = (torch.rand(B, generator=g) + 0.5).cuda()
        acc0 = torch.randn(C, generator=g).cuda()
        # identity rows
        out = torch.empty(n, C, device="cuda", dtype=torch.bfloat16)
        cs = acc0.clone()
        K.cast_rows_bf16(x, out, rscale=sc, rscale_rows=tok, colsum=cs)
        ref = x * sc.repeat_interleave(tok)[:, None]
        assert torch.equal(out, ref.to(torch.bfloat16))
        assert rel_l2(cs - acc0, ref.sum(0)) < 1e-5
        # window order: pad rows are zeros and add nothing
        rows, _, _ = window_row_map(geom)
        rows = rows.cuda()
        M = geom.rows()
        outw = torch.full((M, C), 7.0, device="cuda", dtype=torch.bfloat16)
        cs = acc0.clone()
        K.cast_rows_bf16(x, outw, geom, colsum=cs)
        refw = torch.zeros(M, C, device="cuda")
        live = rows >= 0
        refw[live] = x[rows[live]]
        assert torch.equal(outw, refw.to(torch.bfloat16))
        assert rel_l2(cs - acc0, refw.sum(0)) < 1e-5
        # and the plain cast gives the same rows
        outp = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
        K.cast_rows_bf16(x, outp, geom)
        assert torch.equal(outp, outw)


@pytest.mark.parametrize("stage,heads", [(0, 1), (1, 2)])
def test_pwam_batchnorm_att_norm_backward(stage, heads):
    """--att_norm_layer_type BN in training mode (2-D backbone, reference lib/backbone.py:1297-1316: nn.BatchNorm1d in train(), batch statistics
    over all clips and tokens): the normalisation reaches the PWAM kernels as folded statistics, its adjoint reuses the decoder's BatchNorm
    kernels.  PWAM + gate gradients (incl. the BatchNorm weights / biases), dx, d l_feats and the running statistics vs the oracle."""
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200 import train_engine as T
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    from lavt_rs_b200.weights import load_reference_state_dict
    fh = [heads if i == stage else 1 for i in range(4)]
    cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False, att_norm="BN", fusion_heads=tuple(fh))
    sd = dict(O.random_state_dict(cfg, seed=0))
    bb = MultiModalSwinTransformer(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0,
                                   patch_norm=True, num_heads_fusion=fh, args=default_args(["--att_norm_layer_type", "BN"]))
    C = 128 * 2 ** stage
    pre = f"backbone.layers.{stage}."
    g = torch.Generator().manual_seed(37)
    for k in ("res_gate.0.weight", "res_gate.2.weight"):
        sd[pre + k] = torch.randn(C, C, generator=g) * C ** -0.5
    load_reference_state_dict(bb, sd, "backbone.")
    bb = bb.cuda().train()
    layer = bb.layers[stage]
    B, n, Nl = 3, 520, 12
    x = torch.randn(B, n, C, generator=g)
    l = torch.randn(B, 768, Nl, generator=g)
    m = torch.ones(B, Nl, 1)
    m[1, Nl - 3:] = 0
    gr = torch.randn(B, n, C, generator=g)
    gx = torch.randn(B, n, C, generator=g)

    def fn(sd2, xx, ll):
        r = O.pwam(xx, ll, m, sd2, pre + "fusion.", heads, att_norm="BN", train_norm=True)
        return torch.cat([r, O.language_gate(xx, r, sd2, pre + "res_gate.")], 0)
    (dx_ref, dl_ref), pg_ref = _oracle_grads(fn, sd, pre, [x, l], torch.cat([gr, gx], 0))
    pg_ref = {k: v for k, v in pg_ref.items() if k.startswith("fusion.") or k.startswith("res_gate.")}
    assert "fusion.image_lang_att.f_query.1.weight" in pg_ref and "fusion.image_lang_att.W.1.bias" in pg_ref
    ws = E.workspace("cuda")
    grads = T.GradStore()
    xf = x.cuda().reshape(-1, C).contiguous()
    bnq = layer.fusion.image_lang_att.f_query[1]
    rm0 = bnq.running_mean.clone()
    r32, xg, saved = T.pwam_gate_fwd(xf, xf.to(torch.bfloat16), layer.fusion, layer.res_gate, l.cuda(), m.squeeze(-1).cuda(), B, ws)
    ref = fn(sd, x, l)
    assert rel_l2(r32, ref[:B].reshape(-1, C)) < 1.5e-2
    assert rel_l2(xg, ref[B:].reshape(-1, C)) < 1.5e-2
    # running statistics moved like nn.BatchNorm1d (momentum 0.1) towards the batch mean of f_query's projection
    qraw = torch.nn.functional.conv1d(x.transpose(1, 2), sd[pre + "fusion.image_lang_att.f_query.0.weight"], sd[pre + "fusion.image_lang_att.f_query.0.bias"])
    want = 0.9 * rm0.cpu() + 0.1 * qraw.mean((0, 2))
    assert rel_l2(bnq.running_mean, want) < 2e-2 and int(bnq.num_batches_tracked) == 1
    dl = torch.zeros(B, 768, Nl, device="cuda")
    dx = T.pwam_gate_bwd(layer.fusion, layer.res_gate, saved, gr.cuda().reshape(-1, C).contiguous(), gx.cuda().reshape(-1, C).contiguous(), grads, ws, dl)
    torch.cuda.synchronize()
    assert rel_l2(dx, dx_ref.reshape(-1, C)) < GRAD_L2, ("dx", rel_l2(dx, dx_ref.reshape(-1, C)))
    assert rel_l2(dl, dl_ref) < GRAD_L2, ("dl", rel_l2(dl, dl_ref))
    check_grads(grads.named(layer), pg_ref, f"PWAM with BatchNorm attention norms, stage {stage}")


def test_conv_weight_gradient_tma():
    """lavt_conv3x3_wgrad / lavt_conv3d_wgrad (one launch, 4-D / 5-D TMA boxes as MN-major operands, taps as box offsets, zero padding
    from out-of-bounds fill) vs autograd of F.conv2d / F.conv3d, including sizes with partial pixel tiles."""
    import torch.nn.functional as F
    from lavt_rs_b200 import _cabi as K
    g = torch.Generator().manual_seed(13)
    for n, H, W, Cin, Cout in ((3, 12, 10, 512, 512), (2, 24, 24, 640, 512), (4, 7, 9, 1536, 512), (1, 5, 3, 64, 128)):
        x = torch.randn(n, Cin, H, W, generator=g).to(torch.bfloat16)
        dz = torch.randn(n, Cout, H, W, generator=g).to(torch.bfloat16)
        w = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
        F.conv2d(x.float(), w, padding=1).backward(dz.float())
        ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)
        dw = torch.full((Cout, 9 * Cin), 0.25, device="cuda")
        ws = torch.empty(K.conv3x3_wgrad_workspace_floats(n, H, W, Cin, Cout), device="cuda")
        K.conv3x3_wgrad(dz.permute(0, 2, 3, 1).contiguous().cuda(), x.permute(0, 2, 3, 1).contiguous().cuda(), dw, ws, accumulate=True)
        assert rel_l2(dw - 0.25, ref) < 1e-4, (n, H, W, Cin, Cout, rel_l2(dw - 0.25, ref))
    for B, D, H, W, Cin, Cout in ((2, 4, 6, 5, 128, 128), (1, 3, 9, 10, 64, 256)):
        x = torch.randn(B, Cin, D, H, W, generator=g).to(torch.bfloat16)
        dz = torch.randn(B, Cout, D, H, W, generator=g).to(torch.bfloat16)
        w = torch.zeros(Cout, Cin, 3, 3, 3, requires_grad=True)
        F.conv3d(x.float(), w, padding=1).backward(dz.float())
        ref = w.grad.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin)
        dw = torch.zeros(Cout, 27 * Cin, device="cuda")
        K.conv3d_wgrad(dz.permute(0, 2, 3, 4, 1).contiguous().cuda(), x.permute(0, 2, 3, 4, 1).contiguous().cuda(), dw,
                       lambda nfl: torch.empty(nfl, device="cuda"), accumulate=True)
        assert rel_l2(dw, ref) < 1e-4, (B, D, H, W, Cin, Cout, rel_l2(dw, ref))
