"""tcgen05 GEMM / implicit-GEMM conv (C-ABI lavt_gemm_bf16 / lavt_conv3x3_bf16) vs a plain fp32
PyTorch evaluation of the same bf16 operands.  Tolerance: outputs are compared in fp32 against an
fp32-accumulated reference of identical bf16 inputs, so only accumulation order and the final bf16
rounding differ: |a-b| <= 2e-2*|b| + 2e-2*rms(b) (bf16 outputs), 1e-3 relative-rms for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(out, ref, rtol):
    out, ref = out.float(), ref.float()
    rms = ref.pow(2).mean().sqrt().item() + 1e-12
    err = (out - ref).abs()
    bound = rtol * ref.abs() + rtol * rms
    bad = (err > bound).float().mean().item()
    assert bad == 0.0, f"{bad*100:.4f}% elements out of tolerance; max err {err.max().item():.4e}, rms {rms:.4e}"


@pytest.fixture(scope="module")
def cabi():
    from lavt_rs_b200 import _cabi
    _cabi.check(_cabi.lib().lavt_check_device(), "lavt_check_device")
    return _cabi


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 192), (1000, 128, 128), (4608, 1536, 512),
                                   (77, 384, 1024)])
def test_gemm_bias_gelu(cabi, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, bias=bias, act=cabi.ACT_GELU, out_bf16=out)
    ref = F.gelu(a.float() @ w.float().t() + bias)
    _close(out, ref, 2e-2)


@pytest.mark.parametrize("M,N,K", [(300, 256, 192), (4608, 2048, 512), (1000, 384, 128), (520, 512, 1024)])
def test_gemm_training_epilogues(cabi, M, N, K):
    """Training-path epilogues: (1) out_pre = the pre-activation next to the GELU'd output (fc1 forward), (2) mul_act = GELU:
    the result times GELU'(mul) (fc2 input gradient x derivative of the saved pre-activation; exact-erf GELU, reference :30-36)."""
    g = torch.Generator(device="cuda").manual_seed(3 * M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, bias=bias, act=cabi.ACT_GELU, out_bf16=out, out_pre=pre)
    ref_pre = a.float() @ w.float().t() + bias
    _close(pre, ref_pre, 2e-2)
    _close(out, F.gelu(ref_pre), 2e-2)
    # (2): the derivative of GELU at the saved (bf16) pre-activation, spread over [-6, 6]
    hpre = (3.0 * torch.randn(M, N, device="cuda", generator=g)).bfloat16()
    x = hpre.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    dgelu = x.grad
    dh = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, out_bf16=dh, mul=hpre, mul_act=cabi.ACT_GELU)
    ref = (a.float() @ w.float().t()) * dgelu
    _close(dh, ref, 2e-2)
    # (3) pre_mode = 1: the same launch writes GELU(pre) and GELU'(pre)
    dpre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, bias=bias, act=cabi.ACT_GELU, out_bf16=out, out_pre=dpre, pre_mode=1)
    x = ref_pre.clone().requires_grad_(True)
    F.gelu(x).sum().backward()
    _close(out, F.gelu(ref_pre), 2e-2)
    _close(dpre, x.grad, 2e-2)
    # and the plain mul epilogue is unchanged
    cabi.gemm_bf16(a, w, out_bf16=dh, mul=hpre)
    _close(dh, (a.float() @ w.float().t()) * hpre.float(), 2e-2)


@pytest.mark.parametrize("M,N,K,act", [(40000, 384, 128, 0), (32768, 512, 128, 1), (33000, 192, 256, 0), (36864, 1024, 256, 1), (50000, 64, 64, 2)])
def test_gemm_short_k_four_warpgroup_epilogue(cabi, M, N, K, act):
    """Short-K launches with a plain bf16 output in GEMM-row order (M >= 32768, K <= 256; bias / scale / GELU / ReLU only) run the
    four-epilogue-warpgroup instantiation of the GEMM: partial row tiles, partial column tiles (N = 192: one and a half 128-wide tiles),
    bit-identical to the general epilogue; launches with mul / a wider output pitch stay on the general path."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M + 8, N), 7.0, device="cuda", dtype=torch.bfloat16)     # rows past M must stay untouched
    cabi.gemm_bf16(a, w, bias=bias, act=act, out_bf16=out[:M])
    pre = a.float() @ w.float().t() + bias
    ref = F.gelu(pre) if act == 1 else torch.relu(pre) if act == 2 else pre
    _close(out[:M], ref, 2e-2)
    assert (out[M:] == 7.0).all()
    # scale + mul + residual, output pitch larger than N (column window of a wider buffer)
    cs = torch.rand(N, device="cuda", generator=g) + 0.5
    mul = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    wide = torch.full((M, N + 64), 3.0, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, cscale=cs, mul=mul, out_bf16=wide[:, :N])
    _close(wide[:, :N], (a.float() @ w.float().t()) * cs * mul.float(), 2e-2)
    assert (wide[:, N:] == 3.0).all()
    # the general epilogue gives the same bits (compared through a launch the variant does not take: an fp32 + bf16 dual output)
    o32 = torch.empty(M, N, device="cuda")
    ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, bias=bias, act=act, out_f32=o32, out_bf16=ob)
    assert torch.equal(ob, out[:M])


def test_gemm_scale_mul_tanh_resid(cabi):
    M, N, K = 513, 256, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    cs = torch.rand(N, device="cuda", generator=g) + 0.5
    mul = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    resid = torch.randn(M, N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.gemm_bf16(a, w, cscale=cs, act=cabi.ACT_TANH, mul=mul, resid=resid, out_f32=out, out_bf16=outb)
    ref = torch.tanh((a.float() @ w.float().t()) * cs) * mul.float() + resid
    _close(out, ref, 2e-3)
    _close(outb, ref, 2e-2)


@pytest.mark.parametrize("dims,window,shifted", [((2, 8, 14, 14), (8, 7, 7), True),
                                                  ((1, 8, 10, 13), (8, 7, 7), True),
                                                  ((1, 4, 24, 24), (8, 12, 12), True),
                                                  ((1, 16, 14, 14), (8, 7, 7), True),
                                                  ((2, 8, 12, 12), (8, 12, 12), False)])
def test_gemm_window_scatter(cabi, dims, window, shifted):
    from lavt_rs_b200.geometry import window_geometry, window_row_map
    B, D, H, W = dims
    C = 128
    geom = window_geometry(B, D, H, W, window, shifted)
    rows, _, _ = window_row_map(geom)
    M = geom.rows()
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    bias = torch.randn(C, device="cuda", generator=g)
    x = torch.randn(geom.tokens(), C, device="cuda", generator=g)
    out = torch.full_like(x, float("nan"))
    cabi.gemm_bf16(a, w, bias=bias, resid=x, out_f32=out, win=geom)
    y = a.float() @ w.float().t() + bias
    rows = rows.cuda()
    live = rows >= 0
    ref = x.clone()
    ref[rows[live]] += y[live]
    assert not torch.isnan(out).any(), "some token rows were never written"
    _close(out, ref, 2e-3)


@pytest.mark.parametrize("n_img,H,W,Cin,Cout", [(2, 24, 24, 64, 128), (1, 20, 12, 128, 256), (3, 48, 48, 192, 128),
                                                (1, 96, 96, 640, 512), (2, 15, 30, 64, 128)])
def test_conv3x3_bn_relu(cabi, n_img, H, W, Cin, Cout):
    g = torch.Generator(device="cuda").manual_seed(H * W + Cin)
    x = torch.randn(n_img, H, W, Cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    scale = torch.rand(Cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(Cout, device="cuda", generator=g) * 0.1
    w_taps = wt.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = torch.empty(n_img, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    cabi.conv3x3_bf16(x, w_taps, cscale=scale, bias=shift, act=cabi.ACT_RELU, out_bf16=out.view(-1, Cout))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1)
    _close(out, ref, 2e-2)


@pytest.mark.parametrize("n_clip,D,H,W,Cin,Cout,act", [(1, 4, 24, 24, 128, 128, 0), (2, 8, 12, 12, 256, 256, 1), (1, 8, 7, 10, 64, 128, 0),
                                                        (1, 2, 48, 48, 128, 128, 1), (2, 3, 1, 2, 64, 128, 0)])
def test_conv3d_333(cabi, n_clip, D, H, W, Cin, Cout, act):
    """Conv3d(3,3,3), pad 1 (SepTPWAM branches, reference lib/video_swin_transformer.py:1334-1461) vs torch fp32 on the same
    bf16 operands; frames / rows / columns outside the clip are zero (TMA out-of-bounds fill)."""
    g = torch.Generator(device="cuda").manual_seed(D * H * W + Cin)
    x = torch.randn(n_clip, D, H, W, Cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) / (27 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    w_taps = wt.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous()
    out = torch.empty(n_clip * D * H * W, Cout, device="cuda", dtype=torch.float32)
    cabi.conv3d_bf16(x, w_taps, bias=bias, act=cabi.ACT_GELU if act else cabi.ACT_NONE, out_f32=out)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt.float(), bias, padding=1)
    if act:
        ref = F.gelu(ref)
    _close(out.view(n_clip, D, H, W, Cout), ref.permute(0, 2, 3, 4, 1), 2e-2)


@pytest.mark.parametrize("M,N,K", [(1000, 96, 96), (777, 288, 96), (4096, 192, 384), (300, 1152, 480), (129, 32, 40)])
def test_gemm_tile_remainders(cabi, M, N, K):
    """N % 128 != 0, K % 64 != 0 (Swin-T/S widths): TMA zero fill past the operand edges + column guard in the epilogue."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    guard = torch.full((M + 1, N), 7.0, device="cuda")          # the row after the output must stay untouched
    out = guard[:M]
    cabi.gemm_bf16(a, w, bias=bias, act=cabi.ACT_GELU, resid=res, out_f32=out)
    ref = F.gelu(a.float() @ w.float().t() + bias) + res
    _close(out, ref, 2e-2)
    assert (guard[M] == 7.0).all()


def test_gemm_rejects_bad_shapes(cabi):
    a = torch.zeros(128, 100, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(128, 100, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(128, 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(cabi.LavtError):
        cabi.gemm_bf16(a, w, out_bf16=out)          # K % 8 != 0 (TMA needs 16-byte row pitches)
    with pytest.raises(cabi.LavtError):
        cabi.gemm_bf16(a.cpu(), w, out_bf16=out)    # no CPU fallback


@pytest.mark.parametrize("M,N,K", [(160, 2304, 768), (160, 768, 3072), (20, 3072, 768), (77, 768, 768)])
def test_small_m_splitk_gemm_matches_torch(M, N, K):
    """lavt_gemm_bf16_smallm (the text encoder's dense layers): split-K launch + reduce / epilogue kernel against the fp32 product of the same
    bf16 operands, with the epilogues BERT uses (column scale + bias -> bf16; bias + GELU -> bf16; bias + fp32 residual in place)."""
    from lavt_rs_b200 import _cabi as Kx
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).cuda().to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().to(torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    cs = (torch.rand(N, generator=g) + 0.5).cuda()
    ws = torch.empty(Kx.splitk_workspace_floats(M, N, K), device="cuda")
    ref = a.float() @ w.float().t()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    Kx.gemm_bf16_smallm(a, w, ws, cscale=cs, bias=bias, out_bf16=out)
    assert ((out.float() - (ref * cs + bias)).norm() / (ref * cs + bias).norm()).item() < 4e-3
    Kx.gemm_bf16_smallm(a, w, ws, bias=bias, act=Kx.ACT_GELU, out_bf16=out)
    want = torch.nn.functional.gelu(ref + bias)
    assert ((out.float() - want).norm() / want.norm()).item() < 4e-3
    x = torch.randn(M, N, generator=g).cuda()
    x0 = x.clone()
    Kx.gemm_bf16_smallm(a, w, ws, bias=bias, resid=x, out_f32=x)
    assert ((x - (x0 + ref + bias)).norm() / (x0 + ref + bias).norm()).item() < 1e-4
