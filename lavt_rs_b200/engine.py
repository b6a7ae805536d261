"""Host-side execution engine: sequences the sm_100a kernels for each reference function.

Layout in HBM
  * residual stream  x : fp32 [B*D*H*W, C] channels-last (the reference's (B,D,H,W,C) view)
  * GEMM operands      : bf16 rows (tokens x C), produced by the LN / gather / epilogue kernels
  * weights            : bf16 [out, in] copies of the fp32 parameters (``PreparedWeights``), refreshed
                         whenever a parameter's version counter changes
Scratch buffers come from a per-device ``Workspace`` so that a steady-state forward performs no
allocations (required for CUDA-graph capture in bench.py).

No reference code path is used here; each function cites the reference function it replaces.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _cabi as K
from .geometry import window_geometry

LAUNCHES = 0  # number of our kernel launches issued (bench.py reports it)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


class Workspace:
    """Named scratch buffers, grown on demand and reused across calls (one per device / stream user)."""

    def __init__(self):
        self._bufs: Dict[Tuple[str, torch.dtype], torch.Tensor] = {}

    def get(self, name: str, shape: Sequence[int], dtype: torch.dtype, device) -> torch.Tensor:
        numel = 1
        for s in shape:
            numel *= int(s)
        key = (name, dtype)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < numel or buf.device != torch.device(device):
            buf = torch.empty(max(numel, 1), dtype=dtype, device=device)
            self._bufs[key] = buf
        return buf[:numel].view(*shape)

    def bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self._bufs.values())


_WORKSPACES: Dict[Tuple[int, str], Workspace] = {}


def workspace(device, tag: str = "main") -> Workspace:
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), tag)
    if key not in _WORKSPACES:
        _WORKSPACES[key] = Workspace()
    return _WORKSPACES[key]


class PreparedWeights:
    """bf16 / folded copies of a module's parameters, rebuilt when any source tensor changes."""

    def __init__(self):
        self._cache: Dict[str, Tuple[tuple, object]] = {}

    def get(self, key: str, sources: Sequence[torch.Tensor], build):
        ver = tuple((t.data_ptr(), t._version, t.device) for t in sources)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._cache[key] = (ver, val)
        return val

    def clear(self):
        self._cache.clear()


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise K.LavtError(f"{what} must live on a CUDA device: the B200 path has no CPU fallback")


# ------------------------------------------------------------------------------------------------
# Swin block  (reference SwinTransformerBlock3D.forward, lib/video_swin_transformer.py:253-273)
# ------------------------------------------------------------------------------------------------
def swin_block(x: torch.Tensor, blk, B: int, D: int, H: int, W: int, window, shifted: bool, clamp: bool,
               ws: Workspace, xb_out: Optional[torch.Tensor] = None) -> None:
    """In place on the fp32 residual stream x [B*D*H*W, C].  ``blk`` is a host SwinTransformerBlock3D.
    If ``xb_out`` is given the final value of x is also written there in bf16 (feeds PWAM's GEMMs)."""
    n, C = x.shape
    dev = x.device
    geom = window_geometry(B, D, H, W, window, shifted, clamp)
    rows = geom.rows()
    nH = blk.num_heads
    pw = blk.prepared
    hd = C // nH

    qkv_w = pw.get("qkv_w", [blk.attn.qkv.weight], lambda: _bf16(blk.attn.qkv.weight))
    # q columns are scaled by head_dim^-0.5 (:147) times log2(e) in the epilogue -- the attention kernel's softmax
    # runs in base 2:  (acc + b) * s == acc * s + b * s
    def _qscale():
        s = torch.ones(3 * C, device=dev, dtype=torch.float32)
        s[:C] = hd ** -0.5 * 1.4426950408889634
        b = _f32(blk.attn.qkv.bias) * s if blk.attn.qkv.bias is not None else torch.zeros(3 * C, device=dev)
        return s, b
    qkv_s, qkv_b = pw.get("qkv_sb", [blk.attn.qkv.weight] + ([blk.attn.qkv.bias] if blk.attn.qkv.bias is not None else []), _qscale)
    proj_w = pw.get("proj_w", [blk.attn.proj.weight], lambda: _bf16(blk.attn.proj.weight))
    fc1_w = pw.get("fc1_w", [blk.mlp.fc1.weight], lambda: _bf16(blk.mlp.fc1.weight))
    fc2_w = pw.get("fc2_w", [blk.mlp.fc2.weight], lambda: _bf16(blk.mlp.fc2.weight))
    table_t = pw.get("table_t", [blk.attn.relative_position_bias_table],
                     lambda: _f32(blk.attn.relative_position_bias_table.t()))

    # --- attention half: LN1 + shift + partition gather -> qkv GEMM -> window attention -> proj GEMM + scatter + residual
    xw = ws.get("xw", (rows, C), torch.bfloat16, dev)
    K.layernorm_window_gather(x, geom, blk.norm1.weight, blk.norm1.bias, xw, eps=blk.norm1.eps)
    qkv = ws.get("qkv", (rows, 3 * C), torch.bfloat16, dev)
    K.gemm_bf16(xw, qkv_w, cscale=qkv_s, bias=qkv_b, out_bf16=qkv)
    att = ws.get("att", (rows, C), torch.bfloat16, dev)
    K.window_attention(qkv, table_t, geom, att)
    K.gemm_bf16(att, proj_w, bias=blk.attn.proj.bias.detach(), resid=x, out_f32=x, win=geom)
    # --- MLP half: LN2 -> fc1 + GELU -> fc2 + residual
    h1 = ws.get("ln2", (n, C), torch.bfloat16, dev)
    K.layernorm_rows(x, blk.norm2.weight, blk.norm2.bias, out_bf16=h1, eps=blk.norm2.eps)
    hid = ws.get("hid", (n, fc1_w.shape[0]), torch.bfloat16, dev)
    K.gemm_bf16(h1, fc1_w, bias=blk.mlp.fc1.bias.detach(), act=K.ACT_GELU, out_bf16=hid)
    K.gemm_bf16(hid, fc2_w, bias=blk.mlp.fc2.bias.detach(), resid=x, out_f32=x, out_bf16=xb_out)
    _count(7)


# ------------------------------------------------------------------------------------------------
# PWAM + LanguageGate  (reference PWAM.forward :919-934, SpatialImageLanguageAttention.forward :975-1009,
# res_gate :519-525 applied at :570)
# ------------------------------------------------------------------------------------------------
def _conv1x1_w(conv) -> torch.Tensor:
    return conv.weight[:, :, 0]


def pwam_gate(x: torch.Tensor, xb: torch.Tensor, fusion, res_gate, l: torch.Tensor, mask: torch.Tensor, B: int,
              ws: Workspace, gate_act: str = "tanh", r_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x fp32 [B*n, C] (updated in place with the gated residual), xb = bf16 copy of x, l fp32 [B,768,Nl],
    mask fp32 [B,Nl].  Returns r (= x_residual) as fp32 [B*n, C]."""
    if getattr(fusion, "kind", "pwam") == "gacd":
        return gacd_gate(x, xb, fusion, res_gate, l, mask, B, ws, gate_act=gate_act, r_f32=r_f32)
    if getattr(fusion, "kind", "pwam") == "bcam":
        return bcam_gate(x, xb, fusion, res_gate, l, mask, B, ws, gate_act=gate_act, r_f32=r_f32)
    if getattr(fusion, "kind", "pwam") == "efn":
        return efn_gate(x, xb, fusion, res_gate, l, mask, B, ws, gate_act=gate_act, r_f32=r_f32)
    N_, C = x.shape
    n = N_ // B
    dev = x.device
    pw = fusion.prepared
    att = fusion.image_lang_att
    Nl = l.shape[-1]

    def wprep(name, conv):
        return pw.get(name, [conv.weight], lambda: _bf16(_conv1x1_w(conv)))

    if not fusion.attention:
        # --fuse simple (reference :916-917, :929-930): lang = LangProject(mean-pooled words), one vector per clip
        vis = ws.get("pw_vis", (B, n, C), torch.bfloat16, dev)
        K.gemm_bf16(xb, wprep("vis_w", fusion.vis_project[0]), bias=fusion.vis_project[0].bias.detach(), act=K.ACT_GELU, out_bf16=vis.view(N_, C))
        stats = ws.get("pw_stats", (B, 2, C), torch.float32, dev)
        pr = att.project
        K.lang_project(l, mask, _f32(pr[0].weight), _f32(pr[0].bias), _f32(pr[2].weight), _f32(pr[2].bias), stats)
        zeros = pw.get("zeros_%d_%d" % (B, n), [], lambda: torch.zeros(B, n, C, device=dev, dtype=torch.float32))
        a2 = ws.get("pw_o", (B, n, C), torch.bfloat16, dev)
        K.pwam_mul_norm(vis, zeros, stats, a2)          # vis * (0 - (-lang)) * 1
        r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
        rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
        K.gemm_bf16(a2.view(N_, C), wprep("mm_w", fusion.project_mm[0]), bias=fusion.project_mm[0].bias.detach(), act=K.ACT_GELU, out_f32=r32,
                    out_bf16=rb)
        _count(6)
        if res_gate is not None:
            g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
            g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
            g1 = vis.view(N_, C)
            K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=g1)
            if gate_act != "tanh":
                raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
            K.gemm_bf16(g1, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
            _count(2)
        return r32
    heads = att.num_heads

    vis_w = wprep("vis_w", fusion.vis_project[0])
    q_w = wprep("q_w", att.f_query[0])
    W_w = wprep("W_w", att.W[0])
    mm_w = wprep("mm_w", fusion.project_mm[0])
    k_w = pw.get("k_w", [att.f_key[0].weight], lambda: _f32(_conv1x1_w(att.f_key[0])))
    v_w = pw.get("v_w", [att.f_value[0].weight], lambda: _f32(_conv1x1_w(att.f_value[0])))

    vis = ws.get("pw_vis", (B, n, C), torch.bfloat16, dev)
    K.gemm_bf16(xb, vis_w, bias=fusion.vis_project[0].bias.detach(), act=K.ACT_GELU, out_bf16=vis.view(N_, C))
    qpre = ws.get("pw_q", (B, n, C), torch.float32, dev)      # fp32: feeds InstanceNorm + SIMT attention, never an MMA operand
    K.gemm_bf16(xb, q_w, bias=att.f_query[0].bias.detach(), out_f32=qpre.view(N_, C))
    stats = ws.get("pw_stats", (B, 2, C), torch.float32, dev)
    stw = ws.get("pw_statw", (K.instnorm_workspace_floats(B, n, C),), torch.float32, dev)
    K.instnorm_stats(qpre, stats, stw)
    kk = ws.get("pw_k", (B, Nl, C), torch.float32, dev)
    vv = ws.get("pw_v", (B, Nl, C), torch.float32, dev)
    K.pwam_kv(l, mask, k_w, att.f_key[0].bias.detach(), v_w, att.f_value[0].bias.detach(), kk, vv)
    o = ws.get("pw_o", (B, n, C), torch.bfloat16, dev)
    K.pwam_attend(qpre, stats, kk, vv, mask, o, heads)
    lang = qpre  # q_pre is dead: reuse its buffer for lang_pre
    K.gemm_bf16(o.view(N_, C), W_w, bias=att.W[0].bias.detach(), out_f32=lang.view(N_, C))
    K.instnorm_stats(lang, stats, stw)
    a2 = o  # o is dead after the W projection
    K.pwam_mul_norm(vis, lang, stats, a2)
    r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
    rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
    K.gemm_bf16(a2.view(N_, C), mm_w, bias=fusion.project_mm[0].bias.detach(), act=K.ACT_GELU, out_f32=r32, out_bf16=rb)
    _count(11)
    if res_gate is not None:
        g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
        g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
        g1 = vis.view(N_, C)  # vis is dead
        K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=g1)
        if gate_act != "tanh":
            raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
        K.gemm_bf16(g1, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
        _count(2)
    return r32


def gacd_gate(x: torch.Tensor, xb: torch.Tensor, fusion, res_gate, l: torch.Tensor, mask: torch.Tensor, B: int, ws: Workspace,
              gate_act: str = "tanh", r_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GA-CD fusion (reference lib/bcam.py:78-127, --gacd) + LanguageGate; same contract as ``pwam_gate``."""
    N_, C = x.shape
    n = N_ // B
    dev = x.device
    pw = fusion.prepared
    stats = ws.get("pw_stats", (B, 2, C), torch.float32, dev)
    pr = fusion.lang_gen.project
    K.lang_project(l, mask, _f32(pr[0].weight), _f32(pr[0].bias), _f32(pr[2].weight), _f32(pr[2].bias), stats)      # (-ls, 1)
    zeros = pw.get("zeros_%d_%d" % (B, n), [], lambda: torch.zeros(B, n, C, device=dev, dtype=torch.float32))
    a = ws.get("pw_o", (B, n, C), torch.bfloat16, dev)
    K.pwam_mul_norm(xb.view(B, n, C), zeros, stats, a)                                        # ls * x
    xm = ws.get("pw_q", (B, n, C), torch.float32, dev)
    mm = fusion.mm_gen[0]
    K.gemm_bf16(a.view(N_, C), pw.get("mm_w", [mm.weight], lambda: _bf16(mm.weight)), bias=mm.bias.detach(), act=K.ACT_RELU,
                out_f32=xm.view(N_, C))
    r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
    rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
    K.gacd_fuse(xm, stats, _f32(fusion.query.weight), _f32(fusion.query.bias), _f32(fusion.key_c.weight), _f32(fusion.key_c.bias),
                _f32(fusion.key_d.weight), _f32(fusion.key_d.bias), _f32(fusion.value.weight), _f32(fusion.value.bias),
                lambda nfl: ws.get("gacd_ws", (nfl,), torch.float32, dev), out_f32=r32, out_bf16=rb)
    _count(10)
    if res_gate is not None:
        g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
        g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
        g1 = a.view(N_, C)           # the ls * x operand is dead
        K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=g1)
        if gate_act != "tanh":
            raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
        K.gemm_bf16(g1, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
        _count(2)
    return r32


def bcam_gate(x: torch.Tensor, xb: torch.Tensor, fusion, res_gate, l: torch.Tensor, mask: torch.Tensor, B: int, ws: Workspace,
              gate_act: str = "tanh", r_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """BCAM fusion (reference lib/bcam.py:43-75, --bcam) + LanguageGate; same contract as ``pwam_gate``.  Eleven tcgen05 GEMMs with
    the three ``csrc/bcam_kernels.cu`` kernels between them; sums of two Linear layers run as one GEMM over row-concatenated operands
    (one [n, 3C] buffer holds out2 | out | query2, so cat[out2, out] and [out | query2] are column windows of it, never copies).
    The hw x hw relation map (:61-62) is materialised once as fp32 logits and once as bf16 probabilities (1.2 GB at 120 x 120)."""
    N_, C = x.shape
    n = N_ // B
    dev = x.device
    if n != fusion.hw:
        raise K.LavtError(f"BCAM: a_proj maps to {fusion.hw} positions but the feature map has {n} (the reference pins --bcam to 480 x 480 inputs)")
    pw = fusion.prepared
    Nl = l.shape[-1]
    Nlp = (Nl + 31) // 32 * 32           # N granule of the sim GEMM (also a K granule of the out GEMM)
    hwp = (n + 31) // 32 * 32            # N granule of a_proj
    hw8 = (n + 7) // 8 * 8               # K granule of rel_map @ query3

    def wb(name, lin):
        return pw.get(name, [lin.weight], lambda: _bf16(lin.weight))

    def b_(lin):
        return lin.bias.detach()

    lr = ws.get("bc_lr", (B, Nlp, C), torch.bfloat16, dev)
    lrT = ws.get("bc_lrT", (B, C, Nlp), torch.bfloat16, dev)
    K.bcam_words(l, _f32(fusion.lang_reduce.weight), _f32(fusion.lang_reduce.bias), lr, lrT)
    # VLAM (:51-56): sim = softmax(relu(vis_1 x) lr^T + mask), out = sim lr
    q = ws.get("pw_vis", (N_, C), torch.bfloat16, dev)
    K.gemm_bf16(xb, wb("v1_w", fusion.vis_1[0]), bias=b_(fusion.vis_1[0]), act=K.ACT_RELU, out_bf16=q)
    sim = ws.get("bc_sim", (N_, Nlp), torch.float32, dev)
    for b in range(B):
        K.gemm_bf16(q[b * n:(b + 1) * n], lr[b], out_f32=sim[b * n:(b + 1) * n])
    simp = ws.get("bc_simp", (N_, Nlp), torch.bfloat16, dev)
    K.bcam_softmax_rows(sim, Nl, simp, mask=mask, rows_per_mask=n)
    cat = ws.get("bc_cat", (N_, 3 * C), torch.bfloat16, dev)        # out2 | out | query2
    for b in range(B):
        K.gemm_bf16(simp[b * n:(b + 1) * n], lrT[b], out_bf16=cat[b * n:(b + 1) * n, C:2 * C])
    # LVAM (:59-64): A = tanh(out_1(out) + vis_2_2(relu(vis_2 x))); rel_map = softmax(a_proj(A)); out2 = rel_map relu(vis_3 x)
    K.gemm_bf16(xb, wb("v2_w", fusion.vis_2[0]), bias=b_(fusion.vis_2[0]), act=K.ACT_RELU, out_bf16=cat[:, 2 * C:])
    w_a = pw.get("a_w", [fusion.out_1.weight, fusion.vis_2_2.weight], lambda: _bf16(torch.cat([fusion.out_1.weight, fusion.vis_2_2.weight], 1)))
    b_a = pw.get("a_b", [fusion.out_1.bias, fusion.vis_2_2.bias], lambda: _f32(fusion.out_1.bias + fusion.vis_2_2.bias))
    a = ws.get("pw_o", (N_, C), torch.bfloat16, dev)
    K.gemm_bf16(cat[:, C:], w_a, bias=b_a, act=K.ACT_TANH, out_bf16=a)

    def _aproj():
        w = torch.zeros(hwp, C, device=dev, dtype=torch.bfloat16)
        w[:n] = fusion.a_proj.weight.detach()
        bb = torch.zeros(hwp, device=dev, dtype=torch.float32)
        bb[:n] = fusion.a_proj.bias.detach()
        return w, bb
    ap_w, ap_b = pw.get("ap", [fusion.a_proj.weight, fusion.a_proj.bias], _aproj)
    logits = ws.get("bc_logits", (N_, hwp), torch.float32, dev)
    K.gemm_bf16(a, ap_w, bias=ap_b, out_f32=logits)
    rel = ws.get("bc_rel", (N_, hwp), torch.bfloat16, dev)
    K.bcam_softmax_rows(logits, n, rel)
    K.gemm_bf16(xb, wb("v3_w", fusion.vis_3[0]), bias=b_(fusion.vis_3[0]), act=K.ACT_RELU, out_bf16=q)         # query is dead: query3
    q3T = ws.get("bc_q3T", (B, C, hw8), torch.bfloat16, dev)
    K.bcam_transpose_pad(q, q3T)
    for b in range(B):
        K.gemm_bf16(rel[b * n:(b + 1) * n, :hw8], q3T[b], out_bf16=cat[b * n:(b + 1) * n, :C])
    # out3 = relu(out3_proj(cat[out2, out])) + relu(vis_4 x)  (:65-71)
    q4 = ws.get("pw_q", (N_, C), torch.float32, dev)
    K.gemm_bf16(xb, wb("v4_w", fusion.vis_4[0]), bias=b_(fusion.vis_4[0]), act=K.ACT_RELU, out_f32=q4)
    r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
    rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
    K.gemm_bf16(cat[:, :2 * C], wb("o3_w", fusion.out3_proj[0]), bias=b_(fusion.out3_proj[0]), act=K.ACT_RELU, resid=q4, out_f32=r32, out_bf16=rb)
    _count(11 + 3 * B)
    if res_gate is not None:
        g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
        g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
        K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=q)
        if gate_act != "tanh":
            raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
        K.gemm_bf16(q, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
        _count(2)
    return r32


def efn_gate(x: torch.Tensor, xb: torch.Tensor, fusion, res_gate, l: torch.Tensor, mask: torch.Tensor, B: int, ws: Workspace,
             gate_act: str = "tanh", r_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EFN fusion (reference lib/bcam.py:172-269, --efn) + LanguageGate; same contract as ``pwam_gate``.  The token axis is read as a square
    image (h = sqrt(n)) and 2 x 2 pooled when n > 225, as in the reference; the co-attention maps are (n/4)^2 at most (3600^2 at 480 x 480)."""
    N_, C = x.shape
    n = N_ // B
    dev = x.device
    h = int(round(n ** 0.5))
    pooled = n > 225                         # (:243)
    if h * h != n or (pooled and h % 2):
        raise K.LavtError(f"EFN: {n} tokens are not a square (even-sided when pooled) feature map (reference lib/bcam.py:239-249)")
    pw = fusion.prepared
    att = fusion.image_lang_att
    Nl = l.shape[-1]
    Nlp = (Nl + 31) // 32 * 32
    n2 = n // 4 if pooled else n             # co-attention positions
    n2p = (n2 + 31) // 32 * 32
    scale = C ** -0.5

    def cs(N):
        return pw.get("cs_%d" % N, [], lambda: torch.full((N,), scale, device=dev, dtype=torch.float32))

    # M = gelu(project(cat[x, sentence])) (:179-185): the sentence half of the Conv1d is a per-image bias
    pj = fusion.project[0]
    sb = ws.get("ef_sb", (B, C), torch.float32, dev)
    K.efn_sentence_bias(l, mask, pw.get("pj_wl", [pj.weight], lambda: _f32(pj.weight[:, C:, 0])), _f32(pj.bias), sb)
    wx = pw.get("pj_wx", [pj.weight], lambda: _bf16(pj.weight[:, :C, 0]))
    M = ws.get("pw_vis", (N_, C), torch.bfloat16, dev)
    for b in range(B):
        K.gemm_bf16(xb[b * n:(b + 1) * n], wx, bias=sb[b], act=K.ACT_GELU, out_bf16=M[b * n:(b + 1) * n])
    # lang = gelu(lang_project(l)) * mask; L = softmax(C^-0.5 M lang + pad mask) lang^T (:186-194)
    lp = fusion.lang_project[0]
    lang = ws.get("bc_lr", (B, Nlp, C), torch.bfloat16, dev)
    langT = ws.get("bc_lrT", (B, C, Nlp), torch.bfloat16, dev)
    K.bcam_words(l, pw.get("lp_w", [lp.weight], lambda: _f32(lp.weight[:, :, 0])), _f32(lp.bias), lang, langT, mask=mask, act=K.ACT_GELU)
    sim = ws.get("bc_sim", (N_, Nlp), torch.float32, dev)
    for b in range(B):
        K.gemm_bf16(M[b * n:(b + 1) * n], lang[b], cscale=cs(Nlp), out_f32=sim[b * n:(b + 1) * n])
    # f_key(L) = sum_j p_j f_key(lang_j): the key side stays fp32 up to its InstanceNorm (see csrc/bcam_kernels.cu: efn_word_attend_kernel)
    fk = att.f_key[0]
    G = ws.get("ef_G", (B, Nlp, C), torch.float32, dev)
    K.gemm_bf16(lang.view(B * Nlp, C), pw.get("fk_w", [fk.weight], lambda: _bf16(_conv1x1_w(fk))), bias=fk.bias.detach(), out_f32=G.view(B * Nlp, C))
    # EFNAttention (:236-269): q = pool(IN(f_query M)), k = pool(IN(f_key L))
    stw = ws.get("pw_statw", (K.instnorm_workspace_floats(B, n, C),), torch.float32, dev)
    stats = ws.get("pw_stats", (B, 2, C), torch.float32, dev)
    pre = ws.get("pw_q", (B, n, C), torch.float32, dev)
    qk = []
    for name in ("fq", "fk"):
        if name == "fq":
            conv = att.f_query[0]
            K.gemm_bf16(M, pw.get("fq_w", [conv.weight], lambda: _bf16(_conv1x1_w(conv))), bias=conv.bias.detach(), out_f32=pre.view(N_, C))
        else:
            K.efn_word_attend(sim, mask, G, pre.view(N_, C))
        K.instnorm_stats(pre, stats, stw)
        t = ws.get("ef_" + name, (B, n2p, C), torch.bfloat16, dev)
        K.efn_norm_pool(pre, stats, t, h, pooled)
        qk.append(t)
    q, k = qk
    qT = ws.get("ef_qT", (B, C, n2p), torch.bfloat16, dev)
    kT = ws.get("ef_kT", (B, C, n2p), torch.bfloat16, dev)
    K.bcam_transpose_pad(q.view(B * n2p, C), qT)
    K.bcam_transpose_pad(k.view(B * n2p, C), kT)
    logits = ws.get("bc_logits", (n2, n2p), torch.float32, dev)
    prob = ws.get("bc_rel", (n2, n2p), torch.bfloat16, dev)
    cat = ws.get("bc_cat", (B, n2, 2 * C), torch.bfloat16, dev)            # Lp | Mp
    for b in range(B):
        # Lp = softmax_rows(sim) k ; Mp = softmax_cols(sim)^T q = softmax_rows(sim^T) q   (:251-258)
        for a_, w_, vT, col in ((q, k, kT, 0), (k, q, qT, C)):
            K.gemm_bf16(a_[b, :n2], w_[b], cscale=cs(n2p), out_f32=logits)
            K.bcam_softmax_rows(logits, n2, prob)
            K.gemm_bf16(prob, vT[b], out_bf16=cat[b, :, col:col + C])
    # W: Conv1d(2C -> C, k = 3, pad 1) over the flattened token axis = centre tap + two row-shifted accumulations; then IN (+ upsample)
    Wc = att.W[0]
    taps = pw.get("W_taps", [Wc.weight], lambda: [_bf16(Wc.weight[:, :, t]) for t in range(3)])
    o = ws.get("pw_q", (B, n2, C), torch.float32, dev)                     # q / k pre-activations are dead
    for b in range(B):
        K.gemm_bf16(cat[b], taps[1], bias=Wc.bias.detach(), out_f32=o[b])
        K.gemm_bf16(cat[b, :n2 - 1], taps[0], resid=o[b, 1:], out_f32=o[b, 1:])          # out[i] += W[:, :, 0] cat[i - 1]
        K.gemm_bf16(cat[b, 1:], taps[2], resid=o[b, :n2 - 1], out_f32=o[b, :n2 - 1])      # out[i] += W[:, :, 2] cat[i + 1]
    K.instnorm_stats(o, stats, stw)
    r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
    rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
    K.efn_norm_upsample(o, stats, n, h, pooled, out_f32=r32, out_bf16=rb)
    _count(13 + 11 * B)
    if res_gate is not None:
        g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
        g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
        K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=M)
        if gate_act != "tanh":
            raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
        K.gemm_bf16(M, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
        _count(2)
    return r32


def sep_t_pwam_gate(x: torch.Tensor, xb: torch.Tensor, fusion, res_gate, l: torch.Tensor, mask: torch.Tensor, B: int, D: int,
                    H: int, W: int, ws: Workspace, gate_act: str = "tanh", r_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SepTPWAM + LanguageGate under the README video flags (reference SepTPWAM.forward :1480-1584): every projection of
    ``pwam_gate`` is the sum of a Conv3d(3,3,3) branch (implicit GEMM over NDHWC with 5-D TMA boxes) and a Conv3d(1,1,1)
    branch (plain GEMM).  x fp32 [B*n, C] is updated in place with the gated residual; returns r fp32 [B*n, C]."""
    N_, C = x.shape
    n = N_ // B
    dev = x.device
    pw = fusion.prepared
    heads = fusion.num_heads
    Nl = l.shape[-1]

    def w333(name, conv):     # (Cout, Cin, kz, ky, kx) -> [Cout, ((kz*3+ky)*3+kx)*Cin + ci]
        return pw.get(name, [conv.weight], lambda: _bf16(conv.weight.detach().permute(0, 2, 3, 4, 1).reshape(conv.weight.shape[0], -1)))

    def w111(name, conv):
        return pw.get(name, [conv.weight], lambda: _bf16(conv.weight.detach().reshape(conv.weight.shape[0], -1)))

    def b_(conv):
        return conv.bias.detach()

    xb5 = xb.view(B, D, H, W, C)
    t32 = ws.get("sp_t32", (N_, C), torch.float32, dev)              # temporal-branch result waiting for the spatial one
    # ts_vis = GELU(conv333(x)) + GELU(conv111(x))                     (:1487-1503)
    vt, vs_ = fusion.temporal_vis_project[0], fusion.spatial_vis_project[0]
    K.conv3d_bf16(xb5, w333("vis_t", vt), bias=b_(vt), act=K.ACT_GELU, out_f32=t32)
    vis = ws.get("pw_vis", (B, n, C), torch.bfloat16, dev)
    K.gemm_bf16(xb, w111("vis_s", vs_), bias=b_(vs_), act=K.ACT_GELU, resid=t32, out_bf16=vis.view(N_, C))
    # query = IN3d(conv333(x)) + IN3d(conv111(x))                      (:1512-1524)
    qa = ws.get("pw_q", (B, n, C), torch.float32, dev)
    qb = ws.get("sp_qb", (B, n, C), torch.float32, dev)
    sa = ws.get("pw_stats", (B, 2, C), torch.float32, dev)
    sb = ws.get("sp_stats_b", (B, 2, C), torch.float32, dev)
    stw = ws.get("pw_statw", (K.instnorm_workspace_floats(B, n, C),), torch.float32, dev)
    qt, qs = fusion.f_query_t[0], fusion.f_query_s[0]
    K.conv3d_bf16(xb5, w333("q_t", qt), bias=b_(qt), out_f32=qa.view(N_, C))
    K.gemm_bf16(xb, w111("q_s", qs), bias=b_(qs), out_f32=qb.view(N_, C))
    K.instnorm_stats(qa, sa, stw)
    K.instnorm_stats(qb, sb, stw)
    K.instnorm_sum2(qa, sa, qb, sb, qa)
    ident = pw.get("ident_%d_%s" % (B, dev), [], lambda: torch.stack([torch.zeros(B, C), torch.ones(B, C)], 1).to(dev).contiguous())
    # words: k, v and the masked softmax are those of PWAM             (:1534-1553)
    k_w = pw.get("k_w", [fusion.f_key[0].weight], lambda: _f32(_conv1x1_w(fusion.f_key[0])))
    v_w = pw.get("v_w", [fusion.f_value[0].weight], lambda: _f32(_conv1x1_w(fusion.f_value[0])))
    kk = ws.get("pw_k", (B, Nl, C), torch.float32, dev)
    vv = ws.get("pw_v", (B, Nl, C), torch.float32, dev)
    K.pwam_kv(l, mask, k_w, b_(fusion.f_key[0]), v_w, b_(fusion.f_value[0]), kk, vv)
    o = ws.get("pw_o", (B, n, C), torch.bfloat16, dev)
    K.pwam_attend(qa, ident, kk, vv, mask, o, heads)
    # lang = IN3d(conv333(o)) + IN3d(conv111(o))                       (:1556-1561)
    Wt, Ws = fusion.W_t[0], fusion.W_s[0]
    K.conv3d_bf16(o.view(B, D, H, W, C), w333("W_t", Wt), bias=b_(Wt), out_f32=qa.view(N_, C))
    K.gemm_bf16(o.view(N_, C), w111("W_s", Ws), bias=b_(Ws), out_f32=qb.view(N_, C))
    K.instnorm_stats(qa, sa, stw)
    K.instnorm_stats(qb, sb, stw)
    K.instnorm_sum2(qa, sa, qb, sb, qa)
    a2 = o  # o is dead after the W projections
    K.pwam_mul_norm(vis, qa, ident, a2)
    # r = GELU(conv333(mm)) + GELU(conv111(mm))                         (:1574-1578)
    mt, ms = fusion.project_mm_t[0], fusion.project_mm_s[0]
    K.conv3d_bf16(a2.view(B, D, H, W, C), w333("mm_t", mt), bias=b_(mt), act=K.ACT_GELU, out_f32=t32)
    r32 = r_f32 if r_f32 is not None else ws.get("pw_r32", (N_, C), torch.float32, dev)
    rb = ws.get("pw_rb", (N_, C), torch.bfloat16, dev)
    K.gemm_bf16(a2.view(N_, C), w111("mm_s", ms), bias=b_(ms), act=K.ACT_GELU, resid=t32, out_f32=r32, out_bf16=rb)
    _count(19)
    if res_gate is not None:
        g0 = pw.get("g0", [res_gate[0].weight], lambda: _bf16(res_gate[0].weight))
        g2 = pw.get("g2", [res_gate[2].weight], lambda: _bf16(res_gate[2].weight))
        g1 = vis.view(N_, C)  # vis is dead
        K.gemm_bf16(rb, g0, act=K.ACT_RELU, out_bf16=g1)
        if gate_act != "tanh":
            raise K.LavtError("only the tanh LanguageGate is implemented on the B200 path")
        K.gemm_bf16(g1, g2, act=K.ACT_TANH, mul=rb, resid=x, out_f32=x)
        _count(2)
    return r32


# ------------------------------------------------------------------------------------------------
# PatchMerging (reference :289-311) and PatchEmbed3D (:616-634)
# ------------------------------------------------------------------------------------------------
def patch_merging(x: torch.Tensor, ds, B: int, D: int, H: int, W: int, ws: Workspace, out: torch.Tensor) -> None:
    """x fp32 [B*D*H*W, C] -> out fp32 [B*D*ceil(H/2)*ceil(W/2), 2C]."""
    C = x.shape[1]
    dev = x.device
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    red_w = ds.prepared.get("red_w", [ds.reduction.weight], lambda: _bf16(ds.reduction.weight))
    g = ws.get("merge", (B * D * H2 * W2, 4 * C), torch.bfloat16, dev)
    K.patch_merge_layernorm(x, B, D, H, W, ds.norm.weight, ds.norm.bias, g, eps=ds.norm.eps)
    K.gemm_bf16(g, red_w, out_f32=out)
    _count(2)


def patch_embed(x5: torch.Tensor, pe, ws: Workspace, out: torch.Tensor) -> Tuple[int, int]:
    """x5 fp32 (B,3,T,H,W) strided view; out fp32 [B*T*Hp*Wp, C].  Returns (Hp, Wp)."""
    B, _, T, H, W = x5.shape
    dev = x5.device
    Hp, Wp = (H + 3) // 4, (W + 3) // 4
    C = pe.embed_dim

    def _w():
        w = pe.proj.weight.detach().reshape(C, -1).to(torch.bfloat16)     # [C, 48], col = c*16 + ph*4 + pw
        wp = torch.zeros(C, 64, device=w.device, dtype=torch.bfloat16)
        wp[:, : w.shape[1]] = w
        return wp
    pw_ = pe.prepared.get("w", [pe.proj.weight], _w)
    cols = ws.get("pe_cols", (B * T * Hp * Wp, 64), torch.bfloat16, dev)
    K.patch_embed_im2col(x5, cols)
    if pe.norm is not None:
        K.gemm_bf16(cols, pw_, bias=pe.proj.bias.detach(), out_f32=out)
        K.layernorm_rows(out, pe.norm.weight, pe.norm.bias, out_f32=out, eps=pe.norm.eps)   # row-wise, safe in place
        _count(3)
    else:
        K.gemm_bf16(cols, pw_, bias=pe.proj.bias.detach(), out_f32=out)
        _count(2)
    return Hp, Wp


# ------------------------------------------------------------------------------------------------
# SimpleDecoding (reference lib/mask_predictor.py:56-99) on NHWC bf16 feature maps
# ------------------------------------------------------------------------------------------------
def _bn_fold(bn):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return s.contiguous(), (bn.bias.detach().float() - bn.running_mean.detach().float() * s).contiguous()


def _conv_taps(conv) -> torch.Tensor:
    w = conv.weight.detach()                                    # [Cout, Cin, 3, 3]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()   # [(ky*3+kx)*Cin + ci]


def _cbr(x_nhwc: torch.Tensor, dec, conv_name: str, bn_name: str, out: torch.Tensor) -> None:
    conv, bn = getattr(dec, conv_name), getattr(dec, bn_name)
    if bn.training:
        raise K.LavtError("SimpleDecoding on the B200 path is inference-only (BatchNorm must be in eval mode)")
    w = dec.prepared.get(conv_name, [conv.weight], lambda: _conv_taps(conv))
    s, b = dec.prepared.get(bn_name, [bn.weight, bn.bias, bn.running_mean, bn.running_var], lambda: _bn_fold(bn))
    K.conv3x3_bf16(x_nhwc, w, cscale=s, bias=b, act=K.ACT_RELU, out_bf16=out.view(-1, out.shape[-1]))
    _count(1)


def decoder_nhwc(dec, c4: torch.Tensor, c3: torch.Tensor, c2: torch.Tensor, c1: torch.Tensor, ws: Workspace,
                 logits_nchw: Optional[torch.Tensor], feats: Optional[list] = None) -> torch.Tensor:
    """c_i bf16 NHWC [n_img, H_i, W_i, C_i] (c1 = None under --lazy_pred); optional logits_nchw fp32 [n_img, 2, H_1, W_1].
    Returns the low-resolution logits as an NHWC fp32 workspace view [n_img, H_1, W_1, 2].
    ``feats``: if a list, the three top-down maps after conv2_4 / conv2_3 / conv2_2 (NHWC bf16 workspace views) are
    appended (SimpleDecoding.forward_feats, lib/mask_predictor.py:102-150)."""
    dev = c4.device
    n_img = c4.shape[0]
    hid = dec.conv1_4.weight.shape[0]
    y = c4
    if (c1 is None) != bool(getattr(dec, "lazy_pred", False)):
        raise K.LavtError("decoder: the 1/4-scale map is omitted exactly when the decoder was built with --lazy_pred")
    for skip, (ca, ba, cb, bb) in ((c3, ("conv1_4", "bn1_4", "conv2_4", "bn2_4")),
                                   (c2, ("conv1_3", "bn1_3", "conv2_3", "bn2_3")),
                                   (c1, ("conv1_2", "bn1_2", "conv2_2", "bn2_2"))):
        if skip is None:                      # --lazy_pred: no 1/4-scale level (reference lib/mask_predictor.py:77)
            continue
        _, H, W, Cs = skip.shape
        if y.shape[1] > H or y.shape[2] > W:
            raise K.LavtError("decoder: coarser map is larger than the skip connection")
        cat = ws.get("dec_cat", (n_img, H, W, y.shape[-1] + Cs), torch.bfloat16, dev)
        K.upsample_concat(y, skip, cat)
        t1 = ws.get("dec_t1", (n_img, H, W, hid), torch.bfloat16, dev)
        _cbr(cat, dec, ca, ba, t1)
        t2 = ws.get("dec_t2_%d" % H, (n_img, H, W, hid), torch.bfloat16, dev)
        _cbr(t1, dec, cb, bb, t2)
        y = t2
        if feats is not None:
            feats.append(t2)
        _count(1)
    _, H, W, _ = y.shape
    w11 = dec.prepared.get("w11", [dec.conv1_1.weight], lambda: _f32(dec.conv1_1.weight.reshape(2, -1)))
    lg = ws.get("dec_logits", (n_img, H, W, 2), torch.float32, dev)
    K.conv1x1_logits(y.view(-1, hid), w11, dec.conv1_1.bias.detach(), lg.view(-1, 2))
    _count(1)
    if logits_nchw is not None:
        K.upsample_logits(lg, logits_nchw)      # same size: exact NHWC -> NCHW transposition
        _count(1)
    return lg
