"""Training-step engine: forward with saved activations + hand-written backward on the sm_100a kernels.

The reference trains with autograd over PyTorch ops (train.py:330-360: ``loss.backward()``).  Here every forward unit of
``engine.py`` has a ``*_fwd`` twin that keeps what its adjoint needs (bf16 GEMM operands, the fp32 LayerNorm inputs) and a
``*_bwd`` that sequences the backward kernels of ``include/lavt_b200.h``:

  * activation gradients of a Linear / Conv1d(k=1):  dX = dY W        -> ``lavt_gemm_bf16`` with the transposed bf16 weight
  * weight gradients:                                 dW += dY^T X     -> ``lavt_transpose_bf16`` x 2 + ``lavt_gemm_bf16_splitk``
  * bias gradients:                                   db += colsum dY  -> ``lavt_colsum_accumulate``
  * LayerNorm (+ window gather / PatchMerging gather) -> ``lavt_layernorm_*_bwd``
  * window attention                                  -> ``lavt_window_attention_bwd``

The gradient on the residual stream is fp32 [tokens, C] and is updated in place from block to block.  Parameter gradients
accumulate in fp32 buffers owned by a ``GradStore`` and are handed to ``param.grad`` by ``GradStore.finalize()``.
No PyTorch math runs on this path and there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi as K
from . import engine as E
from .engine import Workspace, _bf16, _f32, _count
from .geometry import window_geometry


class GradStore:
    """fp32 accumulation buffers for parameter gradients (zeroed at creation = ``optimizer.zero_grad()``, train.py:352)."""

    def __init__(self):
        self._g: Dict[int, Tuple[torch.nn.Parameter, torch.Tensor]] = {}
        self._table_t: Dict[int, Tuple[torch.nn.Parameter, torch.Tensor]] = {}

    def of(self, param: torch.Tensor) -> torch.Tensor:
        """Accumulation buffer with the parameter's shape."""
        hit = self._g.get(id(param))
        if hit is None:
            hit = (param, torch.zeros(param.shape, device=param.device, dtype=torch.float32))
            self._g[id(param)] = hit
        return hit[1]

    def table_t(self, param: torch.Tensor) -> torch.Tensor:
        """relative_position_bias_table is [L, nH]; the attention kernels work on its transpose [nH, L]."""
        hit = self._table_t.get(id(param))
        if hit is None:
            hit = (param, torch.zeros(param.shape[1], param.shape[0], device=param.device, dtype=torch.float32))
            self._table_t[id(param)] = hit
        return hit[1]

    def finalize(self) -> None:
        """Hand the accumulated gradients to ``param.grad`` (added to an existing ``.grad`` like autograd does).  Tensor-container
        bookkeeping, not a hot path."""
        with torch.no_grad():
            for param, buf in self._table_t.values():
                self.of(param).add_(buf.t())
            for param, buf in self._g.values():
                if not param.requires_grad:
                    continue
                g = buf.to(param.dtype)
                if param.grad is None:
                    param.grad = g
                else:
                    param.grad.add_(g)
        self._g.clear()
        self._table_t.clear()

    def named(self, module: torch.nn.Module) -> Dict[str, torch.Tensor]:
        """name -> gradient buffer (after folding the transposed tables) for tests."""
        out = {}
        for name, prm in module.named_parameters():
            if id(prm) in self._g or id(prm) in self._table_t:
                g = self.of(prm).clone()
                if id(prm) in self._table_t:
                    g += self._table_t[id(prm)][1].t()
                out[name] = g
        return out


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _transposed(ws: Workspace, name: str, x: torch.Tensor) -> torch.Tensor:
    """bf16 [M, N] -> workspace view [N, M8] (M8 = M rounded up to the 16-byte TMA pitch; pad columns zeroed)."""
    M, N = x.shape
    M8 = _pad8(M)
    buf = ws.get(name, (N, M8), torch.bfloat16, x.device)
    if M8 != M:
        buf[:, M:].zero_()
    K.transpose_bf16(x, buf[:, :M])
    return buf


def linear_bwd(dy: torch.Tensor, x: Optional[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor], grads: GradStore,
               ws: Workspace, prepared: E.PreparedWeights, key: str, *, dx_bf16: Optional[torch.Tensor] = None,
               dx_f32: Optional[torch.Tensor] = None, dx_resid: Optional[torch.Tensor] = None, x_t: Optional[torch.Tensor] = None,
               **dx_epi) -> Optional[torch.Tensor]:
    """Adjoint of y = x W^T + b for bf16 rows dy [M, out], x [M, in]; ``weight`` is the [out, in(,1)] parameter.
    dW += dy^T x, db += colsum(dy); if a dx buffer is given: dx = dy W (through the GEMM epilogue options in ``dx_epi``).
    Returns the transposed x operand so that a caller with several consumers of the same x can reuse it."""
    M, Nout = dy.shape
    w2 = weight.view(weight.shape[0], -1) if weight.dim() != 2 else weight
    Kin = w2.shape[1]
    dev = dy.device
    if weight.requires_grad:
        dy_t = _transposed(ws, "bw_dyT", dy)
        if x_t is None:
            x_t = _transposed(ws, "bw_xT", x)
        wsf = K.splitk_workspace_floats(Nout, Kin, dy_t.shape[1])
        part = ws.get("bw_splitk", (wsf,), torch.float32, dev)
        K.gemm_bf16_splitk(dy_t, x_t, grads.of(weight).view(Nout, Kin), part, accumulate=True)
        _count(4)
    if bias is not None and bias.requires_grad:
        K.colsum_accumulate(dy, grads.of(bias))
        _count(1)
    if dx_bf16 is not None or dx_f32 is not None:
        w_t = prepared.get(key + "_wT", [weight], lambda: _bf16(w2.t()))      # [in, out]
        K.gemm_bf16(dy, w_t, out_bf16=dx_bf16, out_f32=dx_f32, resid=dx_resid, **dx_epi)
        _count(1)
    return x_t


# ------------------------------------------------------------------------------------------------
# Swin block  (reference SwinTransformerBlock3D.forward, lib/video_swin_transformer.py:214-273)
# ------------------------------------------------------------------------------------------------
def swin_block_fwd(x: torch.Tensor, blk, B: int, D: int, H: int, W: int, window, shifted: bool, clamp: bool, ws: Workspace,
                   xb_out: Optional[torch.Tensor] = None):
    """x fp32 [B*D*H*W, C] (not modified) -> (x_out, saved).  Same kernels as ``engine.swin_block`` except that fc1 stores
    its pre-activation (GELU runs as its own kernel) and nothing is overwritten."""
    n, C = x.shape
    dev = x.device
    geom = window_geometry(B, D, H, W, window, shifted, clamp)
    rows = geom.rows()
    nH = blk.num_heads
    pw = blk.prepared
    hd = C // nH
    qkv_w = pw.get("qkv_w", [blk.attn.qkv.weight], lambda: _bf16(blk.attn.qkv.weight))

    def _qscale():
        s = torch.ones(3 * C, device=dev, dtype=torch.float32)
        s[:C] = hd ** -0.5 * 1.4426950408889634
        b = _f32(blk.attn.qkv.bias) * s if blk.attn.qkv.bias is not None else torch.zeros(3 * C, device=dev)
        return s, b
    qkv_s, qkv_b = pw.get("qkv_sb", [blk.attn.qkv.weight] + ([blk.attn.qkv.bias] if blk.attn.qkv.bias is not None else []), _qscale)
    proj_w = pw.get("proj_w", [blk.attn.proj.weight], lambda: _bf16(blk.attn.proj.weight))
    fc1_w = pw.get("fc1_w", [blk.mlp.fc1.weight], lambda: _bf16(blk.mlp.fc1.weight))
    fc2_w = pw.get("fc2_w", [blk.mlp.fc2.weight], lambda: _bf16(blk.mlp.fc2.weight))
    table_t = pw.get("table_t", [blk.attn.relative_position_bias_table], lambda: _f32(blk.attn.relative_position_bias_table.t()))
    hidden = fc1_w.shape[0]

    xw = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    K.layernorm_window_gather(x, geom, blk.norm1.weight, blk.norm1.bias, xw, eps=blk.norm1.eps)
    qkv = torch.empty(rows, 3 * C, device=dev, dtype=torch.bfloat16)
    K.gemm_bf16(xw, qkv_w, cscale=qkv_s, bias=qkv_b, out_bf16=qkv)
    att = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    K.window_attention(qkv, table_t, geom, att)
    x1 = torch.empty_like(x)
    K.gemm_bf16(att, proj_w, bias=blk.attn.proj.bias.detach(), resid=x, out_f32=x1, win=geom)
    h1 = torch.empty(n, C, device=dev, dtype=torch.bfloat16)
    K.layernorm_rows(x1, blk.norm2.weight, blk.norm2.bias, out_bf16=h1, eps=blk.norm2.eps)
    hpre = torch.empty(n, hidden, device=dev, dtype=torch.bfloat16)
    K.gemm_bf16(h1, fc1_w, bias=blk.mlp.fc1.bias.detach(), out_bf16=hpre)
    hid = torch.empty(n, hidden, device=dev, dtype=torch.bfloat16)
    K.gelu_fwd(hpre, hid)
    x2 = torch.empty_like(x)
    K.gemm_bf16(hid, fc2_w, bias=blk.mlp.fc2.bias.detach(), resid=x1, out_f32=x2, out_bf16=xb_out)
    _count(8)
    return x2, (x, xw, qkv, att, x1, h1, hpre, hid, geom, table_t)


def swin_block_bwd(blk, saved, dx: torch.Tensor, grads: GradStore, ws: Workspace) -> torch.Tensor:
    """dx fp32 [n, C] = gradient of the block output; updated IN PLACE to the gradient of the block input and returned."""
    x0, xw, qkv, att, x1, h1, hpre, hid, geom, table_t = saved
    n, C = x0.shape
    dev = x0.device
    rows = geom.rows()
    hidden = hid.shape[1]
    pw = blk.prepared
    # ---- MLP half: x2 = x1 + fc2(GELU(fc1(LN2(x1))))
    dyb = ws.get("bw_dyb", (n, C), torch.bfloat16, dev)
    K.cast_rows_bf16(dx, dyb)
    dhid = ws.get("bw_dhid", (n, hidden), torch.bfloat16, dev)
    linear_bwd(dyb, hid, blk.mlp.fc2.weight, blk.mlp.fc2.bias, grads, ws, pw, "fc2", dx_bf16=dhid)
    K.gelu_bwd(dhid, hpre, dhid)
    dh1 = ws.get("bw_dh1", (n, C), torch.bfloat16, dev)
    linear_bwd(dhid, h1, blk.mlp.fc1.weight, blk.mlp.fc1.bias, grads, ws, pw, "fc1", dx_bf16=dh1)
    K.layernorm_rows_bwd(x1, dh1, blk.norm2.weight, dx, grads.of(blk.norm2.weight), grads.of(blk.norm2.bias), dres=dx,
                         eps=blk.norm2.eps)
    # ---- attention half: x1 = x0 + scatter(proj(attn(qkv(gather(LN1(x0))))))
    dyw = ws.get("bw_dyw", (rows, C), torch.bfloat16, dev)
    K.cast_rows_bf16(dx, dyw, geom)
    datt = ws.get("bw_datt", (rows, C), torch.bfloat16, dev)
    linear_bwd(dyw, att, blk.attn.proj.weight, blk.attn.proj.bias, grads, ws, pw, "proj", dx_bf16=datt)
    dqkv = ws.get("bw_dqkv", (rows, 3 * C), torch.bfloat16, dev)
    tbl = blk.attn.relative_position_bias_table
    K.window_attention_bwd(qkv, att, datt, table_t, geom, dqkv, grads.table_t(tbl) if tbl.requires_grad else None)
    dxw = ws.get("bw_dxw", (rows, C), torch.bfloat16, dev)
    linear_bwd(dqkv, xw, blk.attn.qkv.weight, blk.attn.qkv.bias, grads, ws, pw, "qkv", dx_bf16=dxw)
    K.layernorm_window_gather_bwd(x0, geom, dxw, blk.norm1.weight, dx, grads.of(blk.norm1.weight), grads.of(blk.norm1.bias),
                                  dres=dx, eps=blk.norm1.eps)
    _count(7)
    return dx


# ------------------------------------------------------------------------------------------------
# PatchMerging (reference :289-311)
# ------------------------------------------------------------------------------------------------
def patch_merging_fwd(x: torch.Tensor, ds, B: int, D: int, H: int, W: int, ws: Workspace):
    C = x.shape[1]
    dev = x.device
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    red_w = ds.prepared.get("red_w", [ds.reduction.weight], lambda: _bf16(ds.reduction.weight))
    g = torch.empty(B * D * H2 * W2, 4 * C, device=dev, dtype=torch.bfloat16)
    K.patch_merge_layernorm(x, B, D, H, W, ds.norm.weight, ds.norm.bias, g, eps=ds.norm.eps)
    out = torch.empty(B * D * H2 * W2, 2 * C, device=dev, dtype=torch.float32)
    K.gemm_bf16(g, red_w, out_f32=out)
    _count(2)
    return out, (x, g, (B, D, H, W))


def patch_merging_bwd(ds, saved, dout: torch.Tensor, grads: GradStore, ws: Workspace) -> torch.Tensor:
    """dout fp32 [rows/4, 2C] -> dx fp32 [rows, C] (new tensor)."""
    x, g, (B, D, H, W) = saved
    dev = x.device
    C = x.shape[1]
    dyb = ws.get("bw_dyb2", tuple(dout.shape), torch.bfloat16, dev)
    K.cast_rows_bf16(dout, dyb)
    dg = ws.get("bw_dmerge", tuple(g.shape), torch.bfloat16, dev)
    linear_bwd(dyb, g, ds.reduction.weight, None, grads, ws, ds.prepared, "red", dx_bf16=dg)
    dx = torch.empty_like(x)
    if (H % 2) or (W % 2):
        pass    # every real token still belongs to exactly one merged row; padded positions have no source token
    K.patch_merge_layernorm_bwd(x, B, D, H, W, dg, ds.norm.weight, dx, grads.of(ds.norm.weight), grads.of(ds.norm.bias), eps=ds.norm.eps)
    _count(2)
    return dx
