"""Language-aware Video Swin backbone -- B200 host modules.

Same class names, constructor arguments, attribute names and ``state_dict`` keys as the reference's
``lib/video_swin_transformer.py`` so that checkpoints and call sites carry over, but ``forward`` never
runs PyTorch math: the modules are parameter containers and the work is done by the sm_100a kernels
sequenced in ``lavt_rs_b200.engine`` (fused LN+window gather, tcgen05 GEMMs with fused epilogues,
flash-style window attention, PWAM kernels).  Inference only in this round (no autograd through the
kernels); unsupported ablation flags raise instead of silently falling back.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import _cabi as K
from .. import engine as E
from ..geometry import rel_const, window_geometry


class Mlp(nn.Module):
    """fc1 -> GELU -> fc2 parameters (reference :18-36)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        if drop != 0.0:
            raise NotImplementedError("dropout > 0 is not supported on the B200 path")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)


def _relative_position_index(window: Sequence[int]) -> torch.Tensor:
    """(N,N) int64 buffer kept only for state-dict compatibility (reference :107-127); the kernels use the
    closed form code(i) - code(j) + const instead."""
    Wd, Wh, Ww = window
    d, h, w = torch.meshgrid(torch.arange(Wd), torch.arange(Wh), torch.arange(Ww), indexing="ij")
    code = ((d * (2 * Wh - 1) + h) * (2 * Ww - 1) + w).reshape(-1)
    const = ((Wd - 1) * (2 * Wh - 1) + (Wh - 1)) * (2 * Ww - 1) + (Ww - 1)
    return code[:, None] - code[None, :] + const


class WindowAttention3D(nn.Module):
    """qkv / proj / relative-position table parameters (reference :86-135)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if qk_scale is not None or attn_drop != 0.0 or proj_drop != 0.0:
            raise NotImplementedError("qk_scale / attention dropout are not supported on the B200 path")
        if dim // num_heads != 32 or dim % num_heads:
            raise NotImplementedError(f"head_dim must be 32 (dim={dim}, heads={num_heads})")
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = (dim // num_heads) ** -0.5
        Wd, Wh, Ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * Wd - 1) * (2 * Wh - 1) * (2 * Ww - 1), num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class SwinTransformerBlock3D(nn.Module):
    """One (shifted-)window block (reference :171-273).  ``forward`` works on the channels-last tensor."""

    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size = tuple(window_size), tuple(shift_size)
        for s, w in zip(self.shift_size, self.window_size):
            assert 0 <= s < w, "shift_size must in 0-window_size"
        self.drop_path_rate = float(drop_path)
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, self.window_size, num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                      attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.prepared = E.PreparedWeights()
        self.clamp_window = True

    @property
    def shifted(self) -> bool:
        return any(s > 0 for s in self.shift_size)

    def forward(self, x: torch.Tensor, mask_matrix=None) -> torch.Tensor:
        """x (B,D,H,W,C) fp32 -> same shape (new tensor).  ``mask_matrix`` is accepted for API compatibility
        and ignored: the shifted-window mask is computed in closed form inside the attention kernel."""
        E.require_cuda(x, "x")
        if self.training and self.drop_path_rate > 0:
            raise NotImplementedError("stochastic depth (training) is not implemented on the B200 path")
        B, D, H, W, C = x.shape
        y = x.detach().to(torch.float32).reshape(B * D * H * W, C).clone()
        E.swin_block(y, self, B, D, H, W, self.window_size, self.shifted, self.clamp_window, E.workspace(x.device))
        return y.view(B, D, H, W, C)


class PatchMerging(nn.Module):
    """2x2 gather -> LN(4C) -> Linear(4C, 2C) (reference :276-311)."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,D,H,W,C) -> (B,D,ceil(H/2),ceil(W/2),2C)."""
        E.require_cuda(x, "x")
        B, D, H, W, C = x.shape
        out = torch.empty(B, D, (H + 1) // 2, (W + 1) // 2, 2 * C, device=x.device, dtype=torch.float32)
        E.patch_merging(x.detach().float().reshape(-1, C).contiguous(), self, B, D, H, W, E.workspace(x.device),
                        out.view(-1, 2 * C))
        return out


class SpatialImageLanguageAttention(nn.Module):
    """Pixel-word attention parameters (reference :937-1009)."""

    def __init__(self, v_in_channels, l_in_channels, key_channels, value_channels, out_channels=None, num_heads=1,
                 att_norm_layer_type="IN"):
        super().__init__()
        if att_norm_layer_type not in ("IN", "BN", "LN", "none"):
            raise ValueError(f"unknown --att_norm_layer_type {att_norm_layer_type!r}")
        self.att_norm_layer_type = att_norm_layer_type
        norm = {"IN": nn.InstanceNorm1d, "BN": nn.BatchNorm1d, "LN": nn.LayerNorm, "none": lambda c: nn.Identity()}[att_norm_layer_type]
        self.v_in_channels, self.l_in_channels = v_in_channels, l_in_channels
        self.key_channels, self.value_channels = key_channels, value_channels
        self.out_channels = out_channels if out_channels is not None else value_channels
        self.num_heads = num_heads
        if not (v_in_channels == key_channels == value_channels == self.out_channels):
            raise NotImplementedError("PWAM with differing channel widths is not supported on the B200 path")
        # index 1 of each Sequential is the norm of the reference: parameter-free InstanceNorm1d by default; the 2-D backbone's
        # --att_norm_layer_type offers BatchNorm1d / LayerNorm / Identity (reference lib/backbone.py:1297-1316)
        self.f_key = nn.Sequential(nn.Conv1d(l_in_channels, key_channels, 1))
        self.f_query = nn.Sequential(nn.Conv1d(v_in_channels, key_channels, 1), norm(key_channels))
        self.f_value = nn.Sequential(nn.Conv1d(l_in_channels, value_channels, 1))
        self.W = nn.Sequential(nn.Conv1d(value_channels, self.out_channels, 1), norm(self.out_channels))


class LangProject(nn.Module):
    """Mean-pooled sentence vector -> Linear -> ReLU -> Linear (reference :1012-1039; the --fuse simple ablation)."""

    def __init__(self, l_in_channels, l_out_channels):
        super().__init__()
        self.l_in_channels, self.l_out_channels = l_in_channels, l_out_channels
        self.project = nn.Sequential(nn.Linear(l_in_channels, l_out_channels), nn.ReLU(), nn.Linear(l_out_channels, l_out_channels))


class PWAM(nn.Module):
    """Pixel-word attention module (reference :889-934); ``attention=False`` = the --fuse simple ablation (LangProject)."""

    def __init__(self, dim, v_in_channels, l_in_channels, key_channels, value_channels, num_heads=0, dropout=0.0,
                 attention=True, att_norm_layer_type="IN"):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("fusion dropout > 0 is not supported on the B200 path")
        self.attention = attention
        self.vis_project = nn.Sequential(nn.Conv1d(dim, dim, 1, 1), nn.GELU(), nn.Dropout(dropout))
        if attention:
            self.image_lang_att = SpatialImageLanguageAttention(v_in_channels, l_in_channels, key_channels, value_channels,
                                                                out_channels=value_channels, num_heads=num_heads,
                                                                att_norm_layer_type=att_norm_layer_type)
        else:
            if dim != value_channels:
                raise NotImplementedError("--fuse simple with differing channel widths is not supported on the B200 path")
            self.image_lang_att = LangProject(l_in_channels, value_channels)
        self.project_mm = nn.Sequential(nn.Conv1d(value_channels, value_channels, 1, 1), nn.GELU(), nn.Dropout(dropout))
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor) -> torch.Tensor:
        """x (B,n,C); l (B,768,Nl); l_mask (B,Nl,1) -> x_residual (B,n,C)."""
        E.require_cuda(x, "x")
        B, n, C = x.shape
        xf = x.detach().float().reshape(B * n, C).contiguous()
        xb = xf.to(torch.bfloat16)
        r = torch.empty(B * n, C, device=x.device, dtype=torch.float32)
        E.pwam_gate(xf, xb, self, None, _lang(l), _mask(l_mask), B, E.workspace(x.device), r_f32=r)
        return r.view(B, n, C)


class SepTPWAM(nn.Module):
    """Separated temporal / spatial PWAM (reference :1300-1584) in the configuration the reference README trains the
    video models with: ``--sep_t_pwam --conv3d_kernel_size_t 3-3-3 --conv3d_kernel_size_s 1-1-1 --w_t3x3_s1x1
    --mm_t3x3_s1x1``.  Every PWAM projection is the sum of a Conv3d(3,3,3) and a Conv3d(1,1,1) branch; other kernel-size
    / gate / fuse sub-flags of the reference are rejected (no silent fallback)."""

    def __init__(self, dim, v_in_channels, l_in_channels, key_channels, value_channels, num_heads=0, dropout=0.0,
                 conv3d_kernel_size_t=(3, 1, 1), conv3d_kernel_size_s=(1, 1, 1), w_3x3=False, mm_3x3=False, w_3=False,
                 mm_3=False, sum_3_kernel_size=None, cat_reduce_kernel_size=None, w_t3x3_s1x1=None, mm_t3x3_s1x1=None,
                 args=None):
        super().__init__()
        if tuple(conv3d_kernel_size_t) != (3, 3, 3) or tuple(conv3d_kernel_size_s) != (1, 1, 1):
            raise NotImplementedError("SepTPWAM on the B200 path needs --conv3d_kernel_size_t 3-3-3 --conv3d_kernel_size_s 1-1-1")
        if not (w_t3x3_s1x1 and mm_t3x3_s1x1) or w_3x3 or mm_3x3 or w_3 or mm_3 or sum_3_kernel_size or cat_reduce_kernel_size:
            raise NotImplementedError("SepTPWAM on the B200 path needs --w_t3x3_s1x1 --mm_t3x3_s1x1 and no other W / mm / fuse flag")
        for f in ("s_tanh_plus_1_gate_1_q", "s_tanh_plus_1_gate_1_v", "t_tanh_plus_1_gate_1_q", "t_tanh_plus_1_gate_1_v"):
            if getattr(args, f, False):
                raise NotImplementedError(f"--{f} is not implemented on the B200 path")
        if dropout != 0.0:
            raise NotImplementedError("fusion dropout > 0 is not supported on the B200 path")
        if not (dim == v_in_channels == key_channels == value_channels):
            raise NotImplementedError("SepTPWAM with differing channel widths is not supported on the B200 path")
        self.num_heads = num_heads
        c3 = dict(kernel_size=(3, 3, 3), stride=1, padding=(1, 1, 1))
        c1 = dict(kernel_size=(1, 1, 1), stride=1, padding=0)
        self.temporal_vis_project = nn.Sequential(nn.Conv3d(dim, dim, **c3), nn.GELU(), nn.Dropout(dropout))
        self.spatial_vis_project = nn.Sequential(nn.Conv3d(dim, dim, **c1), nn.GELU(), nn.Dropout(dropout))
        self.f_query_t = nn.Sequential(nn.Conv3d(v_in_channels, key_channels, **c3), nn.InstanceNorm3d(key_channels))
        self.f_query_s = nn.Sequential(nn.Conv3d(v_in_channels, key_channels, **c1), nn.InstanceNorm3d(key_channels))
        self.f_key = nn.Sequential(nn.Conv1d(l_in_channels, key_channels, kernel_size=1, stride=1))
        self.f_value = nn.Sequential(nn.Conv1d(l_in_channels, value_channels, kernel_size=1, stride=1))
        self.W_t = nn.Sequential(nn.Conv3d(value_channels, value_channels, **c3), nn.InstanceNorm3d(value_channels))
        self.W_s = nn.Sequential(nn.Conv3d(value_channels, value_channels, **c1), nn.InstanceNorm3d(value_channels))
        self.project_mm_t = nn.Sequential(nn.Conv3d(value_channels, value_channels, **c3), nn.GELU(), nn.Dropout(dropout))
        self.project_mm_s = nn.Sequential(nn.Conv3d(value_channels, value_channels, **c1), nn.GELU(), nn.Dropout(dropout))
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor) -> torch.Tensor:
        """x (B,D,H,W,C); l (B,768,Nl); l_mask (B,Nl,1) -> x_residual (B,D*H*W,C)  (reference :1480-1584)."""
        E.require_cuda(x, "x")
        B, D, H, W, C = x.shape
        n = D * H * W
        xf = x.detach().float().reshape(B * n, C).contiguous()
        xb = xf.to(torch.bfloat16)
        r = torch.empty(B * n, C, device=x.device, dtype=torch.float32)
        E.sep_t_pwam_gate(xf, xb, self, None, _lang(l), _mask(l_mask), B, D, H, W, E.workspace(x.device), r_f32=r)
        return r.view(B, n, C)


def _lang(l: torch.Tensor) -> torch.Tensor:
    return l.detach().to(torch.float32).contiguous()


def _mask(l_mask: torch.Tensor) -> torch.Tensor:
    m = l_mask.detach()
    if m.dim() == 3:
        m = m.squeeze(-1)
    return m.to(torch.float32).contiguous()


_UNSUPPORTED_FLAGS = ("ts_pwam", "t_pwam", "t_pwam_comp", "seq_t_pwam", "sep_t_pwam_inner",
                      "sep_seq_t_pwam", "sep_seq_t_pwam_inner")


def _ksize(text, default):
    return tuple(int(a) for a in str(text).split("-")) if text else default


def check_args(args) -> None:
    """Reject model-structure flags whose reference behaviour is not implemented (no silent fallback)."""
    if args is None:
        return
    for f in _UNSUPPORTED_FLAGS:
        if getattr(args, f, False):
            raise NotImplementedError(f"--{f} is not implemented on the B200 path yet (SURVEY.md section 8f)")
    if getattr(args, "fuse", "default") not in ("default", "", "simple"):
        raise ValueError(f"unknown --fuse {args.fuse}")
    if getattr(args, "fuse", "default") == "simple" and getattr(args, "sep_t_pwam", False):
        raise NotImplementedError("--fuse simple together with --sep_t_pwam is not implemented on the B200 path")
    if getattr(args, "version", "default") not in ("default", "no_gate", "none"):
        raise ValueError(f"unknown --version {args.version}")


class MMBasicLayer(nn.Module):
    """One stage: Swin blocks -> PWAM -> LanguageGate -> PatchMerging (reference :331-592)."""

    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4.0, qkv_bias=False, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 num_heads_fusion=1, fusion_drop=0.0, sr_ratio=1, args=None):
        super().__init__()
        check_args(args)
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth = depth
        self.dim = dim
        self.use_checkpoint = use_checkpoint
        self.version = getattr(args, "version", "default")
        self.gate_act = "tanh"                      # the 2-D backbone's --lg_act_layer may switch it to sigmoid (lib/backbone.py:552-554)
        self.is_last_layer = num_heads in (24, 32)
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim=dim, num_heads=num_heads, window_size=self.window_size,
                                   shift_size=(0, 0, 0) if i % 2 == 0 else self.shift_size, mlp_ratio=mlp_ratio,
                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                   drop_path=drop_path[i] if isinstance(drop_path, (list, tuple)) else drop_path,
                                   norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.hs = bool(getattr(args, "hs", False))       # stage output = gated x (E_i) instead of the residual (:579-587)
        # --lazy_pred: stage output = features BEFORE fusion (V_i, :556-558); --hs wins when both are set (x_out is overwritten, :579-581)
        self.lazy_pred = bool(getattr(args, "lazy_pred", False)) and not self.hs
        self.sep_t_pwam = bool(getattr(args, "sep_t_pwam", False))
        if self.sep_t_pwam:      # reference :470-479
            self.fusion = SepTPWAM(dim, dim, 768, dim, dim, num_heads=num_heads_fusion, dropout=fusion_drop,
                                   conv3d_kernel_size_t=_ksize(getattr(args, "conv3d_kernel_size_t", ""), (3, 1, 1)),
                                   conv3d_kernel_size_s=_ksize(getattr(args, "conv3d_kernel_size_s", ""), (1, 1, 1)),
                                   w_3x3=getattr(args, "w_3x3", False), mm_3x3=getattr(args, "mm_3x3", False),
                                   w_3=getattr(args, "w_3", False), mm_3=getattr(args, "mm_3", False),
                                   sum_3_kernel_size=getattr(args, "sum_3_kernel_size", None) or None,
                                   cat_reduce_kernel_size=getattr(args, "cat_reduce_kernel_size", None) or None,
                                   w_t3x3_s1x1=getattr(args, "w_t3x3_s1x1", False),
                                   mm_t3x3_s1x1=getattr(args, "mm_t3x3_s1x1", False), args=args)
        else:
            self.fusion = PWAM(dim, dim, 768, dim, dim, num_heads=num_heads_fusion, dropout=fusion_drop,
                               attention=getattr(args, "fuse", "default") != "simple")      # reference :502-511
        self.has_gate = self.version == "default" and not (self.is_last_layer and use_checkpoint)
        if self.has_gate:
            self.res_gate = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.ReLU(), nn.Linear(dim, dim, bias=False), nn.Tanh())
            nn.init.zeros_(self.res_gate[0].weight)
            nn.init.zeros_(self.res_gate[2].weight)
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    # -- engine-level stage: works on the flat fp32 residual stream, returns (r fp32 [B*n,C], x_next, dims) --
    def run(self, x: torch.Tensor, B: int, D: int, H: int, W: int, l: torch.Tensor, mask: torch.Tensor, ws: E.Workspace,
            r_out: torch.Tensor, lang_ready=None, pre_fusion=None):
        """``pre_fusion``: optional callable invoked with the fp32 stream right after the Swin blocks (--lazy_pred reads V_i there,
        before the LanguageGate updates the stream in place)."""
        dev = x.device
        C = self.dim
        xb = ws.get("stage_xb", (B * D * H * W, C), torch.bfloat16, dev)
        if self.training and any(b.drop_path_rate > 0 for b in self.blocks):
            raise NotImplementedError("stochastic depth (training) is not implemented on the B200 path")
        for i, blk in enumerate(self.blocks):
            E.swin_block(x, blk, B, D, H, W, self.window_size, blk.shifted, blk.clamp_window, ws,
                         xb_out=xb if i == self.depth - 1 else None)
        if pre_fusion is not None:
            pre_fusion(x)
        if self.version == "swin":          # plain Swin stage (reference lib/backbone.py BasicLayer, :1409-1510): no language fusion at all
            if self.downsample is not None:
                H2, W2 = (H + 1) // 2, (W + 1) // 2
                nxt = ws.get("stage_x_%d" % (2 * C), (B * D * H2 * W2, 2 * C), torch.float32, dev)
                E.patch_merging(x, self.downsample, B, D, H, W, ws, nxt)
                return nxt, H2, W2
            return x, H, W
        if lang_ready is not None:          # first consumer of the language features
            torch.cuda.current_stream().wait_event(lang_ready)
        if self.sep_t_pwam:
            def fuse(gate):
                E.sep_t_pwam_gate(x, xb, self.fusion, gate, l, mask, B, D, H, W, ws, gate_act=self.gate_act, r_f32=r_out)
        else:
            def fuse(gate):
                E.pwam_gate(x, xb, self.fusion, gate, l, mask, B, ws, gate_act=self.gate_act, r_f32=r_out)
        if self.version == "none":      # fusion still produces the stage output; x is left untouched
            fuse(None)
        elif self.version == "no_gate":  # ablation flag: plain residual add (tensor-container op, not a hot path)
            fuse(None)
            x.add_(r_out)
        else:
            fuse(self.res_gate if self.has_gate else None)
        if self.downsample is not None:
            H2, W2 = (H + 1) // 2, (W + 1) // 2
            nxt = ws.get("stage_x_%d" % (2 * C), (B * D * H2 * W2, 2 * C), torch.float32, dev)
            E.patch_merging(x, self.downsample, B, D, H, W, ws, nxt)
            return nxt, H2, W2
        return x, H, W

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor):
        """Reference signature: x (B,C,D,H,W) -> (x_residual (B,C,D,H,W), x_next (B,C',D,H',W'))."""
        E.require_cuda(x, "x")
        B, C, D, H, W = x.shape
        ws = E.workspace(x.device)
        xf = x.detach().float().permute(0, 2, 3, 4, 1).reshape(B * D * H * W, C).contiguous()
        r = torch.empty(B * D * H * W, C, device=x.device, dtype=torch.float32)
        v_i = torch.empty_like(r) if self.lazy_pred else None           # V_i (:556-558); API-compat copy, not the hot path
        nxt, H2, W2 = self.run(xf, B, D, H, W, _lang(l), _mask(l_mask), ws, r,
                               pre_fusion=(lambda t: v_i.copy_(t)) if self.lazy_pred else None)
        if self.hs:
            r = xf                          # the gated stream E_i (:579-581); xf is this call's private copy
        elif self.lazy_pred:
            r = v_i
        r5 = r.view(B, D, H, W, C).permute(0, 4, 1, 2, 3)
        n5 = nxt.view(B, D, H2, W2, -1).permute(0, 4, 1, 2, 3).clone()
        return r5, n5


class PatchEmbed3D(nn.Module):
    """Conv3d(k = s = patch) + LN (reference :595-634); runs as im2col + tcgen05 GEMM + LN kernel."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        if tuple(patch_size) != (1, 4, 4) or in_chans != 3:
            raise NotImplementedError("the B200 path implements patch_size (1,4,4) with 3 input channels")
        self.patch_size, self.in_chans, self.embed_dim = tuple(patch_size), in_chans, embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,3,T,H,W) -> (B,C,T,H/4,W/4) (reference layout)."""
        E.require_cuda(x, "x")
        B, _, T, H, W = x.shape
        Hp, Wp = (H + 3) // 4, (W + 3) // 4
        out = torch.empty(B * T * Hp * Wp, self.embed_dim, device=x.device, dtype=torch.float32)
        E.patch_embed(_planes(x), self, E.workspace(x.device), out)
        return out.view(B, T, Hp, Wp, self.embed_dim).permute(0, 4, 1, 2, 3)


def _planes(x: torch.Tensor) -> torch.Tensor:
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(-1) != 1 or x.stride(-2) != x.shape[-1]:
        x = x.contiguous()
    return x


class MultiModalSwinTransformer3D(nn.Module):
    """Backbone (reference :637-886): ``forward(x, l, l_mask)`` -> tuple of (B*T, C_i, H_i, W_i) maps."""

    def __init__(self, pretrained=None, pretrained2d=False, patch_size=(4, 4, 4), in_chans=3, embed_dim=96,
                 depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(2, 7, 7), mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=False, out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False,
                 num_heads_fusion=[1, 1, 1, 1], fusion_drop=0.0, args=None):
        super().__init__()
        check_args(args)
        if drop_rate != 0.0:
            raise NotImplementedError("drop_rate > 0 is not supported on the B200 path")
        self.pretrained, self.pretrained2d = pretrained, pretrained2d
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.out_indices = tuple(out_indices)
        self.frozen_stages = frozen_stages
        self.window_size = tuple(window_size)
        self.patch_size = tuple(patch_size)
        self.patch_embed = PatchEmbed3D(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                        norm_layer=norm_layer if patch_norm else None)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(MMBasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint,
                num_heads_fusion=num_heads_fusion[i], fusion_drop=fusion_drop, args=args))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i in self.out_indices:
            self.add_module(f"norm{i}", norm_layer(self.num_features[i]))
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 1:
            for i in range(self.frozen_stages):
                m = self.layers[i]
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def inflate_weights(self):
        """Initialise from an ImageNet 2-D Swin checkpoint at ``self.pretrained`` (reference :759-805; Kinetics-style inflation)."""
        from ..weights import inflate_swin2d_state_dict
        checkpoint = torch.load(self.pretrained, map_location="cpu", weights_only=False)
        sd = inflate_swin2d_state_dict(checkpoint["model"], self.patch_size[0], self.window_size, self.state_dict())
        msg = self.load_state_dict(sd, strict=False)
        print(msg)
        print(f"=> loaded successfully '{self.pretrained}'")
        return msg

    def init_weights(self, pretrained=None):
        """trunc_normal(0.02) on Linear weights, zero biases, LayerNorm 1/0 (reference :811-852).  Loading a
        Video-Swin checkpoint follows the reference: keep ``backbone.*`` keys, sum the patch-embed kernel over time."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str) and self.pretrained:
            self.apply(_init)
            sd = torch.load(self.pretrained, map_location="cpu")["state_dict"]
            sd = {k[9:]: v for k, v in sd.items() if "backbone." in k}
            sd["patch_embed.proj.weight"] = sd["patch_embed.proj.weight"].sum(dim=2, keepdim=True)
            self.load_state_dict(sd, strict=False)
        elif self.pretrained is None or self.pretrained == "":
            self.apply(_init)
        else:
            raise TypeError("pretrained must be a str or None")

    # -- engine-level forward: NHWC outputs for the fused model path ------------------------------
    def run(self, x5: torch.Tensor, l: torch.Tensor, mask: torch.Tensor, want_nchw: bool = True, want_nhwc_bf16: bool = False,
            lang_ready=None):
        """x5: (B,3,T,H,W) strided fp32 view.  Returns (list of NCHW fp32 maps or None, list of NHWC bf16 maps or None).
        ``lang_ready``: optional CUDA event after which ``l`` is valid (the text encoder runs on a side stream while the
        patch embedding and the first Swin blocks -- which do not read the language features -- run on this one)."""
        dev = x5.device
        ws = E.workspace(dev)
        B, _, T, H, W = x5.shape
        Hc, Wc = (H + 3) // 4, (W + 3) // 4
        D = T
        C = self.embed_dim
        x = ws.get("stage_x_%d" % C, (B * D * Hc * Wc, C), torch.float32, dev)
        E.patch_embed(x5, self.patch_embed, ws, x)
        nchw: List[torch.Tensor] = []
        nhwc: List[torch.Tensor] = []
        for i, layer in enumerate(self.layers):
            C = layer.dim
            n = B * D * Hc * Wc
            r = ws.get("stage_r", (n, C), torch.float32, dev)

            def emit(src, i=i, n=n, C=C, Hc=Hc, Wc=Wc):
                norm = getattr(self, f"norm{i}")
                of = ws.get("out_f32", (n, C), torch.float32, dev) if want_nchw else None
                ob = None
                if want_nhwc_bf16:
                    ob = ws.get("out_bf16_%d" % i, (B * D, Hc, Wc, C), torch.bfloat16, dev)
                K.layernorm_rows(src, norm.weight, norm.bias, out_bf16=ob.view(n, C) if ob is not None else None, out_f32=of, eps=norm.eps)
                E._count(1)
                if want_nchw:
                    o = torch.empty(B * D, C, Hc, Wc, device=dev, dtype=torch.float32)
                    K.nhwc_to_nchw(of.view(B * D, Hc * Wc, C), o.view(B * D, C, Hc * Wc))
                    E._count(1)
                    nchw.append(o)
                if want_nhwc_bf16:
                    nhwc.append(ob)
            out_here = i in self.out_indices
            early = out_here and getattr(layer, "lazy_pred", False)          # V_i must be read before the gate rewrites the stream
            x_next, H2, W2 = layer.run(x, B, D, Hc, Wc, l, mask, ws, r, lang_ready=lang_ready if i == 0 else None,
                                       pre_fusion=emit if early else None)
            if out_here and not early:
                # stage output: the PWAM residual, or with --hs the gated features (x is not modified by the downsample)
                emit(x if (layer.hs or layer.version == "swin") else r)
            x, Hc, Wc = x_next, H2, W2
        return (nchw if want_nchw else None), (nhwc if want_nhwc_bf16 else None)

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor):
        """x (B,3,T,H,W); l (B,768,Nl); l_mask (B,Nl,1) -> tuple of (B*T, C_i, H_i, W_i) fp32 (reference :854-881)."""
        E.require_cuda(x, "x")
        nchw, _ = self.run(_planes(x), _lang(l), _mask(l_mask), want_nchw=True, want_nhwc_bf16=False)
        return tuple(nchw)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        return self
