"""Language-aware 2-D Swin backbone (image models ``lavt`` / ``lavt_one``) -- B200 host modules.

Mirror of the reference's lib/backbone.py classes on the hot path (``MultiModalSwinTransformer`` :334-521,
``MMBasicLayer`` :523-686, ``SwinTransformerBlock`` :146-245, ``WindowAttention`` :65-143, ``PatchMerging`` :248-288,
``PatchEmbed`` :291-331, ``PWAM`` :1238-1278).  The math is the 3-D path with a temporal extent of 1, so these are
thin specialisations of the video modules: window (1, w, w), and -- unlike the 3-D backbone -- the window is NEVER
clamped to the feature map (the reference pads H, W up to a window multiple, :205-208, and odd blocks always shift).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine as E
from . import video_swin_transformer as V

_REJECTED = ()


def _check_2d_args(args) -> None:
    V.check_args(args)
    if args is None:
        return
    for f in _REJECTED:
        if getattr(args, f, False):
            raise NotImplementedError(f"--{f} fusion (reference lib/bcam.py) is not implemented on the B200 path yet")
    if getattr(args, "lg_act_layer", "tanh") not in ("tanh", "sigmoid"):
        raise ValueError("--lg_act_layer must be tanh or sigmoid (reference lib/backbone.py:552-554)")
    if getattr(args, "att_norm_layer_type", "IN") not in ("IN", "BN", "LN", "none"):
        raise ValueError("--att_norm_layer_type must be IN, BN, LN or none (reference lib/backbone.py:1297-1302)")


class GACD(nn.Module):
    """GA-CD fusion parameters (reference lib/bcam.py:78-127; --gacd).  forward(x (B,n,C), l (B,768,Nl), l_mask (B,Nl,1)) -> (B,n,C)."""
    kind = "gacd"

    def __init__(self, dim, v_in_channels, l_in_channels, num_heads=0):
        super().__init__()
        if dim != v_in_channels:
            raise NotImplementedError("GA-CD with differing channel widths is not supported on the B200 path")
        self.k, self.dim = num_heads, dim
        self.lang_gen = V.LangProject(l_in_channels, v_in_channels)
        self.mm_gen = nn.Sequential(nn.Linear(v_in_channels, dim), nn.ReLU())
        self.query = nn.Linear(dim, dim)
        self.key_c = nn.Linear(v_in_channels, dim)
        self.key_d = nn.Linear(v_in_channels, dim)
        self.value = nn.Linear(v_in_channels, dim)
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor) -> torch.Tensor:
        E.require_cuda(x, "x")
        B, n, C = x.shape
        xf = x.detach().float().reshape(B * n, C).contiguous()
        r = torch.empty(B * n, C, device=x.device, dtype=torch.float32)
        E.gacd_gate(xf, xf.to(torch.bfloat16), self, None, V._lang(l), V._mask(l_mask), B, E.workspace(x.device), r_f32=r)
        return r.view(B, n, C)


class BCAM(nn.Module):
    """BCAM fusion parameters (reference lib/bcam.py:8-75; --bcam, from BRINet).  forward(x (B,hw,C), l (B,768,Nl), l_mask (B,Nl,1)) -> (B,hw,C).
    ``a_proj`` is a Linear(dim -> hw) with hw tied to the width as in the reference (:11-18), i.e. to the feature maps of a 480 x 480 input."""
    kind = "bcam"
    HW = {128: 120 * 120, 256: 60 * 60, 512: 30 * 30, 1024: 15 * 15}

    def __init__(self, dim, v_in_channels, l_in_channels):
        super().__init__()
        if dim not in self.HW:
            raise NotImplementedError(f"BCAM is defined for widths {sorted(self.HW)} only (reference lib/bcam.py:11-18), got {dim}")
        if dim != v_in_channels:
            raise NotImplementedError("BCAM with differing channel widths is not supported on the B200 path")
        self.dim, self.hw = dim, self.HW[dim]
        self.lang_reduce = nn.Linear(l_in_channels, dim)
        self.vis_1 = nn.Sequential(nn.Linear(v_in_channels, dim), nn.ReLU())
        self.vis_2 = nn.Sequential(nn.Linear(v_in_channels, dim), nn.ReLU())
        self.vis_3 = nn.Sequential(nn.Linear(v_in_channels, dim), nn.ReLU())
        self.vis_4 = nn.Sequential(nn.Linear(v_in_channels, dim), nn.ReLU())
        self.out_1 = nn.Linear(dim, dim)
        self.vis_2_2 = nn.Linear(dim, dim)
        self.a_proj = nn.Linear(dim, self.hw)
        self.out3_proj = nn.Sequential(nn.Linear(2 * dim, dim), nn.ReLU())
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor) -> torch.Tensor:
        E.require_cuda(x, "x")
        B, n, C = x.shape
        xf = x.detach().float().reshape(B * n, C).contiguous()
        r = torch.empty(B * n, C, device=x.device, dtype=torch.float32)
        E.bcam_gate(xf, xf.to(torch.bfloat16), self, None, V._lang(l), V._mask(l_mask), B, E.workspace(x.device), r_f32=r)
        return r.view(B, n, C)


class EFNAttention(nn.Module):
    """Parameter container of the reference's EFNAttention (lib/bcam.py:207-233); the math runs in ``engine.efn_gate``."""

    def __init__(self, in_channels, key_channels):
        super().__init__()
        self.in_channels, self.key_channels = in_channels, key_channels
        self.f_key = nn.Sequential(nn.Conv1d(in_channels, key_channels, kernel_size=1, stride=1), nn.InstanceNorm1d(key_channels))
        self.f_query = nn.Sequential(nn.Conv1d(in_channels, key_channels, kernel_size=1, stride=1), nn.InstanceNorm1d(key_channels))
        self.W = nn.Sequential(nn.Conv1d(2 * in_channels, in_channels, kernel_size=3, stride=1, padding=1), nn.InstanceNorm1d(in_channels))


class EFN(nn.Module):
    """EFN fusion parameters (reference lib/bcam.py:160-204; --efn).  forward(x (B,hw,C), l (B,768,Nl), l_mask (B,Nl,1)) -> (B,hw,C);
    hw must be a square map (the reference reshapes the token axis to sqrt(hw) x sqrt(hw) and pools it when hw > 225)."""
    kind = "efn"

    def __init__(self, dim, v_in_channels, l_in_channels):
        super().__init__()
        if dim != v_in_channels:
            raise NotImplementedError("EFN with differing channel widths is not supported on the B200 path")
        self.dim = dim
        self.project = nn.Sequential(nn.Conv1d(v_in_channels + l_in_channels, dim, 1, 1), nn.GELU())
        self.lang_project = nn.Sequential(nn.Conv1d(l_in_channels, dim, 1, 1), nn.GELU())
        self.image_lang_att = EFNAttention(in_channels=dim, key_channels=dim)
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor) -> torch.Tensor:
        E.require_cuda(x, "x")
        B, n, C = x.shape
        xf = x.detach().float().reshape(B * n, C).contiguous()
        r = torch.empty(B * n, C, device=x.device, dtype=torch.float32)
        E.efn_gate(xf, xf.to(torch.bfloat16), self, None, V._lang(l), V._mask(l_mask), B, E.workspace(x.device), r_f32=r)
        return r.view(B, n, C)


class PatchEmbed(nn.Module):
    """Conv2d(k = s = 4) + LN (reference :291-331)."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        if patch_size not in (4, (4, 4)) or in_chans != 3:
            raise NotImplementedError("the B200 path implements patch_size 4 with 3 input channels")
        self.patch_size, self.in_chans, self.embed_dim = (4, 4), in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=4, stride=4)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None
        self.prepared = E.PreparedWeights()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,3,H,W) -> (B,C,H/4,W/4)."""
        E.require_cuda(x, "x")
        B, _, H, W = x.shape
        Hp, Wp = (H + 3) // 4, (W + 3) // 4
        out = torch.empty(B * Hp * Wp, self.embed_dim, device=x.device, dtype=torch.float32)
        E.patch_embed(V._planes(x).unsqueeze(2), self, E.workspace(x.device), out)
        return out.view(B, Hp, Wp, self.embed_dim).permute(0, 3, 1, 2)


class MMBasicLayer(V.MMBasicLayer):
    """One 2-D stage (reference :523-686): blocks -> PWAM -> gate -> PatchMerging."""

    def __init__(self, dim, depth, num_heads, window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 num_heads_fusion=1, fusion_drop=0.0, args=None):
        _check_2d_args(args)
        super().__init__(dim=dim, depth=depth, num_heads=num_heads, window_size=(1, window_size, window_size),
                         mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                         drop_path=drop_path, norm_layer=norm_layer, downsample=downsample, use_checkpoint=False,
                         num_heads_fusion=num_heads_fusion, fusion_drop=fusion_drop, args=args)
        if getattr(args, "bcam", False):      # reference lib/backbone.py:573-577 (checked before --gacd there too)
            self.fusion = BCAM(dim, dim, 768)
        elif getattr(args, "gacd", False):    # reference lib/backbone.py:578-582
            self.fusion = GACD(dim, dim, 768, num_heads=num_heads_fusion)
        elif getattr(args, "efn", False):     # reference lib/backbone.py:583-588
            self.fusion = EFN(dim, dim, 768)
        elif getattr(args, "att_norm_layer_type", "IN") != "IN":          # reference lib/backbone.py:589-599
            self.fusion = V.PWAM(dim, dim, 768, dim, dim, num_heads=num_heads_fusion, dropout=fusion_drop,
                                 attention=getattr(args, "fuse", "default") != "simple", att_norm_layer_type=args.att_norm_layer_type)
        self.gate_act = getattr(args, "lg_act_layer", "tanh")
        if self.has_gate and self.gate_act == "sigmoid":
            self.res_gate[3] = nn.Sigmoid()
        for blk in self.blocks:
            blk.clamp_window = False          # the 2-D reference always pads to a full window and always shifts
        self.use_checkpoint = use_checkpoint

    def forward(self, x, H, W, l, l_mask):
        """Reference signature: x (B, H*W, C) -> (x_residual, H, W, x_down, Wh, Ww)."""
        E.require_cuda(x, "x")
        B, L, C = x.shape
        ws = E.workspace(x.device)
        xf = x.detach().float().reshape(B * L, C).contiguous().clone()
        r = torch.empty(B * L, C, device=x.device, dtype=torch.float32)
        v_i = torch.empty_like(r) if self.lazy_pred else None           # V_i (reference :662-664); API-compat copy, not the hot path
        nxt, H2, W2 = self.run(xf, B, 1, H, W, V._lang(l), V._mask(l_mask), ws, r,
                               pre_fusion=(lambda t: v_i.copy_(t)) if self.lazy_pred else None)
        if self.hs:
            r = xf                          # the gated stream (reference :678-681)
        elif self.lazy_pred:
            r = v_i
        return r.view(B, L, C), H, W, nxt.view(B, H2 * W2, -1).clone(), H2, W2


class MultiModalSwinTransformer(V.MultiModalSwinTransformer3D):
    """2-D backbone (reference :334-521): ``forward(x[B,3,H,W], l, l_mask)`` -> tuple of (B, C_i, H_i, W_i)."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False, num_heads_fusion=[1, 1, 1, 1],
                 fusion_drop=0.0, args=None):
        nn.Module.__init__(self)
        _check_2d_args(args)
        if ape:
            raise NotImplementedError("absolute position embedding (ape=True) is not implemented on the B200 path")
        if drop_rate != 0.0:
            raise NotImplementedError("drop_rate > 0 is not supported on the B200 path")
        self.pretrain_img_size = pretrain_img_size
        self.pretrained, self.pretrained2d = None, False
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape, self.patch_norm = ape, patch_norm
        self.out_indices = tuple(out_indices)
        self.frozen_stages = frozen_stages
        self.window_size = (1, window_size, window_size)
        self.patch_size = (1, 4, 4)
        self.patch_embed = PatchEmbed(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if patch_norm else None)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(MMBasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=V.PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint,
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
        if self.frozen_stages >= 2:
            for i in range(self.frozen_stages - 1):
                m = self.layers[i]
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def init_weights(self, pretrained=None):
        """trunc_normal(0.02) on Linear, zero bias, LayerNorm 1/0 (reference :460-488).  Loading ImageNet Swin weights
        goes through the reference's OpenMMLab loader, which is outside the hot path: load a state dict instead."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        if isinstance(pretrained, str) and pretrained:
            # reference :476-486: init, then load_checkpoint(self, pretrained, strict=False) -- the fusion / gate / per-stage norm
            # parameters are not in an ImageNet Swin checkpoint and keep their initialisation
            from ..weights import checkpoint_state_dict
            self.apply(_init)
            msg = self.load_state_dict(checkpoint_state_dict(pretrained), strict=False)
            print(f"=> loaded '{pretrained}': {len(msg.missing_keys)} keys kept their initialisation, {len(msg.unexpected_keys)} unused")
        elif pretrained is None or pretrained == "":
            self.apply(_init)
        else:
            raise TypeError("pretrained must be a str or None")

    def forward(self, x: torch.Tensor, l: torch.Tensor, l_mask: torch.Tensor):
        E.require_cuda(x, "x")
        nchw, _ = self.run(V._planes(x).unsqueeze(2), V._lang(l), V._mask(l_mask), want_nchw=True, want_nhwc_bf16=False)
        return tuple(nchw)


class SwinTransformer(MultiModalSwinTransformer):
    """Plain 2-D Swin backbone without language fusion (reference lib/backbone.py:1512-1650; the encoder of the ``vlt`` model,
    lib/segmentation.py:299-352): ``forward(x[B,3,H,W])`` -> tuple of LayerNorm-ed stage outputs (B, C_i, H_i, W_i).  Same
    state-dict keys as the reference class: ``patch_embed.*``, ``layers.{i}.blocks.*``, ``layers.{i}.downsample.*``, ``norm{i}.*``."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                 window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True, out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False):
        super().__init__(pretrain_img_size=pretrain_img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, depths=depths,
                         num_heads=num_heads, window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                         drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer, ape=ape,
                         patch_norm=patch_norm, out_indices=out_indices, frozen_stages=frozen_stages, use_checkpoint=use_checkpoint,
                         num_heads_fusion=[1, 1, 1, 1], fusion_drop=0.0, args=None)
        for layer in self.layers:               # a BasicLayer owns blocks + downsample only
            del layer.fusion
            if hasattr(layer, "res_gate"):
                del layer.res_gate
            layer.has_gate = False
            layer.version = "swin"

    def forward(self, x: torch.Tensor):
        E.require_cuda(x, "x")
        nchw, _ = self.run(V._planes(x).unsqueeze(2), None, None, want_nchw=True, want_nhwc_bf16=False)
        return tuple(nchw)
