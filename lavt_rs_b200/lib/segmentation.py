"""Model builders -- same entry points as the reference's lib/segmentation.py
(``segmentation.__dict__[args.model](pretrained=..., args=args)``, train.py:572 / test_ytvos.py:173).

Returned modules expose ``.backbone``, ``.classifier`` and ``.text_encoder`` with the reference's
``state_dict`` keys; their forward passes run on the sm_100a kernels.
"""
from __future__ import annotations

from ._utils import LAVT, LAVT_VLT, LAVTOne, LAVTVideo, VLT
from .mask_predictor import SimpleDecoding
from .video_swin_transformer import MultiModalSwinTransformer3D

__all__ = ["lavt", "lavt_one", "lavt_video", "vlt", "lavt_vlt"]

# swin_type -> (embed_dim, depths, num_heads, drop_path_rate)   (reference :156-172; image model :100-123)
_SWIN = {
    "tiny": (96, [2, 2, 6, 2], [3, 6, 12, 24], {"video": 0.1, "image": 0.3}),
    "small": (96, [2, 2, 18, 2], [3, 6, 12, 24], {"video": 0.2, "image": 0.3}),
    "base": (128, [2, 2, 18, 2], [4, 8, 16, 32], {"video": 0.3, "image": 0.3}),
    "large": (192, [2, 2, 18, 2], [6, 12, 24, 48], {"video": 0.3, "image": 0.3}),
}


def _fusion_heads(args):
    mha = getattr(args, "mha", "")
    return [int(a) for a in mha.split("-")] if mha else [1, 1, 1, 1]


def _out_indices(args):
    """--lazy_pred drops the 1/4-scale stage from the decoder's inputs (reference :116-118, :183-185)."""
    return (1, 2, 3) if getattr(args, "lazy_pred", False) else (0, 1, 2, 3)


def _swin_cfg(args, kind):
    st = getattr(args, "swin_type", "base")
    if st not in _SWIN or (kind == "video" and st == "large"):
        raise ValueError(f"unknown swin_type {st!r}")
    embed_dim, depths, heads, dpr = _SWIN[st]
    if embed_dim % 32:
        raise NotImplementedError(f"swin_type {st!r} (embed_dim {embed_dim}): the sm_100a kernels need channel counts that "
                                  "are multiples of 32")
    return embed_dim, list(depths), list(heads), dpr[kind]


def lavt_video(pretrained="", args=None):
    """Video model (reference _segm_lavt_video, :154-211): Video Swin + PWAM + SimpleDecoding + BERT."""
    embed_dim, depths, heads, dpr = _swin_cfg(args, "video")
    w = 12 if getattr(args, "window12", False) else 7
    backbone = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=embed_dim, depths=depths, num_heads=heads,
                                           window_size=(8, w, w), drop_path_rate=dpr, patch_norm=True,
                                           out_indices=_out_indices(args), use_checkpoint=getattr(args, "use_checkpoint", False),
                                           num_heads_fusion=_fusion_heads(args),
                                           fusion_drop=getattr(args, "fusion_drop", 0.0), args=args)
    backbone.init_weights(pretrained=pretrained if pretrained else None)
    classifier = SimpleDecoding(8 * embed_dim, args)
    return LAVTVideo(backbone, classifier, args)


def _image_backbone(pretrained, args):
    from .backbone import MultiModalSwinTransformer
    embed_dim, depths, heads, dpr = _swin_cfg(args, "image")
    w = 12 if (getattr(args, "window12", False) or "window12" in (pretrained or "")) else 7
    backbone = MultiModalSwinTransformer(embed_dim=embed_dim, depths=depths, num_heads=heads, window_size=w,
                                         drop_path_rate=dpr, patch_norm=True, out_indices=_out_indices(args),
                                         use_checkpoint=False, num_heads_fusion=_fusion_heads(args),
                                         fusion_drop=getattr(args, "fusion_drop", 0.0), args=args)
    backbone.init_weights(pretrained=pretrained if pretrained else None)
    return backbone, embed_dim


def lavt_one(pretrained="", args=None):
    """Image model with BERT inside (reference _segm_lavt_one, :100-150)."""
    backbone, embed_dim = _image_backbone(pretrained, args)
    return LAVTOne(backbone, SimpleDecoding(8 * embed_dim, args), args)


def lavt(pretrained="", args=None):
    """Image model taking precomputed language features (reference _segm_lavt, :14-60)."""
    backbone, embed_dim = _image_backbone(pretrained, args)
    return LAVT(backbone, SimpleDecoding(8 * embed_dim, args))


def _vlt_head(args, embed_dim):
    from .vlt import VLTFuseAndClassify
    if embed_dim != 128:
        raise NotImplementedError("VLTFuseAndClassify hard-codes the Swin-B stage widths 256 / 512 / 1024 (reference lib/vlt.py:16-18): "
                                  "use --swin_type base")
    return VLTFuseAndClassify(args=args)


def vlt(pretrained="", args=None):
    """Swin encoder (no language fusion, stages 1-3) + VLT head (reference _vlt, :299-352)."""
    from .backbone import SwinTransformer
    embed_dim, depths, heads, dpr = _swin_cfg(args, "image")
    w = 12 if (getattr(args, "window12", False) or "window12" in (pretrained or "")) else 7
    backbone = SwinTransformer(embed_dim=embed_dim, depths=depths, num_heads=heads, window_size=w, ape=False, drop_path_rate=dpr,
                               patch_norm=True, out_indices=(1, 2, 3), use_checkpoint=False)
    backbone.init_weights(pretrained=pretrained if pretrained else None)
    return VLT(backbone, _vlt_head(args, embed_dim), args=args)


def lavt_vlt(pretrained="", args=None):
    """LAVT encoder (stages 1-3) + VLT head (reference _lavt_vlt, :368-422)."""
    from .backbone import MultiModalSwinTransformer
    embed_dim, depths, heads, dpr = _swin_cfg(args, "image")
    w = 12 if (getattr(args, "window12", False) or "window12" in (pretrained or "")) else 7
    backbone = MultiModalSwinTransformer(embed_dim=embed_dim, depths=depths, num_heads=heads, window_size=w, ape=False, drop_path_rate=dpr,
                                         patch_norm=True, out_indices=(1, 2, 3), use_checkpoint=False, num_heads_fusion=_fusion_heads(args),
                                         fusion_drop=getattr(args, "fusion_drop", 0.0), args=args)
    backbone.init_weights(pretrained=pretrained if pretrained else None)
    return LAVT_VLT(backbone, _vlt_head(args, embed_dim), args=args)
