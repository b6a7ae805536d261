"""Argument namespace for the model builders.

The reference reads model-structure switches straight from its argparse namespace inside ``lib/``
(``args.swin_type``, ``args.window12``, ``args.mha``, ``args.version`` ...; reference args.py).  This
parser declares the switches the hot path consults, with the reference's names and defaults, so
``segmentation.lavt_video(pretrained, args)`` can be driven the same way.  Flags whose behaviour is not
implemented on the B200 path are still declared and rejected by ``check_args`` (no silent fallback).
"""
from __future__ import annotations

import argparse

_STRUCTURE_FLAGS = [
    # name, type/action, default, help
    ("--model", str, "lavt_video", "lavt | lavt_one | lavt_video"),
    ("--swin_type", str, "base", "tiny | small | base | large"),
    ("--window12", "store_true", False, "window 12 instead of 7 (video: (8,12,12))"),
    ("--mha", str, "", "PWAM heads per stage, e.g. 1-1-1-1"),
    ("--fusion_drop", float, 0.0, "dropout inside PWAM (must be 0 here)"),
    ("--version", str, "default", "default | no_gate | none"),
    ("--fuse", str, "default", "default | simple"),
    ("--use_checkpoint", "store_true", False, "activation checkpointing flag of the reference (affects last-stage gate)"),
    ("--ck_bert", str, "bert-base-uncased", "BERT checkpoint directory"),
    ("--bert_tokenizer", str, "bert-base-uncased", "tokenizer name"),
    ("--pretrained_swin_weights", str, "", "Video-Swin checkpoint"),
    ("--img_size", int, 480, "input size"),
    ("--lg_act_layer", str, "tanh", "LanguageGate activation (2-D backbone)"),
    ("--att_norm_layer_type", str, "IN", "PWAM attention norm (2-D backbone)"),
    ("--hs", "store_true", False, "stage outputs = gated features E_i instead of the PWAM residuals"),
    ("--lazy_pred", "store_true", False, "stage outputs = features before fusion at stages 1-3; decoder stops at 1/8 scale (inference)"),
    ("--gacd", "store_true", False, "2-D image backbone: GA-CD fusion (lib/bcam.py) instead of PWAM"),
    ("--bcam", "store_true", False, "2-D image backbone: BCAM fusion (lib/bcam.py) instead of PWAM; 480 x 480 inputs only, inference"),
    ("--efn", "store_true", False, "2-D image backbone: EFN fusion (lib/bcam.py) instead of PWAM; square feature maps, inference"),
    ("--sep_t_pwam", "store_true", False, "SepTPWAM fusion: temporal Conv3d + spatial Conv3d branches, summed"),
    ("--conv3d_kernel_size_t", str, "3-1-1", "temporal-branch Conv3d kernel (B200 path: 3-3-3)"),
    ("--conv3d_kernel_size_s", str, "1-1-1", "spatial-branch Conv3d kernel (B200 path: 1-1-1)"),
    ("--w_t3x3_s1x1", "store_true", False, "SepTPWAM: W = IN(conv_t) + IN(conv_s)"),
    ("--mm_t3x3_s1x1", "store_true", False, "SepTPWAM: project_mm = GELU(conv_t) + GELU(conv_s)"),
    ("--loss", str, "ce", "training criterion: ce (weighted cross-entropy) | mc_dice | dice_focal | dice_boundary (lavt_rs_b200/losses.py)"),
    ("--loss_focal_rate", float, 3.0, "weight of the focal term of --loss dice_focal"),
    ("--loss_dice_rate", float, 1.0, "weight of the Dice term"),
    ("--loss_boundary_rate", float, 0.05, "weight of the boundary term of --loss dice_boundary"),
    ("--interpolate_before_seg", "store_true", False, "decoder: one more conv3x3 level at 1/2 scale before the classifier (inference)"),
    ("--seg_last", "store_true", False, "with --interpolate_before_seg: a further conv3x3 level at full scale; the video model then returns "
                                        "the classifier output without the final interpolation (inference)"),
]
_REJECTED_BOOL_FLAGS = ["ts_pwam", "t_pwam", "t_pwam_comp", "seq_t_pwam",
                        "sep_t_pwam_inner", "sep_seq_t_pwam", "sep_seq_t_pwam_inner"]


def get_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="LAVT-RS B200 hot path")
    for name, typ, default, helptext in _STRUCTURE_FLAGS:
        if typ == "store_true":
            p.add_argument(name, action="store_true", help=helptext)
        else:
            p.add_argument(name, type=typ, default=default, help=helptext)
    for name in _REJECTED_BOOL_FLAGS:
        p.add_argument("--" + name, action="store_true", help="declared for compatibility; rejected at model build")
    return p


def default_args(argv=()):
    return get_parser().parse_args(list(argv))
