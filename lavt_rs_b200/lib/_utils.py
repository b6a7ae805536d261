"""Top-level LAVT modules -- B200 host side (reference lib/_utils.py:10-238).

``LAVTVideo.forward(x[B,T,3,H,W], text[B,Nl], l_mask[B,Nl]) -> [B*T,2,H,W]`` and
``LAVTOne.forward(x[B,3,H,W], text, l_mask)`` keep the reference signatures.  The stock ``transformers`` BertModel is
kept as the parameter container of ``text_encoder`` (the reference's own copy is not in its
tree) but its forward runs on the sm_100a kernels too (``lavt_rs_b200/bert.py``); backbone and decoder hand NHWC bf16.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _cabi as K
from .. import bert as BERT
from .. import engine as E
from .video_swin_transformer import _lang, _mask, _planes


def _build_text_encoder(args):
    from transformers import BertConfig, BertModel
    ck = getattr(args, "ck_bert", None)
    if ck and os.path.exists(ck):
        enc = BertModel.from_pretrained(ck)
    else:  # offline: BERT-base architecture, random init (weights are loaded later from the LAVT checkpoint)
        enc = BertModel(BertConfig())
    enc.pooler = None
    return enc


class _Segmenter(nn.Module):
    video = False

    def _train_mode(self) -> bool:
        """model.train() with autograd on: the call is part of a training step (reference train.py:330-360)."""
        return self.training and torch.is_grad_enabled()

    def _encode_text_async(self, text: torch.Tensor, l_mask: torch.Tensor):
        """Text encoder on a high-priority side stream: its ~90 small launches overlap the patch embedding and the first
        Swin blocks (which do not read the language features).  Returns (l_feats, event that marks them valid)."""
        cur = torch.cuda.current_stream()
        side = getattr(self, "_text_stream", None)
        if side is None or side.device != cur.device:
            side = torch.cuda.Stream(device=cur.device, priority=-1)
            object.__setattr__(self, "_text_stream", side)
        B, Nl = text.shape
        buf = E.workspace(text.device).get("bert_out_cf", (B, self.text_encoder.config.hidden_size, Nl), torch.float32, text.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            l_feats = BERT.bert_forward(self.text_encoder, text, l_mask, out_cf=buf)
            ev = torch.cuda.Event()
            ev.record(side)
        return l_feats, ev

    def _segment(self, x5: torch.Tensor, l_feats: torch.Tensor, l_mask: torch.Tensor, size, lang_ready=None) -> torch.Tensor:
        """x5 (B,3,T,H,W) strided view; l_feats (B,768,Nl); l_mask (B,Nl[,1])."""
        _, nhwc = self.backbone.run(x5, _lang(l_feats), _mask(l_mask), want_nchw=False, want_nhwc_bf16=True,
                                    lang_ready=lang_ready)
        c1, c2, c3, c4 = nhwc if len(nhwc) == 4 else (None, *nhwc)              # --lazy_pred: three maps (lib/_utils.py:101-105)
        ws = E.workspace(c4.device)
        lg = E.decoder_nhwc(self.classifier, c4, c3, c2, c1, ws, None)       # (n_img, H/4, W/4, 2) NHWC fp32 (H/8 under --lazy_pred)
        if getattr(self, "seg_last", False):          # the video model skips the final interpolation (reference lib/_utils.py:105-106)
            size = lg.shape[1:3]
        out = torch.empty(lg.shape[0], 2, size[0], size[1], device=lg.device, dtype=torch.float32)
        K.upsample_logits(lg, out)                                             # bilinear x4 (x8) + NCHW (lib/_utils.py:106)
        E._count(1)
        return out


class LAVT(_Segmenter):
    """Backbone + decoder with precomputed language features (reference :10-30)."""

    def __init__(self, backbone, classifier):
        super().__init__()
        self.backbone, self.classifier = backbone, classifier

    def forward(self, x, l_feats, l_mask):
        E.require_cuda(x, "x")
        if self._train_mode():
            from ..training import train_forward
            return train_forward(self, x, l_feats, l_mask)
        x5 = _planes(x).unsqueeze(2)
        return self._segment(x5, l_feats, l_mask, x.shape[-2:])


class LAVTOne(_Segmenter):
    """BERT inside the model (reference :36-66)."""

    def __init__(self, backbone, classifier, args):
        super().__init__()
        self.backbone, self.classifier = backbone, classifier
        self.text_encoder = _build_text_encoder(args)
        self.lazy_pred = bool(getattr(args, "lazy_pred", False))

    def forward(self, x, text, l_mask):
        E.require_cuda(x, "x")
        if self._train_mode():
            from ..training import train_forward
            return train_forward(self, x, text, l_mask)
        l_feats, ev = self._encode_text_async(text, l_mask)               # (B, 768, Nl): [0].permute(0, 2, 1) of the reference
        x5 = _planes(x).unsqueeze(2)
        return self._segment(x5, l_feats, l_mask, x.shape[-2:], lang_ready=ev)


class _VLTBase(_Segmenter):
    """Encoder with three outputs (strides 8 / 16 / 32) + VLTFuseAndClassify + BERT (reference lib/_utils.py:279-343)."""
    fused_backbone = True

    def __init__(self, backbone, classifier, args=None, link=None):
        super().__init__()
        self.backbone, self.classifier, self.link = backbone, classifier, link
        self.model = getattr(args, "model", None)
        self.text_encoder = _build_text_encoder(args)

    def forward(self, x, text, l_mask):
        E.require_cuda(x, "x")
        if self._train_mode():
            raise NotImplementedError("the VLT head is inference-only on the B200 path")
        l_feats, ev = self._encode_text_async(text, l_mask)
        x5 = _planes(x).unsqueeze(2)
        if self.fused_backbone:
            _, nhwc = self.backbone.run(x5, _lang(l_feats), _mask(l_mask), want_nchw=False, want_nhwc_bf16=True, lang_ready=ev)
        else:
            _, nhwc = self.backbone.run(x5, None, None, want_nchw=False, want_nhwc_bf16=True)
            torch.cuda.current_stream().wait_event(ev)
        if len(nhwc) != 3:
            raise K.LavtError("the VLT head reads three stage outputs: build the backbone with out_indices=(1, 2, 3)")
        c2, c3, c4 = nhwc
        ws = E.workspace(c4.device)
        lg = E.vlt_head(self.classifier, c4, c3, c2, _lang(l_feats), _mask(l_mask), ws)
        out = torch.empty(lg.shape[0], 2, x.shape[-2], x.shape[-1], device=lg.device, dtype=torch.float32)
        K.upsample_logits(lg, out)                       # F.interpolate(..., size=input_shape, bilinear, align_corners=True) (:301, :336)
        E._count(1)
        return out


class VLT(_VLTBase):
    """Plain Swin encoder + VLT head (reference :279-307): the image branch never sees the language features."""
    fused_backbone = False


class LAVT_VLT(_VLTBase):
    """LAVT encoder (PWAM + gates) + VLT head (reference :314-342)."""


class LAVTVideo(_Segmenter):
    """Video model (reference :76-131): x (B,T,3,H,W) -> (B*T,2,H,W)."""
    video = True

    def __init__(self, backbone, classifier, args):
        super().__init__()
        self.backbone, self.classifier = backbone, classifier
        self.text_encoder = _build_text_encoder(args)
        self.lazy_pred = bool(getattr(args, "lazy_pred", False))
        self.seg_last = bool(getattr(args, "seg_last", False))

    def encode_text(self, text, l_mask):
        """BertModel(text, attention_mask=l_mask)[0].permute(0, 2, 1) on the sm_100a kernels (lavt_rs_b200/bert.py)."""
        return BERT.bert_forward(self.text_encoder, text, l_mask)

    def forward(self, x, text, l_mask):
        E.require_cuda(x, "x")
        if self._train_mode():          # training step: saved activations + hand-written backward behind autograd
            from ..training import train_forward
            return train_forward(self, x, text, l_mask)
        l_feats, ev = self._encode_text_async(text, l_mask)
        x5 = _planes(x).permute(0, 2, 1, 3, 4)
        return self._segment(x5, l_feats, l_mask, x.shape[-2:], lang_ready=ev)

    def forward_feats(self, x, text, l_mask):
        """Reference lib/_utils.py:110-131: (logits (B*T,2,H,W), decoder feature maps) -- used by test.py:155."""
        E.require_cuda(x, "x")
        l_feats, ev = self._encode_text_async(text, l_mask)
        x5 = _planes(x).permute(0, 2, 1, 3, 4)
        nchw, _ = self.backbone.run(x5, _lang(l_feats), _mask(l_mask), want_nchw=True, want_nhwc_bf16=False, lang_ready=ev)
        x_c1, x_c2, x_c3, x_c4 = nchw
        low, feats = self.classifier.forward_feats(x_c4, x_c3, x_c2, x_c1)
        out = torch.empty(low.shape[0], 2, x.shape[-2], x.shape[-1], device=low.device, dtype=torch.float32)
        K.upsample_logits(low.permute(0, 2, 3, 1).contiguous(), out)
        return out, feats

    def forward_with_lang(self, x, l_feats, l_mask):
        """Hot path after BERT: x (B,T,3,H,W), l_feats (B,768,Nl), l_mask (B,Nl)."""
        x5 = _planes(x).permute(0, 2, 1, 3, 4)     # (B,3,T,H,W) view; read in place by the im2col kernel
        return self._segment(x5, l_feats, l_mask, x.shape[-2:])

    def _load_inflated(self, pretrained, drop_fusion: bool):
        from ..weights import inflate_lavt2d_state_dict
        checkpoint = torch.load(pretrained, map_location="cpu")
        sd = inflate_lavt2d_state_dict(checkpoint["model"], self.backbone.window_size, self.state_dict(), drop_fusion=drop_fusion)
        msg = self.load_state_dict(sd, strict=False)    # not strict: the index buffers were dropped from the source dict
        print(msg)
        print(f"=> loaded successfully '{pretrained}'")
        return msg

    def load_from_pretrained2d_lavt_weights(self, pretrained):
        """Initialise the video model from a 2-D LAVT checkpoint (reference lib/_utils.py:133-181; train.py:575,
        test_ytvos.py:176): patch-embed weight unsqueezed in time, bias tables resized + tiled over the temporal offsets."""
        return self._load_inflated(pretrained, drop_fusion=False)

    def load_from_pretrained2d_lavt_weights_into_a_3d_model(self, pretrained):
        """Same, but the 2-D ``.fusion`` weights are dropped (reference lib/_utils.py:183-238)."""
        return self._load_inflated(pretrained, drop_fusion=True)
