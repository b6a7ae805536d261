"""SimpleDecoding mask decoder -- B200 host module (reference lib/mask_predictor.py:7-99).

Parameter names match the reference (``conv{1,2}_{4,3,2}``, ``bn{1,2}_{4,3,2}``, ``conv1_1``).  The three
upsample+concat+(conv3x3+BN+ReLU)x2 levels run as: one bandwidth kernel that writes the concatenated NHWC bf16
conv input, and the tcgen05 implicit-GEMM convolution with eval-mode BatchNorm folded into the epilogue.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _cabi as K
from .. import engine as E


class SimpleDecoding(nn.Module):
    def __init__(self, c4_dims, args=None, factor=2):
        super().__init__()
        # --interpolate_before_seg: one more conv3x3 level at twice the 1/4-scale resolution; --seg_last (only read inside that branch
        # by the reference, lib/mask_predictor.py:88-97): a further level at four times it, i.e. full resolution
        self.interpolate_before_seg = bool(getattr(args, "interpolate_before_seg", False))
        self.seg_last = bool(getattr(args, "seg_last", False))
        if self.interpolate_before_seg and getattr(args, "lazy_pred", False):
            raise ValueError("--interpolate_before_seg needs the 1/4-scale map (the reference dereferences x_c1, which --lazy_pred drops)")
        self.lazy_pred = bool(getattr(args, "lazy_pred", False))      # no 1/4-scale level (reference :32, :77): logits at 1/8 scale
        hidden = c4_dims // factor
        c4, c3, c2, c1 = c4_dims, c4_dims // factor, c4_dims // factor ** 2, c4_dims // factor ** 3
        levels = (("1_4", c4 + c3), ("2_4", hidden), ("1_3", hidden + c2), ("2_3", hidden), ("1_2", hidden + c1), ("2_2", hidden))
        for name, cin in (levels[:4] if self.lazy_pred else levels):
            setattr(self, "conv" + name, nn.Conv2d(cin, hidden, 3, padding=1, bias=False))
            setattr(self, "bn" + name, nn.BatchNorm2d(hidden))
        if self.interpolate_before_seg:           # reference :40-43 (conv2_1 is followed by bn1_1)
            self.conv2_1 = nn.Conv2d(hidden, hidden, 3, padding=1, bias=False)
            self.bn1_1 = nn.BatchNorm2d(hidden)
        if self.seg_last:                         # reference :45-48
            self.conv1_0 = nn.Conv2d(hidden, hidden, 3, padding=1, bias=False)
            self.bn1_0 = nn.BatchNorm2d(hidden)
        self.conv1_1 = nn.Conv2d(hidden, 2, 1)
        self.prepared = E.PreparedWeights()

    def _out_scale(self) -> int:
        """Resolution of the logits relative to the finest input map."""
        if not self.interpolate_before_seg:
            return 1
        return 4 if self.seg_last else 2

    def run_nhwc(self, c4, c3, c2, c1) -> torch.Tensor:
        """NHWC bf16 maps -> logits (n_img, 2, H1, W1) fp32 NCHW (H2, W2 and c1 = None under --lazy_pred)."""
        fine = c2 if self.lazy_pred else c1
        n_img, H, W, _ = fine.shape
        H, W = H * self._out_scale(), W * self._out_scale()
        logits = torch.empty(n_img, 2, H, W, device=fine.device, dtype=torch.float32)
        E.decoder_nhwc(self, c4, c3, c2, None if self.lazy_pred else c1, E.workspace(fine.device), logits)
        return logits

    def forward_feats(self, x_c4, x_c3, x_c2, x_c1):
        """Reference lib/mask_predictor.py:102-150: (logits, [x_c4, Y3, Y2, Y1]) with the top-down maps after each
        conv2_* + BN + ReLU as NCHW fp32 (API-compat copies of the NHWC bf16 buffers; not a hot path)."""
        maps = self._to_nhwc(x_c4, x_c3, x_c2, x_c1)
        n_img, H, W, _ = (maps[2] if self.lazy_pred else maps[3]).shape
        H, W = H * self._out_scale(), W * self._out_scale()
        logits = torch.empty(n_img, 2, H, W, device=x_c4.device, dtype=torch.float32)
        inter = []
        E.decoder_nhwc(self, *maps, E.workspace(x_c4.device), logits, feats=inter)
        return logits, [x_c4] + [t.permute(0, 3, 1, 2).float().contiguous() for t in inter]

    def forward(self, x_c4, x_c3, x_c2, x_c1) -> torch.Tensor:
        """Reference signature: four NCHW fp32 maps (coarse -> fine) -> (n_img, 2, H1, W1) logits."""
        return self.run_nhwc(*self._to_nhwc(x_c4, x_c3, x_c2, x_c1))

    def _to_nhwc(self, x_c4, x_c3, x_c2, x_c1):
        if self.lazy_pred:
            x_c1 = None                     # the reference ignores it too (lib/_utils.py:57-58, 103-104 pass None)
        elif x_c1 is None:
            raise K.LavtError("SimpleDecoding: x_c1 is required unless the model was built with --lazy_pred")
        for t in (x_c4, x_c3, x_c2, x_c1):
            if t is not None:
                E.require_cuda(t, "feature map")
        maps = []
        ws = E.workspace(x_c4.device)
        for i, t in enumerate((x_c4, x_c3, x_c2, x_c1)):
            if t is None:
                maps.append(None)
                continue
            n, C, H, W = t.shape
            t = t.detach().float().contiguous()
            o = ws.get("dec_in_%d" % i, (n, H, W, C), torch.bfloat16, t.device)
            K.nchw_to_nhwc_bf16(t.view(n, C, H * W), o.view(n, H * W, C))
            E._count(1)
            maps.append(o)
        return maps
