"""VLT fuse-and-classify head on the B200 path (reference lib/vlt.py:12-485; models ``vlt`` / ``lavt_vlt``,
lib/segmentation.py:299-433).

The classes below are PARAMETER CONTAINERS with the reference's attribute names -- a ``lavt_vlt`` / ``vlt`` checkpoint loads with
``load_state_dict`` unchanged (``nn.MultiheadAttention`` / ``nn.TransformerEncoder`` / ``nn.TransformerDecoder`` are instantiated for their
parameter layout only).  ``VLTFuseAndClassify.forward`` runs ``engine.vlt_head``: every Conv / Linear is a tcgen05 GEMM or implicit-GEMM
conv with the eval BatchNorm folded into its epilogue, the four attention shapes (16 x Nl, s^2 x s^2, 16 x 16, 16 x s^2) run on
``lavt_mha_small``, the rest on the kernels of ``csrc/vlt_kernels.cu``.  Inference only (BatchNorm in eval mode); no fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _cabi as K
from .. import engine as E


def _cbr(cin: int, cout: int, k: int):
    """Conv2d (no bias) + BatchNorm2d + ReLU: three consecutive entries of an nn.Sequential."""
    return [nn.Conv2d(cin, cout, k, padding=k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU()]


class PositionalEncoding(nn.Module):
    """Interleaved sin / cos table (:204-222), registered as the buffer ``pe`` [max_len, 1, dim] like the reference so that the
    state-dict keys match; the kernels read its first rows as a [positions, dim] table."""

    def __init__(self, dim, max_len=5000):
        super().__init__()
        import math
        pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
        pe = torch.zeros(max_len, 1, dim)
        pe[:, 0, 0::2] = torch.sin(pos * div)
        pe[:, 0, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)

    def table(self, n: int) -> torch.Tensor:
        return self.pe[:n, 0].contiguous()


class TransformerModel(nn.Module):
    """Parameter layout of the reference's encoder-decoder fusion (:225-264): post-norm layers, ReLU, d_hid feed-forward."""

    def __init__(self, d_model, nhead, d_hid, nlayers, dropout=0.0, h=26, w=26):
        super().__init__()
        self.d_model, self.nhead, self.nlayers, self.h, self.w = d_model, nhead, nlayers, h, w
        self.pos_encoder = PositionalEncoding(d_model)
        self.transformer_encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward=d_hid, dropout=dropout), nlayers,
                                                         enable_nested_tensor=False)
        self.transformer_decoder = nn.TransformerDecoder(nn.TransformerDecoderLayer(d_model, nhead, dim_feedforward=d_hid, dropout=dropout), nlayers)


class QueryGenerationModule(nn.Module):
    """(:295-356) coordinates + three 3x3 convs -> 16 query maps -> Conv1d over the h*w axis -> cross-attention to the words."""

    def __init__(self, visual_dim, dim, h=26, w=26, lang_dim=768, num_queries=16):
        super().__init__()
        self.visual_dim, self.dim, self.h, self.w, self.lang_dim, self.num_queries = visual_dim, dim, h, w, lang_dim, num_queries
        self.project_1 = nn.Sequential(*_cbr(visual_dim + 6, visual_dim, 3), *_cbr(visual_dim, visual_dim, 3), *_cbr(visual_dim, visual_dim, 3))
        self.project_2 = nn.Conv2d(visual_dim, num_queries, 1, bias=False)
        self.project_query = nn.Sequential(nn.Conv1d(h * w, dim, 1, bias=False), nn.ReLU())
        self.project_lang = nn.Sequential(nn.Conv1d(lang_dim, dim, 1, bias=False), nn.ReLU())
        self.pos_encoder = PositionalEncoding(dim)
        self.query_gen = nn.MultiheadAttention(dim, 8)


class QueryBalancingModule(nn.Module):
    """(:379-405) confidence gate on the decoded queries."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.not_decoded_query_proj = nn.Sequential(nn.Conv1d(dim, dim, 1, bias=False), nn.ReLU())
        self.decoded_query_proj = nn.Sequential(nn.Conv1d(dim, dim, 1, bias=False), nn.ReLU())
        self.gate_proj = nn.Sequential(nn.Conv1d(2 * dim, dim, 1, bias=False), nn.ReLU(), nn.Conv1d(dim, 1, 1, bias=False), nn.Sigmoid())


class ProgressiveDecoding(nn.Module):
    """(:428-485) conv-BN-ReLU x 2, three (x2 bilinear upsample, conv-BN-ReLU), 1x1 classifier."""

    def __init__(self, c4_dim, hidden_size):
        super().__init__()
        for name, cin in (("1_4", c4_dim), ("2_4", hidden_size), ("1_3", hidden_size), ("1_2", hidden_size), ("1_1", hidden_size)):
            setattr(self, "conv" + name, nn.Conv2d(cin, hidden_size, 3, padding=1, bias=False))
            setattr(self, "bn" + name, nn.BatchNorm2d(hidden_size))
        self.classifier = nn.Conv2d(hidden_size, 2, 1)


class VLTFuseAndClassify(nn.Module):
    """The ``classifier`` of the vlt / lavt_vlt models: ``forward(x_c4, x_c3, x_c2, l, l_mask) -> (B, 2, img/2, img/2)`` (:129-199)."""

    def __init__(self, d_model=256, nhead=8, d_hid=256, nlayers=2, args=None):
        super().__init__()
        top, mid, bot = 1024, 512, 256                    # Swin-B stage widths, hard-coded by the reference (:16-18)
        self.d_model, self.nhead, self.d_hid, self.nlayers = d_model, nhead, d_hid, nlayers
        self.num_queries = 16
        self.size = args.img_size // 16
        self.joint_dim = top
        j = self.joint_dim
        self.vis_reduce_chann_1 = nn.Sequential(*_cbr(top, top // 2, 1), *_cbr(top // 2, top, 3))
        self.vis_reduce_chann_2 = nn.Sequential(*_cbr(mid, mid, 1))
        self.fuse_1_2 = nn.Sequential(*_cbr(j + mid, j // 2, 1))
        self.vis_reduce_chann_3 = nn.Sequential(*_cbr(bot, bot, 1))
        self.fuse_2_3 = nn.Sequential(*_cbr(j // 2 + bot, j // 2, 1))
        self.hallucinate_result_of_23 = nn.Sequential(*_cbr(j // 2, j // 4, 1), *_cbr(j // 4, j // 2, 3))
        self.project_again = nn.Sequential(*_cbr(j, j // 2, 1))
        self.fuse_again = nn.Sequential(*_cbr(j + j // 2, d_model, 1))
        self.last_project = nn.Sequential(*_cbr(d_model, d_model, 1))
        self.lang_proj = nn.Sequential(nn.Linear(768, j), nn.BatchNorm1d(j), nn.ReLU())
        self.joint_threshold = nn.Sequential(nn.BatchNorm2d(j), nn.ReLU())
        self.query_generation = QueryGenerationModule(j // 2, d_model, h=self.size, w=self.size, num_queries=self.num_queries)
        self.transformer_fusion = TransformerModel(d_model, nhead, d_hid, nlayers, getattr(args, "fusion_drop", 0.0), self.size, self.size)
        self.query_balancing = QueryBalancingModule(d_model)
        self.q_to_spatial = nn.Sequential(nn.Conv1d(d_model, self.size * self.size, 1, bias=False), nn.ReLU())
        self.spatial_refine = nn.Sequential(*_cbr(self.num_queries, d_model, 3))
        self.decoding = ProgressiveDecoding(d_model, d_model)
        self.prepared = E.PreparedWeights()

    def forward(self, x_c4, x_c3, x_c2, l, l_mask):
        """NCHW fp32 feature maps (the reference's layout) -> logits (B, 2, 8 s, 8 s) with s = img_size / 16."""
        from .video_swin_transformer import _lang, _mask
        E.require_cuda(x_c4, "x_c4")
        ws = E.workspace(x_c4.device)
        maps = []
        for i, t in enumerate((x_c4, x_c3, x_c2)):
            n, C, H, W = t.shape
            ob = ws.get("vlt_in_%d" % i, (n, H, W, C), torch.bfloat16, t.device)
            K.nchw_to_nhwc_bf16(t.float().contiguous().view(n, C, H * W), ob.view(n, H * W, C))
            maps.append(ob)
        lg = E.vlt_head(self, maps[0], maps[1], maps[2], _lang(l), _mask(l_mask), ws)
        n, H, W, _ = lg.shape
        out = torch.empty(n, 2, H, W, device=lg.device, dtype=torch.float32)
        K.upsample_logits(lg, out)                  # same size: NHWC -> NCHW
        return out
