"""CPU oracle of the VLT fuse-and-classify head (reference lib/vlt.py) -- TEST INFRASTRUCTURE ONLY.

Plain-torch fp32 restatement, in eval mode, of ``VLTFuseAndClassify.forward`` (lib/vlt.py:129-199) and the modules it owns:
``QueryGenerationModule`` (:295-356), ``TransformerModel`` (:225-264: nn.TransformerEncoder / nn.TransformerDecoder, post-norm, ReLU),
``QueryBalancingModule`` (:379-405), ``ProgressiveDecoding`` (:428-485), ``PositionalEncoding`` (:204-222), ``vlt_concat_coords`` (:267-292).
Driven by a state dict with the reference's parameter names, so a reference checkpoint of the ``lavt_vlt`` / ``vlt`` models applies as is.
Pinned against the unmodified reference module in tests/test_oracle_vs_reference.py::test_vlt_head_matches_reference.  There is no CUDA
path for this head yet ("what comes next" item 1): nothing in the product imports this file.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _bn(x: Tensor, sd, name: str) -> Tensor:
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)


def _cbr(x: Tensor, sd, seq: str, i: int = 0) -> Tensor:
    """Conv2d (no bias, 'same' padding) + BatchNorm2d (eval) + ReLU = entries i, i+1, i+2 of the nn.Sequential ``seq``."""
    w = sd[f"{seq}.{i}.weight"]
    return F.relu(_bn(F.conv2d(x, w, padding=w.shape[-1] // 2), sd, f"{seq}.{i + 1}"))


def _pos(x: Tensor) -> Tensor:
    """PositionalEncoding (:204-222): x (len, B, dim) + interleaved sin / cos of the position."""
    n, _, dim = x.shape
    position = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros(n, 1, dim)
    pe[:, 0, 0::2] = torch.sin(position * div)
    pe[:, 0, 1::2] = torch.cos(position * div)
    return x + pe


def _mha(q: Tensor, k: Tensor, v: Tensor, sd, pre: str, heads: int, key_padding_mask: Tensor = None) -> Tensor:
    """nn.MultiheadAttention forward (sequence-first): q (L,B,E), k / v (S,B,E); key_padding_mask (B,S) True = ignore."""
    L, B, E = q.shape
    S = k.shape[0]
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    qp = q @ w[:E].t() + b[:E]
    kp = k @ w[E:2 * E].t() + b[E:2 * E]
    vp = v @ w[2 * E:].t() + b[2 * E:]
    hd = E // heads
    qh = qp.reshape(L, B, heads, hd).permute(1, 2, 0, 3)           # (B, h, L, hd)
    kh = kp.reshape(S, B, heads, hd).permute(1, 2, 0, 3)
    vh = vp.reshape(S, B, heads, hd).permute(1, 2, 0, 3)
    s = (qh @ kh.transpose(-1, -2)) * hd ** -0.5
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = (s.softmax(-1) @ vh).permute(2, 0, 1, 3).reshape(L, B, E)
    return o @ sd[pre + "out_proj.weight"].t() + sd[pre + "out_proj.bias"]


def _ln(x: Tensor, sd, name: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _ffn(x: Tensor, sd, pre: str) -> Tensor:
    h = F.relu(x @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"])
    return h @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"]


def transformer_fusion(src: Tensor, tgt: Tensor, sd, pre: str, heads: int, nlayers: int) -> Tensor:
    """TransformerModel.forward (:244-264): src (B,dim,h,w) -> encoder memory; tgt (Q,B,dim) queries -> decoder output (Q,B,dim)."""
    B, dim = src.shape[:2]
    mem = _pos(src.reshape(B, dim, -1).permute(2, 0, 1))
    for i in range(nlayers):                                        # nn.TransformerEncoderLayer, norm_first=False
        p = f"{pre}transformer_encoder.layers.{i}."
        mem = _ln(mem + _mha(mem, mem, mem, sd, p + "self_attn.", heads), sd, p + "norm1")
        mem = _ln(mem + _ffn(mem, sd, p), sd, p + "norm2")
    out = _pos(tgt)
    for i in range(nlayers):                                        # nn.TransformerDecoderLayer
        p = f"{pre}transformer_decoder.layers.{i}."
        out = _ln(out + _mha(out, out, out, sd, p + "self_attn.", heads), sd, p + "norm1")
        out = _ln(out + _mha(out, mem, mem, sd, p + "multihead_attn.", heads), sd, p + "norm2")
        out = _ln(out + _ffn(out, sd, p), sd, p + "norm3")
    return out


def concat_coords(x: Tensor) -> Tensor:
    """vlt_concat_coords (:267-292): append the x coordinate three times and the y coordinate three times, both in [-1, 1]."""
    B, _, h, w = x.shape
    ys = (2.0 * torch.arange(h, dtype=torch.float32) / (h - 1.0) - 1.0)[:, None].expand(h, w)
    xs = (2.0 * torch.arange(w, dtype=torch.float32) / (w - 1.0) - 1.0)[None, :].expand(h, w)
    xs, ys = xs[None, None].expand(B, 1, h, w), ys[None, None].expand(B, 1, h, w)
    return torch.cat([x, xs, xs, xs, ys, ys, ys], 1)


def query_generation(x: Tensor, l: Tensor, l_mask_row: Tensor, sd, pre: str, num_queries: int = 16) -> Tensor:
    """QueryGenerationModule.forward (:329-356): x (B,C,h,w); l (B,768,Nl); l_mask_row (B,1,Nl) -> (num_queries, B, dim)."""
    B = x.shape[0]
    y = concat_coords(x)
    for i in (0, 3, 6):
        y = _cbr(y, sd, pre + "project_1", i)
    y = F.conv2d(y, sd[pre + "project_2.weight"])                                  # (B, Q, h, w)
    y = y.reshape(B, num_queries, -1).permute(0, 2, 1)                              # (B, h*w, Q)
    vis = F.relu(F.conv1d(y, sd[pre + "project_query.0.weight"]))                   # (B, dim, Q)
    q = _pos(vis.permute(2, 0, 1))
    lang = _pos(F.relu(F.conv1d(l, sd[pre + "project_lang.0.weight"])).permute(2, 0, 1))
    pad = (1 - l_mask_row.squeeze(1)).bool()
    return _mha(q, lang, lang, sd, pre + "query_gen.", 8, key_padding_mask=pad) + vis.permute(2, 0, 1)


def query_balancing(nd: Tensor, d: Tensor, sd, pre: str) -> Tensor:
    """QueryBalancingModule.forward (:396-405): (Q,B,dim) x 2 -> gated decoded queries (B, dim, Q)."""
    x = F.relu(F.conv1d(nd.permute(1, 2, 0), sd[pre + "not_decoded_query_proj.0.weight"]))
    y = F.relu(F.conv1d(d.permute(1, 2, 0), sd[pre + "decoded_query_proj.0.weight"]))
    g = F.relu(F.conv1d(torch.cat([y, x], 1), sd[pre + "gate_proj.0.weight"]))
    return torch.sigmoid(F.conv1d(g, sd[pre + "gate_proj.2.weight"])) * y


def progressive_decoding(x: Tensor, sd, pre: str) -> Tensor:
    """ProgressiveDecoding.forward (:459-485): conv-BN-ReLU x 2, then three (x 2 bilinear align_corners upsample, conv-BN-ReLU), 1 x 1 classifier."""
    def cbr(t, c, b):
        return F.relu(_bn(F.conv2d(t, sd[pre + c + ".weight"], padding=1), sd, pre + b))

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    x = cbr(cbr(x, "conv1_4", "bn1_4"), "conv2_4", "bn2_4")
    for lvl in ("3", "2", "1"):
        x = cbr(up(x), "conv1_" + lvl, "bn1_" + lvl)
    return F.conv2d(x, sd[pre + "classifier.weight"], sd[pre + "classifier.bias"])


def vlt_fuse_and_classify(sd: Dict[str, Tensor], x_c4: Tensor, x_c3: Tensor, x_c2: Tensor, l: Tensor, l_mask: Tensor, pre: str = "",
                          nhead: int = 8, nlayers: int = 2, num_queries: int = 16) -> Tensor:
    """VLTFuseAndClassify.forward (:129-199).  x_c4 (B,1024,s/2,s/2), x_c3 (B,512,s,s), x_c2 (B,256,2s,2s) with s = img_size / 16;
    l (B,768,Nl); l_mask (B,Nl,1) -> logits (B, 2, 8s, 8s)."""
    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    B = x_c4.shape[0]
    m = l_mask.permute(0, 2, 1).to(l.dtype)                                          # (B, 1, Nl)
    sent = (l * m).sum(-1) / m.sum(-1)                                               # (B, 768)
    sent = sent @ sd[pre + "lang_proj.0.weight"].t() + sd[pre + "lang_proj.0.bias"]
    sent = F.relu(_bn(sent, sd, pre + "lang_proj.1"))[:, :, None, None]             # BatchNorm1d (eval)
    x4 = x_c4 + _cbr(_cbr(x_c4, sd, pre + "vis_reduce_chann_1", 0), sd, pre + "vis_reduce_chann_1", 3)
    mm4 = F.relu(_bn(x4 * sent, sd, pre + "joint_threshold.0"))
    mid_q = _cbr(torch.cat([up(mm4), _cbr(x_c3, sd, pre + "vis_reduce_chann_2")], 1), sd, pre + "fuse_1_2")
    c2 = _cbr(F.avg_pool2d(x_c2, 2), sd, pre + "vis_reduce_chann_3")
    fm_q = _cbr(torch.cat([mid_q, c2], 1), sd, pre + "fuse_2_3")
    t3 = _cbr(_cbr(fm_q, sd, pre + "hallucinate_result_of_23", 0), sd, pre + "hallucinate_result_of_23", 3)
    mid_tf = torch.cat([t3, mid_q], 1)
    f_tf = _cbr(torch.cat([up(mm4), _cbr(mid_tf, sd, pre + "project_again")], 1), sd, pre + "fuse_again")
    f_tf = _cbr(f_tf, sd, pre + "last_project")
    nd = query_generation(fm_q, l, m, sd, pre + "query_generation.", num_queries)
    dq = transformer_fusion(f_tf, nd, sd, pre + "transformer_fusion.", nhead, nlayers)
    bal = query_balancing(nd, dq, sd, pre + "query_balancing.")                      # (B, dim, Q)
    size = x_c3.shape[-1]
    out = F.relu(F.conv1d(bal, sd[pre + "q_to_spatial.0.weight"]))                   # (B, size*size, Q)
    out = out.permute(0, 2, 1).reshape(B, num_queries, size, size)
    out = _cbr(out, sd, pre + "spatial_refine")
    return progressive_decoding(out, sd, pre + "decoding.")
