"""Text encoder forward on the sm_100a kernels (reference lib/_utils.py:52-54, 98-100: ``BertModel(text, attention_mask)[0]``;
the reference's bert/ package = HF Transformers v3.0.2 modeling_bert.py).

The stock ``transformers.BertModel`` stays the PARAMETER CONTAINER (its ``text_encoder.*`` state-dict keys are the reference's,
197 keys), but its forward is not used on the GPU path: dense layers run on the tcgen05 GEMM kernel (fused bias / GELU /
residual epilogues), LayerNorms on ``lavt_layernorm_rows`` (eps 1e-12), and the embedding gather, the per-head attention
over the <= 128 tokens of a sentence and the final (B,Nl,C) -> (B,C,Nl) layout change on the three kernels of
``csrc/bert_kernels.cu``.  6 + 7 * layers launches per call, no cuBLAS, no CPU fallback.
"""
from __future__ import annotations

import math

import torch

from . import _cabi as K
from . import engine as E

_LOG2E = 1.4426950408889634
# "bf16" (default): plain bf16 operands, 4.5e-3 rel-L2 on the 12-layer output.
# "split3": every dense layer runs as ONE tcgen05 GEMM over split-precision operands ([hi | lo | hi] x [W_hi | W_hi | W_lo], fp32 accumulate)
# with fp32 activations in between: 2.5e-5 rel-L2 (tests/test_model_gpu.py::test_bert_text_encoder).  Measured on the whole model
# (bench.py parity leg, 8x384^2): logits rel-L2 1.55e-2 with split3 vs 1.59e-2 with bf16 -- the text encoder is NOT what the model-level
# error is made of (the 24 Swin blocks' bf16 operands are: stage outputs c3 / c4 sit at 1.4e-2 / 1.8e-2 with exact language features),
# and its 145 launches on the side stream cost 4 % of throughput; so it is an option (validation of the text side), not the default.
PRECISION = "bf16"
import os as _os
SMALL_M = int(_os.environ.get("LAVT_BERT_SMALLM", "512"))          # rows up to which the dense layers run as split-K launches (gemm_bf16_smallm)


def set_precision(mode: str) -> str:
    global PRECISION
    if mode not in ("split3", "bf16"):
        raise ValueError("text-encoder precision must be 'split3' or 'bf16'")
    prev, PRECISION = PRECISION, mode
    return prev


def _w3(w: torch.Tensor) -> torch.Tensor:
    """nn.Linear weight [N, K] fp32 -> bf16 [N, 3K] = W_hi | W_hi | W_lo."""
    w = w.detach().float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], 1).contiguous()


def _bert_forward_split3(enc, ids, maskf, out_cf, B, Nl):
    cfg = enc.config
    H, heads = cfg.hidden_size, cfg.num_attention_heads
    M, dev = B * Nl, ids.device
    ws, pw, emb, eps = E.workspace(dev), _prepared(enc), enc.embeddings, float(cfg.layer_norm_eps)
    I = cfg.intermediate_size
    x = ws.get("bert_x", (M, H), torch.float32, dev)
    x3 = ws.get("bert_x3", (M, 3 * H), torch.bfloat16, dev)
    type0 = pw.get("type0", [emb.token_type_embeddings.weight], lambda: emb.token_type_embeddings.weight.detach()[0].float().contiguous())
    K.bert_embed(ids, emb.word_embeddings.weight.detach(), emb.position_embeddings.weight.detach(), type0, x)
    K.layernorm_rows(x, emb.LayerNorm.weight.detach(), emb.LayerNorm.bias.detach(), out_f32=x, eps=eps)
    K.split3_bf16(x, x3)
    qkv = ws.get("bert_qkv32", (M, 3 * H), torch.float32, dev)
    ctx = ws.get("bert_ctx32", (M, H), torch.float32, dev)
    ctx3 = ws.get("bert_ctx3", (M, 3 * H), torch.bfloat16, dev)
    hid = ws.get("bert_hid32", (M, I), torch.float32, dev)
    hid3 = ws.get("bert_hid3", (M, 3 * I), torch.bfloat16, dev)
    qs = pw.get("qscale", [], lambda: torch.cat([torch.full((H,), 64 ** -0.5 * _LOG2E), torch.ones(2 * H)]).to(dev))
    for i, layer in enumerate(enc.encoder.layer):
        sa, so = layer.attention.self, layer.attention.output
        w_qkv = pw.get(f"qkv_w3_{i}", [sa.query.weight, sa.key.weight, sa.value.weight],
                       lambda: _w3(torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0)))
        b_qkv = pw.get(f"qkv_b{i}", [sa.query.bias, sa.key.bias, sa.value.bias],
                       lambda: (torch.cat([sa.query.bias, sa.key.bias, sa.value.bias]).detach().float() * qs).contiguous())
        K.gemm_bf16(x3, w_qkv, cscale=qs, bias=b_qkv, out_f32=qkv)
        K.bert_attention_f32(qkv, maskf, ctx, heads)
        K.split3_bf16(ctx, ctx3)
        K.gemm_bf16(ctx3, pw.get(f"o_w3_{i}", [so.dense.weight], lambda: _w3(so.dense.weight)), bias=so.dense.bias.detach(), resid=x, out_f32=x)
        K.layernorm_rows(x, so.LayerNorm.weight.detach(), so.LayerNorm.bias.detach(), out_f32=x, eps=eps)
        K.split3_bf16(x, x3)
        K.gemm_bf16(x3, pw.get(f"fc1_w3_{i}", [layer.intermediate.dense.weight], lambda: _w3(layer.intermediate.dense.weight)),
                    bias=layer.intermediate.dense.bias.detach(), act=K.ACT_GELU, out_f32=hid)
        K.split3_bf16(hid, hid3)
        K.gemm_bf16(hid3, pw.get(f"fc2_w3_{i}", [layer.output.dense.weight], lambda: _w3(layer.output.dense.weight)),
                    bias=layer.output.dense.bias.detach(), resid=x, out_f32=x)
        K.layernorm_rows(x, layer.output.LayerNorm.weight.detach(), layer.output.LayerNorm.bias.detach(), out_f32=x, eps=eps)
        K.split3_bf16(x, x3)
    K.rows_to_channels_first(x.view(B, Nl, H), out_cf)
    E._count(4 + 11 * cfg.num_hidden_layers)
    return out_cf


def _prepared(enc) -> E.PreparedWeights:
    pw = getattr(enc, "_lavt_prepared", None)
    if pw is None:
        pw = E.PreparedWeights()
        object.__setattr__(enc, "_lavt_prepared", pw)     # plain attribute: keeps state_dict() untouched
    return pw


def bert_forward(enc, ids: torch.Tensor, mask: torch.Tensor, out_cf: torch.Tensor = None) -> torch.Tensor:
    """enc: transformers.BertModel (eval); ids (B,Nl) int64; mask (B,Nl) -> l_feats fp32 (B, H, Nl) (= [0].permute(0,2,1))."""
    E.require_cuda(ids, "text")
    if enc.training:
        raise NotImplementedError("text-encoder dropout (training) is not implemented on the B200 path")
    cfg = enc.config
    H, heads, layers = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
    if H // heads != 64 or H % 128 != 0:
        raise K.LavtError("BERT on the B200 path needs head_dim 64 and hidden size % 128 == 0 (BERT-base: 768 / 12)")
    if getattr(cfg, "hidden_act", "gelu") != "gelu":
        raise K.LavtError("BERT on the B200 path implements the erf GELU only")
    B, Nl = ids.shape
    M = B * Nl
    dev = ids.device
    ws = E.workspace(dev)
    pw = _prepared(enc)
    emb = enc.embeddings
    eps = float(cfg.layer_norm_eps)
    ids = ids.detach().to(torch.int64).contiguous()
    maskf = mask.detach().reshape(B, Nl).to(torch.float32).contiguous()

    if PRECISION == "split3":
        if out_cf is None:
            out_cf = torch.empty(B, H, Nl, device=dev, dtype=torch.float32)
        return _bert_forward_split3(enc, ids, maskf, out_cf, B, Nl)

    x = ws.get("bert_x", (M, H), torch.float32, dev)          # residual stream
    xb = ws.get("bert_xb", (M, H), torch.bfloat16, dev)
    type0 = pw.get("type0", [emb.token_type_embeddings.weight], lambda: emb.token_type_embeddings.weight.detach()[0].float().contiguous())
    K.bert_embed(ids, emb.word_embeddings.weight.detach(), emb.position_embeddings.weight.detach(), type0, x)
    K.layernorm_rows(x, emb.LayerNorm.weight.detach(), emb.LayerNorm.bias.detach(), out_bf16=xb, out_f32=x, eps=eps)
    qkv = ws.get("bert_qkv", (M, 3 * H), torch.bfloat16, dev)
    ctx = ws.get("bert_ctx", (M, H), torch.bfloat16, dev)
    hid = ws.get("bert_hid", (M, cfg.intermediate_size), torch.bfloat16, dev)
    qs = pw.get("qscale", [], lambda: torch.cat([torch.full((H,), 64 ** -0.5 * _LOG2E), torch.ones(2 * H)]).to(dev))
    # A sentence batch is a few rows (M = clips x words = 160 at 8 clips): as plain launches these GEMMs are 2 x 3..12 output tiles that each
    # walk up to 48 k-blocks alone (26-43 us per launch, on SMs the backbone's persistent kernels are waiting for); split over K they fill the
    # GPU and a reduce kernel applies the epilogue.
    small = M <= SMALL_M
    I = cfg.intermediate_size
    skw = ws.get("bert_splitk", (max(K.splitk_workspace_floats(M, n_, k_) for n_, k_ in ((3 * H, H), (H, H), (I, H), (H, I))),),
                 torch.float32, dev) if small else None

    def gemm(a_, w_, **epi):
        if small:
            K.gemm_bf16_smallm(a_, w_, skw, **epi)
        else:
            K.gemm_bf16(a_, w_, **epi)
    for i, layer in enumerate(enc.encoder.layer):
        sa, so = layer.attention.self, layer.attention.output
        w_qkv = pw.get(f"qkv_w{i}", [sa.query.weight, sa.key.weight, sa.value.weight],
                       lambda: torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0).detach().to(torch.bfloat16).contiguous())
        b_qkv = pw.get(f"qkv_b{i}", [sa.query.bias, sa.key.bias, sa.value.bias],
                       lambda: (torch.cat([sa.query.bias, sa.key.bias, sa.value.bias]).detach().float() * qs).contiguous())
        gemm(xb, w_qkv, cscale=qs, bias=b_qkv, out_bf16=qkv)         # q pre-scaled by 64^-0.5 * log2(e)
        K.bert_attention(qkv, maskf, ctx, heads)
        w_o = pw.get(f"o_w{i}", [so.dense.weight], lambda: so.dense.weight.detach().to(torch.bfloat16).contiguous())
        gemm(ctx, w_o, bias=so.dense.bias.detach(), resid=x, out_f32=x)
        K.layernorm_rows(x, so.LayerNorm.weight.detach(), so.LayerNorm.bias.detach(), out_bf16=xb, out_f32=x, eps=eps)
        w_1 = pw.get(f"fc1_w{i}", [layer.intermediate.dense.weight],
                     lambda: layer.intermediate.dense.weight.detach().to(torch.bfloat16).contiguous())
        gemm(xb, w_1, bias=layer.intermediate.dense.bias.detach(), act=K.ACT_GELU, out_bf16=hid)
        w_2 = pw.get(f"fc2_w{i}", [layer.output.dense.weight], lambda: layer.output.dense.weight.detach().to(torch.bfloat16).contiguous())
        gemm(hid, w_2, bias=layer.output.dense.bias.detach(), resid=x, out_f32=x)
        K.layernorm_rows(x, layer.output.LayerNorm.weight.detach(), layer.output.LayerNorm.bias.detach(), out_bf16=xb, out_f32=x, eps=eps)
    if out_cf is None:
        out_cf = torch.empty(B, H, Nl, device=dev, dtype=torch.float32)
    K.rows_to_channels_first(x.view(B, Nl, H), out_cf)
    E._count(3 + (11 if small else 7) * layers)
    return out_cf
