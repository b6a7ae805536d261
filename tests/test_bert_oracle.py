"""Pin oracle/bert_oracle.py against the installed ``transformers`` BertModel (the reference's bert/ package is a copy of
HF v3.0.2 modeling_bert.py and is not in its tree -- SURVEY.md section 8c): same state-dict keys, random init."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bert_oracle as BO  # noqa: E402


@pytest.mark.parametrize("layers,B,Nl", [(2, 3, 20), (12, 2, 22)])
def test_bert_oracle_matches_transformers(layers, B, Nl):
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.BertConfig(num_hidden_layers=layers)
    enc = transformers.BertModel(cfg).eval()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1000, 5000, (B, Nl), generator=g)
    mask = torch.zeros(B, Nl, dtype=torch.int64)
    for b in range(B):
        mask[b, : max(1, Nl - 3 * b - 2)] = 1
    with torch.no_grad():
        ref = enc(ids, attention_mask=mask)[0]
        sd = {"text_encoder." + k: v for k, v in enc.state_dict().items()}
        got = BO.bert_forward(sd, ids, mask)
    # padded QUERY positions differ by construction only if a release masks them differently; compare real tokens exactly
    # and all positions loosely
    live = mask.bool()
    assert (got[live] - ref[live]).abs().max().item() < 2e-4
    assert (got - ref).abs().max().item() < 2e-3
