"""Oracle (CPU) vs golden outputs produced by the UNMODIFIED reference (oracle/make_golden.py).  This is the pin that
travels: it runs wherever the repo is, with no access to /root/reference."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lavt_oracle as O  # noqa: E402
from oracle.make_golden import CASES, OUT, case_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_outputs(name):
    gold = np.load(os.path.join(OUT, name + ".npz"))
    cfg, sd, x, l, m = case_inputs(CASES[name])
    cap = {}
    with torch.no_grad():
        logits = O.model_forward(sd, cfg, x, l, m, capture=cap)
    assert np.abs(logits.numpy() - gold["logits"]).max() < 2e-4
    assert np.abs(cap["logits_lowres"].numpy() - gold["logits_lowres"]).max() < 2e-4
    for i in range(4):
        key = f"c{i + 1}"
        if key in gold:
            assert np.abs(cap[key].numpy() - gold[key]).max() < 2e-4, key
        assert abs(cap[key].abs().mean().item() - float(gold[key + "_absmean"])) < 1e-4
