"""State-dict plumbing: load reference-format checkpoints into the B200 host modules."""
from __future__ import annotations

from typing import Dict, Iterable

import torch

# buffers the reference registers but that carry no information for the kernels
_BENIGN = ("relative_position_index", "num_batches_tracked", "attn_mask")


def load_reference_state_dict(module: torch.nn.Module, sd: Dict[str, torch.Tensor], prefix: str = "",
                              ignore_prefixes: Iterable[str] = ()) -> None:
    """Load ``sd`` (reference key names) into ``module``; every non-benign key must match exactly."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    result = module.load_state_dict(sd, strict=False)
    ign = tuple(ignore_prefixes)
    missing = [k for k in result.missing_keys if not k.endswith(_BENIGN) and not k.startswith(ign)]
    unexpected = [k for k in result.unexpected_keys if not k.endswith(_BENIGN) and not k.startswith(ign)]
    if missing or unexpected:
        raise KeyError(f"state dict mismatch: missing={missing[:8]} unexpected={unexpected[:8]}")
