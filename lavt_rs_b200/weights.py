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


def inflate_lavt2d_state_dict(sd: Dict[str, torch.Tensor], window_size, current: Dict[str, torch.Tensor],
                              drop_fusion: bool = False) -> Dict[str, torch.Tensor]:
    """2-D LAVT checkpoint -> video model state dict (reference lib/_utils.py:133-181, :183-238 with ``drop_fusion``).

    * ``relative_position_index`` / ``attn_mask`` entries are dropped (buffers are rebuilt by the modules)
    * the patch-embedding Conv2d weight (C,3,4,4) gains a temporal axis of length 1 -> Conv3d weight (C,3,1,4,4)
    * every 2-D bias table ((2w-1)^2, nH) is bicubically resized to the video model's (2Wh-1) x (2Ww-1) when the spatial
      window differs, then tiled (2Wd-1) times along the leading (temporal-offset-major) axis; a head-count mismatch
      leaves the tiled pretrained table as is (the reference prints an error and passes)
    * ``drop_fusion``: the ``.fusion`` entries are removed (2-D PWAM vs 3-D fusion modules differ)
    """
    Wd, Wh, Ww = (int(w) for w in window_size)
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if "relative_position_index" in k or "attn_mask" in k:
            continue
        if drop_fusion and ".fusion" in k:
            continue
        out[k] = v
    if "backbone.patch_embed.proj.weight" not in out:
        raise KeyError("2-D LAVT checkpoint without backbone.patch_embed.proj.weight")
    out["backbone.patch_embed.proj.weight"] = out["backbone.patch_embed.proj.weight"].unsqueeze(2)
    for k in [k for k in out if "relative_position_bias_table" in k]:
        tab = out[k]
        L1, nH1 = tab.shape
        nH2 = current[k].shape[1]
        L2 = (2 * Wh - 1) * (2 * Ww - 1)
        if nH1 == nH2 and L1 != L2:
            S1 = int(L1 ** 0.5)
            grid = tab.permute(1, 0).reshape(1, nH1, S1, S1)
            grid = torch.nn.functional.interpolate(grid, size=(2 * Wh - 1, 2 * Ww - 1), mode="bicubic")
            tab = grid.reshape(nH2, L2).permute(1, 0)
        out[k] = tab.repeat(2 * Wd - 1, 1)
    return out


def inflate_swin2d_state_dict(sd: Dict[str, torch.Tensor], patch_t: int, window_size, current: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """ImageNet 2-D Swin checkpoint -> Video Swin backbone state dict (reference MultiModalSwinTransformer3D.inflate_weights,
    lib/video_swin_transformer.py:759-805): index / mask buffers dropped, the patch-embedding kernel repeated ``patch_t`` times along a new
    temporal axis and divided by ``patch_t``, every bias table bicubically resized to the video window when needed and tiled over the
    2 Wd - 1 temporal offsets (a head-count mismatch leaves the tiled pretrained table as is, like the reference)."""
    Wd, Wh, Ww = (int(w) for w in window_size)
    out = {k: v for k, v in sd.items() if "relative_position_index" not in k and "attn_mask" not in k}
    out["patch_embed.proj.weight"] = out["patch_embed.proj.weight"].unsqueeze(2).repeat(1, 1, patch_t, 1, 1) / patch_t
    for k in [k for k in out if "relative_position_bias_table" in k]:
        tab = out[k]
        L1, nH1 = tab.shape
        nH2 = current[k].shape[1]
        L2 = (2 * Wh - 1) * (2 * Ww - 1)
        if nH1 == nH2 and L1 != L2:
            S1 = int(L1 ** 0.5)
            grid = torch.nn.functional.interpolate(tab.permute(1, 0).reshape(1, nH1, S1, S1), size=(2 * Wh - 1, 2 * Ww - 1), mode="bicubic")
            tab = grid.reshape(nH2, L2).permute(1, 0)
        out[k] = tab.repeat(2 * Wd - 1, 1)
    return out


def checkpoint_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """What the OpenMMLab ``load_checkpoint`` the reference's 2-D backbone uses (lib/backbone.py:476-486) extracts from a file: the
    ``state_dict`` / ``model`` entry if present (else the object itself), with a leading ``module.`` (DataParallel) stripped."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = ck.get("state_dict", ck.get("model", ck)) if isinstance(ck, dict) else ck
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
