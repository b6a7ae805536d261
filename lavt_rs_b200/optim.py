"""Optimizer side of the reference's training loop on the B200 path (train.py:614-699).

* ``reference_param_groups`` restates the reference's grouping: backbone parameters whose name contains ``norm`` /
  ``absolute_pos_embed`` / ``relative_position_bias_table`` without weight decay, the rest of the backbone, the classifier, and the
  text-encoder layers selected by ``--lang_enc_params`` (train.py:615-684).
* ``FusedAdamW`` is a ``torch.optim.Optimizer`` (so ``LambdaLR`` and ``state_dict`` work unchanged, state keys = torch.optim.AdamW's)
  whose ``step()`` updates a whole parameter group with ONE multi-tensor kernel launch (``lavt_adamw_step``) instead of PyTorch's
  per-tensor foreach kernels.
* ``poly_lr_lambda`` is the reference's schedule ``(1 - it / total) ** 0.9`` (train.py:698-699).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _cabi as K


def reference_param_groups(model, lang_enc_params: str = "encoder-10", bert_model=None) -> List[dict]:
    """``bert_model``: the separate text encoder of the non-integrated ``lavt`` model (train.py:624-633): when the model has no
    ``text_encoder`` of its own, the first 10 encoder layers of ``bert_model`` form the fourth group, as in the reference."""
    no_decay, decay = [], []
    for name, prm in model.backbone.named_parameters():
        (no_decay if ("norm" in name or "absolute_pos_embed" in name or "relative_position_bias_table" in name) else decay).append(prm)
    groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay},
              {"params": [p for p in model.classifier.parameters() if p.requires_grad]}]
    enc = getattr(model, "text_encoder", None)
    if enc is not None:
        def layers(n):
            return [p for i in range(n) for p in enc.encoder.layer[i].parameters() if p.requires_grad]
        if lang_enc_params == "encoder-10":
            groups.append({"params": layers(10)})
        elif lang_enc_params == "encoder-all":
            groups.append({"params": [p for p in enc.encoder.parameters() if p.requires_grad]})
        elif lang_enc_params == "embeddings+encoder-10":
            groups.append({"params": [p for p in enc.embeddings.parameters() if p.requires_grad]})
            groups.append({"params": layers(10)})
        elif lang_enc_params == "embeddings+encoder-all":
            groups.append({"params": [p for p in enc.embeddings.parameters() if p.requires_grad]})
            groups.append({"params": [p for p in enc.encoder.parameters() if p.requires_grad]})
        else:
            raise ValueError(f"unknown --lang_enc_params {lang_enc_params}")
    elif bert_model is not None:
        groups.append({"params": [p for i in range(10) for p in bert_model.encoder.layer[i].parameters() if p.requires_grad]})
    return groups


def poly_lr_lambda(total_iters: int):
    return lambda it: (1 - it / total_iters) ** 0.9


def _bump_versions(tensors) -> None:
    """Advance the autograd version counter of tensors that were written outside of PyTorch (no kernel launch)."""
    try:
        torch._C._autograd._unsafe_set_version_counter(tuple(tensors), tuple(t._version + 1 for t in tensors))
    except (AttributeError, TypeError):          # older / newer torch without the batched setter: a real (cheap) in-place op
        torch._foreach_add_(list(tensors), 0.0)


class _Tensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("vmax", C.c_void_p), ("n", C.c_int64),
                ("bc1", C.c_float), ("bc2", C.c_float)]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._chunk = int(K.lib().lavt_adamw_chunk_elems())
        self._host = {}       # group index -> (pinned table bytes, pinned prefix, device buffers)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            entries, prefix, blocks = [], [0], 0
            b1, b2 = group["betas"]
            for p in live:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise K.LavtError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                # torch.optim.AdamW keeps one step counter per parameter: a Python int in checkpoints written by torch 1.x (the
                # reference's era), a tensor today -- and load_state_dict may have moved it to the GPU.  Normalise to a CPU
                # scalar tensor so that .item() never synchronises the device.
                s_old = st["step"]
                step_no = int(s_old.item() if torch.is_tensor(s_old) else s_old) + 1
                if torch.is_tensor(s_old) and not s_old.is_cuda:
                    s_old.fill_(float(step_no))
                else:
                    st["step"] = torch.tensor(float(step_no), dtype=torch.float32)
                e = _Tensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                            st["max_exp_avg_sq"].data_ptr() if group["amsgrad"] else None, p.numel(),
                            1.0 - b1 ** step_no, 1.0 - b2 ** step_no)
                entries.append((e, g))            # keep g alive until the launch is enqueued
                blocks += (p.numel() + self._chunk - 1) // self._chunk
                prefix.append(blocks)
            n = len(entries)
            raw = (_Tensor * n)(*[e for e, _ in entries])
            tab = torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8).pin_memory().to(live[0].device, non_blocking=True)
            pre = torch.tensor(prefix, dtype=torch.int32).pin_memory().to(live[0].device, non_blocking=True)
            K.check(K.lib().lavt_adamw_step(tab.data_ptr(), pre.data_ptr(), n, blocks, float(group["lr"]), float(b1), float(b2),
                                            float(group["eps"]), float(group["weight_decay"]), K.stream_ptr()), "lavt_adamw_step")
            self._host[gi] = (tab, pre, [g for _, g in entries])      # lifetime: until the next step of this group
            # the kernel wrote the parameters through raw pointers: bump their autograd version counters so that every cache keyed on
            # (data_ptr, _version) -- engine.PreparedWeights' bf16 / folded copies -- is rebuilt before the next forward
            _bump_versions(live)
        return loss
