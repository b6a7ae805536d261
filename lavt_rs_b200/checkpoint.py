"""Checkpoint save / resume in the reference's dict format (train.py:752-762 save, :608-612 and :720-727 resume):
``{'model': state_dict, 'optimizer': ..., 'epoch': int, 'args': namespace, 'lr_scheduler': ...}`` written with ``torch.save``;
only rank 0 writes (utils.save_on_master).  Files are interchangeable with the reference's: the model keys are identical and
``FusedAdamW`` keeps torch.optim.AdamW's state layout."""
from __future__ import annotations

import os

import torch


def is_main_process() -> bool:
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def save_checkpoint(path: str, model, optimizer, lr_scheduler, epoch: int, args=None, bert_model=None) -> None:
    if not is_main_process():
        return
    single = model.module if hasattr(model, "module") else model
    d = {"model": single.state_dict(), "optimizer": optimizer.state_dict(), "epoch": epoch, "args": args,
         "lr_scheduler": lr_scheduler.state_dict()}
    if bert_model is not None:      # non-integrated text encoder (the `lavt` model, train.py:752-755)
        d["bert_model"] = (bert_model.module if hasattr(bert_model, "module") else bert_model).state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(d, path)


def resume_checkpoint(path: str, model, optimizer=None, lr_scheduler=None, bert_model=None) -> int:
    """Returns the epoch to resume from (checkpoint epoch, the reference continues at epoch + 1, train.py:720-727)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    single = model.module if hasattr(model, "module") else model
    single.load_state_dict(ck["model"])
    if bert_model is not None and "bert_model" in ck:
        (bert_model.module if hasattr(bert_model, "module") else bert_model).load_state_dict(ck["bert_model"])
    if optimizer is not None and "optimizer" in ck:
        optimizer.load_state_dict(ck["optimizer"])
    if lr_scheduler is not None and "lr_scheduler" in ck:
        lr_scheduler.load_state_dict(ck["lr_scheduler"])
    return int(ck.get("epoch", -1))
