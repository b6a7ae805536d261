"""Image-model bench: BASELINE.json configs[0] -- ``lavt_one`` (Swin-B, window 12, BERT-base inside), batch 1, one 480 x 480 image and a
20-token sentence, random init -- through the reference's public call ``LAVTOne.forward(x, text, l_mask)`` on one B200, for the default PWAM
fusion and the lib/bcam.py ablations (``--gacd / --bcam / --efn``).

    python tools/bench_image.py [--steps 20] [--warmup 3] [--fusions pwam,gacd,bcam,efn] [--cpu-baseline]

One JSON line per fusion: ``value`` = images/s with the inputs resident in HBM, ``e2e`` = the same call fed from pinned host memory with the
logits read back, both timed with CUDA events; inputs (2.8 MB) are smaller than L2, so a 256 MB buffer is rewritten between timed steps.
``--cpu-baseline`` also times the CPU oracle port (oracle/lavt_oracle.py, PWAM, fp32, all host threads, BERT excluded) on the same image.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fusions", default="pwam,gacd,bcam,efn")
    ap.add_argument("--cpu-baseline", action="store_true")
    a = ap.parse_args()
    from lavt_rs_b200 import engine as E
    from lavt_rs_b200.args import default_args
    from lavt_rs_b200.lib import segmentation
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x_h = torch.randn(1, 3, 480, 480, generator=g).pin_memory()
    ids_h = torch.randint(1000, 5000, (1, 20), generator=g).pin_memory()
    mask_h = torch.zeros(1, 20, dtype=torch.int64)
    mask_h[:, :14] = 1
    mask_h = mask_h.pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for fusion in a.fusions.split(","):
        flags = ["--model", "lavt_one", "--swin_type", "base", "--window12"] + ([] if fusion == "pwam" else ["--" + fusion])
        torch.manual_seed(0)
        model = segmentation.lavt_one(pretrained="", args=default_args(flags)).to(dev).eval()
        x, ids, mask = x_h.to(dev), ids_h.to(dev), mask_h.to(dev)
        with torch.no_grad():
            for _ in range(a.warmup):
                out = model(x, ids, mask)
            torch.cuda.synchronize()
            E.LAUNCHES = 0
            model(x, ids, mask)
            launches = E.LAUNCHES
            res = {}
            for mode in ("resident", "e2e"):
                ms = 0.0
                for _ in range(a.steps):
                    flush.fill_(1)
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record()
                    if mode == "e2e":
                        out = model(x_h.to(dev, non_blocking=True), ids_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True))
                        host = out.to("cpu", non_blocking=True)
                    else:
                        out = model(x, ids, mask)
                    t1.record()
                    torch.cuda.synchronize()
                    ms += t0.elapsed_time(t1)
                res[mode] = ms / a.steps
        assert tuple(out.shape) == (1, 2, 480, 480) and bool(torch.isfinite(out).all())
        line = {"metric": "LAVT image model (lavt_one, Swin-B window12 + BERT-base) forward, 1 x 480x480 + 20 tokens", "value": 1000.0 / res["resident"],
                "unit": "images/s", "ms_per_step": res["resident"], "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
                "dtype": "bf16", "data": "synthetic", "fusion": fusion, "gpu_launches_per_step": launches,
                "config": {"workload": "BASELINE configs[0]: lavt_one base window12, batch 1, 480x480, Nl=20, fusion=" + fusion,
                           "l2": "256 MB buffer rewritten between timed steps"},
                "e2e": {"value": 1000.0 / res["e2e"], "unit": "images/s", "ms_per_step": res["e2e"],
                        "h2d_bytes_per_step": x_h.numel() * 4 + ids_h.numel() * 8 + mask_h.numel() * 8, "d2h_bytes_per_step": out.numel() * 4}}
        print(json.dumps(line), flush=True)
        del model
        torch.cuda.empty_cache()
    if a.cpu_baseline:
        from oracle import lavt_oracle as O          # checker-side code, timed here only as the reported CPU baseline
        cfg = O.OracleConfig(depths=(2, 2, 18, 2), window=(1, 12, 12), clamp_window=False, video=False)
        sd = O.random_state_dict(cfg, seed=0)
        l = torch.randn(1, 768, 20, generator=g)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            O.model_forward(sd, cfg, x_h, l, mask_h)
            t = time.perf_counter()
            O.model_forward(sd, cfg, x_h, l, mask_h)
            dt = time.perf_counter() - t
        print(json.dumps({"cpu_baseline": {"value": 1.0 / dt, "unit": "images/s", "s_per_image": dt, "cores": torch.get_num_threads(), "kind": "port",
                                           "sample": "1 warm-up + 1 timed forward of the oracle port, PWAM, fp32, BERT excluded"}}), flush=True)


if __name__ == "__main__":
    main()
