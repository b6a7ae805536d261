"""Per-shape CUDA-event breakdown of one eager forward (8 clips): python tools/profile_step.py [--window12] [--clips N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lavt_rs_b200 import _cabi as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--window12", action="store_true")
ap.add_argument("--clips", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = bench.build_model(a.window12, dev)
x, l, m = (t.to(dev) for t in bench.synth_batch(a.clips, 1))
with torch.no_grad():
    for _ in range(2):
        model(x, l, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(x, l, m)
    e1.record()
    torch.cuda.synchronize()
    print("eager forward ms", e0.elapsed_time(e1))
    K.TIMER.enabled = True
    model(x, l, m)
    rows = K.TIMER.by_tag()
tot = sum(v["ms"] for v in rows.values())
for (fam, tag), v in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{fam[:12]:12s} {tag:38s} x{v['launches']:3d} {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%  {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TF {v['bytes']/max(v['ms'],1e-9)/1e6:7.0f} GB/s")
print("timed kernels total ms", tot)
