"""One LayerNorm launch at a bench shape, timed warm and with a flushed L2 (ncu target): python tools/run_ln_once.py [stage] [window 0|1]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi as K
from lavt_rs_b200.geometry import window_geometry
s = int(sys.argv[1]) if len(sys.argv) > 1 else 0
win = len(sys.argv) > 2 and sys.argv[2] == "1"
C, HW = 128 * 2 ** s, 96 // 2 ** s
M = 8 * 8 * HW * HW
x = torch.randn(M, C, device="cuda")
g, b = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
geom = window_geometry(8, 8, HW, HW, (8, 7, 7), True, True)
out = torch.empty(geom.rows() if win else M, C, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def run():
    if win: K.layernorm_window_gather(x, geom, g, b, out)
    else: K.layernorm_rows(x, g, b, out_bf16=out)
ts = []
for _ in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e-3)
t = sorted(ts)[len(ts) // 2]
byt = M * C * 4 + out.numel() * 2
print(f"stage {s} C {C} rows {M} window {win}: {t*1e6:.1f} us  {byt/t/1e12:.2f} TB/s  ({byt/1e6:.0f} MB)")
