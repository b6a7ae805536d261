"""One window-attention configuration a few times (ncu target): python tools/run_attn_once.py [stage] [w7|w12] [shift]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi as K
from lavt_rs_b200.geometry import window_geometry
s = int(sys.argv[1]) if len(sys.argv) > 1 else 2
window = (8, 12, 12) if len(sys.argv) > 2 and sys.argv[2] == "w12" else (8, 7, 7)
shift = not (len(sys.argv) > 3 and sys.argv[3] == "0")
C, nH, HW = 128 * 2 ** s, 4 * 2 ** s, 96 // 2 ** s
geom = window_geometry(8, 8, HW, HW, window, shift, True)
rows = geom.rows()
qkv = (torch.randn(rows, 3 * C, device="cuda") * 0.3).bfloat16()
L = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
tt = (torch.randn(nH, L, device="cuda") * 0.5).contiguous()
o = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    K.window_attention(qkv, tt, geom, o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    K.window_attention(qkv, tt, geom, o)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
print(f"stage {s} window {window} shift {shift}: {t*1e6:.1f} us  {4.0*rows*geom.N*C/t/1e12:.1f} TF")
