"""Pretty-print the clock64 timeline written by a -DT3_TRACE build of attn_tc3.cu (LAVT_ATTN_TRACE=file): per item, cycles spent in each phase
for the warps of warpgroup 0.   python tools/tc3_trace.py gpurun_out/tc3_trace.txt [warpgroup]"""
import sys
import numpy as np
rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip()]
a = np.array(rows).reshape(16, 8, -1)
g = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = a[a > 0].min()
names = ["wait S", "softmax", "epi+st wait", "(issuer) wait P", "(issuer) PV issue", "(issuer) QK issue"]
for w in range(4 * g, 4 * g + 4):
    print(f"softmax warp {w}:  item | start | wait S | softmax (2 pieces) | epilogue + st-wait | chunk total")
    ev = a[w]
    for n in range(8, min(40, ev.shape[1])):
        if ev[0, n] == 0: break
        e = ev[:, n]
        nxt = ev[0, n + 1] if n + 1 < ev.shape[1] and ev[0, n + 1] else 0
        tail = f" | tile end -> next item {nxt - e[7]:6d}" if e[7] and nxt else ""
        print(f"   {n:3d} {e[0]-t0:8d} | {e[1]-e[0]:6d} | {e[2]-e[1]:6d} | {e[3]-e[2]:6d} | {(nxt - e[0]) if nxt else 0:6d}{tail}")
ev = a[12 + g]
print(f"issuer warp {12 + g}:  item | P wait begins | wait P | P.V issue | look-ahead Q.K^T issue | S(n+2) ready after P(n) ready (softmax warp {4*g} view)")
for n in range(8, min(40, ev.shape[1])):
    e = ev[:, n]
    if e[3] == 0: break
    print(f"   {n:3d} {e[3]-t0:8d} | {e[4]-e[3]:6d} | {e[5]-e[4]:6d} | {e[6]-e[5]:6d}")
