"""Print the per-tile event timeline of CTA 0 of the tcgen05 GEMM kernel (LAVT_GEMM_TRACE dump)."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
names = ["accfree", "opsrdy", "mmaiss", "accrdy", "epiend", "tma0"]
t0 = min(v for r in rows for v in r if v > 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
print("tile " + " ".join(f"{x:>8s}" for x in names))
for t in range(n):
    print(f"{t:4d} " + " ".join(f"{(r[t]-t0) if r[t] else -1:8d}" for r in rows))
