"""Print the per-tile event timeline of CTA 0 of the tcgen05 attention kernel (LAVT_ATTN_TRACE dump)."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
names = ["QK0", "QK1", "QK2", "PV0", "PV1", "PV2", "S0", "S1", "S2", "P1_0", "P1_1", "P1_2", "P2_0", "P2_1", "P2_2", "EPIb", "EPIe"]
t0 = min(v for r in rows for v in r if v > 0)
ntile = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print("tile " + " ".join(f"{n:>6s}" for n in names))
for t in range(ntile):
    print(f"{t:4d} " + " ".join(f"{(r[t]-t0) if r[t] else -1:6d}" for r in rows))
