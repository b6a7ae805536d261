"""Summarise the source page of an .ncu-rep: stall reasons and the hottest SASS instructions.  python tools/ncu_src.py file.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
idx = {k: i for i, k in enumerate(hdr)}
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("samples", tot, "instructions", sum(f(r, "Instructions Executed") for r in data))
for s, v in sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda x: -x[1])[:10]:
    print(f"  {s:26s} {v:9.0f} {v / tot:.3f}")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print(f"{f(r, '# Samples'):7.0f} {f(r, 'Instructions Executed'):9.0f}  {r[idx['Source']][:64]:64s} {st[0][1][6:]}:{st[0][0]:.0f} {st[1][1][6:]}:{st[1][0]:.0f}")
if len(sys.argv) > 3:
    import collections
    ops = collections.Counter()
    for r in data:
        src = r[idx['Source']].split()
        op = next((t for t in src if not t.startswith('@')), '?').split('.')[0]
        ops[op] += f(r, 'Instructions Executed')
    tot_i = sum(ops.values())
    print("opcode mix (warp instructions executed):")
    for op, v in ops.most_common(30):
        print(f"  {op:12s} {v:12.0f} {v / tot_i:.3f}")
