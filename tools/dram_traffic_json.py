"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> the JSON bench.py reads for roofline.traffic:
    python tools/dram_traffic_json.py gpurun_out/r2_gemm_traffic.csv <launches per step> > profiles/r2_dram_traffic.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
per_step = int(sys.argv[2])
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
iid, im, iv, iu = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0}
ids, tot = set(), {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
tens = []
for r in rows[start:]:
    if len(r) <= iv:
        continue
    ids.add(r[iid])
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    if r[im] in tot:
        tot[r[im]] += v
    elif r[im].startswith("sm__pipe_tensor"):
        tens.append(v)
n = len(ids)
print(json.dumps({
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (tools/ncu_r2.sh): the "
              "gemm_bf16_tc_kernel launches of one eager step of `bench.py --steps 2 --warmup 3 --no-graph` (8 clips, window 8x7x7)",
    "gemm_bf16_tc_kernel": {
        "launches": n, "launches_per_step": per_step, "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
        "dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
        "ncu_time_ms": tot["gpu__time_duration.sum"], "tensor_pipe_pct_mean": sum(tens) / max(len(tens), 1)}}, indent=1))
