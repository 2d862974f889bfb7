"""profiles/ncu_conv_umma_traffic.json (what bench.py reports as roofline.traffic) from an ncu launch list taken with
tools/gpu/r2_final.sh: DRAM bytes and tensor-pipe activity of every conv launch (k_conv_umma + k_conv1x1_umma) of one
cfg2 B=64 forward.   usage: python tools/traffic_json.py profiles/X.csv > profiles/ncu_conv_umma_traffic.json"""
import collections, csv, json, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: j for j, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows:
    if len(r) < len(hdr) or r is hdr or "k_conv" not in r[ix["Kernel Name"]]:
        continue
    d = per.setdefault(r[ix["ID"]], {"k1": "1x1" in r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
n = len(per)
byt = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values())
t = sum(d["gpu__time_duration.sum"] for d in per.values())
tp = sum(d["gpu__time_duration.sum"] * d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for d in per.values())
print(json.dumps({
    "kernel": "k_conv_umma + k_conv1x1_umma",
    "launches_per_forward": n,
    "launches_conv1x1": sum(1 for d in per.values() if d["k1"]),
    "dram_bytes_per_forward": byt,
    "dram_bytes_per_launch_mean": byt / n,
    "tensor_pipe_pct_time_weighted": tp / t,
    "source": sys.argv[1] + ": ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... over every launch of one cfg2 B=64 forward",
}, indent=1))
