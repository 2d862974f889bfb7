"""Per-kernel totals of an ncu launch list taken with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file X.csv ...
usage: python tools/launch_metrics_summary.py X.csv"""
import collections, csv, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ix = {h: j for j, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])


def val(d, key, scale):
    v, u = d.get(key, (0.0, ""))
    return v * scale.get(u, 1.0)


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for d in per.values():
    name = d["name"].replace("mcvd::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0][:28]
    us = val(d, "gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3})
    a = agg[name]
    a[0] += 1
    a[1] += us
    a[2] += val(d, "dram__bytes_read.sum", {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3})
    a[3] += val(d, "dram__bytes_write.sum", {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3})
    a[4] += us * val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", {})
tot = sum(a[1] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
      "sm__pipe_tensor_cycles_active...pct --clock-control none")
print("# one forward, every launch; per kernel: launches, total us (cold-cache, serialised), DRAM read MB, "
      "DRAM write MB, time-weighted tensor-pipe %")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:30s} n={a[0]:3d} {a[1]:9.1f} us ({100 * a[1] / tot:4.1f}%)  dram_rd {a[2]:9.1f} MB  "
          f"dram_wr {a[3]:9.1f} MB  tensor {a[4] / max(a[1], 1e-9):5.1f}%")
print(f"total {tot:.1f} us, {sum(a[0] for a in agg.values())} launches, "
      f"DRAM {sum(a[2] + a[3] for a in agg.values()) / 1e3:.2f} GB")
