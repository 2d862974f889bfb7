"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + top launches."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0.0, 0]); allr = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = row["Kernel Name"].replace("mcvd::", "").replace("<unnamed>::", "")[:44]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
    tot[name][0] += v; tot[name][1] += 1; allr.append((name, v, row["Grid Size"]))
s = sum(v[0] for v in tot.values())
print(f"{path}: {sum(v[1] for v in tot.values())} launches, {s:.3f} ms total (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:46s} {v[0]:8.3f} ms {v[1]:4d} launches {100 * v[0] / s:5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("top launches:")
for r in sorted(allr, key=lambda r: -r[1])[:n]: print("  %-44s %8.3f ms grid %s" % r)
