"""Warp-stall samples per CUDA source line of an `ncu --set full --import-source on` report (needs -lineinfo).
usage: python tools/ncu_source_top.py X.ncu-rep [N]"""
import collections, csv, io, subprocess, sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur, hdr, agg, tot = None, None, collections.OrderedDict(), 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= 8 and r[2] == "-":          # per-source-line summary rows carry no SASS address
        try:
            s, ex = int(r[6]), int(r[7])
        except ValueError:
            continue
        agg[(cur, int(r[0]))] = (s, ex, r[1].strip()[:110])
        tot += s
print(f"# {rep}: {tot} warp-stall samples; the {n} source lines with the most samples (share, warp-instructions executed)")
for (f, l), (s, ex, src) in sorted(sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]):
    print(f"{f}:{l:<4d} {100 * s / max(tot, 1):5.1f}%  ex={ex:>9d}  {src}")
