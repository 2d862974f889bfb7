"""Top stall locations of a kernel from an .ncu-rep (source page).  usage: ncu_top.py rep [kernel-id] [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "1"; n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
body = rows[2:]
data = [(int(r[isamp] or 0), int(r[iex] or 0), r[ia].strip(), i) for i, r in enumerate(body) if len(r) > isamp]
tot = sum(d[0] for d in data)
print(rows[0][:2], "samples", tot, "instr", sum(d[1] for d in data))
for d in sorted(data, key=lambda d: -d[0])[:n]:
    print("%6d %5.1f%% exec %9d  line %4d  %s" % (d[0], 100 * d[0] / tot, d[1], d[3], d[2][:100]))
if len(sys.argv) > 4:
    a, b = map(int, sys.argv[4].split(":"))
    for i in range(a, b): print("%5d %6s  %s" % (i, body[i][isamp], body[i][ia].strip()[:110]))
