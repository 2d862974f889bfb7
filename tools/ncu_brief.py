"""Side-by-side digest of `ncu --set full` reports: duration, pipes, issue, instruction cache, top stall reasons, memory.
usage: python tools/ncu_brief.py a.ncu-rep [b.ncu-rep ...]      (first profiled launch of each report)"""
import csv, io, subprocess, sys

FIXED = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
         "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "sm__warps_active.avg.per_cycle_active",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
         "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
reps = sys.argv[1:]
data = []
for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    data.append({h: (v, u) for h, u, v in zip(hdr, units, r)})
print("# ncu --set full --clock-control none --import-source on; one launch per column (cold-cache replays, never a bench value)")
print("# " + " | ".join(f"{d.get('Kernel Name', ('?', ''))[0].split('(')[0][-28:]} grid {d.get('launch__grid_size', ('?', ''))[0]}" for d in data))
def row(name):
    vals = []
    for d in data:
        v, u = d.get(name, ("", ""))
        vals.append(f"{v:>18s} {u}".rstrip() if v else f"{'-':>18s}")
    print(f"{name:96s} " + " ".join(vals))
for m in FIXED:
    row(m)
stall = sorted({h for d in data for h in d if "issue_stalled" in h and h.endswith("per_issue_active.ratio")},
               key=lambda h: -max(float(d.get(h, ("0", ""))[0] or 0) for d in data))
for m in stall[:9]:
    row(m)
