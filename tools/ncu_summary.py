"""One line per profiled launch of an .ncu-rep: duration, tensor pipe, DRAM bytes, occupancy, registers."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
cols = [("Grid Size", "grid"), ("gpu__time_duration.sum", "dur"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem")]
print(f"# {rep}: ncu --set full --clock-control none (kernel replays are cold-cache)")
for r in rows[2:]:
    print(" | ".join(f"{n}={r[idx[c]]}{units[idx[c]]}" for c, n in cols if c in idx))
