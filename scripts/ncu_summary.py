"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.md]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
]
out = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    out.append(f"### {name[:110]}")
    for k in want:
        if k in hdr:
            out.append(f"- {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
    pipes = []
    for i, h in enumerate(hdr):
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum") or (h.startswith("sm__pipe_") and "pct_of_peak_sustained_active" in h):
            try:
                if float(r[i].replace(",", "")) > 0:
                    pipes.append(f"{h.replace('sm__inst_executed_pipe_', 'inst:').replace('.avg.pct_of_peak_sustained_active', '%').replace('sm__pipe_', 'busy:')}={r[i]}")
            except ValueError:
                pass
    out.append("- pipes: " + ", ".join(pipes))
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1
    out.append("- pc-sampling stall reasons: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:7]))
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
