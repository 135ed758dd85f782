"""Summarise an ncu report (read here, no GPU): python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
print(f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none)\n")
names = [r[col["Kernel Name"]] for r in data]
for n, r in zip(names, data):
    short = n.split("(")[0]
    print(f"## {short}  (ID {r[col['ID']]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in col:
            print(f"| {w} | {r[col[w]]} | {units[col[w]]} |")
    tot = sum(float(r[col[s]] or 0) for s in stalls)
    print("\nwarp-state samples (pc sampling): " + ", ".join(
        f"{s.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * float(r[col[s]] or 0) / max(tot, 1):.1f}%"
        for s in sorted(stalls, key=lambda s: -float(r[col[s]] or 0))[:9]) + "\n")
