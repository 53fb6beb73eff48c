"""Summarise an .ncu-rep here (no GPU): headline metrics + hottest SASS lines.  python dev/ncu_summary.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg"]
for k in range(2, len(rows)):
    print("---- kernel", rows[k][h.index("Kernel Name")][:80] if "Kernel Name" in h else k)
    for n in want:
        if n in h:
            print(f"  {n:70s} {rows[1][h.index(n)]:>10s} {rows[k][h.index(n)]}")
    for i, c in enumerate(h):
        if c.startswith("smsp__pcsamp_warps_issue_stalled") and not c.endswith("_not_issued"):
            v = rows[k][i]
            if v and float(v) > 0:
                print(f"  {c[33:]:40s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ia = h.index("Instructions Executed"); isamp = h.index("# Samples"); s = h.index("Source")
data = [(i, r[s].strip(), int(r[ia]), int(r[isamp])) for i, r in enumerate(rows[hi + 1:]) if len(r) > ia and r[ia].isdigit()]
tot = sum(d[2] for d in data); ts = max(1, sum(d[3] for d in data))
print("total warp-instr", tot, "samples", ts)
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"  {d[0]:5d} inst {100 * d[2] / tot:5.2f}% samp {100 * d[3] / ts:5.2f}%  {d[1][:100]}")
