"""Key counters of an `ncu --set full` report (.ncu-rep), one block per captured launch: python scripts/ncu_full_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units = rows[0], rows[1]
for r in rows[2:]:
    rd, wr = None, None
    for w in WANT:
        if w in head:
            i = head.index(w)
            print(f"{w} = {r[i]} {units[i]}")
            if w == "dram__bytes_read.sum": rd = (float(r[i].replace(',', '')), units[i])
            if w == "dram__bytes_write.sum": wr = (float(r[i].replace(',', '')), units[i])
    print()
