"""Tensor-pipe / TMEM / DRAM counters of the tcgen05 GEMM launches in an `ncu --set full` report:
python scripts/ncu_gemm_summary.py file.ncu-rep [role names in launch order]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "registers"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (% of elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor hmma-subpipe cycles active"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe inst (% of peak)"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe inst (% of peak)"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput (% of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput (% of peak)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active"),
    ("sm__cycles_elapsed.avg", "sm cycles elapsed"),
]


def main():
    rep = sys.argv[1]
    roles = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]

    def col(name):
        for i, h in enumerate(head):
            if h == name or h.endswith("." + name):
                return i
        return None

    for n, r in enumerate(rows[2:]):
        role = roles[n] if n < len(roles) else ""
        print(f"--- launch {n} {role}: {r[col('Kernel Name')][:80]}")
        for k, label in KEYS:
            i = col(k)
            if i is not None:
                print(f"  {label:38s} {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    main()
