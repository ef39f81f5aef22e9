"""Per captured launch of an `ncu --set full` report: duration, DRAM bytes, occupancy and the top warp-stall reasons.
python scripts/ncu_stalls.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units = rows[0], rows[1]
def col(r, name):
    return float(r[head.index(name)].replace(',', '')) if name in head and r[head.index(name)] not in ('', 'n/a') else float('nan')
stall = [i for i, h in enumerate(head) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    name = r[head.index("Kernel Name")][:70]
    t = col(r, "gpu__time_duration.sum"); tu = units[head.index("gpu__time_duration.sum")]
    rd = col(r, "dram__bytes_read.sum"); ru = units[head.index("dram__bytes_read.sum")]
    wr = col(r, "dram__bytes_write.sum"); wu = units[head.index("dram__bytes_write.sum")]
    print(f"{name} grid {r[head.index('launch__grid_size')]} block {r[head.index('launch__block_size')]} regs {r[head.index('launch__registers_per_thread')]}")
    print(f"  time {t} {tu}  dram R {rd} {ru} W {wr} {wu}  dram% {col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}"
          f"  warps_active% {col(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}  issue_active% {col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}"
          f"  L2hit% {col(r,'lts__t_sector_hit_rate.pct'):.1f}  inst {col(r,'smsp__inst_executed.sum'):.0f}")
    st = sorted(((float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0, head[i]) for i in stall), reverse=True)[:4]
    print("  stalls: " + ", ".join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')} {v:.2f}" for v, h in st))
