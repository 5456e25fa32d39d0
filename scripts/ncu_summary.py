"""Summarise `ncu --set full` reports into the plain-text tables kept under profiles/.
Usage: python scripts/ncu_summary.py report.ncu-rep [...] > profiles/rN_xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    print(f"# {rep}")
    for r in data:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print(f"## {d.get('Kernel Name', '?')[:150]}")
        for k in WANT:
            if k in d:
                print(f"   {k:80s} {d[k]:>16s} {u.get(k, '')}")
    print()
