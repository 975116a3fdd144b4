"""Text summary of an `ncu --set full` report for profiles/: per captured launch the duration, DRAM bytes, achieved DRAM throughput,
tensor-pipe activity, occupancy and register count.   python tools/ncu_kernel_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: {len(data)} captured launches (ncu --set full --clock-control none)")
for r in data:
    print(f"\n{r[idx['Kernel Name']][:110]}")
    for key, label in WANT:
        if key in idx:
            print(f"    {label:34s} {r[idx[key]]:>18s} {units[idx[key]]}")
