"""Markdown table of the key counters of every kernel in an `ncu --page raw --csv` export:
    python scripts/ncu_table.py gpurun_out/r2_kernels_raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(d, name, scale=1.0, fmt="{:.1f}"):
    if name not in col or d[col[name]] in ("", "n/a"):
        return "-"
    try:
        return fmt.format(float(d[col[name]].replace(",", "")) * scale)
    except ValueError:
        return d[col[name]]


units = rows[1]
print("| kernel | us | grid | regs | warps active % | issue active % | tensor pipe % | DRAM rd MB | DRAM wr MB | DRAM % of peak | L2 % | achieved GB/s (DRAM) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for d in data:
    name = d[col["Kernel Name"]].replace("void ", "").split("(")[0]
    us = float(d[col["gpu__time_duration.sum"]].replace(",", ""))
    if units[col["gpu__time_duration.sum"]] == "ns":
        us /= 1e3

    def mb(n):
        if n not in col:
            return 0.0
        v = float(d[col[n]].replace(",", ""))
        u = units[col[n]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    print(f"| `{name}` | {us:.1f} | {get(d, 'launch__grid_size', fmt='{:.0f}')} | {get(d, 'launch__registers_per_thread', fmt='{:.0f}')} | "
          f"{get(d, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | {get(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{get(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {rd:.1f} | {wr:.1f} | "
          f"{get(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {get(d, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{(rd + wr) / us * 1e3:.0f} |")
