#!/usr/bin/env python
"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) kernel by kernel: duration, DRAM bytes, tensor-pipe activity,
L2 traffic, registers.  Optional per-launch work (TFLOP or GB) to print the achieved rate:

    python tools/summarize_ncu.py gpurun_out/r2p/conv_full.ncu-rep [--work name=tflop ...] [--skip N]
"""
import csv
import io
import subprocess
import sys

WANT = ["launch__grid_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def main():
    path = sys.argv[1]
    work, skip = {}, 0
    args = sys.argv[2:]
    while args:
        a = args.pop(0)
        if a == "--skip":
            skip = int(args.pop(0))
        elif a == "--work":
            while args and not args[0].startswith("--"):
                k, v = args.pop(0).split("=")
                work[k] = float(v)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(data[skip:]):
        name = r[col["Kernel Name"]]
        print(f"## launch {n + skip}: {name[:110]}")
        for m in WANT:
            if m in col:
                print(f"{m:80s} {r[col[m]]} {units[col[m]]}")
        dur_us = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1)
        rd = to_bytes(r[col["dram__bytes_read.sum"]].replace(",", ""), units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]].replace(",", ""), units[col["dram__bytes_write.sum"]])
        print(f"{'-> DRAM traffic (read + write) per launch':80s} {(rd + wr) / 1e9:.4f} GB  = {(rd + wr) / 1e3 / dur_us:.1f} GB/s")
        for k, v in work.items():
            if k in name:
                print(f"{'-> ' + str(v) + ' TFLOP per launch / duration':80s} {v / (dur_us * 1e-6):.1f} TFLOP/s")
        print()


if __name__ == "__main__":
    main()
