#!/usr/bin/env python
"""Summarise .ncu-rep captures (read with `ncu -i ... --page raw --csv`, no GPU needed) into the text kept
under profiles/.  Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/summary.txt
With --traffic, prints {kernel: dram bytes read+written per launch} as JSON instead."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = name.split("(")[0]
    for junk in ("void ", "snpg::", "<unnamed>::", "unnamed>::"):
        name = name.replace(junk, "")
    return name


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    traffic = {}
    for path in args:
        hdr, units, rows = rows_of(path)
        seen = set()
        for r in rows:
            k = short(r[hdr.index("Kernel Name")])
            to_b = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            tb = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(m)
                tb += float(r[i].replace(",", "")) * to_b[units[i]]
            traffic.setdefault(k, []).append(tb)
            if k in seen or "--traffic" in sys.argv:
                continue
            seen.add(k)
            print(f"== {k}   (from {path.split('/')[-1]}, ncu --set full --clock-control none)")
            for key in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"   {key:92s} {r[i]:>16s} {units[i]}")
            for i, h in enumerate(hdr):
                if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                    v = float(r[i].replace(",", ""))
                    if v > 0.3:
                        print(f"   {h:92s} {v:16.3f}")
    if "--traffic" in sys.argv:
        print(json.dumps({k: sum(v) / len(v) for k, v in traffic.items()}, indent=1))


if __name__ == "__main__":
    main()
