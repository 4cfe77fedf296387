#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key metrics of each captured kernel.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [substring ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_alu",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_uniform",
        "sm__warps_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem", "sm__throughput.avg.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg ",
        "sm__cycles_active.avg", "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__issue_active.avg", "smsp__inst_issued.avg", "tensor", "tmem", "smsp__pcsamp_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = dict(zip(hdr, r)).get("Kernel Name", "?")
        print("==", name)
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS + extra) and v not in ("", "0", "n/a"):
                print("  %-90s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
