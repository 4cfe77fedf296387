#!/bin/bash
# ncu --set full of the decoder's weight-gradient launch with the tap-split jobs (single-launch backward for the capture)
mkdir -p gpurun_out
TRAIN_STEPS=1 TURBOAE_B200_WGRAD_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 \
    -o gpurun_out/r02_wgrad_tapsplit -f python scripts/train_bench.py > gpurun_out/r3o_wgrad.log 2>&1
tail -2 gpurun_out/r3o_wgrad.log
ncu -i gpurun_out/r02_wgrad_tapsplit.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h, u, r = rows[0], rows[1], rows[2]
want = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
for a, b, c in zip(h, u, r):
    if a in want: print(a, c, b)
"
