#!/bin/bash
# round 2 evidence: new GPU tests, ncu launch list + full capture of the fused decoder, full capture of the GRU recurrence kernel
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_tc.py -m gpu -q -x -k "rnn or gru or flags or norm_stats or variable or two_forwards" -s 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/${TAG}_newtests.log
export BENCH_SKIP_SECONDARY=1
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1; tail -2 gpurun_out/${TAG}_launches.log | cut -c1-200
echo "== dec full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_pair -s 4 -c 1 -o gpurun_out/${TAG}_dec -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_dec_full.log 2>&1; tail -2 gpurun_out/${TAG}_dec_full.log | cut -c1-200
echo "== gru full"; RNN_B=18944 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gru_pair -s 30 -c 1 -o gpurun_out/${TAG}_gru -f \
    python scripts/rnn_bench.py > gpurun_out/${TAG}_gru_full.log 2>&1; tail -2 gpurun_out/${TAG}_gru_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
