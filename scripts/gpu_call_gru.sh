#!/bin/bash
# GRU recurrence with the input part of step t+1 issued during the epilogue of step t (libturboae_b200_e.so): parity tests + bench
mkdir -p gpurun_out
export TURBOAE_B200_LIB=$PWD/turboae_b200/lib/libturboae_b200_e.so
echo "== gru tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rnn or gru" 2>&1 | tail -6 | tee gpurun_out/r02_gru_e_tests.log
echo "== rnn bench 18944"; RNN_B=18944 timeout 300 python scripts/rnn_bench.py 2>/dev/null | tail -1 | tee gpurun_out/r02_gru_e_bench.json
echo "== rnn bench 2368"; RNN_B=2368 timeout 300 python scripts/rnn_bench.py 2>/dev/null | tail -1
echo "== rnn bench 40000 (multi-pair)"; RNN_B=40000 RNN_L=200 timeout 300 python scripts/rnn_bench.py 2>/dev/null | tail -1
