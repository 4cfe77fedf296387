#!/bin/bash
# README.md:94 / :98 of the reference through the launcher: classical turbo encoder (the reference's own, CommPy on the host) with
# this package's DEC_LargeRNN / dense DEC_LargeCNN; tiny runs (1 epoch)
mkdir -p gpurun_out/dropin_work && cd gpurun_out/dropin_work && ln -sfn $OLDPWD/baseline/_ref/models models
export PYTHONPATH=$OLDPWD
for DEC in TurboAE_rate3_cnn TurboAE_rate3_rnn; do
echo "== -encoder Turbo_rate3_757 -decoder $DEC"
timeout 900 python -m turboae_b200.launch --seed 5 --reference $OLDPWD/baseline/_ref main.py -encoder Turbo_rate3_757 -decoder $DEC -dec_num_unit 100 -dec_num_layer 5 -num_iter_ft 5 -channel awgn -num_train_dec 2 -code_rate_k 1 -code_rate_n 3 -snr_test_start 0.0 -snr_test_end 2.0 -snr_points 2 -num_iteration 2 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -is_same_interleaver 1 -dec_lr 0.0001 -num_block 400 -batch_size 100 -block_len 100 -num_epoch 1 --print_test_traj 2>&1 | grep -i "Epoch\|Test set\|^BER\|rror\|Warn\|Traceback" | head -10
done
