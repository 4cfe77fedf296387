#!/bin/bash
mkdir -p gpurun_out/dropin_work && cd gpurun_out/dropin_work && ln -sfn $OLDPWD/baseline/_ref/models models
export PYTHONPATH=$OLDPWD
C="-encoder TurboAE_rate3_cnn -decoder TurboAE_rate3_cnn -enc_num_unit 100 -enc_num_layer 2 -dec_num_unit 100 -dec_num_layer 5 -num_iter_ft 5 -channel awgn -num_train_dec 2 -num_train_enc 1 -code_rate_k 1 -code_rate_n 3 -train_enc_channel_low 2.0 -train_enc_channel_high 2.0 -snr_test_start 0.0 -snr_test_end 2.0 -snr_points 2 -num_iteration 6 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -dec_lr 0.0001 -enc_lr 0.0001 -num_block 2000 -batch_size 500 -train_channel_mode block_norm -test_channel_mode block_norm --print_test_traj -loss bce -num_epoch 1 -is_same_interleaver 1"
run() { echo "== $1"; shift; timeout 600 python -m turboae_b200.launch --seed 4 --reference $OLDPWD/baseline/_ref main.py $C "$@" 2>&1 | grep -i "Epoch\|Test set\|^BER\|rror\|Warn\|Traceback" | head -8; }
run "optimizer lookahead" -optimizer lookahead
run "variable block length" --is_variable_block_len -block_len_low 50 -block_len_high 150
run "sgd, mse loss, bec channel" -optimizer sgd -loss mse -channel bec -bec_p 0.1
