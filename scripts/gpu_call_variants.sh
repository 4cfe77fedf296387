#!/bin/bash
# the unmodified reference main.py through the launcher with NON-default shapes: fused kernels at other sizes, and a shape that falls to
# the fp32 kernels (block length 1000)
mkdir -p gpurun_out/dropin_work && cd gpurun_out/dropin_work && ln -sfn $OLDPWD/baseline/_ref/models models
export PYTHONPATH=$OLDPWD
COMMON="-encoder TurboAE_rate3_cnn -decoder TurboAE_rate3_cnn -channel awgn -num_train_dec 2 -num_train_enc 1 -code_rate_k 1 -code_rate_n 3 -train_enc_channel_low 2.0 -train_enc_channel_high 2.0 -snr_test_start 0.0 -snr_test_end 3.0 -snr_points 3 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -is_same_interleaver 1 -dec_lr 0.0001 -enc_lr 0.0001 -train_channel_mode block_norm -test_channel_mode block_norm --print_test_traj -loss bce"
echo "== variant 1: L=200, 64 units, 3 layers, 4 iterations, F=3, training from scratch 1 epoch"
timeout 600 python -m turboae_b200.launch --seed 5 --reference $OLDPWD/baseline/_ref main.py $COMMON -enc_num_unit 64 -enc_num_layer 3 -dec_num_unit 64 -dec_num_layer 3 -num_iter_ft 3 -num_iteration 4 -block_len 200 -num_block 2000 -batch_size 500 -num_epoch 1 2>&1 | grep -i "Epoch\|Test set\|^BER\|error\|Warn" | head -12
echo "== variant 2: L=1000 evaluation (fp32 kernels), 1 iteration-pair model from scratch"
timeout 600 python -m turboae_b200.launch --seed 5 --reference $OLDPWD/baseline/_ref main.py $COMMON -enc_num_unit 100 -enc_num_layer 2 -dec_num_unit 100 -dec_num_layer 5 -num_iter_ft 5 -num_iteration 2 -block_len 1000 -num_block 200 -batch_size 100 -num_epoch 0 2>&1 | grep -i "^BER\|^BLER\|error\|Warn" | head -8
echo "== variant 3: README line 98 style: classical turbo encoder + CNN decoder (dense variant), tiny run"
timeout 600 python -m turboae_b200.launch --seed 5 --reference $OLDPWD/baseline/_ref main.py -encoder Turbo_rate3_757 -decoder TurboAE_rate3_cnn -dec_num_unit 100 -dec_num_layer 5 -num_iter_ft 5 -channel awgn -num_train_dec 1 -code_rate_k 1 -code_rate_n 3 -snr_test_start 0.0 -snr_test_end 2.0 -snr_points 2 -num_iteration 2 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -is_same_interleaver 1 -dec_lr 0.0001 -num_block 200 -batch_size 100 -block_len 100 -num_epoch 1 --print_test_traj 2>&1 | grep -i "Epoch\|Test set\|^BER\|rror\|Warn" | head -10
