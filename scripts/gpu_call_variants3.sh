#!/bin/bash
# README (1b): enc5/dec5 checkpoint evaluation; README (4): binarised codes (block_norm_ste) evaluation + 1 fine-tuning epoch
mkdir -p gpurun_out/dropin_work && cd gpurun_out/dropin_work && ln -sfn $OLDPWD/baseline/_ref/models models
export PYTHONPATH=$OLDPWD
C="-encoder TurboAE_rate3_cnn -decoder TurboAE_rate3_cnn -enc_num_unit 100 -dec_num_unit 100 -dec_num_layer 5 -num_iter_ft 5 -channel awgn -num_train_dec 5 -num_train_enc 1 -code_rate_k 1 -code_rate_n 3 -snr_test_start -1.5 -snr_test_end 4.0 -snr_points 12 -num_iteration 6 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -is_same_interleaver 1 --print_test_traj -loss bce"
echo "== (1b) enc5_dec5_cont_1dBenc.pt, -num_epoch 0"
timeout 900 python -m turboae_b200.launch --seed 9 --reference $OLDPWD/baseline/_ref main.py $C -enc_num_layer 5 -enc_kernel_size 5 -dec_kernel_size 5 -train_enc_channel_low 1.0 -train_enc_channel_high 1.0 -dec_lr 0.00005 -enc_lr 0.00005 -num_block 100000 -batch_size 1000 -train_channel_mode block_norm -test_channel_mode block_norm -optimizer adam -init_nw_weight ./models/enc5_dec5_cont_1dBenc.pt -num_epoch 0 2>&1 | grep -i "^BER\|^BLER\|rror\|Warn\|Traceback" | head -6
echo "== (4) block_norm_ste, dta_steq2 checkpoint: evaluation + 1 epoch"
timeout 900 python -m turboae_b200.launch --seed 9 --reference $OLDPWD/baseline/_ref main.py $C -enc_num_layer 2 -train_enc_channel_low 2.0 -train_enc_channel_high 2.0 -dec_lr 0.0001 -enc_lr 0.0001 -num_block 5000 -batch_size 500 -train_channel_mode block_norm_ste -test_channel_mode block_norm_ste -init_nw_weight ./models/dta_steq2_cnn2_cnn5_enctrain2_dectrainneg15_2.pt -num_epoch 1 2>&1 | grep -i "Epoch\|Test set\|^BER\|^BLER\|rror\|Warn\|Traceback" | head -14
