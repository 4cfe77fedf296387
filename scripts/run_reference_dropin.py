#!/usr/bin/env python
"""Run the UNMODIFIED reference main.py (staged by scripts/stage_reference.py into baseline/_ref/) on top of
turboae_b200 through `python -m turboae_b200.launch` on a B200, and record what trainer.test / trainer.train print.

    python scripts/run_reference_dropin.py [--mode eval|train|scratch|all] [--out profiles/r02_dropin_*.json]

eval    : README command (1) (`-num_epoch 0`, shipped checkpoint), BER / BLER lists printed by the reference's own
          trainer.test (main.py:98-260 -> trainer.py:157-248), compared per point with tests/golden/ber_c1.json.
train   : README command (3) (fine-tune from the checkpoint), 1 epoch, -num_block 5000: trainer.train's loss.backward()
          / optimizer.step() on this package's autograd path; loss / BER trajectory recorded.
scratch : README command (2) (from scratch), 2 epochs, -num_block 5000.
c4      : BASELINE config 4 through the reference's own trainer: from scratch, batch 1000, -num_block 20000, 1 epoch (1 encoder + 5
          decoder passes of 20 steps, trainer.py:33-76 with its CPU bit / noise generation, host-to-device copies and per-step
          loss.item()); codewords/s of every pass from the "running time" trainer.train prints.  With --stock the same command on the
          reference's own classes (torch eager on the same GPU).
Everything the reference prints goes to gpurun_out/dropin_<mode>.log."""
import argparse
import ast
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("-encoder TurboAE_rate3_cnn -decoder TurboAE_rate3_cnn -enc_num_unit 100 -enc_num_layer 2 -dec_num_unit 100 "
          "-dec_num_layer 5 -num_iter_ft 5 -channel awgn -num_train_dec 5 -num_train_enc 1 -code_rate_k 1 -code_rate_n 3 "
          "-train_enc_channel_low 2.0 -train_enc_channel_high 2.0 -snr_test_start -1.5 -snr_test_end 4.0 -snr_points 12 "
          "-num_iteration 6 -is_parallel 1 -train_dec_channel_low -1.5 -train_dec_channel_high 2.0 -is_same_interleaver 1 "
          "-dec_lr 0.0001 -enc_lr 0.0001 -train_channel_mode block_norm -test_channel_mode block_norm --print_test_traj "
          "-loss bce").split()
CKPT = "./models/dta_cont_cnn2_cnn5_enctrain2_dectrainneg15_2.pt"


def run(mode, ref, extra, workdir, log_path, nproc=1, launch_opts=()):
    os.makedirs(workdir, exist_ok=True)
    link = os.path.join(workdir, "models")
    if not os.path.exists(link):
        os.symlink(os.path.join(ref, "models"), link)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
               "127.0.0.1", "--master-port", "29577", "-m", "turboae_b200.launch"]
    else:
        cmd = [sys.executable, "-m", "turboae_b200.launch"]
    cmd += list(launch_opts) + ["--reference", ref, "main.py"] + COMMON + extra
    t0 = time.time()
    with open(log_path, "w") as f:
        f.write("$ " + " ".join(cmd) + "\n")
        f.flush()
        rc = subprocess.call(cmd, cwd=workdir, env=env, stdout=f, stderr=subprocess.STDOUT)
    return rc, time.time() - t0


def parse_lists(text):
    out = {}
    for key in ("BER", "BLER"):
        m = re.search(r"^%s (\[.*\])$" % key, text, re.M)          # first occurrence = un-punctured list (trainer.py:238-239)
        if m:
            out[key] = ast.literal_eval(m.group(1))
    for key, tag in (("test loss trajectory", "loss_traj"), ("test ber trajectory", "ber_traj")):
        m = re.search(r"^%s (\[.*\])$" % key, text, re.M)
        if m:
            out[tag] = ast.literal_eval(m.group(1))
    out["epoch_lines"] = re.findall(r"^====> Epoch.*$", text, re.M)
    out["validate_lines"] = re.findall(r"^====> Test set.*$", text, re.M)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="eval", choices=["eval", "train", "scratch", "c4", "all"])
    ap.add_argument("--reference", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--num-block", type=int, default=1000000, help="eval: blocks per SNR point")
    ap.add_argument("--batch-size", type=int, default=50000, help="eval batch size")
    ap.add_argument("--nproc", type=int, default=1)
    ap.add_argument("--seed", type=int, default=None, help="launcher seed (same bits and noise in both arms)")
    ap.add_argument("--stock", action="store_true", help="run the reference's OWN classes (torch eager on the GPU): comparison arm")
    ap.add_argument("--no-tf32", action="store_true", help="torch's own CUDA kernels in true fp32 (the --stock arm then computes the reference's fp32 arithmetic)")
    ap.add_argument("--device-channel", action="store_true", help="launcher --device-channel: AWGN noise drawn on the GPU")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    ref = os.path.abspath(a.reference)
    if not os.path.isfile(os.path.join(ref, "main.py")):
        raise SystemExit("no staged reference at %s (run scripts/stage_reference.py in the build container)" % ref)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    results = {}
    modes = ["eval", "train", "scratch"] if a.mode == "all" else [a.mode]
    for mode in modes:
        if mode == "eval":
            extra = ["-num_block", str(a.num_block), "-batch_size", str(a.batch_size), "-init_nw_weight", CKPT, "-num_epoch", "0"]
        elif mode == "train":
            extra = ["-num_block", "5000", "-batch_size", "500", "-init_nw_weight", CKPT, "-num_epoch", "1"]
        elif mode == "c4":
            extra = ["-num_block", "20000", "-batch_size", "1000", "-num_epoch", "1"]
        else:
            extra = ["-num_block", "5000", "-batch_size", "500", "-num_epoch", "2"]
        tag = ("stock_" if a.stock else "dropin_") + mode + ("" if a.nproc == 1 else "_n%d" % a.nproc)
        log = os.path.join(out_dir, tag + ".log")
        opts = (["--seed", str(a.seed)] if a.seed is not None else []) + (["--stock"] if a.stock else []) + (["--no-tf32"] if a.no_tf32 else []) + (["--device-channel"] if a.device_channel else [])
        rc, secs = run(mode, ref, extra, os.path.join(out_dir, "dropin_work"), log, a.nproc, opts)
        text = open(log).read()
        r = parse_lists(text)
        r.update(rc=rc, seconds=secs, args=" ".join(extra), nproc=a.nproc, seed=a.seed, classes="reference (stock)" if a.stock else "turboae_b200", device_channel=bool(a.device_channel))
        if mode == "c4":
            secs_pass = [float(l.rsplit(" ", 1)[1]) for l in r["epoch_lines"]]
            r["pass_seconds"] = secs_pass                     # pass 0 = encoder mode (includes one-time setup), 1..5 = decoder mode
            r["train_cw_per_s_decoder_passes"] = [20000.0 / t for t in secs_pass[1:]]
            if len(secs_pass) > 2:
                r["train_cw_per_s_median"] = sorted(r["train_cw_per_s_decoder_passes"])[len(secs_pass[1:]) // 2]
        if mode == "eval" and "BER" in r:
            g = json.load(open(os.path.join(ROOT, "tests", "golden", "ber_c1.json")))
            gold = [e / (g["blocks"] * 100.0) for e in g["bit_errors"]]
            r["golden_ber"] = gold
            r["abs_diff"] = [abs(x - y) for x, y in zip(r["BER"], gold)]
            r["within_1e-4"] = [d <= 1e-4 for d in r["abs_diff"]]
            r["ber_0db"] = r["BER"][3]
        results[mode] = r
        print(mode, "rc", rc, "%.1f s" % secs, {k: v for k, v in r.items() if k in ("BER", "BLER", "abs_diff", "loss_traj", "ber_traj", "pass_seconds", "train_cw_per_s_median")}, flush=True)
        if rc != 0:
            print(text[-3000:])
    out = a.out or os.path.join(out_dir, "%s_%s.json" % ("stock" if a.stock else "dropin", a.mode))
    json.dump(results, open(out, "w"), indent=1)
    return 0 if all(r["rc"] == 0 for r in results.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
