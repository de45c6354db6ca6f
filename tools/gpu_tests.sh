#!/bin/bash
# Runs the GPU test groups separately, each under its own timeout, so one hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 240 python -m pytest "$@" -m gpu -q -p no:cacheprovider > gpurun_out/t_$name.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/t_$name.log)"; }
run self tests/test_umma_selftest.py
run simt tests/test_gpu_blocks.py -k "fconv_fwd_bwd or full or lg_tc or encoder or stft"
run fconv_tc tests/test_gpu_blocks.py -k "fconv_tc"
run ffn_fwd tests/test_gpu_blocks.py -k "ffn_fwd"
run mhsa_fwd tests/test_gpu_blocks.py -k "mhsa_fwd"
run ffn_bwd tests/test_gpu_blocks.py -k "ffn_bwd"
run mhsa_bwd tests/test_gpu_blocks.py -k "mhsa_bwd"
if [ "$1" == "model" ]; then run model tests/test_gpu_model.py; fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
