#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one training step (device time per launch; compare SHARES),
# (2) one --set full capture of the kernels named in $1 (regex), 1 GPU only.
mkdir -p gpurun_out
PAT=${1:-ffn_bwd_kernel}
BATCH=${2:-32}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile --batch $BATCH > gpurun_out/prof_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$PAT -c 2 \
    -o gpurun_out/prof_top -f python bench.py --profile --batch $BATCH > gpurun_out/prof_full.log 2>&1
echo "full capture rc=$?"
