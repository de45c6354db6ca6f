#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one training step (device time per launch; compare SHARES),
# (2) --set full captures of selected kernels (1 GPU only).  Usage: tools/profile.sh <regex> <batch> <tag>
mkdir -p gpurun_out
PAT=${1:-ffn_bwd_kernel}
BATCH=${2:-32}
TAG=${3:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --profile --batch ${LBATCH:-$BATCH} > gpurun_out/prof_launch.log 2>&1
echo "launch list rc=$?"
# full captures on a ONE-layer network: one step then holds exactly one launch of every kernel (fwd and bwd)
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$PAT" -c 24 \
    -o gpurun_out/prof_$TAG -f python bench.py --profile --batch $BATCH --layers 1 > gpurun_out/prof_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
