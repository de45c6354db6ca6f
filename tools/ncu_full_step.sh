mkdir -p gpurun_out
PAT="ffn_|mhsa_|fconv_tc|wgrad|lg_|squeeze|unsqueeze|stft"
s=$(date +%s)
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$PAT" -c 24 \
    -o gpurun_out/prof_r02ab -f python bench.py --profile --batch 32 --layers 1 > gpurun_out/prof_full.log 2>&1
echo "full capture rc=$? [$(( $(date +%s) - s )) s]"
ncu -i gpurun_out/prof_r02ab.ncu-rep --page raw --csv > gpurun_out/prof_r02ab_raw.csv 2>/dev/null
python tools/traffic_from_ncu.py gpurun_out/prof_r02ab_raw.csv 32 > gpurun_out/r02ab_traffic.json; cat gpurun_out/r02ab_traffic.json | head -30
python tools/ncu_summary.py gpurun_out/prof_r02ab_raw.csv > gpurun_out/r02ab_ncu_summary.md 2>&1; head -40 gpurun_out/r02ab_ncu_summary.md
ls -la gpurun_out/prof_r02ab.ncu-rep
