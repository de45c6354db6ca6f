timeout 900 python -m pytest tests/test_gpu_online.py -m gpu -q --tb=short 2>&1 | tail -15
for B in 1 64; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/online_launches_b$B.csv python tools/online_step.py $B > gpurun_out/online_prof_b$B.log 2>&1
python tools/launch_summary.py gpurun_out/online_launches_b$B.csv | head -6
done
timeout 300 python - <<'PY'
import json, torch, bench
print(json.dumps(bench.bench_online(torch.device("cuda", 0)), indent=1))
PY
