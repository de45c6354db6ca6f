timeout 600 python -m pytest tests/test_gpu_predict.py tests/test_gpu_online.py -m gpu -q --tb=short 2>&1 | tail -5
for B in 1 64; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/online_launches_b$B.csv python tools/online_step.py $B > gpurun_out/online_prof_b$B.log 2>&1
python tools/launch_summary.py gpurun_out/online_launches_b$B.csv
done
