mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_b4.csv python bench.py --profile --batch 4 > gpurun_out/prof_b4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b4.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_b32.csv python bench.py --profile --batch 32 > gpurun_out/prof_b32.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b32.csv
timeout 300 python bench.py --batch 4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-nbc2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch4 ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
