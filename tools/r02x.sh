timeout 900 python -m pytest tests/test_gpu_blocks.py -m gpu -q -p no:cacheprovider -k "ffn or persistent" 2>&1 | tail -3
timeout 300 python tools/time_kernels.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-nbc2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
timeout 300 python bench.py --batch 4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-nbc2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch 4 ms/step', d['ms_per_step'])"
