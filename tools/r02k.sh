timeout 900 python -m pytest tests/test_gpu_blocks.py -m gpu -q -p no:cacheprovider -k "mhsa" 2>&1 | tail -3
timeout 300 python tools/time_kernels.py 2>&1 | tail -1
timeout 300 python - <<'PY'
import json, torch, bench
print(json.dumps(bench.bench_online(torch.device("cuda", 0)), indent=1))
PY
