timeout 600 python - <<'PY'
import json, torch, bench
print(json.dumps(bench.bench_long(torch.device("cuda", 0)), indent=1))
PY
