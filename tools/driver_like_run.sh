# what the driver runs at round end, in its order: GPU tests, smoke, reference arm, bench
mkdir -p gpurun_out
TAG=${1:-r02f}
s=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest.log) [$(( $(date +%s) - s )) s]"
s=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/${TAG}_smoke.log) [$(( $(date +%s) - s )) s]"
s=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "ref arm rc=$? [$(( $(date +%s) - s )) s] $(head -c 400 gpurun_out/${TAG}_bench_reference.json)"
s=$(date +%s); timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$? [$(( $(date +%s) - s )) s] lines=$(wc -l < gpurun_out/${TAG}_bench.json)"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
    print(k, json.dumps(d.get(k)))
print("eager", json.dumps(d.get("gpu_eager_baseline"))[:600])
print("extra", json.dumps(d.get("extra_workloads"))[:1500])
PY
