# One short GPU call that validates the head: the tests added last first, then the whole -m gpu suite (what the driver runs), smoke, and a
# quick bench line (no eager / extra / CPU legs).  usage (from the repo root on the GPU box): bash tools/final_check.sh TAG
mkdir -p gpurun_out
TAG=${1:-r02ai}
s=$(date +%s); timeout 150 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider -k "known_answer or resume_across or clip_adam" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/${TAG}_pytest_new.log) [$(( $(date +%s) - s )) s]"
s=$(date +%s); timeout 330 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${TAG}_pytest_gpu.log) [$(( $(date +%s) - s )) s]"
s=$(date +%s); timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/${TAG}_smoke.log) [$(( $(date +%s) - s )) s]"
s=$(date +%s); timeout 150 python bench.py --no-eager-baseline --no-nbc2 --no-cpu-baseline > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench_quick.log; echo "bench rc=$? [$(( $(date +%s) - s )) s] $(head -c 600 gpurun_out/${TAG}_bench_quick.json)"
