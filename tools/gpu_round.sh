#!/bin/bash
# One gpurun call of a build/measure round: GPU tests (grouped, each under its own timeout), the per-phase profile of the
# slab kernels (profiling build), and the default bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r02}
run() { name=$1; shift; echo "=== $name"; timeout 600 python -m pytest "$@" -m gpu -q -p no:cacheprovider -s > gpurun_out/t_$name.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/t_$name.log)"; }
run self tests/test_umma_selftest.py
run blocks tests/test_gpu_blocks.py
run model tests/test_gpu_model.py
run parity tests/test_gpu_parity_r2.py
if [ -f nbss_b200/lib/libnbss_b200_prof.so ]; then
  echo "=== phases"; NBSS_LIB=nbss_b200/lib/libnbss_b200_prof.so timeout 300 python tools/phase_profile.py --batch 8 --out gpurun_out/${TAG}_phases.json > gpurun_out/${TAG}_phases.txt 2>&1; echo "rc=$?"
fi
if [ -f nbss_b200/lib/libnbss_b200_exact.so ]; then
  echo "=== A/B sigmoid"; for L in "" _exact; do NBSS_LIB=nbss_b200/lib/libnbss_b200$L.so timeout 300 python tools/ab_forward.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_ab.txt
fi
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "rc=$? $(head -c 300 gpurun_out/${TAG}_bench.json)"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
