#!/bin/bash
# One gpurun call of a build/measure round.  Everything lands in gpurun_out/.
#   TAG (arg 1)         prefix of the output files
#   TESTS=all|fast|none GPU test groups (fast skips the fp64-oracle parity group, ~2.5 min of host time)
#   AB="_nopf _exact"   extra library variants for tools/time_kernels.py (A/B against the default build)
#   BENCH=1|0, BENCH_ARGS
mkdir -p gpurun_out
TAG=${1:-r02}
TESTS=${TESTS:-all}; BENCH=${BENCH:-1}
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q -p no:cacheprovider -s > gpurun_out/t_$name.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/t_$name.log)"; }
if [ "$TESTS" != "none" ]; then
  run self tests/test_umma_selftest.py
  run blocks tests/test_gpu_blocks.py
  run model tests/test_gpu_model.py
  if [ "$TESTS" == "all" ]; then run parity tests/test_gpu_parity_r2.py; fi
  for f in $EXTRA_TESTS; do run $(basename $f .py) $f; done
fi
if [ -f nbss_b200/lib/libnbss_b200_prof.so ]; then
  echo "=== phases"; NBSS_LIB=nbss_b200/lib/libnbss_b200_prof.so timeout 300 python tools/phase_profile.py --batch 8 --out gpurun_out/${TAG}_phases.json > gpurun_out/${TAG}_phases.txt 2>&1; echo "rc=$?"
fi
echo "=== kernel times"; for L in "" $AB; do NBSS_LIB=nbss_b200/lib/libnbss_b200$L.so timeout 300 python tools/time_kernels.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_kernels.txt
if [ "$BENCH" == "1" ]; then
  echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "rc=$? $(head -c 300 gpurun_out/${TAG}_bench.json)"
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
