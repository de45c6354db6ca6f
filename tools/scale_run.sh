N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-nbc2 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.log
echo "rc=$?"; tail -3 gpurun_out/scale_n$N.log; python -c "
import json; d=json.load(open('gpurun_out/scale_n$N.json')); print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','e2e','gpu_launches')})"
