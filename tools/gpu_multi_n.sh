# usage: gpu_multi_n.sh N  -- weak-scaling bench line (+ reference arm) under torchrun on N GPUs of one box
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r01_bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 100 --warmup 5 --scaling strong --no-cpu-baseline --no-e2e > gpurun_out/r01_bench_cfg3_strong_n$N.json 2> gpurun_out/bench_cfg3_strong_n$N.err; echo rc=$?
for f in cfg3_n$N cfg3_strong_n$N; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r01_bench_$f.json').read().strip().splitlines()[-1])
    print("$f", "value=%.3e"%d['value'], "ms/step=%.4f"%d['ms_per_step'], d['scaling'], d['n_gpus'], d['config']['envs_per_gpu'], d['episode_stats']['episodes'])
except Exception as e: print("fail $f", e)
PY
done
