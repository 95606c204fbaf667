mkdir -p gpurun_out
for f in "$@"; do
  export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so
  for c in ${CFGS:-cfg5}; do
    timeout 300 python bench.py --config $c --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_${f}_$c.json 2> gpurun_out/ab_${f}_$c.err || tail -3 gpurun_out/ab_${f}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_${f}_$c.json').read().strip().splitlines()[-1])
    print("$f $c", "value=%.3e"%d['value'], "ms/step=%.4f"%d['ms_per_step'], "kernel_ms=%.4f"%d['roofline']['kernel_ms'])
except Exception as e: print("fail", e)
PY
  done
done
