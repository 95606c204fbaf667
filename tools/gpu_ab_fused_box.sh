mkdir -p gpurun_out
for v in 0 1; do
  export SWARM_FUSED_BOX=$v
  for c in ${CFGS:-cfg3 cfg2}; do
    timeout 300 python bench.py --config $c --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_box${v}_$c.json 2> gpurun_out/ab_box${v}_$c.err || tail -3 gpurun_out/ab_box${v}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_box${v}_$c.json').read().strip().splitlines()[-1])
    print("box=$v $c", "value=%.3e"%d['value'], "ms/step=%.4f"%d['ms_per_step'], "kernel_ms=%.4f"%d['roofline']['kernel_ms'], d['config']['kernel_path'])
except Exception as e: print("fail", e)
PY
  done
done
unset SWARM_FUSED_BOX
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo tests rc=$?; tail -8 gpurun_out/gpu_tests.log
