mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/gpu_tests.log
for c in ${CFGS:-cfg5 cfg3 cfg4}; do timeout 300 python bench.py --config $c --steps 96 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1])
    print("$c", "value=%.3e"%d['value'], "ms/step=%.4f"%d['ms_per_step'], "kernel_ms=%.4f"%d['roofline']['kernel_ms'], d['config']['kernel_path'], d['config']['timed_loop'])
except Exception as e: print("fail", e)
PY
done
