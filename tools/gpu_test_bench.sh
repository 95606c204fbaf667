mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo tests rc=$?; tail -8 gpurun_out/gpu_tests.log
for c in ${CFGS:-cfg3 cfg2 cfg5 cfg4}; do timeout 300 python bench.py --config $c --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo $c rc=$?; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1])
    print("$c", "value=%.3e"%d['value'], "ms/step=%.4f"%d['ms_per_step'], "kernel_ms=%.4f"%d['roofline']['kernel_ms'], "e2e=%.3e"%d['e2e']['value'], "e2e_ms=%.3f"%d['e2e']['ms_per_step'], d['config']['kernel_path'])
except Exception as e: print("fail", e)
PY
tail -3 gpurun_out/bench_$c.err; done
B4="python bench.py --config cfg4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B4 > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg4.csv $B4 > gpurun_out/ncu_l4.log 2>&1; echo rc=$?
