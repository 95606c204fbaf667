# predator kernel A/B: kernel tests on the default build, then tools/kernel_roofline.py predator per library variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k predator 2>&1 | tail -5 > gpurun_out/pred_tests.txt; tail -3 gpurun_out/pred_tests.txt
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  echo "== $f"
  timeout 300 python tools/kernel_roofline.py predator 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['kernel'], '|', d['shape'], '|', round(d['ms_mean'], 4), 'ms | frac', round(d['frac'], 3))
"
done 2>&1 | tee gpurun_out/pred_ab.txt
