# packed thread-per-env kernel: parity first, then cfg5 / stream timings per library variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tiny or golden_default" 2>&1 | tail -5 | tee gpurun_out/tiny_tests.txt
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  for c in cfg5 stream; do
    timeout 300 python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-subconfigs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $f $c', round(d['ms_per_step'],5), d['value'])"
  done
done 2>&1 | tee gpurun_out/tiny_ab.txt
