# A/B of differently built libraries (tools/build_variant.sh): parity tests on the default build first, then timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q 2>&1 | tail -5 > gpurun_out/r02_ab_tests.txt; tail -3 gpurun_out/r02_ab_tests.txt
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  for c in ${CFGS:-cfg3}; do
    python bench.py --config $c --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-subconfigs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $f $c', round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5))"
  done
done 2>&1 | tee gpurun_out/r02_ab.txt
