mkdir -p gpurun_out
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  for c in ${CFGS:-cfg3}; do
    python bench.py --config $c --steps 200 --warmup 10 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $f $c', d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done
