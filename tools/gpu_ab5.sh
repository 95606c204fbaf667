for v in 0 1 0 1; do
  if [ $v = 1 ]; then export SWARM_NO_FAR_BORDER=1; else unset SWARM_NO_FAR_BORDER; fi
  python bench.py --config cfg3 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB no_far_border=$v', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
