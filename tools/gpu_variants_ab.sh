# headline timing (bench.py --config $CFG, default cfg3) per library variant built by tools/build_variant.sh
mkdir -p gpurun_out
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  timeout 300 python bench.py --config ${CFG:-cfg3} --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-subconfigs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $f', round(d['ms_per_step'],5), d['value'])"
done 2>&1 | tee gpurun_out/variants_ab.txt
