mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -k "predator or staged or golden or tiny or properties" 2>&1 | tail -3
for f in base vpl2; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  python tools/kernel_roofline.py predator 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('PRED $f', d['kernel'][:40], d.get('shape',''), round(d.get('ms_mean',0),4), round(d.get('frac',0),3))"
done
for f in base t6 t8; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  for c in cfg5 stream; do
  python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-subconfigs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $f $c', round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5), round(d['roofline']['frac'],3))"
  done
done
