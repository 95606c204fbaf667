mkdir -p gpurun_out
for f in "$@"; do
  if [ "$f" = "base" ]; then unset SWARM_B200_LIB; else export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$f.so; fi
  python tools/kernel_roofline.py predator 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('PRED $f', d['kernel'][:40], d.get('shape',''), round(d.get('ms_mean',0),4), round(d.get('frac',0),3))"
done
