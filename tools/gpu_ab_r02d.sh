mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "predator or staged or large or cfg4" 2>&1 | tail -3
python tools/kernel_roofline.py predator 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('PRED', d['kernel'][:40], d.get('shape',''), round(d.get('ms_mean',0),4), round(d.get('frac',0),3))"
