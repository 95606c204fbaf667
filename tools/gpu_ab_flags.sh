# usage: gpu_ab_flags.sh "<cfg> <flags> [extra bench args]" ...   (A/B of kernel switches on the default library)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_host_features.py -x -q 2>&1 | tail -5 > gpurun_out/r02_ab_tests.txt; tail -3 gpurun_out/r02_ab_tests.txt
for spec in "$@"; do
  set -- $spec; c=$1; f=$2; shift 2
  python bench.py --config $c --flags $f --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-subconfigs "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $c flags=$f $*', d['run']['kernel_path'], round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5))"
done 2>&1 | tee gpurun_out/r02_ab_flags.txt
