# pytest -m gpu, then the default bench line (headline + sub-records), the reference arm, and the kernel rooflines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_gputest.txt; tail -3 gpurun_out/r02_gputest.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo bench rc=$?
tail -5 gpurun_out/r02_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo ref rc=$?
python tools/kernel_roofline.py predator > gpurun_out/r02_kernel_roofline_pred.jsonl 2> gpurun_out/kr.err; echo kr rc=$?
