mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/gpu_tests.log
python tools/kernel_roofline.py predator > gpurun_out/kernel_roofline_pred.jsonl 2>&1; cat gpurun_out/kernel_roofline_pred.jsonl
