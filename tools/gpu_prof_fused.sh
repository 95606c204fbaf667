mkdir -p gpurun_out
B="python bench.py --config ${1:-cfg3} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:fused_step|tiny_step" -s 3 -c 1 -f -o gpurun_out/prof_${1:-cfg3}_fused $B > gpurun_out/ncu_full.log 2>&1; echo rc=$?
