mkdir -p gpurun_out
set -x
python tools/kernel_roofline.py > gpurun_out/kernel_roofline.jsonl 2> gpurun_out/kernel_roofline.err; echo rc=$?; cat gpurun_out/kernel_roofline.jsonl; tail -3 gpurun_out/kernel_roofline.err
B4="python bench.py --config cfg4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B4 > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg4.csv $B4 > gpurun_out/ncu_l4.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:pairwise_kernel -s 3 -c 1 -f -o gpurun_out/prof_cfg4_pairwise $B4 > gpurun_out/ncu_f4.log 2>&1; echo rc=$?
KR="python tools/kernel_roofline.py move predator"
$KR > gpurun_out/plainkr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'move_kernel|predator_kernel' -s 5 -c 3 -f -o gpurun_out/prof_move $KR > gpurun_out/ncu_kr.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:predator_kernel -s 5 -c 1 -f -o gpurun_out/prof_predator $KR > gpurun_out/ncu_kr2.log 2>&1; echo rc=$?
B5="python bench.py --config cfg5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B5 > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_cfg5_fused $B5 > gpurun_out/ncu_f5.log 2>&1; echo rc=$?
