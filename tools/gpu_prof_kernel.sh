# usage: gpu_prof_kernel.sh <config> <kernel regex> [skip]
mkdir -p gpurun_out
B="python bench.py --config $1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --graph off"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${3:-3} -c 1 -f -o gpurun_out/prof_k $B > gpurun_out/ncu_full.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/prof_k.ncu-rep > gpurun_out/prof_k_full.txt; rm -f gpurun_out/prof_k.ncu-rep
