mkdir -p gpurun_out
SKIP=4 COUNT=1 bash tools/gpu_prof_variant.sh base cfg5 tiny_step r02_cfg5_tiny
KR="python tools/kernel_roofline.py move predator"
$KR > gpurun_out/r02_kernel_roofline.jsonl 2> gpurun_out/kr.err; echo kr rc=$?
ncu --set full --clock-control none --import-source on -k "regex:predator_group" -s 5 -c 1 -f -o gpurun_out/pp128 $KR > gpurun_out/ncu_pp.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/pp128.ncu-rep > gpurun_out/r02_predator_n128_full.txt
ncu -i gpurun_out/pp128.ncu-rep --page raw --csv > gpurun_out/r02_predator_n128_raw.csv; rm -f gpurun_out/pp128.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:predator_group" -s 25 -c 1 -f -o gpurun_out/pp8 $KR > gpurun_out/ncu_pp8.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/pp8.ncu-rep > gpurun_out/r02_predator_n8_full.txt
ncu -i gpurun_out/pp8.ncu-rep --page raw --csv > gpurun_out/r02_predator_n8_raw.csv; rm -f gpurun_out/pp8.ncu-rep
