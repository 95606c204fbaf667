# usage: gpu_prof_variant.sh <variant|base> <config> <kernel regex> <tag>: one ncu --set full capture, summarised (+ raw csv)
mkdir -p gpurun_out
if [ "$1" != "base" ]; then export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_$1.so; fi
B="python bench.py --config $2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$3" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_$4 $B > gpurun_out/ncu_$4.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/prof_$4.ncu-rep > gpurun_out/$4_full.txt
python tools/ncu_regions.py gpurun_out/prof_$4.ncu-rep > gpurun_out/$4_regions.txt 2>/dev/null
ncu -i gpurun_out/prof_$4.ncu-rep --page raw --csv > gpurun_out/$4_raw.csv
ncu -i gpurun_out/prof_$4.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$4_source.csv 2>/dev/null
gzip -f gpurun_out/$4_source.csv
rm -f gpurun_out/prof_$4.ncu-rep
