# ncu --set full of one standalone kernel of tools/kernel_roofline.py: <which> <tag> <kernel regex> <skip>
mkdir -p gpurun_out
KR="python tools/kernel_roofline.py $1"
ncu --set full --clock-control none --import-source on -k "regex:$3" -s $4 -c 1 -f -o gpurun_out/prof_$2 $KR > gpurun_out/ncu_$2.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/prof_$2.ncu-rep > gpurun_out/$2_full.txt
ncu -i gpurun_out/prof_$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv
rm -f gpurun_out/prof_$2.ncu-rep
