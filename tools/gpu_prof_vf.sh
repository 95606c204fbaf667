# ncu capture of the cfg3 step kernel with vf_size = 5 (sparse visual-field mode); summaries only come back
mkdir -p gpurun_out
B="python bench.py --config cfg3 --vf 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --graph off"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:fused_step" -s 3 -c 1 -f -o gpurun_out/prof_vf $B > gpurun_out/ncu_full.log 2>&1; echo rc=$?
python tools/ncu_summary.py full gpurun_out/prof_vf.ncu-rep > gpurun_out/r01_cfg3_vf5_full.txt
python tools/ncu_regions.py gpurun_out/prof_vf.ncu-rep > gpurun_out/r01_cfg3_vf5_regions.txt
rm -f gpurun_out/prof_vf.ncu-rep
