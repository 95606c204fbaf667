# Summarise what tools/gpu_final_profiles.sh brought back into tracked files under profiles/.
R=${1:-r01}
for f in gpurun_out/${R}_bench_*.json gpurun_out/${R}_kernel_roofline.jsonl; do [ -s "$f" ] && cp "$f" profiles/; done
for c in cfg3 cfg4 cfg5; do [ -s gpurun_out/${R}_${c}_launches.csv ] && cp gpurun_out/${R}_${c}_launches.csv profiles/ && python tools/ncu_summary.py launches gpurun_out/${R}_${c}_launches.csv > profiles/${R}_${c}_launches.txt; done
for n in cfg3_fused cfg5_tiny cfg4_staged move predator_n128 predator_n16384; do [ -s gpurun_out/${R}_${n}_full.txt ] && cp gpurun_out/${R}_${n}_full.txt profiles/; done
[ -s gpurun_out/${R}_cfg3_fused_regions.txt ] && cp gpurun_out/${R}_cfg3_fused_regions.txt profiles/
ls -la profiles
