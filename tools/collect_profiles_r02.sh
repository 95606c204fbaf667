# Copy what tools/gpu_final_profiles_r02.sh brought back in gpurun_out/ into tracked files under profiles/ and refresh
# profiles/ncu_traffic.json / ncu_pipe.json (the entries bench.py reads) from the raw ncu csv files.
R=r02
for f in gpurun_out/${R}_bench_*.json gpurun_out/${R}_kernel_roofline.jsonl; do [ -s "$f" ] && cp "$f" profiles/; done
for c in cfg3 cfg4 cfg5; do [ -s gpurun_out/${R}_${c}_launches.csv ] && cp gpurun_out/${R}_${c}_launches.csv profiles/ && python tools/ncu_summary.py launches gpurun_out/${R}_${c}_launches.csv > profiles/${R}_${c}_launches.txt; done
for n in cfg3_fused cfg3_vf5 cfg5_tiny cfg4_staged move predator_n128 predator_n16384 predator_n8 observe_vf; do [ -s gpurun_out/${R}_${n}_full.txt ] && cp gpurun_out/${R}_${n}_full.txt profiles/; done
[ -s gpurun_out/${R}_cfg3_fused_regions.txt ] && cp gpurun_out/${R}_cfg3_fused_regions.txt profiles/
python tools/refresh_profile_json.py
ls profiles | wc -l
