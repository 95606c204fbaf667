# Round artefacts: bench lines for every config, launch lists, one ncu --set full capture per dominant kernel,
# standalone kernel rooflines.  Everything lands in gpurun_out/ (copied to profiles/ by tools/collect_profiles.sh).
mkdir -p gpurun_out
R=${1:-r01}
python bench.py --steps 200 --warmup 10 > gpurun_out/${R}_bench_cfg3.json 2> gpurun_out/${R}_bench_cfg3.err; echo cfg3 rc=$?
python bench.py --steps 100 --warmup 10 --vf 5 > gpurun_out/${R}_bench_cfg3_vf5.json 2> gpurun_out/${R}_bench_cfg3_vf5.err; echo cfg3 vf5 rc=$?
for c in cfg2 cfg4 cfg5 stream; do python bench.py --config $c --steps 100 --warmup 10 > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err; echo $c rc=$?; done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_cfg3.json 2>/dev/null; echo ref rc=$?
python tools/kernel_roofline.py > gpurun_out/${R}_kernel_roofline.jsonl 2> gpurun_out/kernel_roofline.err; echo kr rc=$?
for c in cfg3 cfg4 cfg5; do
  B="python bench.py --config $c --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $B > gpurun_out/plain_$c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${R}_${c}_launches.csv $B > gpurun_out/ncu_l_$c.log 2>&1; echo launches $c rc=$?
done
B="python bench.py --config cfg3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:fused_step" -s 3 -c 1 -f -o gpurun_out/${R}_cfg3_fused $B > gpurun_out/ncu_f3.log 2>&1; echo rc=$?
B="python bench.py --config cfg5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:tiny_step" -s 3 -c 1 -f -o gpurun_out/${R}_cfg5_tiny $B > gpurun_out/ncu_f5.log 2>&1; echo rc=$?
B="python bench.py --config cfg4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k "regex:pairwise_sym|collide_bitmap|boxcount" -s 6 -c 2 -f -o gpurun_out/${R}_cfg4_staged $B > gpurun_out/ncu_f4.log 2>&1; echo rc=$?
KR="python tools/kernel_roofline.py move predator"
$KR > gpurun_out/plainkr.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:move_kernel|predator_group|predator_block" -s 5 -c 1 -f -o gpurun_out/${R}_move $KR > gpurun_out/ncu_kr.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k "regex:predator_group" -s 5 -c 1 -f -o gpurun_out/${R}_predator_n128 $KR > gpurun_out/ncu_kr2.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k "regex:predator_block" -s 5 -c 1 -f -o gpurun_out/${R}_predator_n16384 $KR > gpurun_out/ncu_kr3.log 2>&1; echo rc=$?

# summarise on the box (ncu is installed there too) and drop the bulky reports: gpurun_out/ may not exceed 64 MiB
for n in cfg3_fused cfg5_tiny cfg4_staged move predator_n128 predator_n16384; do
  if [ -s gpurun_out/${R}_$n.ncu-rep ]; then python tools/ncu_summary.py full gpurun_out/${R}_$n.ncu-rep > gpurun_out/${R}_${n}_full.txt; fi
done
[ -s gpurun_out/${R}_cfg3_fused.ncu-rep ] && python tools/ncu_regions.py gpurun_out/${R}_cfg3_fused.ncu-rep > gpurun_out/${R}_cfg3_fused_regions.txt
for n in cfg5_tiny cfg4_staged move predator_n128 predator_n16384; do rm -f gpurun_out/${R}_$n.ncu-rep; done
du -sh gpurun_out
