# Round-2 artefacts: the default bench line (headline + every named config as a sub-record), per-config lines, the
# reference arm, standalone kernel rooflines, ncu launch lists and one ncu --set full capture per dominant kernel
# (summarised on the box; reports deleted: gpurun_out/ is capped at 64 MiB).  tools/collect_profiles_r02.sh copies
# the summaries into profiles/.
mkdir -p gpurun_out
R=r02
python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo default rc=$?
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference_cfg3.json 2>/dev/null; echo ref rc=$?
python bench.py --steps 100 --warmup 10 --vf 5 --no-cpu-baseline --no-subconfigs > gpurun_out/${R}_bench_cfg3_vf5.json 2>/dev/null; echo vf5 rc=$?
python bench.py --steps 200 --warmup 10 --flags 8 --no-cpu-baseline --no-subconfigs --no-e2e > gpurun_out/${R}_bench_cfg3_two_launch.json 2>/dev/null; echo two-launch rc=$?
python bench.py --steps 200 --warmup 10 --flags 48 --no-cpu-baseline --no-subconfigs --no-e2e > gpurun_out/${R}_bench_cfg3_one_launch_ring.json 2>/dev/null; echo one-launch-ring rc=$?
python bench.py --steps 200 --warmup 10 --flags 32 --no-cpu-baseline --no-subconfigs --no-e2e > gpurun_out/${R}_bench_cfg3_ring.json 2>/dev/null; echo ring rc=$?
for c in cfg2 cfg4 cfg5 stream n256; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err; echo $c rc=$?; done
python tools/kernel_roofline.py > gpurun_out/${R}_kernel_roofline.jsonl 2> gpurun_out/kernel_roofline.err; echo kr rc=$?
for c in cfg3 cfg4 cfg5; do
  B="python bench.py --config $c --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off"
  $B > gpurun_out/plain_$c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${R}_${c}_launches.csv $B > gpurun_out/ncu_l_$c.log 2>&1; echo launches $c rc=$?
done
full() {  # <tag> <kernel regex> <skip> <count> <command...>
  tag=$1; rx=$2; sk=$3; cn=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $sk -c $cn -f -o gpurun_out/${R}_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1; echo full $tag rc=$?
  if [ -s gpurun_out/${R}_$tag.ncu-rep ]; then
    python tools/ncu_summary.py full gpurun_out/${R}_$tag.ncu-rep > gpurun_out/${R}_${tag}_full.txt
    ncu -i gpurun_out/${R}_$tag.ncu-rep --page raw --csv > gpurun_out/${R}_${tag}_raw.csv
  fi
}
B3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off"
full cfg3_fused fused_step 6 1 $B3
[ -s gpurun_out/${R}_cfg3_fused.ncu-rep ] && python tools/ncu_regions.py gpurun_out/${R}_cfg3_fused.ncu-rep > gpurun_out/${R}_cfg3_fused_regions.txt
full cfg3_vf5 fused_step 6 2 $B3 --vf 5
full cfg5_tiny "tiny8_packed|tiny_step" 3 1 python bench.py --config cfg5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off
full cfg4_staged "collide_bitmap|boxcount|predator_block" 6 3 python bench.py --config cfg4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off
KR="python tools/kernel_roofline.py move predator observe"
full move move_kernel 5 1 $KR
full predator_n128 predator_async_kernel 5 1 $KR
full predator_n16384 predator_block_async 5 1 $KR
full predator_n8 predator_async_kernel 18 1 $KR
full observe_vf observe_vf_warp 4 1 $KR
rm -f gpurun_out/${R}_*.ncu-rep
du -sh gpurun_out
