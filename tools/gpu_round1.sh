mkdir -p gpurun_out
set -x
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo rc=$?
tail -c 3000 gpurun_out/bench_cfg3.json; tail -5 gpurun_out/bench_cfg3.err
for c in cfg2 cfg5 cfg4 stream; do timeout 300 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo $c rc=$?; tail -c 1500 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err; done
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 2 -f -o gpurun_out/prof_cfg3_fused python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo ncu2 rc=$?
tail -5 gpurun_out/ncu_full.log
