# Round-2 evidence that needs no code change: compute-sanitizer (memcheck / racecheck) over every kernel path and
# ncu --set full captures of the two tiled pairwise kernels (the INT32-roofline contract kernels).
mkdir -p gpurun_out
which compute-sanitizer; compute-sanitizer --version | head -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo memcheck rc=$?
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo racecheck rc=$?
tail -5 gpurun_out/r02_sanitizer_memcheck.txt gpurun_out/r02_sanitizer_racecheck.txt
KR="python tools/kernel_roofline.py pairwise"
$KR > gpurun_out/r02_pairwise_roofline.jsonl 2> gpurun_out/kr.err; echo kr rc=$?
ncu --set full --clock-control none --import-source on -k "regex:pairwise_sym_kernel" -s 3 -c 1 -f -o gpurun_out/r02_pairwise_sym $KR > gpurun_out/ncu_ps.log 2>&1; echo ncu sym rc=$?
B="python bench.py --config cfg4 --vf 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --graph off"
$B > gpurun_out/plain_cfg4vf.log 2>&1; echo plain rc=$?
ncu --set full --clock-control none --import-source on -k "regex:pairwise_kernel" -s 3 -c 1 -f -o gpurun_out/r02_pairwise_vf $B > gpurun_out/ncu_pv.log 2>&1; echo ncu vf rc=$?
for n in pairwise_sym pairwise_vf; do
  [ -s gpurun_out/r02_$n.ncu-rep ] && python tools/ncu_summary.py full gpurun_out/r02_$n.ncu-rep > gpurun_out/r02_${n}_full.txt
  [ -s gpurun_out/r02_$n.ncu-rep ] && ncu -i gpurun_out/r02_$n.ncu-rep --page raw --csv > gpurun_out/r02_${n}_raw.csv
  rm -f gpurun_out/r02_$n.ncu-rep
done
du -sh gpurun_out
