mkdir -p gpurun_out
export SWARM_B200_LIB=$PWD/gym-swarm_b200/libswarm_b200_jitter.so
( echo "# race canary: library built with -DSWARM_RACE_JITTER (lane-dependent sleeps after every lane-group barrier)";
  timeout 1500 python tools/race_canary.py 1000;
  echo "# the GPU parity suite through the same library:";
  timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 ) > gpurun_out/r02_race_canary.txt 2>&1
tail -8 gpurun_out/r02_race_canary.txt
unset SWARM_B200_LIB
( echo "# control: the production library on the same canary"; timeout 900 python tools/race_canary.py 200 ) >> gpurun_out/r02_race_canary.txt 2>&1
tail -3 gpurun_out/r02_race_canary.txt
