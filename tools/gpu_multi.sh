mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; echo rc=$?
tail -c 2500 gpurun_out/bench_cfg3_n$N.json; tail -5 gpurun_out/bench_cfg3_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong --no-cpu-baseline > gpurun_out/bench_cfg3_strong_n$N.json 2> gpurun_out/bench_cfg3_strong_n$N.err; echo rc=$?
tail -c 1200 gpurun_out/bench_cfg3_strong_n$N.json; tail -5 gpurun_out/bench_cfg3_strong_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo rc=$?
tail -c 600 gpurun_out/bench_ref_n$N.json; tail -3 gpurun_out/bench_ref_n$N.err
