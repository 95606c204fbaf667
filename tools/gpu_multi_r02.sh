# usage: gpu_multi_r02.sh N : the default bench line (headline + sub-records incl. NCCL stats inside cfg5's timed loop) on N GPUs
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo rc=$?
tail -3 gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2>/dev/null; echo ref rc=$?
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/lscpu_n$N.txt
