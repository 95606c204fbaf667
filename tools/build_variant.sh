#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags...]   ->  gym-swarm_b200/libswarm_b200_NAME.so
# A differently built copy of the library for kernel A/B runs (select it with SWARM_B200_LIB=...) and for the race
# canary (NAME=jitter, -DSWARM_RACE_JITTER).  Same four translation units as __graft_entry__.build().
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/gym-swarm_b200/csrc
OBJ=$ROOT/build/obj_$NAME
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress=128 $@"
pids=()
for u in swarm_abi fused_launch fused_vf_launch tiny_launch; do
  (cd $CSRC && nvcc $FLAGS -c $u.cu -o $OBJ/$u.o) & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/gym-swarm_b200/libswarm_b200_$NAME.so $OBJ/*.o -ldl
echo built libswarm_b200_$NAME.so
