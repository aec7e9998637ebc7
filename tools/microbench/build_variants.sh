#!/bin/bash
# builds poseidon_bench variants into ./bin/<name>; usage: build_variants.sh name "flags" ...
set -e
mkdir -p bin
CS=../../plonky2_goldibear_b200/csrc
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -I$CS $flags poseidon_bench.cu $CS/poseidon_tables.cpp -o bin/$name &
done
wait
