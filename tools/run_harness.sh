#!/bin/bash
# build + run one CPU harness (tests/cpu_harness/<name>.cpp) outside pytest: tools/run_harness.sh box5_host_check
set -e
cd "$(dirname "$0")/.."
B=tests/cpu_harness/_build
mkdir -p $B
for f in hex8_me elem_kernels linsys; do
  [ $B/$f.o -nt oracle/$f.c ] && [ $B/$f.o -nt oracle/nalu_oracle.h ] || gcc -O2 -std=c99 -ffp-contract=off -c oracle/$f.c -o $B/$f.o
done
[ $B/graph.o -nt nalu_b200/csrc/graph.cpp ] && [ $B/graph.o -nt nalu_b200/csrc/graph.h ] || g++ -O2 -std=c++17 -fopenmp -c nalu_b200/csrc/graph.cpp -o $B/graph.o
g++ -O1 -std=c++17 -fopenmp tests/cpu_harness/$1.cpp $B/hex8_me.o $B/elem_kernels.o $B/linsys.o $B/graph.o -o $B/$1 -lm
$B/$1
