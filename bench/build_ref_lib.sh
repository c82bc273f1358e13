#!/bin/bash
# Builds particlesmc_b200/lib/libpmc_b200_<name>.so from the sources of git revision $1 (A/B runs on one box:
# PMC_B200_LIB=libpmc_b200_<name>.so python bench.py ...).  Usage: bench/build_ref_lib.sh <rev> <name>
set -e
rev=$1; name=$2; tmp=$(mktemp -d)
git archive "$rev" particlesmc_b200/csrc include | tar -x -C "$tmp"
for f in api chains chains_fast chains_spec box; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -w \
       -I "$tmp/include" -I "$tmp/particlesmc_b200/csrc" -c "$tmp/particlesmc_b200/csrc/$f.cu" -o "$tmp/$f.o" &
done
wait
nvcc -shared -ccbin /usr/bin/g++ -o "particlesmc_b200/lib/libpmc_b200_$name.so" "$tmp"/*.o 2>/dev/null
rm -rf "$tmp"; echo "particlesmc_b200/lib/libpmc_b200_$name.so"
