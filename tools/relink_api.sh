#!/bin/bash
# Recompile only gmpe_api.cu (host side + tile kernel) and relink libgmpe_b200.so with the lane-per-agent
# objects of the last full build (contracts-marl-aam-corridors_b200/csrc/build.py): 25 s instead of 2.5 min
# while iterating on the tile kernel.  Extra arguments go to nvcc (e.g. -DGMPE_EDGE_LINEAR=0).
set -e
cd "$(dirname "$0")/../contracts-marl-aam-corridors_b200/csrc"
OBJ=$(ls -td ../../build/obj_*/ | while read d; do [ -f "$d/gmpe_coop_wide_f64.o" ] && echo "$d" && break; done)
OUT=${OUT:-libgmpe_b200.so}
nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -O3 "$@" -c gmpe_api.cu -o "$OBJ/gmpe_api.o"
nvcc -ccbin /usr/bin/g++ -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/gmpe_api.o "$OBJ"/gmpe_coop_f32.o "$OBJ"/gmpe_coop_f64.o "$OBJ"/gmpe_coop_wide_f32.o "$OBJ"/gmpe_coop_wide_f64.o
echo "$OUT relinked from $OBJ"
