#!/bin/bash
# Rebuilds libb2a.so only (the full build, including the oracle and the drop-in programs, is __graft_entry__.build()).
# usage: tools/build_lib.sh [extra nvcc flags...]
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
     -Iinclude -Igmx2020postanalysis_b200/csrc -o gmx2020postanalysis_b200/libb2a.so gmx2020postanalysis_b200/csrc/b2a.cu
