#!/bin/bash
# A/B build of the engine library with extra nvcc flags: tools/build_alt.sh <name> "-DDR_MIN_BLOCKS=3 ..."
# -> sugarscape_utilicalc_b200/csrc/alt_<name>.so (git-ignored, travels with gpurun); use with tools/dr_time.py ... lib=<name>
set -e
cd "$(dirname "$0")/../sugarscape_utilicalc_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -Xptxas -v \
  $2 -o alt_$1.so ss_engine.cu ss_serial.cu ss_lattice.cu ss_rounds.cu > alt_$1.log 2>&1
grep -A3 "ss_dr_rounds_kernel\|ss_dr_build_kernel" alt_$1.log | grep "registers\|spill" | head -12
