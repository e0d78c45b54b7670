#!/bin/sh
# instrumented build of the engine (phase clocks of the move passes -> ss_get_counters): tools/dbg_pass.py, tools/dbg_solo.py
cd "$(dirname "$0")/../sugarscape_utilicalc_b200/csrc" && \
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  -DSS_PASS_TIMING -o libsugarscape_b200_dbg.so ss_engine.cu ss_serial.cu ss_lattice.cu ss_rounds.cu
