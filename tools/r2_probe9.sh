#!/bin/bash
# round-2 call 9: ncu of the residual 1x1 convs of the 55 M ERes2Net (N=256 K=72 with residual vs N=256 K=64 without)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)0, \(bool\)1' -s 0 -c 4 -o gpurun_out/r2_p9_prof_c5 python tools/model_times.py --only c5 > gpurun_out/r2_p9_ncu.log 2>&1
tail -5 gpurun_out/r2_p9_ncu.log
timeout 300 python tools/e2e_timeline.py c2 2>&1 | head -12
