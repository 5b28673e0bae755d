#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x \
  -k "small_models or predictor_dropin or short_and_long or c2_full or speaker_diar" > gpurun_out/r2_p2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p2_tests.log
VPB_TC_F16=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<0, true' -s 9 -c 9 -o gpurun_out/r2_prof_f16 python tools/prof_run.py 2 > gpurun_out/r2_ncu_f16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<1' -s 18 -c 2 -o gpurun_out/r2_prof_res2 python tools/prof_run.py 2 > gpurun_out/r2_ncu_res2.log 2>&1
timeout 900 python tools/model_times.py --dump gpurun_out/r2_ops_tf32_ > gpurun_out/r2_model_times_tf32.log 2>&1
VPB_TC_F16=1 timeout 900 python tools/model_times.py --dump gpurun_out/r2_ops_f16_ > gpurun_out/r2_model_times_f16.log 2>&1
tail -n 4 gpurun_out/r2_p2_tests.log gpurun_out/r2_ncu_f16.log gpurun_out/r2_model_times_*.log
