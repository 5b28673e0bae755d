#!/bin/bash
# round-2 call 3: full GPU test suite (incl. BASELINE-size + sharded tests), all four bench configs, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/r2_p3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p3_pytest.log
for c in c2 c3 c4 c5; do
  timeout 1200 python bench.py --config $c --dump-ops gpurun_out/r2_p3_ops_$c.json > gpurun_out/r2_p3_bench_$c.json 2> gpurun_out/r2_p3_bench_$c.err
  echo "bench $c rc=$?" >> gpurun_out/r2_p3_pytest.log
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_p3_ref_c2.json 2> gpurun_out/r2_p3_ref_c2.err
VPB_TC_F16=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)0, \(bool\)1' -s 9 -c 9 -o gpurun_out/r2_prof_f16 python tools/prof_run.py 2 > gpurun_out/r2_ncu_f16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)1' -s 18 -c 2 -o gpurun_out/r2_prof_res2 python tools/prof_run.py 2 > gpurun_out/r2_ncu_res2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)0, \(bool\)0' -s 10 -c 10 -o gpurun_out/r2_prof_tf32 python tools/prof_run.py 2 > gpurun_out/r2_ncu_tf32.log 2>&1
tail -n 5 gpurun_out/r2_p3_pytest.log; tail -c 600 gpurun_out/r2_p3_bench_c2.err; for c in c2 c3 c4 c5; do head -c 400 gpurun_out/r2_p3_bench_$c.json; echo; done
