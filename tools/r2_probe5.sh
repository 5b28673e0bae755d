#!/bin/bash
# round-2 call 5: full suite on the current tree, e2e timeline, 4- vs 8-warp epilogue A/B, all four bench configs, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
E8=$PWD/voiceprintrecognition-pytorch_b200/libvpb200_e8.so
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 -s > gpurun_out/r2_p5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p5_pytest.log
VPB_LIB=$E8 timeout 900 python -m pytest tests/test_gpu_conv_engines.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x \
  -k "conv_tc or small_models or full_size or c2_full" > gpurun_out/r2_p5_pytest_e8.log 2>&1
echo "pytest e8 rc=$?" >> gpurun_out/r2_p5_pytest_e8.log
timeout 600 python tools/e2e_timeline.py c2 > gpurun_out/r2_p5_e2e_timeline.log 2>&1
for setting in "X=1" "VPB_LIB=$E8"; do
  echo "== $setting" >> gpurun_out/r2_p5_models.log
  env $setting timeout 900 python tools/model_times.py 2>&1 | grep -E "^c[2345]|^tdnn|^eres" >> gpurun_out/r2_p5_models.log
done
timeout 900 python tools/model_times.py --dump gpurun_out/r2_p5_ops_ > /dev/null 2>&1
VPB_LIB=$E8 timeout 900 python tools/model_times.py --dump gpurun_out/r2_p5_ops_e8_ > /dev/null 2>&1
for c in c2 c3 c4 c5; do
  timeout 1200 python bench.py --config $c --dump-ops gpurun_out/r2_p5_benchops_$c.json > gpurun_out/r2_p5_bench_$c.json 2> gpurun_out/r2_p5_bench_$c.err
  echo "bench $c rc=$?" >> gpurun_out/r2_p5_pytest.log
done
VPB_LIB=$E8 timeout 600 python bench.py --config c2 --no-baselines > gpurun_out/r2_p5_bench_c2_e8.json 2> gpurun_out/r2_p5_bench_c2_e8.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_p5_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 240 --csv --log-file gpurun_out/r2_p5_launches.csv python bench.py --light --steps 4 --warmup 2 > gpurun_out/r2_p5_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)0, \(bool\)1' -s 9 -c 9 -o gpurun_out/r2_p5_prof_f16 python tools/prof_run.py 2 > gpurun_out/r2_p5_ncu_f16.log 2>&1
tail -n 5 gpurun_out/r2_p5_pytest.log gpurun_out/r2_p5_pytest_e8.log; cat gpurun_out/r2_p5_e2e_timeline.log gpurun_out/r2_p5_models.log gpurun_out/r2_p5_smoke.log; for c in c2 c3 c4 c5; do head -c 300 gpurun_out/r2_p5_bench_$c.json; echo; done
