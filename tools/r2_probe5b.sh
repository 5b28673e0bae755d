#!/bin/bash
# round-2 call 5b: full suite + all four bench configs + smoke + launch list + ncu on the fixed tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 -s > gpurun_out/r2_p5b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p5b_pytest.log
timeout 600 python tools/e2e_timeline.py c2 > gpurun_out/r2_p5b_e2e_timeline.log 2>&1
timeout 900 python tools/model_times.py --dump gpurun_out/r2_p5b_ops_ > gpurun_out/r2_p5b_models.log 2>&1
for c in c2 c3 c4 c5; do
  timeout 1200 python bench.py --config $c --dump-ops gpurun_out/r2_p5b_benchops_$c.json > gpurun_out/r2_p5b_bench_$c.json 2> gpurun_out/r2_p5b_bench_$c.err
  echo "bench $c rc=$?" >> gpurun_out/r2_p5b_pytest.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_p5b_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 240 --csv --log-file gpurun_out/r2_p5b_launches.csv python bench.py --light --steps 4 --warmup 2 > gpurun_out/r2_p5b_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'conv_tc_kernel<\(int\)0, \(bool\)1' -s 9 -c 9 -o gpurun_out/r2_p5b_prof_f16 python tools/prof_run.py 2 > gpurun_out/r2_p5b_ncu_f16.log 2>&1
tail -n 8 gpurun_out/r2_p5b_pytest.log; cat gpurun_out/r2_p5b_e2e_timeline.log gpurun_out/r2_p5b_models.log gpurun_out/r2_p5b_smoke.log | grep -v INFO; for c in c2 c3 c4 c5; do head -c 300 gpurun_out/r2_p5b_bench_$c.json; echo; done
