#!/bin/bash
# round-2 call 10: validation of the Hardtanh fast path + staging limits: full suite, c5 / c2 bench lines, model times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 -s > gpurun_out/r2_p10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p10_pytest.log
timeout 900 python tools/model_times.py --dump gpurun_out/r2_p10_ops_ > gpurun_out/r2_p10_models.log 2>&1
for c in c5 c2; do
  timeout 1200 python bench.py --config $c --dump-ops gpurun_out/r2_p10_benchops_$c.json > gpurun_out/r2_p10_bench_$c.json 2> gpurun_out/r2_p10_bench_$c.err
  echo "bench $c rc=$?" >> gpurun_out/r2_p10_pytest.log
done
timeout 300 python tools/e2e_timeline.py c2 2>&1 | head -10 > gpurun_out/r2_p10_e2e_timeline.log
tail -n 6 gpurun_out/r2_p10_pytest.log; grep -E "^c[2345]|^tdnn|^eres" gpurun_out/r2_p10_models.log; cat gpurun_out/r2_p10_e2e_timeline.log; for c in c5 c2; do grep '^{' gpurun_out/r2_p10_bench_$c.json | head -c 300; echo; done
