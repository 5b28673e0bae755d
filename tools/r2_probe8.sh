#!/bin/bash
# round-2 call 8: final single-GPU validation: full suite, all four bench configs, smoke, per-op dumps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 -s > gpurun_out/r2_p8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p8_pytest.log
for c in c2 c3 c4 c5; do
  timeout 1200 python bench.py --config $c --dump-ops gpurun_out/r2_p8_benchops_$c.json > gpurun_out/r2_p8_bench_$c.json 2> gpurun_out/r2_p8_bench_$c.err
  echo "bench $c rc=$?" >> gpurun_out/r2_p8_pytest.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_p8_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_p8_ref_c2.json 2> gpurun_out/r2_p8_ref_c2.err
timeout 300 python tools/e2e_timeline.py c2 2>&1 | head -14 > gpurun_out/r2_p8_e2e_timeline.log
tail -n 8 gpurun_out/r2_p8_pytest.log; grep -v INFO gpurun_out/r2_p8_smoke.log; cat gpurun_out/r2_p8_e2e_timeline.log; for c in c2 c3 c4 c5; do head -c 300 gpurun_out/r2_p8_bench_$c.json; echo; done
