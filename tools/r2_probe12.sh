#!/bin/bash
mkdir -p gpurun_out
VPB_TIMELINE_SWEEP=0 python tools/e2e_timeline.py c2 > gpurun_out/r2_p12_timeline.log 2>&1
python tools/e2e_multi.py c2 > gpurun_out/r2_p12_multi1.log 2>&1
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r2_p12_pytest.log 2>&1
tail -2 gpurun_out/r2_p12_pytest.log
for c in c2 c3; do python bench.py --config $c --no-baselines > gpurun_out/r2_p12_bench_$c.json 2> gpurun_out/r2_p12_bench_$c.err; done
cat gpurun_out/r2_p12_timeline.log; cat gpurun_out/r2_p12_multi1.log | tail -30
python - <<'PY'
import json
for f in ['r2_p12_bench_c2','r2_p12_bench_c3']:
    l=[x for x in open(f'gpurun_out/{f}.json') if x.startswith('{')][-1]; d=json.loads(l)
    print(f, round(d['value']), d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))
PY
