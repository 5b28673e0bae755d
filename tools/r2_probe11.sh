#!/bin/bash
# e2e timeline (host + device marks), front-end stream A/B, then the parity tests that touch the staging loop and a c2 bench line
mkdir -p gpurun_out
python tools/e2e_timeline.py c2 > gpurun_out/r2_p11_timeline.log 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_p11_pytest.log 2>&1
tail -3 gpurun_out/r2_p11_pytest.log
python bench.py --config c2 --no-baselines > gpurun_out/r2_p11_bench_c2.json 2> gpurun_out/r2_p11_bench_c2.err
VPB_FE_STREAM=copy python bench.py --config c2 --no-baselines > gpurun_out/r2_p11_bench_c2_fecopy.json 2>> gpurun_out/r2_p11_bench_c2.err
cat gpurun_out/r2_p11_timeline.log
python - <<'PY'
import json
for f in ['r2_p11_bench_c2','r2_p11_bench_c2_fecopy']:
    l=[x for x in open(f'gpurun_out/{f}.json') if x.startswith('{')][-1]; d=json.loads(l)
    print(f, round(d['value']), d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))
PY
