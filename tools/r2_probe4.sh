#!/bin/bash
# round-2 call 4: validate fp64 front-end / PDL / graphs / fast epilogue / cosine / wide stem, c5 precision study, A/B timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 -s > gpurun_out/r2_p4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p4_pytest.log
# c5 precision: which accumulate-chunk setting gives margin at T = 998
for setting in "VPB_TC_F16=1" "VPB_TC_F16=0" "VPB_TC_CHUNK_K=768" "VPB_TC_CHUNK_K=512 VPB_TC_KC=256" "VPB_TC_CHUNK_K=1024 VPB_TC_KC=256"; do
  echo "== $setting" >> gpurun_out/r2_p4_c5prec.log
  env $setting timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -p no:cacheprovider -s -k "full_batch and c5" 2>&1 | grep -E "c5:|passed|failed|rel-L2|Error" >> gpurun_out/r2_p4_c5prec.log
done
# A/B timings (resident, ECAPA)
for setting in "X=1" "VPB_PDL=0" "VPB_GRAPH=0" "VPB_PDL=0 VPB_GRAPH=0" "VPB_TC_NARROW_KB=96" "VPB_TC_NARROW_KB=128" "VPB_TC_F16=0" "VPB_C1_WIDE=0"; do
  echo "== $setting" >> gpurun_out/r2_p4_ab.log
  env $setting timeout 300 python bench.py --light --steps 30 2>&1 | tail -1 >> gpurun_out/r2_p4_ab.log
done
for setting in "X=1" "VPB_TC_NARROW_KB=96" "VPB_TC_NARROW_KB=128" "VPB_PDL=0 VPB_GRAPH=0" "VPB_C1_WIDE=0"; do
  echo "== $setting" >> gpurun_out/r2_p4_models.log
  env $setting timeout 900 python tools/model_times.py 2>&1 | grep -E "^c[2345]|^tdnn|^eres" >> gpurun_out/r2_p4_models.log
done
timeout 900 python tools/model_times.py --dump gpurun_out/r2_p4_ops_ > /dev/null 2>&1
timeout 900 python bench.py --config c2 --dump-ops gpurun_out/r2_p4_ops_bench_c2.json > gpurun_out/r2_p4_bench_c2.json 2> gpurun_out/r2_p4_bench_c2.err
tail -n 6 gpurun_out/r2_p4_pytest.log; cat gpurun_out/r2_p4_c5prec.log gpurun_out/r2_p4_ab.log gpurun_out/r2_p4_models.log; head -c 1500 gpurun_out/r2_p4_bench_c2.json
