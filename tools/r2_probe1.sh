#!/bin/bash
# round-2 probe 1: run the staged (never executed) variants on hardware and time them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/r2_gpu.txt
VPB_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
  -k "experimental or c_host" > gpurun_out/r2_exp_tests.log 2>&1
echo "exp tests rc=$?" >> gpurun_out/r2_exp_tests.log
timeout 300 python bench.py --light --steps 30 > gpurun_out/r2_light_base.log 2>&1
VPB_TC_F16=1 timeout 300 python bench.py --light --steps 30 > gpurun_out/r2_light_f16.log 2>&1
VPB_TC_RING=1 timeout 300 python bench.py --light --steps 30 > gpurun_out/r2_light_ring.log 2>&1
VPB_POOL_V2=1 timeout 300 python bench.py --light --steps 30 > gpurun_out/r2_light_poolv2.log 2>&1
VPB_TC_F16=1 timeout 600 python bench.py --steps 20 --dump-ops gpurun_out/r2_ops_f16.json > gpurun_out/r2_bench_f16.log 2>&1
VPB_TC_RING=1 timeout 600 python bench.py --steps 20 --dump-ops gpurun_out/r2_ops_ring.json > gpurun_out/r2_bench_ring.log 2>&1
tail -n 3 gpurun_out/r2_exp_tests.log gpurun_out/r2_light_*.log
