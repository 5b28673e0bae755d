#!/bin/bash
# round-2 call 6 (gpurun --gpus 2): NCCL tests of the product sharding path + 2-GPU bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_p6_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -q -p no:cacheprovider -k "sharded" -s > gpurun_out/r2_p6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_p6_pytest.log
port=29611
for c in c2 c3 c5; do
  steps=30; [ $c = c5 ] && steps=8
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
    bench.py --config $c --gpus 2 --steps $steps --warmup 3 > gpurun_out/r2_p6_bench_${c}_n2.json 2> gpurun_out/r2_p6_bench_${c}_n2.err
  echo "bench $c n2 rc=$?" >> gpurun_out/r2_p6_pytest.log
  port=$((port+1))
done
# single-GPU A/B on the same box: residual L2 prefetch (default) and the experimental A-row prefetch (VPB_TC_DEBUG=8)
for setting in "X=1" "VPB_TC_DEBUG=8"; do
  echo "== $setting" >> gpurun_out/r2_p6_models.log
  env $setting timeout 900 python tools/model_times.py 2>&1 | grep -E "^c[2345]|^tdnn|^eres" >> gpurun_out/r2_p6_models.log
  env $setting timeout 300 python bench.py --light --steps 30 2>&1 | tail -1 >> gpurun_out/r2_p6_models.log
done
timeout 600 python tools/e2e_timeline.py c2 2>&1 | head -14 > gpurun_out/r2_p6_e2e_timeline.log
cat gpurun_out/r2_p6_models.log gpurun_out/r2_p6_e2e_timeline.log
tail -n 4 gpurun_out/r2_p6_pytest.log; for c in c2 c3 c5; do head -c 400 gpurun_out/r2_p6_bench_${c}_n2.json; echo; tail -n 2 gpurun_out/r2_p6_bench_${c}_n2.err; done
