#!/bin/bash
# round-2 call 7 (gpurun --gpus 8): C3 = CAM++ 2048 over 8 GPUs, C5 = ERes2Net-55M ragged 256 over 4 GPUs, C2 at 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_p7_gpus.txt 2>&1
run() {  # config n steps port
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $4 \
    bench.py --config $1 --gpus $2 --steps $3 --warmup 3 > gpurun_out/r2_p7_bench_$1_n$2.json 2> gpurun_out/r2_p7_bench_$1_n$2.err
  echo "bench $1 n$2 rc=$?" >> gpurun_out/r2_p7_status.log
}
run c3 8 20 29711
run c5 4 5 29713
run c2 8 30 29712
cat gpurun_out/r2_p7_status.log; for f in gpurun_out/r2_p7_bench_*.json; do grep '^{' $f | head -c 500; echo; done; tail -n 3 gpurun_out/r2_p7_bench_c3_n8.err
