#!/bin/bash
# 8 ranks staging at once: where the end-to-end time goes, then the c2 bench line with the quota-aware thread count
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/e2e_multi.py c2 > gpurun_out/r2_p13_multi8.log 2>&1
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config c2 --no-baselines > gpurun_out/r2_p13_bench_c2_n8.json 2> gpurun_out/r2_p13_bench_c2_n8.err
grep -v "^\[W\|Warning\|warn" gpurun_out/r2_p13_multi8.log | tail -40
grep '^{' gpurun_out/r2_p13_bench_c2_n8.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['e2e'])"
