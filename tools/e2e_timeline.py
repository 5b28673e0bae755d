"""Host-side timeline of MVectorPredictor.predict_batch (ECAPA c2: 256 x 3 s host arrays) + a sweep of the staging knobs.
Dev tool: prints where the end-to-end time goes; not a bench value."""
import json, os, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from loguru import logger
logger.remove()
import bench
from mvector.predict import MVectorPredictor

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'c2']
td = tempfile.mkdtemp()
bench.save_weights(cfg, td)
pred = MVectorPredictor(configs=bench.yml_config(cfg), model_path=td, use_gpu=True)
B = cfg['per_gpu']
pools = [bench.synth_waves(bench.batch_lens(cfg, B, 10 + i), 20 + i) for i in range(4)]


def run(n, trace=False):
    ts = []
    for i in range(n):
        torch.cuda.synchronize()
        pred._trace = [] if trace else None
        t0 = time.perf_counter()
        pred.predict_batch(pools[i % 4])
        t1 = time.perf_counter()
        ts.append((t1 - t0) * 1e3)
        if trace and i == n - 1:
            prev = t0
            for label, t in pred._trace:
                print(f'    +{(t - prev) * 1e3:7.3f} ms  {label}')
                prev = t
            print(f'    +{(t1 - prev) * 1e3:7.3f} ms  return')
    pred._trace = None
    return ts


run(5)
print('default knobs: per-call ms', [round(x, 2) for x in run(6, trace=True)])
for rows, sl, thr, mb in [(64, 16, 8, 256), (32, 16, 8, 256), (128, 16, 8, 256), (64, 8, 8, 256), (64, 32, 8, 256), (64, 16, 4, 256),
                          (64, 16, 16, 256), (64, 16, 24, 256), (64, 16, 8, 128), (32, 16, 16, 128), (64, 16, 16, 64)]:
    MVectorPredictor.STAGE_ROWS, MVectorPredictor.COPY_SLICE, MVectorPredictor.MAX_BATCH = rows, sl, mb
    os.environ['VPB_GATHER_THREADS'] = str(thr)
    run(3)
    ts = run(8)
    print(f'STAGE_ROWS={rows:3d} COPY_SLICE={sl:2d} threads={thr:2d} MAX_BATCH={mb:3d}: median {np.median(ts):6.2f} ms  min {min(ts):6.2f}')
