"""Where the end-to-end time of the sharded call goes when N ranks of one host stage their shards AT THE SAME TIME
(torchrun --nproc-per-node N tools/e2e_multi.py [c2]): per-rank pinned->device bandwidth with all ranks copying at once,
the native gather alone, the whole host-staged call, and rank 0's host + device timeline; then a sweep of the gather
thread count.  Dev tool: nothing printed here is a bench value."""
import ctypes as C
import os
import sys
import tempfile
import time

sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
import __graft_entry__ as ge
ge.build()
from loguru import logger
logger.remove()
import bench
from mvector import _lib as L
from mvector import distributed as mdist
from mvector.predict import MVectorPredictor

bound = mdist.bind_rank_to_local_cpus(local, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'c2']
td = tempfile.mkdtemp()
bench.save_weights(cfg, td)
pred = MVectorPredictor(configs=bench.yml_config(cfg), model_path=td, use_gpu=True)
B = cfg['per_gpu']
pools = [bench.synth_waves(bench.batch_lens(cfg, B, 10 + i), 20 + i + 100 * rank) for i in range(4)]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allranks(x):
    """-> list of one float per rank (on every rank)"""
    t = torch.tensor([float(x)], device=dev)
    if world == 1:
        return [float(x)]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def say(msg):
    if rank == 0:
        print(msg, flush=True)


say(f'world {world}; rank 0 bound to {len(bound) if bound else "all"} cpus; cgroup cpus {MVectorPredictor._cgroup_cpus()}; '
    f'gather threads {MVectorPredictor._gather_threads()}')

# 1. pinned -> device, all ranks at once
nbytes = sum(w.nbytes for w in pools[0])
hp = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
dp = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
for _ in range(3):
    dp.copy_(hp, non_blocking=True)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dp.copy_(hp, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = allranks(nbytes * 10 / e0.elapsed_time(e1) / 1e6)
say(f'H2D {nbytes / 1e6:.1f} MB pinned->device, all {world} ranks copying at once: GB/s per rank {[round(g, 1) for g in gbs]}')

# 2. native gather alone (no copies), all ranks at once
lib = L.lib()
w0 = pools[0]
lmax = max(w.shape[0] for w in w0)
ptrs = np.fromiter((w.__array_interface__['data'][0] for w in w0), dtype=np.uint64, count=B)
lens = np.fromiter(map(len, w0), dtype=np.int32, count=B)
stage = hp[:B * lmax]
for thr in (8, 4, 2, 1):
    for _ in range(3):
        lib.vp_host_stage_h2d(C.c_void_p(ptrs.ctypes.data), C.c_void_p(lens.ctypes.data), B, lmax, C.c_void_p(stage.data_ptr()),
                              C.c_void_p(), 4, thr, C.c_void_p())
    barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        lib.vp_host_stage_h2d(C.c_void_p(ptrs.ctypes.data), C.c_void_p(lens.ctypes.data), B, lmax, C.c_void_p(stage.data_ptr()),
                              C.c_void_p(), 4, thr, C.c_void_p())
    ms = (time.perf_counter() - t0) * 100
    say(f'gather only, {thr} threads, all ranks at once: ms per {nbytes / 1e6:.0f} MB per rank {[round(m, 2) for m in allranks(ms)]}')
del hp, dp


# 3. the whole call, all ranks at once (barrier before every call, like the bench's step)
def run(n, trace=False):
    ts = []
    for i in range(n):
        barrier()
        pred._trace = [] if trace else None
        pred._trace_dev = [] if trace else None
        t0 = time.perf_counter()
        pred.predict_batch(pools[i % 4])
        t1 = time.perf_counter()
        ts.append((t1 - t0) * 1e3)
        if trace and i == n - 1 and rank == 0:
            prev = t0
            for label, t in pred._trace:
                print(f'    +{(t - prev) * 1e3:7.3f} ms  {label}')
                prev = t
            print(f'    +{(t1 - prev) * 1e3:7.3f} ms  return')
            torch.cuda.synchronize()
            ev0 = pred._trace_dev[0][1]
            for label, e in pred._trace_dev[1:]:
                print(f'      {ev0.elapsed_time(e):7.3f}  {label}')
    pred._trace = pred._trace_dev = None
    return ts


run(5)
ts = run(12, trace=True)
say(f'whole call, default knobs: median ms per rank {[round(m, 2) for m in allranks(np.median(ts))]}')
for thr, sl in [(8, 4), (4, 4), (3, 4), (2, 4), (1, 4), (2, 8)]:
    os.environ['VPB_GATHER_THREADS'] = str(thr)
    MVectorPredictor.COPY_SLICE = sl
    run(3)
    ts = run(10)
    say(f'threads={thr} COPY_SLICE={sl}: median ms per rank {[round(m, 2) for m in allranks(np.median(ts))]}')
barrier()
if world > 1:
    dist.destroy_process_group()
