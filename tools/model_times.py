"""Device time per backbone forward at the BASELINE sizes + optional per-op profile (dev tool; not a bench value).

    python tools/model_times.py [--dump gpurun_out/ops_] [--only CAMPPlus]"""
import argparse, json, sys
sys.path.insert(0, '.')
import torch
import __graft_entry__ as ge
ge.build()
from mvector.models import build_model
from mvector.utils.utils import dict_to_object
from oracle import models as om
from loguru import logger
logger.remove()
ap = argparse.ArgumentParser()
ap.add_argument('--dump', default=None)
ap.add_argument('--only', default=None)
a = ap.parse_args()
cases = [('EcapaTdnn', 80, dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536]), 256, 298, 3.090, 'c2'),
         ('TDNN', 80, dict(embd_dim=192, channels=512, pooling_type='ASP'), 256, 298, 1.45, 'tdnn'),
         ('CAMPPlus', 80, dict(embd_dim=192), 256, 298, 3.355, 'c3'),
         ('ResNetSE', 64, dict(embd_dim=192, pooling_type='ASP'), 128, 251, 7.462, 'c4'),
         ('ERes2Net', 80, dict(embd_dim=192, m_channels=32), 128, 298, 10.1, 'eres2net'),
         ('ERes2Net', 80, dict(embd_dim=192, m_channels=64, mul_channel=2, expansion=4, base_width=24, scale=3), 64, 998, 312.17, 'c5')]
for name, fd, margs, B, T, gflop, tag in cases:
    if a.only and a.only not in (name, tag):
        continue
    m = build_model(fd, dict_to_object({'model_conf': {'model': name, 'model_args': margs}}))
    m.load_state_dict(om.random_state_dict(name, fd, seed=0, **margs))
    prog = m.program(B, T)
    f = torch.randn(B, T, fd, device='cuda')
    e = torch.empty(B, 192, device='cuda')
    for _ in range(2): prog.run(f, e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): prog.run(f, e)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{tag:8s} {name:10s} B={B:4d} T={T} ops={prog.n_ops:4d} ws={prog.ws_bytes/2**30:5.2f} GiB  {ms:8.3f} ms  {B/ms*1e3:9.0f} emb/s  {B*gflop/ms:8.1f} TFLOP/s", flush=True)
    if a.dump:
        ops = prog.run_profiled(f, e)
        with open(f'{a.dump}{tag}.json', 'w') as fh:
            json.dump(ops, fh)
    m.engine.close()
