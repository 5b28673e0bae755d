#!/usr/bin/env python
"""bench.py -- embeddings/sec of the waveform -> Fbank -> EcapaTdnn -> 192-d embedding path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch: 256 synthetic 3 s @ 16 kHz utterances PER GPU (BASELINE
configs[1]; weak scaling: the global batch is 256*N, sharded by contiguous utterance ranges, followed by one NCCL
all-gather of the [256, 192] embeddings so every rank holds all of them).

Timed quantities (all on the device with CUDA events, max over ranks, barrier + synchronize on both sides):
  value  : whole-job embeddings/s with the waveforms already resident in HBM (vp_embed_wave + all-gather)
  e2e    : the same through the public API MVectorPredictor.predict_batch(list of host float32 arrays): host staging,
           pinned H2D copy, kernels, D2H of the embeddings (+ all-gather) inside the timed region
  roofline : the dominant kernel (largest share of device time among the program's ops, measured live with per-op CUDA
           events by vp_embed_profiled on extra steps after the timed region): algorithmic FLOPs / launch time vs the
           measured dense bf16 tensor peak of MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (port of the reference's predict_batch flow: per-utterance Kaldi fbank, CMN, model in
           chunks of 32, no_grad) on the box's host cores, bounded sample
`--impl reference` times that CPU path alone and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 256
SAMPLES = 48000
MODEL = 'EcapaTdnn'
MODEL_ARGS = dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536])
FBANK_ARGS = dict(sample_frequency=16000, num_mel_bins=80)
GFLOP_PER_UTT = 3.090          # BASELINE.md section 3 (2 x conv/linear MACs of the reference backbone at T=298)
N_POOL = 4                     # rotating input batches: 4 x 49 MB = 196 MB > 126 MB L2


def bench_config(n_gpus):
    return {'workload': 'EcapaTdnn+Fbank80, batch 256 x 3 s @ 16 kHz per GPU (BASELINE configs[1])',
            'global_batch': BATCH_PER_GPU * n_gpus, 'samples_per_utt': SAMPLES, 'frames': 298,
            'parallelism': f'utterance-sharded x{n_gpus} + all-gather of embeddings',
            'l2': f'inputs rotate over {N_POOL} distinct batches ({N_POOL * BATCH_PER_GPU * SAMPLES * 4 >> 20} MiB > L2); '
                  'per-step activation traffic is several GB',
            'weights': 'seeded random init + randomised BN statistics (oracle.models.random_state_dict seed 0)',
            'precision': 'fp32 in/out; tensor-core GEMMs use error-compensated split-TF32 (fp32-grade, 1e-4 parity)'}


def synth_waves(n, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, SAMPLES, generator=g) * 0.1


def yml_config():
    return {'dataset_conf': {'dataset': {'min_duration': 0.3, 'max_duration': 3, 'sample_rate': 16000,
                                         'use_dB_normalization': False, 'target_dB': -20},
                             'eval_conf': {'batch_size': 16, 'max_duration': 20}},
            'preprocess_conf': {'use_hf_model': False, 'feature_method': 'Fbank', 'method_args': dict(FBANK_ARGS)},
            'model_conf': {'model': MODEL, 'model_args': dict(MODEL_ARGS)}}


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's flow)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(sd, waves):
    """predict.py:244-264 on CPU: featurizer on the whole list (per-utterance kaldi fbank loop + CMN), model in chunks of
    32 (the reference's default batch_size).  Under no_grad (the shipped code leaves autograd on; this is the faster,
    i.e. stronger, baseline)."""
    import torch
    from oracle import frontend as ofe
    from oracle import models as om
    with torch.no_grad():
        feats = ofe.featurize(waves, torch.ones(waves.shape[0]), 'Fbank', FBANK_ARGS)
        out = [om.forward(MODEL, sd, feats[i:i + 32], **MODEL_ARGS) for i in range(0, feats.shape[0], 32)]
    return torch.cat(out)


def pick_cpu_threads(sd, waves):
    """torchrun pins OMP_NUM_THREADS=1 and oversubscribed hosts are slower with every hardware thread: try the whole
    machine, half and a quarter of it on one pass each and keep the fastest (the strongest CPU baseline)."""
    import torch
    ncpu = os.cpu_count() or 1
    best, best_t = None, None
    for n in sorted({ncpu, max(ncpu // 2, 1), max(ncpu // 4, 1)}, reverse=True):
        torch.set_num_threads(n)
        cpu_reference_pass(sd, waves[:8])
        t0 = time.perf_counter()
        cpu_reference_pass(sd, waves[:16])
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = n, t
    torch.set_num_threads(best)
    return best


def time_cpu_reference(n_utts, budget_s, min_reps=1, warmup=1):
    import torch
    from oracle import models as om
    sd = om.random_state_dict(MODEL, 80, seed=0, **MODEL_ARGS)
    waves = synth_waves(n_utts, 4321)
    pick_cpu_threads(sd, waves)
    for _ in range(warmup):
        cpu_reference_pass(sd, waves)
    reps, t0 = 0, time.perf_counter()
    while True:
        cpu_reference_pass(sd, waves)
        reps += 1
        el = time.perf_counter() - t0
        if reps >= min_reps and el >= budget_s:
            break
    return n_utts * reps / el, reps, el


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = 64
    waves = synth_waves(n, 4321)
    from oracle import models as om
    sd = om.random_state_dict(MODEL, 80, seed=0, **MODEL_ARGS)
    cores = pick_cpu_threads(sd, waves)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(sd, waves)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(sd, waves)
    el = time.perf_counter() - t0
    v = n * args.steps / el
    sample = f'{n} utterances x 3 s per step (bounded sample of the 256-utterance batch), oracle port of predict_batch, no_grad'
    line = {'impl': 'reference', 'metric': 'embeddings/sec (3s@16kHz) ECAPA-TDNN', 'value': v, 'unit': 'emb/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': bench_config(args.gpus),
            'cpu_baseline': {'value': v, 'unit': 'emb/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'emb/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line): started before the warm-up so that it is
    already streaming, every sample carries a timestamp and only those inside [t0, t1] (the timed region) are used."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '20'], stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(1.0)                      # let it start streaming
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(',')]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(c[1]), float(c[2]), c[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        window = 'timed region'
        if len(inside) < 2:                      # region shorter than the sampling period: widen to warm-up + region
            inside, window = rows, 'warm-up + timed region'
        reasons = set()
        for _, _, _, flags in inside:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), flags):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if inside:
            sm = sorted(r[1] for r in inside)
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(r[2] for r in inside), 'reasons': sorted(reasons),
                   'samples': len(inside), 'window': window}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback on the product path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    def mark(msg):
        if os.environ.get('VPB_BENCH_TRACE'):
            sys.stderr.write(f'[bench rank {rank} {time.strftime("%H:%M:%S")}] {msg}\n'); sys.stderr.flush()
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
        mark('process group up')
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'

    import __graft_entry__ as ge
    ge.build()
    from mvector.predict import MVectorPredictor
    from oracle import models as om

    sd = om.random_state_dict(MODEL, 80, seed=0, **MODEL_ARGS)
    with tempfile.TemporaryDirectory() as td:
        torch.save({'0.' + k: v for k, v in sd.items()}, os.path.join(td, 'model.pth'))
        import logging
        from loguru import logger
        logger.remove()
        pred = MVectorPredictor(configs=yml_config(), model_path=td, use_gpu=True)
    B = BATCH_PER_GPU
    fz = pred._audio_featurizer
    T = fz.num_frames(SAMPLES)
    prog = pred.predictor.program(B, T)
    from mvector import _lib as L
    pool_host = [synth_waves(B, 1234 + 100 * rank + i).pin_memory() for i in range(N_POOL)]
    pool_dev = [w.to(dev) for w in pool_host]
    # predict_batch input: a list of INDEPENDENTLY allocated pageable numpy arrays (like decoded audio files), not views
    # of one pinned matrix
    pool_np = [[np.array(w[i].numpy(), copy=True) for i in range(B)] for w in pool_host]
    feats = torch.empty(B * T * 80, dtype=torch.float32, device=dev)
    scratch = torch.empty(max(int(L.lib().vp_frontend_scratch_floats(fz.engine.handle, B, SAMPLES)), 1),
                          dtype=torch.float32, device=dev)
    emb = torch.empty(B, 192, dtype=torch.float32, device=dev)
    emb_all = torch.empty(B * world, 192, dtype=torch.float32, device=dev)

    def step_resident(i):
        prog.run_wave(pool_dev[i % N_POOL], None, feats, scratch, emb)
        if world > 1:
            dist.all_gather_into_tensor(emb_all, emb)

    def step_e2e(i):
        e = pred.predict_batch(pool_np[i % N_POOL])
        if world > 1:
            dist.all_gather_into_tensor(emb_all, torch.from_numpy(e).to(dev))
        return e

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, sampler=None):
        for i in range(warmup):
            step_fn(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.mark_start()
        e0.record()
        for i in range(steps):
            step_fn(warmup + i)
        e1.record()
        sync_all()
        if sampler:
            sampler.mark_end()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    mark('programs built, starting timed region')
    clocks = ClockSampler(local) if rank == 0 else None
    ms_res = timed(step_resident, args.steps, args.warmup, clocks)
    clk = clocks.stop() if clocks else None
    mark(f'resident timing done: {ms_res:.2f} ms')
    if os.environ.get('VPB_E2E_ONLY'):
        ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
        if rank == 0:
            print(json.dumps({'e2e_only': True, 'ms_per_step': ms_e2e / args.steps, 'emb_per_s': B * world * args.steps / (ms_e2e * 1e-3)}))
        return
    if args.light:
        if rank == 0:
            print(json.dumps({'light': True, 'ms_per_step': ms_res / args.steps, 'note': 'not a bench value'}))
        if world > 1:
            dist.destroy_process_group()
        return
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))

    mark(f'e2e timing done: {ms_e2e:.2f} ms')
    # sanity inside the bench: the embeddings of the last step agree with the CPU oracle on 2 utterances
    if rank == 0:
        from oracle import frontend as ofe
        i_last = (args.warmup + args.steps - 1) % N_POOL
        prog.run_wave(pool_dev[i_last], None, feats, scratch, emb)     # local only: no collective on a single rank
        torch.cuda.synchronize()
        ref = om.forward(MODEL, sd, ofe.featurize(pool_host[i_last][:2], None, 'Fbank', FBANK_ARGS), **MODEL_ARGS)
        err = float(((emb[:2].cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max())
        assert err < 1e-4, f'parity check inside bench failed: rel-L2 {err}'

    # ---- live per-op timing for the roofline (extra steps, not part of `value`) ----
    roof = None
    if rank == 0:
        fz(pool_dev[0])                        # features for the profiled backbone pass
        f_in = fz(pool_dev[0]).contiguous()
        acc = None
        nprof = 3
        for _ in range(nprof):
            ops = prog.run_profiled(f_in, emb)
            if acc is None:
                acc = ops
            else:
                for a, o in zip(acc, ops):
                    a['ms'] += o['ms']
        for a in acc:
            a['ms'] /= nprof
        total = sum(a['ms'] for a in acc)
        if args.dump_ops:
            with open(args.dump_ops, 'w') as f:
                json.dump(acc, f, indent=0)
        conv = [a for a in acc if a['kind'] == L.OP_CONV and a['K'] > 0]
        top = max(conv, key=lambda a: a['ms'])
        peaks = {}
        ppath = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(ppath):
            with open(ppath) as f:
                peaks = json.load(f)
        peak = peaks.get('bf16_tflops_sustained')
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)'
        if peak is None:
            peak, peak_src = 1400.0, 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md; of fallback)'
        flops = 2.0 * top['M'] * top['N'] * top['K']
        achieved = flops / (top['ms'] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"{top['M']}x{top['N']}x{top['K']}")
        gemm_ms = sum(a['ms'] for a in conv)
        gemm_flops = sum(2.0 * a['M'] * a['N'] * a['K'] for a in conv)
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'traffic': traffic, 'peak_source': peak_src,
                'kernel': ('conv_tc (tcgen05 split-TF32)' if top['engine'] == L.ENGINE_TC else 'conv_ffma_kernel<128> (fp32 FFMA)')
                          + f" M={top['M']} N={top['N']} K={top['K']}",
                'kernel_ms': top['ms'], 'kernel_share_of_backbone': top['ms'] / total,
                'all_conv': {'tflops': gemm_flops / (gemm_ms * 1e-3) / 1e12, 'ms': gemm_ms, 'share_of_backbone': gemm_ms / total},
                'backbone_ms_profiled': total}

    if rank == 0:
        n_emb = B * world * args.steps
        value = n_emb / (ms_res * 1e-3)
        e2e_v = n_emb / (ms_e2e * 1e-3)
        cpu_v, reps, el = time_cpu_reference(32, budget_s=12.0) if world == 1 else (None, 0, 0.0)
        cores = torch.get_num_threads()
        line = {'metric': 'embeddings/sec (3s@16kHz) ECAPA-TDNN', 'value': value, 'unit': 'emb/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_res / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': bench_config(world), 'clocks': clk,
                'e2e': {'value': e2e_v, 'unit': 'emb/s', 'h2d_bytes_per_step': B * SAMPLES * 4,
                        'd2h_bytes_per_step': B * 192 * 4, 'ms_per_step': ms_e2e / args.steps,
                        'api': 'MVectorPredictor.predict_batch(list of 256 host float32 arrays)'},
                'gpu_launches': args.steps * (prog.launches + 2),
                'launches_per_step': prog.launches + 2,
                'roofline': roof,
                'tensor_frac_whole_step': (value * GFLOP_PER_UTT * 1e9 / 1e12) / (roof['peak'] * world) if roof else None}
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': 'emb/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'32 utterances x 3 s, {reps} passes in {el:.1f} s; oracle port of '
                                              'predict_batch (kaldi fbank per utterance, model chunks of 32, no_grad)'}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--light', action='store_true', help='resident-input timing only (for runs under ncu): no e2e / '
                    'cpu baseline / per-op profile; the JSON line is NOT a bench value')
    ap.add_argument('--dump-ops', default=None, help='write the per-op device-time profile (JSON) to this path')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
