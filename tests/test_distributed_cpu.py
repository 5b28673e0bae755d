"""Host-side sharding / all-gather logic of the multi-GPU path on CPU: world_size 2 and 3 over gloo, with a fake
(deterministic, per-utterance) embed function standing in for the GPU kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvector.distributed import embed_sharded, pad_to_global_max, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 256, 2048, 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_embed(x, ratio):
    """Depends on the padded length (like the reference's ragged semantics) and on each utterance's own samples."""
    x = torch.from_numpy(np.asarray(x))
    feats = torch.stack([x.sum(1), (x * x).sum(1), torch.from_numpy(np.asarray(ratio)) * x.shape[1],
                         torch.full((x.shape[0],), float(x.shape[1]))], dim=1)
    return feats.float()


def _worker(rank, world, port, n):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # identical list on every rank
    waves = [rng.standard_normal(int(rng.integers(50, 400))).astype(np.float32) for _ in range(n)]
    got = embed_sharded(_fake_embed, waves, 4, torch.device('cpu'))
    x, ratio = pad_to_global_max(waves)
    ref = _fake_embed(x, ratio)                         # single-process, whole batch
    assert got.shape == ref.shape
    assert torch.equal(got, ref), (rank, (got - ref).abs().max())
    dist.barrier()
    dist.destroy_process_group()


class _FakePredictor:
    """Host-side stand-in with the three things predict_batch_sharded touches: configs, _load_audio, _embed_waves."""
    class _Seg:
        def __init__(self, x):
            self.samples = x

    def __init__(self):
        from mvector.utils.utils import dict_to_object
        self.configs = dict_to_object({'dataset_conf': {'dataset': {'sample_rate': 16000, 'min_duration': 0.001,
                                                                    'use_dB_normalization': False}}})
        self.loaded = []

    def _load_audio(self, audio_data, sample_rate=16000):
        self.loaded.append(len(audio_data))
        return self._Seg(np.asarray(audio_data, dtype=np.float32))

    def _embed_waves(self, waves, lmax, masked, to_numpy=True, group=None):
        if not waves:
            return torch.zeros(0, 4)
        x = np.zeros((len(waves), lmax), dtype=np.float32)
        for i, w in enumerate(waves):
            x[i, :len(w)] = w
        return _fake_embed(x, np.asarray([len(w) / lmax for w in waves], dtype=np.float32))


def _worker_predict(rank, world, port, n):
    from mvector.distributed import predict_batch_sharded, shard_range
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    waves = [rng.standard_normal(int(rng.integers(50, 400))).astype(np.float32) for _ in range(n)]
    x, ratio = pad_to_global_max(waves)
    lo, hi = shard_range(n, rank, world)
    # float64 arrays need _load_audio's conversion: a rank decodes only its own shard, the rest contributes lengths
    pred = _FakePredictor()
    got = predict_batch_sharded(pred, [w.astype(np.float64) for w in waves], as_numpy=False)
    assert torch.equal(got, _fake_embed(x, ratio))               # global Lmax / ratios although only the shard was staged
    assert sorted(pred.loaded) == sorted(len(w) for w in waves[lo:hi])
    # raw float32 arrays at the model's rate (dB normalisation off) are used in place: nothing is decoded at all
    pred = _FakePredictor()
    got = predict_batch_sharded(pred, waves, as_numpy=False)
    assert torch.equal(got, _fake_embed(x, ratio)) and pred.loaded == []
    dist.barrier()
    dist.destroy_process_group()


def test_predict_batch_sharded_gloo_world2_and_more_ranks_than_items():
    mp.spawn(_worker_predict, args=(2, _free_port(), 7), nprocs=2, join=True)
    mp.spawn(_worker_predict, args=(3, _free_port(), 2), nprocs=3, join=True)     # one rank gets an empty shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_embed_sharded_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port(), 7), nprocs=2, join=True)


def test_embed_sharded_gloo_world3_uneven():
    mp.spawn(_worker, args=(3, _free_port(), 5), nprocs=3, join=True)
