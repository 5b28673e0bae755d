"""The oracle (CPU restatement) against the committed golden vectors produced by the real reference
(tests/golden/make_golden.py).  Runs everywhere (no GPU, no /root/reference)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import frontend as fe
from oracle import models as om

SMALL = ['ecapa_small', 'tdnn_small', 'campplus_small', 'resnetse_small', 'eres2net_small', 'eres2net_wide_small',
         'ecapa_sap_small', 'tdnn_tsp_small', 'resnetse_tap_small', 'res2net_small', 'eres2netv2_small', 'tdnn_spec_small',
         'resnetse_mfcc_small', 'ecapa_mfcc400_small']


@pytest.mark.parametrize('name', SMALL)
def test_oracle_matches_reference_golden(name, manifest):
    m = manifest[name]
    z, sd = load_golden(name)
    waves = [z['wave%d' % i] for i in range(len(m['lens']))]
    x, ratio = fe.pad_batch(waves)
    feats = fe.featurize(x, ratio, m['preprocess']['feature_method'], m['preprocess']['method_args'])
    # front-end: same torch CPU ops as the reference -> bit exact
    assert np.array_equal(feats.numpy(), z['feats'])
    emb = om.forward(m['model'], sd, feats, **m['model_args']).numpy()
    assert rel_l2(emb, z['emb']).max() < 1e-6
    # single-utterance path (predict.py:214-229): no padding, no mask
    f1 = fe.featurize(waves[-1], None, m['preprocess']['feature_method'], m['preprocess']['method_args'])
    e1 = om.forward(m['model'], sd, f1, **m['model_args']).numpy()[0]
    assert rel_l2(e1, z['emb_single_last']).max() < 1e-6


def test_seeded_weights_reproduce_golden_state_dict(manifest):
    """random_state_dict must regenerate the exact weights the goldens were made with (bench/tests on the GPU box
    rely on the seeded generator, not on files)."""
    m = manifest['ecapa_small']
    _, sd = load_golden('ecapa_small')
    sd2 = om.random_state_dict(m['model'], m['feature_dim'], seed=m['seed'], **m['model_args'])
    assert list(sd) == list(sd2)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k


def test_c1_contrast_golden(manifest):
    """Config #1: TDNN + Fbank on dataset/a_1.wav vs a_2.wav through the infer_contrast.py flow."""
    m = manifest['c1_tdnn_contrast']
    z = np.load('tests/golden/c1_tdnn_contrast.npz') if False else load_golden('c1_tdnn_contrast')[0]
    sd = om.random_state_dict('TDNN', 80, seed=m['seed'], **m['model_args'])
    embs = []
    for nm in ('a_1', 'a_2'):
        w = z['pcm_' + nm].astype(np.float32) / 32768.0
        rms_db = 10.0 * np.log10(np.mean(w.astype(np.float64) ** 2))
        w = (w * (10.0 ** ((-20 - rms_db) / 20.0))).astype(np.float32)
        f = fe.featurize(w, None, 'Fbank', dict(sample_frequency=16000, num_mel_bins=80))
        assert f.shape[1] == {'a_1': 365, 'a_2': 218}[nm]
        e = om.forward('TDNN', sd, f, **m['model_args']).numpy()[0]
        assert rel_l2(e, z['emb_' + nm]).max() < 1e-6
        embs.append(e)
    sim = float(np.dot(embs[0], embs[1]) / (np.linalg.norm(embs[0]) * np.linalg.norm(embs[1])))
    assert abs(sim - float(z['sim'])) < 1e-6


def test_param_shape_digests(manifest):
    """oracle.param_shapes enumerates exactly the reference's state_dict (names, order, shapes) at the
    default / BASELINE configurations."""
    for key, d in manifest['_param_digests'].items():
        model = 'ERes2Net' if key in ('ERes2Net', 'ERes2Net55M') else key
        shapes = om.param_shapes(model, d['input_size'], **d['model_args'])
        s = ';'.join(f'{k}:{tuple(v)}' for k, v in shapes.items())
        assert hashlib.sha256(s.encode()).hexdigest() == d['sha256'], key
        assert len(shapes) == d['n_tensors']


def test_fbank_matches_torchaudio():
    """torchaudio (third-party, in the image) is what the reference calls (featurizer.py:128)."""
    ka = pytest.importorskip('torchaudio.compliance.kaldi')
    g = torch.Generator().manual_seed(3)
    for n in (400, 559, 560, 16000, 48000):
        w = torch.randn(n, generator=g) * 0.1
        a = ka.fbank(w[None], sample_frequency=16000, num_mel_bins=80)
        b = fe.kaldi_fbank(w, sample_frequency=16000, num_mel_bins=80)
        assert a.shape == b.shape == (1 + (n - 400) // 160, 80)
        assert torch.equal(a, b)
    with pytest.raises(TypeError):
        fe.kaldi_fbank(torch.zeros(1000), n_mels=3)


def test_featurize_ragged_semantics():
    """CMN mean over ALL padded frames, then frames >= round(ratio*T) zeroed."""
    g = torch.Generator().manual_seed(5)
    a = (torch.randn(16000, generator=g) * 0.1).numpy()
    b = (torch.randn(48000, generator=g) * 0.1).numpy()
    x, ratio = fe.pad_batch([a, b])
    f = fe.featurize(x, ratio, 'Fbank', dict(sample_frequency=16000, num_mel_bins=80))
    T = f.shape[1]
    keep = int(torch.round(torch.tensor(ratio[0]) * T))
    assert T == 298 and keep == 99
    assert torch.all(f[0, keep:] == 0) and not torch.all(f[0, keep - 1] == 0)
    alone = fe.featurize(a, None, 'Fbank', dict(sample_frequency=16000, num_mel_bins=80))
    assert alone.shape[1] == 98
    assert (alone[0, :90] - f[0, :90]).abs().max() > 1.0   # batch-composition dependent (different CMN mean)
