"""Stand-alone check (authoring container only): the oracle against the UNMODIFIED reference modules imported from
/root/reference, at the reference's default / BASELINE model sizes.  Run by tests/test_oracle_vs_reference.py in a
subprocess (the reference package is also called ``mvector``, so it cannot share an interpreter with the mirror).

    PYTHONPATH=/root/reference:<repo> python tests/oracle_vs_reference_check.py
"""
import sys
import warnings

import torch

warnings.simplefilter('ignore')
from loguru import logger  # noqa: E402

logger.remove()

from mvector.data_utils.featurizer import AudioFeaturizer  # noqa: E402  (reference)
from mvector.models import build_model  # noqa: E402  (reference)
from mvector.utils.utils import dict_to_object  # noqa: E402  (reference)
from oracle import frontend as ofe, models as om  # noqa: E402

assert '/root/reference' in sys.modules['mvector'].__file__, 'the reference package must be first on PYTHONPATH'

FULL = [
    ('EcapaTdnn', 80, dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536]), 2, 120),
    ('TDNN', 80, dict(embd_dim=192, channels=512, pooling_type='ASP'), 2, 120),
    ('CAMPPlus', 80, dict(embd_dim=192), 2, 150),
    ('ResNetSE', 64, dict(embd_dim=192, pooling_type='ASP'), 2, 80),
    ('ERes2Net', 80, dict(embd_dim=192, m_channels=32), 2, 72),
    ('ERes2Net', 80, dict(embd_dim=192, m_channels=64, mul_channel=2, expansion=4, base_width=24, scale=3), 1, 40),
    ('ERes2NetV2', 80, dict(embd_dim=192, m_channels=32), 2, 72),
    ('Res2Net', 80, dict(embd_dim=192, pooling_type='ASP', m_channels=32), 2, 120),
    ('EcapaTdnn', 80, dict(embd_dim=192, pooling_type='SAP'), 2, 100),
    ('TDNN', 80, dict(embd_dim=192, pooling_type='TSP'), 2, 100),
    ('ResNetSE', 64, dict(embd_dim=192, pooling_type='TAP'), 2, 80),
]
worst = 0.0
for name, fdim, margs, B, T in FULL:
    sd = om.random_state_dict(name, fdim, seed=5, gain=om.CONDITIONED_GAIN[name], **margs)
    ref = torch.nn.Sequential(build_model(fdim, dict_to_object({'model_conf': {'model': name, 'model_args': margs}})))
    assert [k[2:] for k in ref.state_dict()] == list(sd), name          # same names, same order
    missing, unexpected = ref.load_state_dict({'0.' + k: v for k, v in sd.items()}, strict=True)
    ref.eval()
    x = torch.randn(B, T, fdim, generator=torch.Generator().manual_seed(9)) * 2.0
    with torch.no_grad():
        a = ref(x)
        b = om.forward(name, sd, x, **margs)
    err = float(((a - b).norm(dim=1) / a.norm(dim=1)).max())
    worst = max(worst, err)
    assert err < 1e-6, (name, margs, err)
    print(f'model {name:11s} {str(margs)[:60]:60s} rel-L2 {err:.2e}')

g = torch.Generator().manual_seed(1)
w = torch.randn(3, 20000, generator=g) * 0.1
ratio = torch.tensor([1.0, 0.63, 0.3])
FRONT = [
    ('Fbank', dict(sample_frequency=16000, num_mel_bins=80)),
    ('Fbank', dict(sample_frequency=16000, num_mel_bins=40, window_type='hamming', use_power=False)),
    ('MelSpectrogram', dict()),
    ('MelSpectrogram', dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=320, f_min=50.0, f_max=14000.0,
                            n_mels=64)),
    ('Spectrogram', dict()),
    ('Spectrogram', dict(n_fft=512, hop_length=160, power=1.0)),
    ('MFCC', dict()),
    ('MFCC', dict(n_mfcc=24, melkwargs=dict(n_fft=512, hop_length=160, n_mels=64, f_min=20.0))),
    ('MFCC', dict(log_mels=True, norm=None)),
]
for method, args in FRONT:
    fz = AudioFeaturizer(method, method_args=args)
    assert torch.equal(fz(w, ratio), ofe.featurize(w, ratio, method, args)), (method, args)
    assert torch.equal(fz(w[1]), ofe.featurize(w[1], None, method, args)), (method, args)
    assert fz.feature_dim == ofe.feature_dim(method, args)
    print(f'front-end {method:14s} {str(args)[:70]:70s} bit-exact')
print('ORACLE_VS_REFERENCE_OK worst model rel-L2 %.2e' % worst)
