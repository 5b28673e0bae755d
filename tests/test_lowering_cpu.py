"""Host-side lowering (mvector/models/*.py mirrors: weight folding, layout permutations, static memory plan, op
parameters) checked on CPU: the lowered vp_op program is run by the numpy interpreter tests/plan_sim.py and compared
with the reference's golden embeddings.  No GPU needed."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from plan_sim import simulate

from mvector.models import build_model
from mvector.utils.utils import dict_to_object

SMALL = ['ecapa_small', 'tdnn_small', 'campplus_small', 'resnetse_small', 'eres2net_small', 'eres2net_wide_small',
         'ecapa_sap_small', 'tdnn_tsp_small', 'resnetse_tap_small', 'res2net_small', 'eres2netv2_small', 'tdnn_spec_small',
         'resnetse_mfcc_small', 'ecapa_mfcc400_small']


def build(m, sd):
    model = build_model(m['feature_dim'], dict_to_object({'model_conf': {'model': m['model'],
                                                                       'model_args': m['model_args']}}))
    model.load_state_dict({'0.' + k: v for k, v in sd.items()})
    return model


@pytest.mark.parametrize('name', SMALL)
def test_lowered_program_matches_reference_golden(name, manifest):
    m = manifest[name]
    z, sd = load_golden(name)
    model = build(m, sd)
    emb, pb = simulate(model, z['feats'])
    err = rel_l2(emb, z['emb']).max()
    assert err < 2e-5, err
    # memory plan sanity: nothing live at the end except what the model chose to leak, peak below sum of allocations
    assert pb.peak > 0 and len(pb.ops) > 5


def test_unknown_model_arg_raises_typeerror():
    with pytest.raises(TypeError):
        build_model(80, dict_to_object({'model_conf': {'model': 'EcapaTdnn', 'model_args': {'bogus': 1}}}))
    with pytest.raises(AttributeError):
        build_model(80, dict_to_object({'model_conf': {'model': 'NoSuchNet', 'model_args': {}}}))


_TINY = {
    'EcapaTdnn': dict(embd_dim=32, channels=[64, 64, 64, 64, 192], attention_channels=32, res2net_scale=4, se_channels=16),
    'TDNN': dict(embd_dim=32, channels=64),
    'CAMPPlus': dict(embd_dim=32, growth_rate=8, bn_size=4, init_channels=32),
    'ResNetSE': dict(embd_dim=32, layers=[1, 1, 1, 1], num_filters=[16, 16, 32, 32]),
    'ERes2Net': dict(embd_dim=32, num_blocks=[1, 1, 1, 1], m_channels=8),
    'ERes2NetV2': dict(embd_dim=32, num_blocks=[1, 1, 1, 1], m_channels=8),
    'Res2Net': dict(embd_dim=32, m_channels=8, layers=[1, 1, 1, 1]),
}


@pytest.mark.parametrize('name,fdim', [('EcapaTdnn', 201), ('EcapaTdnn', 13), ('TDNN', 257), ('CAMPPlus', 201),
                                       ('CAMPPlus', 13), ('ResNetSE', 40), ('ERes2Net', 24), ('ERes2NetV2', 40),
                                       ('Res2Net', 64)])
def test_lowering_over_feature_dims(name, fdim):
    """Feature dims other than 80: odd dims (Spectrogram's n_fft/2+1 bins, 13 MFCCs) go through the 1-D models' padded
    input and CAM++'s 2-D head; the 2-D nets accept whatever the reference's own shape arithmetic accepts."""
    import torch
    from oracle import models as om
    margs = _TINY[name]
    sd = om.random_state_dict(name, fdim, seed=1, gain=0.8, **margs)
    x = torch.randn(2, 70, fdim, generator=torch.Generator().manual_seed(2)) * 2
    ref = om.forward(name, sd, x, **margs).numpy()
    model = build(dict(feature_dim=fdim, model=name, model_args=margs), sd)
    got, _ = simulate(model, x.numpy())
    assert rel_l2(got, ref).max() < 1e-5


@pytest.mark.parametrize('name,fdim', [('ResNetSE', 201), ('ERes2Net', 13), ('ERes2NetV2', 201), ('Res2Net', 24)])
def test_feature_dims_the_reference_rejects_fail_loudly(name, fdim):
    """Where the reference's own forward dies on a shape mismatch (freq bins not matching the pooled-width formula) the
    mirror must refuse at lowering time instead of computing something."""
    import torch
    margs = _TINY[name]
    model = build_model(fdim, dict_to_object({'model_conf': {'model': name, 'model_args': margs}}))
    model.load_state_dict({'0.' + k: torch.zeros(v) for k, v in model.param_shapes().items()})
    with pytest.raises((AssertionError, ValueError)):
        model.lower(2, 70)
