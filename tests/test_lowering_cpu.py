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
