"""Tiny workload for compute-sanitizer (memcheck / racecheck): small-config ECAPA + CAM++ from waveforms."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import json, numpy as np, torch
from loguru import logger
logger.remove()
from conftest import load_golden
from mvector.models import build_model
from mvector.utils.utils import dict_to_object
from mvector.data_utils.featurizer import AudioFeaturizer
man = json.load(open('tests/golden/manifest.json'))
for name in ('ecapa_small', 'campplus_small', 'eres2net_small'):
    m = man[name]
    z, sd = load_golden(name)
    model = build_model(m['feature_dim'], dict_to_object({'model_conf': {'model': m['model'], 'model_args': m['model_args']}}))
    model.load_state_dict(sd)
    fz = AudioFeaturizer(m['preprocess']['feature_method'], method_args=m['preprocess']['method_args'])
    w = np.stack([z['wave0'][:7360], z['wave1'][:7360]])
    emb = model(fz(torch.from_numpy(w))).cpu().numpy()
    # big enough M for the tcgen05 engine too
    feats = torch.from_numpy(np.tile(z['feats'][:1], (24, 1, 1))).cuda()
    emb2 = model(feats).cpu().numpy()
    print(name, emb.shape, float(np.abs(emb).max()), emb2.shape, flush=True)
print('done')
