"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (/root/reference).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned on outputs of the
reference's own modules:
  * mvector.data_utils.featurizer.AudioFeaturizer           (featurizer.py:9-111)
  * mvector.models.build_model -> nn.Sequential(backbone)    (models/__init__.py:15-21, predict.py:54-55)
  * mvector.utils.checkpoint.load_pretrained                 (checkpoint.py:11-51)
  * mvector.predict.MVectorPredictor.predict / predict_batch / contrast (predict.py:214-279), with a test-only
    stub of the absent third-party ``yeaudio.audio.AudioSegment`` (requirements.txt:12).
Weights are seeded random with randomised BN statistics (oracle.models.random_state_dict) at SMALL model
configurations so the fixtures stay small; full-size parity is oracle-vs-CUDA (the oracle itself is checked
bit-for-bit against the reference at full size by tests/test_oracle_vs_reference.py when the reference is present).
"""
import hashlib
import json
import os
import sys
import tempfile
import types
import wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)


def install_yeaudio_stub():
    """Minimal stand-in for yeaudio.audio.AudioSegment: fields samples/sample_rate/duration, from_ndarray,
    from_file (16-bit PCM wav via the stdlib), normalize (gain to target dB), resample unsupported."""
    class AudioSegment:
        def __init__(self, samples, sample_rate):
            self.samples = np.asarray(samples, dtype=np.float32)
            self.sample_rate = sample_rate

        @property
        def duration(self):
            return self.samples.shape[0] / float(self.sample_rate)

        @classmethod
        def from_ndarray(cls, data, sample_rate=16000):
            return cls(data, sample_rate)

        @classmethod
        def from_file(cls, path):
            with wave.open(path, 'rb') as w:
                assert w.getsampwidth() == 2 and w.getnchannels() == 1
                pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
                return cls(pcm.astype(np.float32) / 32768.0, w.getframerate())

        def resample(self, sr):
            raise NotImplementedError

        def crop(self, duration, mode='eval'):
            # yeaudio is absent here; eval-mode crop keeps the leading `duration` seconds (train mode is random)
            assert mode != 'train'
            self.samples = self.samples[:int(duration * self.sample_rate)]

        def normalize(self, target_db=-20, max_gain_db=300.0):
            rms_db = 10.0 * np.log10(np.mean(self.samples.astype(np.float64) ** 2))
            gain = target_db - rms_db
            self.samples = (self.samples * (10.0 ** (gain / 20.0))).astype(np.float32)

    mod = types.ModuleType('yeaudio')
    sub = types.ModuleType('yeaudio.audio')
    sub.AudioSegment = AudioSegment
    mod.audio = sub
    sys.modules['yeaudio'] = mod
    sys.modules['yeaudio.audio'] = sub
    # import-only stand-ins for what mvector/trainer.py and reader.py pull in but the eval path never calls
    aug = types.ModuleType('yeaudio.augmentation')
    for nm in ('ReverbPerturbAugmentor', 'SpecAugmentor', 'SpeedPerturbAugmentor', 'VolumePerturbAugmentor',
               'NoisePerturbAugmentor'):
        setattr(aug, nm, type(nm, (), {}))
    sys.modules['yeaudio.augmentation'] = aug
    ti = types.ModuleType('torchinfo')
    ti.summary = lambda *a, **k: None
    sys.modules['torchinfo'] = ti
    vd = types.ModuleType('visualdl')
    vd.LogWriter = type('LogWriter', (), {})
    sys.modules['visualdl'] = vd
    return AudioSegment


FBANK80 = dict(feature_method='Fbank', method_args=dict(sample_frequency=16000, num_mel_bins=80))
FBANK24 = dict(feature_method='Fbank', method_args=dict(sample_frequency=16000, num_mel_bins=24))
MEL16 = dict(feature_method='MelSpectrogram',
             method_args=dict(sample_rate=16000, n_fft=512, win_length=512, hop_length=160, f_min=50.0,
                              f_max=7600.0, n_mels=16))
MEL64 = dict(feature_method='MelSpectrogram',
             method_args=dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=320, f_min=50.0,
                              f_max=14000.0, n_mels=64))

SPEC400 = dict(feature_method='Spectrogram', method_args=dict())                      # torchaudio defaults: n_fft 400 -> 201 bins
MFCC24 = dict(feature_method='MFCC', method_args=dict(n_mfcc=24, melkwargs=dict(n_fft=512, hop_length=160, n_mels=64,
                                                                                 f_min=20.0)))
MFCC40 = dict(feature_method='MFCC', method_args=dict())                              # defaults: n_fft 400, 128 mels, 40 coeffs

# name -> (model, model_args, preprocess, lengths in samples)
CASES = {
    'ecapa_small': ('EcapaTdnn', dict(embd_dim=32, pooling_type='ASP', channels=[64, 64, 64, 64, 192],
                                      attention_channels=32, res2net_scale=4, se_channels=16), FBANK80,
                    [12000, 12000, 7360]),
    'tdnn_small': ('TDNN', dict(embd_dim=32, channels=64, pooling_type='ASP'), FBANK80, [9600, 8000]),
    'campplus_small': ('CAMPPlus', dict(embd_dim=32, growth_rate=8, bn_size=4, init_channels=32), FBANK24,
                       [35200, 20000]),
    'resnetse_small': ('ResNetSE', dict(embd_dim=32, layers=[1, 2, 1, 1], num_filters=[16, 16, 32, 32],
                                        pooling_type='ASP'), MEL16, [16000, 11000]),
    'eres2net_small': ('ERes2Net', dict(embd_dim=32, num_blocks=[1, 1, 2, 1], m_channels=8), FBANK24,
                       [14400, 9000]),
    'ecapa_sap_small': ('EcapaTdnn', dict(embd_dim=32, pooling_type='SAP', channels=[64, 64, 64, 64, 192],
                                          res2net_scale=4, se_channels=16), FBANK80, [9000, 6400]),
    'tdnn_tsp_small': ('TDNN', dict(embd_dim=32, channels=64, pooling_type='TSP'), FBANK24, [9600, 8000]),
    'resnetse_tap_small': ('ResNetSE', dict(embd_dim=32, layers=[1, 1, 1, 1], num_filters=[16, 16, 32, 32],
                                            pooling_type='TAP'), MEL16, [12000, 11000]),
    'res2net_small': ('Res2Net', dict(embd_dim=32, m_channels=8, layers=[1, 2, 1, 1], base_width=32, scale=2,
                                      pooling_type='ASP'), MEL64, [16000, 12000]),
    'eres2net_wide_small': ('ERes2Net', dict(embd_dim=32, num_blocks=[1, 1, 1, 1], m_channels=8, mul_channel=2,
                                             expansion=4, base_width=32, scale=3), FBANK24, [12000]),
    # base_width 26 -> group widths 3, 6, 13, 26: none a multiple of 4, exercises the mirror's channel padding
    'eres2netv2_small': ('ERes2NetV2', dict(embd_dim=32, num_blocks=[1, 1, 2, 1], m_channels=8, base_width=26, scale=2),
                         FBANK24, [14400, 9000]),
    # Spectrogram / MFCC front-ends (featurizer.py:43-46); 201 input features exercise the 1-D models' input padding,
    # n_fft = 400 the radix-5 FFT passes, the ragged MFCC batch the call-wide top_db clamp
    'tdnn_spec_small': ('TDNN', dict(embd_dim=32, channels=64, pooling_type='ASP'), SPEC400, [9600, 8000]),
    'resnetse_mfcc_small': ('ResNetSE', dict(embd_dim=32, layers=[1, 1, 1, 1], num_filters=[16, 16, 32, 32],
                                             pooling_type='ASP'), MFCC24, [12000, 11000, 5600]),
    'ecapa_mfcc400_small': ('EcapaTdnn', dict(embd_dim=32, pooling_type='ASP', channels=[64, 64, 64, 64, 192],
                                              attention_channels=32, res2net_scale=4, se_channels=16), MFCC40,
                            [9000, 6400]),
}


def base_config(model, model_args, prep):
    return {
        'dataset_conf': {'dataset': {'min_duration': 0.3, 'max_duration': 3, 'sample_rate': 16000,
                                     'use_dB_normalization': False, 'target_dB': -20},
                         'eval_conf': {'batch_size': 16, 'max_duration': 20}},
        'preprocess_conf': {'use_hf_model': False, 'feature_method': prep['feature_method'],
                            'method_args': dict(prep['method_args'])},
        'model_conf': {'model': model, 'model_args': dict(model_args)},
    }


def synth_wave(n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    w = torch.randn(n, generator=g) * 0.1 + 0.05 * torch.sin(2 * np.pi * 220.0 * t) \
        + 0.05 * torch.sin(2 * np.pi * 1000.0 * t)
    return w.numpy().astype(np.float32)


def main():
    install_yeaudio_stub()
    from mvector.predict import MVectorPredictor
    from mvector.data_utils.featurizer import AudioFeaturizer
    from mvector.models import build_model
    from mvector.utils.utils import dict_to_object
    from oracle import models as om

    manifest = {}
    for ci, (name, (model, margs, prep, lens)) in enumerate(CASES.items()):
        cfg = base_config(model, margs, prep)
        fdim = AudioFeaturizer(**prep).feature_dim
        sd = om.random_state_dict(model, fdim, seed=100 + ci, **margs)
        # the reference's own shapes must agree with the oracle's enumeration
        ref_model = torch.nn.Sequential(build_model(fdim, dict_to_object(cfg)))
        ref_sd = ref_model.state_dict()
        assert [k[2:] for k in ref_sd] == list(sd.keys())
        with tempfile.TemporaryDirectory() as td:
            torch.save({'0.' + k: v for k, v in sd.items()}, os.path.join(td, 'model.pth'))
            pred = MVectorPredictor(configs=cfg, model_path=td, use_gpu=False)
        waves = [synth_wave(n, 1000 * ci + i) for i, n in enumerate(lens)]
        emb = pred.predict_batch(waves)                                         # predict.py:231-265
        emb_single = pred.predict(waves[-1])                                    # predict.py:214-229
        # features exactly as predict_batch computes them (predict.py:244-258)
        lmax = max(lens)
        x = np.zeros((len(lens), lmax), dtype=np.float32)
        for i, w in enumerate(waves):
            x[i, :len(w)] = w
        ratio = torch.tensor([n / lmax for n in lens], dtype=torch.float32)
        feats = pred._audio_featurizer(torch.tensor(x), ratio).numpy()
        out = {'wave%d' % i: w for i, w in enumerate(waves)}
        out.update(feats=feats.astype(np.float32), emb=np.asarray(emb, dtype=np.float32),
                   emb_single_last=np.asarray(emb_single, dtype=np.float32))
        out.update({'sd/' + k: v.numpy() for k, v in sd.items()})
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        manifest[name] = dict(model=model, model_args=margs, preprocess=prep, lens=lens, feature_dim=fdim,
                              seed=100 + ci)
        print(name, feats.shape, emb.shape, os.path.getsize(os.path.join(HERE, name + '.npz')) // 1024, 'KiB')

    # ---- config #1: TDNN + Fbank on the reference's own wav fixtures via the infer_contrast.py flow ----
    import yaml
    with open(os.path.join(REF, 'configs/tdnn.yml'), 'r', encoding='utf-8') as f:
        cfg = yaml.load(f.read(), Loader=yaml.FullLoader)
    margs = cfg['model_conf']['model_args']
    sd = om.random_state_dict('TDNN', 80, seed=7, **margs)
    with tempfile.TemporaryDirectory() as td:
        torch.save({'0.' + k: v for k, v in sd.items()}, os.path.join(td, 'model.pth'))
        pred = MVectorPredictor(configs=os.path.join(REF, 'configs/tdnn.yml'), model_path=td, use_gpu=False)
    a1, a2 = os.path.join(REF, 'dataset/a_1.wav'), os.path.join(REF, 'dataset/a_2.wav')
    e1, e2 = pred.predict(a1), pred.predict(a2)
    sim = float(pred.contrast(a1, a2))                                          # infer_contrast.py:19-23
    pcm = {}
    for nm, p in (('a_1', a1), ('a_2', a2)):
        with wave.open(p, 'rb') as w:
            pcm[nm] = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).copy()
    np.savez_compressed(os.path.join(HERE, 'c1_tdnn_contrast.npz'), pcm_a_1=pcm['a_1'], pcm_a_2=pcm['a_2'],
                        emb_a_1=e1.astype(np.float32), emb_a_2=e2.astype(np.float32),
                        sim=np.float32(sim))
    manifest['c1_tdnn_contrast'] = dict(model='TDNN', model_args=margs, seed=7, config='configs/tdnn.yml',
                                        note='weights = oracle.models.random_state_dict(TDNN, 80, seed=7); '
                                             'dB-normalised to -20 dB by the yeaudio stub')
    print('c1 sim', sim)

    # ---- evaluate caller (trainer.py:403-485) + metrics (metric/metrics.py) on a tiny 3-speaker wav set ----
    import mvector.trainer as ref_trainer
    eargs = dict(embd_dim=32, channels=64, pooling_type='ASP')
    esd = om.random_state_dict('TDNN', 80, seed=21, **eargs)
    rng = np.random.RandomState(5)
    spk_f = [(180.0, 900.0), (260.0, 1500.0), (330.0, 2300.0)]

    def spk_wave(spk, n, seed):
        t = np.arange(n, dtype=np.float64) / 16000.0
        f1, f2 = spk_f[spk]
        w = 0.2 * np.sin(2 * np.pi * f1 * t) + 0.15 * np.sin(2 * np.pi * f2 * t) \
            + 0.05 * np.random.RandomState(seed).randn(n)
        return np.clip(w * 20000, -32767, 32767).astype(np.int16)

    ev = {}
    with tempfile.TemporaryDirectory() as td:
        lists = {}
        for nm, count in (('enroll', 6), ('trials', 9)):
            lines = []
            for i in range(count):
                spk = i % 3
                n = int(rng.randint(6400, 24000))            # 0.4 .. 1.5 s; eval max_duration 1.2 s crops the longest
                pcm = spk_wave(spk, n, 50 * len(lines) + (0 if nm == 'enroll' else 1000) + i)
                path = os.path.join(td, f'{nm}_{i}.wav')
                with wave.open(path, 'wb') as w:
                    w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
                ev[f'{nm}_pcm{i}'] = pcm
                ev[f'{nm}_label{i}'] = np.int32(spk)
                lines.append(f'{path}\t{spk}\n')
            lists[nm] = os.path.join(td, f'{nm}_list.txt')
            with open(lists[nm], 'w') as f:
                f.writelines(lines)
        ecfg = base_config('TDNN', eargs, FBANK80)
        ecfg['dataset_conf'].update(enroll_list=lists['enroll'], trials_list=lists['trials'],
                                    dataLoader={'num_workers': 0})
        ecfg['dataset_conf']['eval_conf'] = {'batch_size': 4, 'max_duration': 1.2}
        ecfg['dataset_conf']['dataset']['use_dB_normalization'] = True
        ecfg['train_conf'] = {'use_compile': False}
        mdir = os.path.join(td, 'model')
        os.makedirs(mdir)
        torch.save({'0.' + k: v for k, v in esd.items()}, os.path.join(mdir, 'model.pth'))
        captured = {}
        real_fnr_fpr = ref_trainer.compute_fnr_fpr

        def spy(scores, labels, weights=None):
            captured['scores'], captured['labels'] = scores.copy(), labels.copy()
            out = real_fnr_fpr(scores, labels, weights)
            captured['fnr'], captured['fpr'], captured['thresholds'] = out
            return out

        ref_trainer.compute_fnr_fpr = spy
        tr = ref_trainer.MVectorTrainer(configs=ecfg, use_gpu=False)
        eer, min_dcf, thr = tr.evaluate(resume_model=mdir)
        ref_trainer.compute_fnr_fpr = real_fnr_fpr
    ev.update(eer=np.float64(eer), min_dcf=np.float64(min_dcf), threshold=np.float64(thr),
              scores=captured['scores'], labels=captured['labels'], fnr=captured['fnr'], fpr=captured['fpr'],
              thresholds=captured['thresholds'])
    ev.update({'sd/' + k: v.numpy() for k, v in esd.items()})
    np.savez_compressed(os.path.join(HERE, 'evaluate_small.npz'), **ev)
    manifest['evaluate_small'] = dict(model='TDNN', model_args=eargs, preprocess=FBANK80, seed=21, n_enroll=6, n_trials=9,
                                      eval_conf={'batch_size': 4, 'max_duration': 1.2},
                                      note='reference MVectorTrainer.evaluate on CPU; yeaudio stubbed (normalize to '
                                           '-20 dB, eval crop = leading max_duration seconds)')
    print('evaluate', eer, min_dcf, thr)
    # metrics alone, on a larger random score list with ties
    from mvector.metric.metrics import compute_fnr_fpr, compute_eer, compute_dcf
    mr = np.random.RandomState(11)
    lab = (mr.rand(4000) < 0.2).astype(np.int32)
    sc = np.round((mr.randn(4000) * 0.25 + lab * 0.45).astype(np.float32), 3)
    fnr, fpr, th = compute_fnr_fpr(sc, lab)
    m_eer, m_thr = compute_eer(fnr, fpr, sc)
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), scores=sc, labels=lab, fnr=fnr, fpr=fpr, thresholds=th,
                        eer=np.float64(m_eer), threshold=np.float64(m_thr), min_dcf=np.float64(compute_dcf(fnr, fpr)),
                        min_dcf_05=np.float64(compute_dcf(fnr, fpr, p_target=0.05, c_miss=10, c_fa=1)))
    manifest['metrics'] = dict(n=4000, note='mvector.metric.metrics on seeded random scores (rounded: ties present)')

    # ---- diarization glue (infer_utils/speaker_diarization.py): chunking, spectral clustering, post-processing ----
    from mvector.infer_utils.speaker_diarization import SpeakerDiarization as RefSD
    dr = np.random.RandomState(3)
    sd_ref = RefSD()
    sr = 16000
    vad_segments = []
    t0 = 0.0
    for dur in (4.1, 0.9, 6.35, 2.0, 3.77):                       # seconds of "speech", 0.4 s gaps
        n = int((t0 + dur) * sr) - int(t0 * sr)
        vad_segments.append([round(t0, 3), round(t0 + dur, 3), dr.randn(n).astype(np.float32)])
        t0 = round(t0 + dur + 0.4, 3)
    for seg in vad_segments:                                       # the reference's own consistency rule
        seg[2] = seg[2][:int(seg[1] * sr) - int(seg[0] * sr)]
        if seg[2].shape[0] < int(seg[1] * sr) - int(seg[0] * sr):
            seg[2] = np.pad(seg[2], (0, int(seg[1] * sr) - int(seg[0] * sr) - seg[2].shape[0]))
    sd_ref._check_audio_list(vad_segments)
    chunks = sd_ref._chunk(vad_segments)
    # embeddings: 3 speakers taking turns over the chunks (+ noise), two of them close enough to test the cosine merge
    centres = dr.randn(3, 32).astype(np.float32)
    centres[2] = centres[1] + 0.35 * dr.randn(32).astype(np.float32)
    turn = np.array([(i // 5) % 3 for i in range(len(chunks))])
    emb = (centres[turn] + 0.15 * dr.randn(len(chunks), 32)).astype(np.float32)
    dz = {'n_vad': np.int32(len(vad_segments)), 'chunk_times': np.array([[c[0], c[1]] for c in chunks]),
          'chunk_sums': np.array([float(np.abs(c[2]).sum()) for c in chunks]), 'emb': emb}
    for i, seg in enumerate(vad_segments):
        dz[f'vad{i}_t'] = np.array([seg[0], seg[1]])
        dz[f'vad{i}_x'] = seg[2]
    for tag, k in (('auto', None), ('k2', 2), ('k3', 3)):
        np.random.seed(0)
        labels, cen = sd_ref.clustering(emb.copy(), speaker_num=k)
        out = sd_ref.postprocess([list(c) for c in chunks], labels)
        dz[f'labels_{tag}'] = np.asarray(labels, dtype=np.int64)
        dz[f'centres_{tag}'] = np.asarray(cen, dtype=np.float32)
        dz[f'out_{tag}'] = np.array([[o['speaker'], o['start'], o['end']] for o in out], dtype=np.float64)
        print('diarization', tag, labels.max() + 1, len(out))
    np.savez_compressed(os.path.join(HERE, 'diarization.npz'), **dz)
    manifest['diarization'] = dict(note='reference SpeakerDiarization._chunk / clustering / postprocess on seeded synthetic '
                                        'VAD segments and 32-d chunk embeddings (np.random.seed(0) before every k_means)')

    # ---- default-config parameter-name/shape digests, for checking oracle.param_shapes on the GPU box ----
    defaults = {
        'EcapaTdnn': (80, dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536])),
        'TDNN': (80, dict(embd_dim=192, channels=512, pooling_type='ASP')),
        'CAMPPlus': (80, dict(embd_dim=192)),
        'ResNetSE': (64, dict(embd_dim=192, pooling_type='ASP')),
        'ERes2Net': (80, dict(embd_dim=192, m_channels=32)),
        'ERes2Net55M': (80, dict(embd_dim=192, m_channels=64, mul_channel=2, expansion=4, base_width=24, scale=3)),
        'Res2Net': (80, dict(embd_dim=192, pooling_type='ASP', m_channels=32)),
        'ERes2NetV2': (80, dict(embd_dim=192, m_channels=32)),
    }
    digests = {}
    for key, (fdim, margs) in defaults.items():
        model = 'ERes2Net' if key in ('ERes2Net', 'ERes2Net55M') else key
        cfg = dict_to_object({'model_conf': {'model': model, 'model_args': margs}})
        rsd = build_model(fdim, cfg).state_dict()
        s = ';'.join(f'{k}:{tuple(v.shape)}' for k, v in rsd.items())
        digests[key] = dict(input_size=fdim, model_args=margs, n_tensors=len(rsd),
                            n_params=int(sum(v.numel() for k, v in rsd.items() if 'num_batches' not in k)),
                            sha256=hashlib.sha256(s.encode()).hexdigest())
    manifest['_param_digests'] = digests
    with open(os.path.join(HERE, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
