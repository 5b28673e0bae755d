"""Oracle front-end: Kaldi Fbank / MelSpectrogram / CMN+mask, torch fp32 on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in this repo's own words:
  * torchaudio.compliance.kaldi.fbank            (kaldi.py:514-645; framing :44-83, window :154-217,
                                                  mel banks :436-511)
  * torchaudio.transforms.MelSpectrogram          (functional.py:52-144 spectrogram, :518-587 melscale_fbanks)
  * torchaudio.transforms.Spectrogram / MFCC      (transforms: Spectrogram.forward, MFCC.forward; functional.py:352-403
                                                  amplitude_to_DB, :640-667 create_dct)
  * mvector AudioFeaturizer.forward               (mvector/data_utils/featurizer.py:53-91)
  * mvector KaldiFbank.forward per-utterance loop (mvector/data_utils/featurizer.py:119-132)
"""
import math

import numpy as np
import torch

_FBANK_KEYS = dict(blackman_coeff=0.42, channel=-1, dither=0.0, energy_floor=1.0, frame_length=25.0,
                   frame_shift=10.0, high_freq=0.0, htk_compat=False, low_freq=20.0, min_duration=0.0,
                   num_mel_bins=23, preemphasis_coefficient=0.97, raw_energy=True, remove_dc_offset=True,
                   round_to_power_of_two=True, sample_frequency=16000.0, snip_edges=True,
                   subtract_mean=False, use_energy=False, use_log_fbank=True, use_power=True,
                   vtln_high=-500.0, vtln_low=100.0, vtln_warp=1.0, window_type='povey')

F32_EPS = float(np.finfo(np.float32).eps)  # kaldi.py:_get_epsilon -> torch.finfo(float32).eps


def fbank_args(**kwargs):
    """Fill defaults exactly like the keyword signature of kaldi.fbank (kaldi.py:514-541).
    Unknown keys raise TypeError, as ``Kaldi.fbank(waveform, **self.kwargs)`` would (featurizer.py:128)."""
    for k in kwargs:
        if k not in _FBANK_KEYS:
            raise TypeError(f"fbank() got an unexpected keyword argument '{k}'")
    a = dict(_FBANK_KEYS)
    a.update(kwargs)
    return a


def frame_geometry(sample_frequency, frame_shift, frame_length, round_to_power_of_two=True):
    """kaldi.py:126-150: shift/size in samples and padded FFT size."""
    shift = int(sample_frequency * frame_shift * 0.001)
    size = int(sample_frequency * frame_length * 0.001)
    padded = (1 if size == 0 else 2 ** (size - 1).bit_length()) if round_to_power_of_two else size
    return shift, size, padded


def num_frames(num_samples, size, shift):
    """kaldi.py:63-67 (snip_edges=True)."""
    if num_samples < size:
        return 0
    return 1 + (num_samples - size) // shift


def feature_window(window_type, size, blackman_coeff=0.42):
    """kaldi.py:86-113."""
    if window_type == 'hanning':
        return torch.hann_window(size, periodic=False, dtype=torch.float32)
    if window_type == 'hamming':
        return torch.hamming_window(size, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    if window_type == 'povey':
        return torch.hann_window(size, periodic=False, dtype=torch.float32).pow(0.85)
    if window_type == 'rectangular':
        return torch.ones(size, dtype=torch.float32)
    if window_type == 'blackman':
        a = 2 * math.pi / (size - 1)
        n = torch.arange(size, dtype=torch.float32)
        return blackman_coeff - 0.5 * torch.cos(a * n) + (0.5 - blackman_coeff) * torch.cos(2 * a * n)
    raise Exception('Invalid window type ' + window_type)


def kaldi_mel_banks(num_bins, padded_size, sample_freq, low_freq, high_freq):
    """kaldi.py:436-511 with vtln_warp == 1.  Returns [num_bins, padded_size//2] float32."""
    assert num_bins > 3 and padded_size % 2 == 0
    num_fft_bins = padded_size / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    assert 0.0 <= low_freq < nyquist and 0.0 < high_freq <= nyquist and low_freq < high_freq
    bin_width = sample_freq / padded_size
    mel_lo = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_hi - mel_lo) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down))


def _kaldi_fbank_fp32(waveform, **kwargs):
    """One utterance: waveform [L] (or [1, L]) float32 -> [m, num_mel_bins] (kaldi.py:514-645), fp32 like the reference.

    Supported subset: dither == 0, vtln_warp == 1, snip_edges True, use_energy False, subtract_mean False
    (everything the shipped configs use; other values raise NotImplementedError)."""
    a = fbank_args(**kwargs)
    if a['dither'] != 0.0 or a['vtln_warp'] != 1.0 or not a['snip_edges'] or a['use_energy'] \
            or a['subtract_mean'] or a['min_duration'] != 0.0:
        raise NotImplementedError('oracle fbank: unsupported option')
    w = torch.as_tensor(waveform, dtype=torch.float32)
    if w.dim() == 2:
        ch = max(a['channel'], 0)
        w = w[ch]
    shift, size, padded = frame_geometry(a['sample_frequency'], a['frame_shift'], a['frame_length'],
                                         a['round_to_power_of_two'])
    assert 2 <= size <= w.numel(), f'choose a window size {size} that is [2, {w.numel()}]'
    m = num_frames(w.numel(), size, shift)
    frames = w.as_strided((m, size), (shift, 1))                               # kaldi.py:82-83
    if a['remove_dc_offset']:
        frames = frames - frames.mean(dim=1, keepdim=True)                     # kaldi.py:183-186
    c = a['preemphasis_coefficient']
    if c != 0.0:
        prev = torch.cat([frames[:, :1], frames[:, :-1]], dim=1)               # replicate pad, kaldi.py:193-198
        frames = frames - c * prev
    frames = frames * feature_window(a['window_type'], size, a['blackman_coeff']).unsqueeze(0)
    if padded != size:
        frames = torch.nn.functional.pad(frames, (0, padded - size))           # kaldi.py:207-211
    spec = torch.fft.rfft(frames).abs()                                        # kaldi.py:616
    if a['use_power']:
        spec = spec.pow(2.0)
    banks = kaldi_mel_banks(a['num_mel_bins'], padded, a['sample_frequency'], a['low_freq'], a['high_freq'])
    banks = torch.nn.functional.pad(banks.to(torch.float32), (0, 1))           # last FFT bin weight 0, :627
    mel = torch.mm(spec, banks.T)                                              # kaldi.py:630
    if a['use_log_fbank']:
        mel = torch.max(mel, torch.tensor(F32_EPS)).log()                      # kaldi.py:633
    return mel


def kaldi_fbank(waveform, exact_spectrum=False, **kwargs):
    """One utterance: waveform [L] (or [1, L]) float32 -> [m, num_mel_bins] (kaldi.py:514-645).

    ``exact_spectrum=True`` evaluates the SAME formula on the SAME fp32 windowed frames, but the rFFT, the power spectrum
    and the mel projection in float64 (rounded to fp32 once, before the log): the exact value of what the reference's
    fp32 pipeline approximates.  Tests use it to measure the reference's own rounding error -- the floor below which no
    implementation can agree with the reference (tools/fbank_precision_study.py)."""
    if not exact_spectrum:
        return _kaldi_fbank_fp32(waveform, **kwargs)
    a = fbank_args(**kwargs)
    w = torch.as_tensor(waveform, dtype=torch.float32)
    if w.dim() == 2:
        w = w[max(a['channel'], 0)]
    shift, size, padded = frame_geometry(a['sample_frequency'], a['frame_shift'], a['frame_length'], a['round_to_power_of_two'])
    m = num_frames(w.numel(), size, shift)
    frames = w.as_strided((m, size), (shift, 1))
    if a['remove_dc_offset']:
        frames = frames - frames.mean(dim=1, keepdim=True)
    c = a['preemphasis_coefficient']
    if c != 0.0:
        frames = frames - c * torch.cat([frames[:, :1], frames[:, :-1]], dim=1)
    frames = frames * feature_window(a['window_type'], size, a['blackman_coeff']).unsqueeze(0)
    if padded != size:
        frames = torch.nn.functional.pad(frames, (0, padded - size))
    spec = torch.fft.rfft(frames.double()).abs()
    if a['use_power']:
        spec = spec.pow(2.0)
    banks = kaldi_mel_banks(a['num_mel_bins'], padded, a['sample_frequency'], a['low_freq'], a['high_freq'])
    banks = torch.nn.functional.pad(banks.to(torch.float32), (0, 1)).double()
    mel = torch.mm(spec, banks.T).float()
    if a['use_log_fbank']:
        mel = torch.max(mel, torch.tensor(F32_EPS)).log()
    return mel


def kaldi_fbank_batch(waveforms, exact_spectrum=False, **kwargs):
    """mvector KaldiFbank.forward (featurizer.py:119-132): per-utterance loop -> [B, F, T]."""
    outs = [kaldi_fbank(w, exact_spectrum=exact_spectrum, **kwargs).transpose(0, 1) for w in waveforms]
    return torch.stack(outs)


# ---------------------------------------------------------------------------------------------
# MelSpectrogram (torchaudio.transforms.MelSpectrogram as constructed by featurizer.py:41-42)
# ---------------------------------------------------------------------------------------------
_MELSPEC_KEYS = dict(sample_rate=16000, n_fft=400, win_length=None, hop_length=None, f_min=0.0, f_max=None,
                     pad=0, n_mels=128, power=2.0, normalized=False, center=True, pad_mode='reflect',
                     onesided=None, norm=None, mel_scale='htk')


def melspec_args(**kwargs):
    for k in kwargs:
        if k not in _MELSPEC_KEYS and k not in ('window_fn', 'wkwargs'):
            raise TypeError(f"MelSpectrogram.__init__() got an unexpected keyword argument '{k}'")
    if 'window_fn' in kwargs or 'wkwargs' in kwargs:
        raise NotImplementedError('oracle melspec: custom window_fn not supported')
    a = dict(_MELSPEC_KEYS)
    a.update(kwargs)
    if a['win_length'] is None:
        a['win_length'] = a['n_fft']
    if a['hop_length'] is None:
        a['hop_length'] = a['win_length'] // 2
    if a['f_max'] is None:
        a['f_max'] = float(a['sample_rate'] // 2)
    return a


def _hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + (f / 700.0))                               # functional.py:437-438


def htk_mel_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """functional.py:518-587 for mel_scale='htk', norm=None.  Returns [n_freqs, n_mels] float32."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)                           # functional.py:473-474
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))                       # functional.py:507-513


def _power_stft(w, n_fft, win, hop, power, pad=0):
    """functional.spectrogram (functional.py:52-144) for center=True, reflect, onesided, not normalized."""
    w = torch.as_tensor(w, dtype=torch.float32)
    if pad > 0:
        w = torch.nn.functional.pad(w, (pad, pad))
    window = torch.hann_window(win)                                             # periodic hann, transforms:62
    spec = torch.stft(w, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True,
                      pad_mode='reflect', normalized=False, onesided=True, return_complex=True)
    return spec.abs() if power == 1.0 else spec.abs().pow(power)                # functional.py:139-142


def mel_spectrogram(waveforms, **kwargs):
    """waveforms [B, L] -> [B, n_mels, T]  (no log: featurizer.py:76 uses the raw transform)."""
    a = melspec_args(**kwargs)
    if a['mel_scale'] != 'htk' or a['norm'] is not None or a['normalized'] or a['pad_mode'] != 'reflect' \
            or not a['center']:
        raise NotImplementedError('oracle melspec: unsupported option')
    spec = _power_stft(waveforms, a['n_fft'], a['win_length'], a['hop_length'], a['power'], a['pad'])
    fb = htk_mel_fbanks(a['n_fft'] // 2 + 1, a['f_min'], a['f_max'], a['n_mels'], a['sample_rate'])
    return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)          # MelScale.forward


_SPEC_KEYS = dict(n_fft=400, win_length=None, hop_length=None, pad=0, power=2.0, normalized=False, center=True,
                  pad_mode='reflect', onesided=True)


def spectrogram(waveforms, **kwargs):
    """torchaudio.transforms.Spectrogram as built by featurizer.py:43-44: waveforms [B, L] -> [B, n_fft//2+1, T]."""
    for k in kwargs:
        if k not in _SPEC_KEYS:
            raise TypeError(f"Spectrogram.__init__() got an unexpected keyword argument '{k}'")
    a = dict(_SPEC_KEYS)
    a.update(kwargs)
    win = a['win_length'] if a['win_length'] is not None else a['n_fft']
    hop = a['hop_length'] if a['hop_length'] is not None else win // 2
    if a['normalized'] or a['pad_mode'] != 'reflect' or not a['center'] or not a['onesided'] or a['power'] is None:
        raise NotImplementedError('oracle spectrogram: unsupported option')
    return _power_stft(waveforms, a['n_fft'], win, hop, a['power'], a['pad'])


_MFCC_KEYS = dict(sample_rate=16000, n_mfcc=40, dct_type=2, norm='ortho', log_mels=False, melkwargs=None)


def dct_matrix(n_mfcc, n_mels, norm):
    """functional.create_dct (functional.py:640-667): DCT-II basis, returned [n_mels, n_mfcc]."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    if norm is None:
        dct *= 2.0
    else:
        assert norm == 'ortho'
        dct[0] *= 1.0 / math.sqrt(2.0)
        dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t()


def power_to_db(x, top_db=80.0, amin=1e-10):
    """AmplitudeToDB('power', top_db) (functional.amplitude_to_DB, functional.py:352-403) with ref 1.0: 10*log10(max(x,
    amin)); the top_db clamp folds a 3-D input's leading (batch) axis into the channel axis, so ONE maximum is shared
    by the whole batch."""
    x_db = 10.0 * torch.log10(torch.clamp(x, min=amin))
    x_db -= 10.0 * math.log10(max(amin, 1.0))
    if top_db is not None:
        shape = x_db.size()
        packed = shape[-3] if x_db.dim() > 2 else 1
        x4 = x_db.reshape(-1, packed, shape[-2], shape[-1])
        x4 = torch.max(x4, (x4.amax(dim=(-3, -2, -1)) - top_db).view(-1, 1, 1, 1))
        x_db = x4.reshape(shape)
    return x_db


def mfcc(waveforms, **kwargs):
    """torchaudio.transforms.MFCC as built by featurizer.py:45-46: waveforms [B, L] -> [B, n_mfcc, T]."""
    for k in kwargs:
        if k not in _MFCC_KEYS:
            raise TypeError(f"MFCC.__init__() got an unexpected keyword argument '{k}'")
    a = dict(_MFCC_KEYS)
    a.update(kwargs)
    if a['dct_type'] != 2:
        raise ValueError('DCT type not supported: {}'.format(a['dct_type']))
    mk = dict(a['melkwargs'] or {})
    mel = mel_spectrogram(waveforms, sample_rate=a['sample_rate'], **mk)
    n_mels = mel.shape[-2]
    if a['n_mfcc'] > n_mels:
        raise ValueError('Cannot select more MFCC coefficients than # mel bins')
    mel = torch.log(mel + 1e-6) if a['log_mels'] else power_to_db(mel)
    return torch.matmul(mel.transpose(-1, -2), dct_matrix(a['n_mfcc'], n_mels, a['norm'])).transpose(-1, -2)


# ---------------------------------------------------------------------------------------------
# AudioFeaturizer.forward
# ---------------------------------------------------------------------------------------------
def featurize(waveforms, input_lens_ratio=None, feature_method='Fbank', method_args=None, exact_spectrum=False):
    """featurizer.py:53-91: feat [B,F,T] -> transpose -> subtract time-mean over ALL T frames -> zero
    frames t >= round(ratio*T).  Returns [B, T, F] float32.  ``exact_spectrum`` (Fbank only): see kaldi_fbank."""
    method_args = dict(method_args or {})
    w = torch.as_tensor(waveforms, dtype=torch.float32)
    if w.dim() == 1:
        w = w.unsqueeze(0)
    if feature_method == 'Fbank':
        feat = kaldi_fbank_batch(w, exact_spectrum=exact_spectrum, **method_args)
    elif feature_method == 'MelSpectrogram':
        feat = mel_spectrogram(w, **method_args)
    elif feature_method == 'Spectrogram':
        feat = spectrogram(w, **method_args)
    elif feature_method == 'MFCC':
        feat = mfcc(w, **method_args)
    else:
        raise Exception(f'预处理方法 {feature_method} 不存在!')
    feat = feat.transpose(2, 1)
    feat = feat - feat.mean(1, keepdim=True)
    if input_lens_ratio is not None:
        ratio = torch.as_tensor(input_lens_ratio, dtype=torch.float32)
        keep = torch.round(ratio * feat.shape[1]).long().unsqueeze(1)          # featurizer.py:82-84
        idx = torch.arange(feat.shape[1]).repeat(feat.shape[0], 1)
        feat = torch.where((idx < keep).unsqueeze(-1), feat, torch.zeros_like(feat))
    return feat


def feature_dim(feature_method, method_args):
    """featurizer.py:93-111."""
    method_args = method_args or {}
    if feature_method == 'MelSpectrogram':
        return method_args.get('n_mels', 128)
    if feature_method == 'Fbank':
        return method_args.get('num_mel_bins', 23)
    if feature_method == 'Spectrogram':
        return method_args.get('n_fft', 400) // 2 + 1
    if feature_method == 'MFCC':
        return method_args.get('n_mfcc', 40)
    raise Exception('没有{}预处理方法'.format(feature_method))


def pad_batch(waves):
    """predict.py:244-255: zero-pad to the longest item, ratio = len / Lmax (python float -> float32)."""
    lmax = max(len(w) for w in waves)
    x = np.zeros((len(waves), lmax), dtype=np.float32)
    ratio = []
    for i, w in enumerate(waves):
        x[i, :len(w)] = w
        ratio.append(len(w) / lmax)
    return x, np.asarray(ratio, dtype=np.float32)
