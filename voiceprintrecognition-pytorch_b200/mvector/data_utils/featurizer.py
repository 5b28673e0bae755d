"""AudioFeaturizer mirror (reference: mvector/data_utils/featurizer.py:9-132), backed by the fused sm_100a front-end.

Host side only prepares constants once (window, sparse mel bank, DCT matrix -- the reference rebuilds the Kaldi ones on
every call, kaldi.py:201,621-627) and hands device pointers to ``vp_fbank`` / ``vp_melspec`` / ``vp_mfcc``; the
per-utterance Python loop (featurizer.py:124-131), the transpose, the CMN and the length mask (featurizer.py:77-90) all
run in two kernels (three for MFCC).  All four ``feature_method`` values of the reference are lowered: Fbank,
MelSpectrogram, Spectrogram (pass-through bank over the n_fft/2+1 bins) and MFCC.
"""
import ctypes as C
import math

import numpy as np
import torch
from loguru import logger

from .. import _lib as L
from ..engine import Engine, _check

_FBANK_DEFAULTS = dict(blackman_coeff=0.42, channel=-1, dither=0.0, energy_floor=1.0, frame_length=25.0,
                       frame_shift=10.0, high_freq=0.0, htk_compat=False, low_freq=20.0, min_duration=0.0,
                       num_mel_bins=23, preemphasis_coefficient=0.97, raw_energy=True, remove_dc_offset=True,
                       round_to_power_of_two=True, sample_frequency=16000.0, snip_edges=True, subtract_mean=False,
                       use_energy=False, use_log_fbank=True, use_power=True, vtln_high=-500.0, vtln_low=100.0,
                       vtln_warp=1.0, window_type='povey')
_MELSPEC_DEFAULTS = dict(sample_rate=16000, n_fft=400, win_length=None, hop_length=None, f_min=0.0, f_max=None,
                         pad=0, n_mels=128, power=2.0, normalized=False, center=True, pad_mode='reflect',
                         onesided=None, norm=None, mel_scale='htk')
_SPEC_DEFAULTS = dict(n_fft=400, win_length=None, hop_length=None, pad=0, power=2.0, normalized=False, center=True,
                      pad_mode='reflect', onesided=True)
_MFCC_DEFAULTS = dict(sample_rate=16000, n_mfcc=40, dct_type=2, norm='ortho', log_mels=False, melkwargs=None)


def _fft_size_ok(n):
    """The front-end FFT handles N = 2^a 3^b 5^c, 4 | N, 64 <= N <= 2048 (torchaudio's default n_fft = 400 included)."""
    if n < 64 or n > 2048 or n % 4:
        return False
    for f in (2, 3, 5):
        while n % f == 0:
            n //= f
    return n == 1


def _stft_window(n_fft, win_length):
    """Periodic Hann of win_length taps, centred inside n_fft like torch.stft does for a short window."""
    win = torch.hann_window(win_length)
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        win = torch.nn.functional.pad(win, (left, n_fft - win_length - left))
    return win.numpy().astype(np.float32)


def _sparse_bank(dense):
    """dense [n_filters, n_bins] -> (start, count, offset, weights) over each filter's non-zero support."""
    start, count, off, w = [], [], [], []
    for row in dense:
        nz = np.nonzero(row)[0]
        if nz.size == 0:
            start.append(0); count.append(0); off.append(len(w))
            continue
        s, e = int(nz[0]), int(nz[-1]) + 1
        start.append(s); count.append(e - s); off.append(len(w))
        w.extend(row[s:e].tolist())
    return (np.asarray(start, np.int32), np.asarray(count, np.int32), np.asarray(off, np.int32),
            np.asarray(w if w else [0.0], np.float32))


class KaldiFbank:
    """kwargs of torchaudio.compliance.kaldi.fbank (featurizer.py:114-117); constants follow kaldi.py:86-113
    (window) and kaldi.py:436-511 (mel banks), evaluated once in fp32 with the same torch ops."""

    def __init__(self, **kwargs):
        for k in kwargs:
            if k not in _FBANK_DEFAULTS:
                raise TypeError(f"fbank() got an unexpected keyword argument '{k}'")
        a = dict(_FBANK_DEFAULTS)
        a.update(kwargs)
        self.kwargs = kwargs
        unsupported = [k for k, bad in (('dither', a['dither'] != 0.0), ('vtln_warp', a['vtln_warp'] != 1.0),
                                        ('snip_edges', not a['snip_edges']), ('use_energy', a['use_energy']),
                                        ('subtract_mean', a['subtract_mean']),
                                        ('min_duration', a['min_duration'] != 0.0),
                                        ('round_to_power_of_two', not a['round_to_power_of_two']),
                                        ('channel', a['channel'] not in (-1, 0))) if bad]
        if unsupported:
            raise NotImplementedError('Fbank options not lowered to the sm_100a front-end: ' + ', '.join(unsupported))
        sf = a['sample_frequency']
        self.hop = int(sf * a['frame_shift'] * 0.001)
        self.win_length = int(sf * a['frame_length'] * 0.001)
        self.n_fft = 1 if self.win_length == 0 else 2 ** (self.win_length - 1).bit_length()
        self.n_mels = a['num_mel_bins']
        assert self.n_mels > 3, 'Must have at least 3 mel bins'
        wt = a['window_type']
        n = self.win_length
        if wt == 'povey':
            win = torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85)
        elif wt == 'hanning':
            win = torch.hann_window(n, periodic=False, dtype=torch.float32)
        elif wt == 'hamming':
            win = torch.hamming_window(n, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
        elif wt == 'rectangular':
            win = torch.ones(n, dtype=torch.float32)
        elif wt == 'blackman':
            c = 2 * math.pi / (n - 1)
            i = torch.arange(n, dtype=torch.float32)
            win = a['blackman_coeff'] - 0.5 * torch.cos(c * i) + (0.5 - a['blackman_coeff']) * torch.cos(2 * c * i)
        else:
            raise Exception('Invalid window type ' + wt)
        self.window = win.numpy().astype(np.float32)
        # mel banks
        nyq = 0.5 * sf
        lo, hi = a['low_freq'], a['high_freq']
        if hi <= 0.0:
            hi += nyq
        assert 0.0 <= lo < nyq and 0.0 < hi <= nyq and lo < hi, 'Bad values in options: low-freq / high-freq'
        bw = sf / self.n_fft
        mlo, mhi = 1127.0 * math.log(1.0 + lo / 700.0), 1127.0 * math.log(1.0 + hi / 700.0)
        delta = (mhi - mlo) / (self.n_mels + 1)
        b = torch.arange(self.n_mels).unsqueeze(1)
        left, center, right = mlo + b * delta, mlo + (b + 1.0) * delta, mlo + (b + 2.0) * delta
        mel = (1127.0 * (1.0 + (bw * torch.arange(self.n_fft / 2)) / 700.0).log()).unsqueeze(0)
        banks = torch.max(torch.zeros(1), torch.min((mel - left) / (center - left), (right - mel) / (right - center)))
        self.bank = _sparse_bank(banks.to(torch.float32).numpy())   # bin n_fft/2 has weight 0 (kaldi.py:627)
        self.desc = L.FrontendDesc(kind=0, n_fft=self.n_fft, win_length=self.win_length, hop=self.hop,
                                   n_mels=self.n_mels, remove_dc=1 if a['remove_dc_offset'] else 0,
                                   preemph=float(a['preemphasis_coefficient']), power=2 if a['use_power'] else 1,
                                   use_log=1 if a['use_log_fbank'] else 0, log_floor=float(np.finfo(np.float32).eps))


class MelSpectrogram:
    """kwargs of torchaudio.transforms.MelSpectrogram (featurizer.py:41-42): periodic Hann, centred reflect-padded
    STFT (functional.py:123-135), |X|^power, HTK triangular bank (functional.py:518-587).  No log (featurizer.py:76)."""

    def __init__(self, **kwargs):
        for k in kwargs:
            if k not in _MELSPEC_DEFAULTS and k not in ('window_fn', 'wkwargs'):
                raise TypeError(f"MelSpectrogram.__init__() got an unexpected keyword argument '{k}'")
        a = dict(_MELSPEC_DEFAULTS)
        a.update(kwargs)
        self.kwargs = kwargs
        n_fft = a['n_fft']
        win_length = a['win_length'] if a['win_length'] is not None else n_fft
        hop = a['hop_length'] if a['hop_length'] is not None else win_length // 2
        f_max = a['f_max'] if a['f_max'] is not None else float(a['sample_rate'] // 2)
        bad = [k for k, b in (('window_fn', 'window_fn' in kwargs or 'wkwargs' in kwargs), ('pad', a['pad'] != 0),
                              ('normalized', bool(a['normalized'])), ('center', not a['center']),
                              ('pad_mode', a['pad_mode'] != 'reflect'), ('norm', a['norm'] is not None),
                              ('mel_scale', a['mel_scale'] != 'htk'), ('power', a['power'] not in (1.0, 2.0, 1, 2)),
                              ('n_fft (2^a 3^b 5^c, multiple of 4, in [64, 2048])', not _fft_size_ok(n_fft)),
                              ('n_mels (<= 128)', a['n_mels'] > 128),
                              ('win_length', win_length > n_fft)) if b]
        if bad:
            raise NotImplementedError('MelSpectrogram options not lowered to the sm_100a front-end: ' + ', '.join(bad))
        self.n_fft, self.hop, self.n_mels = n_fft, hop, a['n_mels']
        self.window = _stft_window(n_fft, win_length)
        self.win_length = n_fft
        n_freqs = n_fft // 2 + 1
        all_freqs = torch.linspace(0, a['sample_rate'] // 2, n_freqs)
        m_min = 2595.0 * math.log10(1.0 + (a['f_min'] / 700.0))
        m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
        f_pts = 700.0 * (10.0 ** (torch.linspace(m_min, m_max, self.n_mels + 2) / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
        self.bank = _sparse_bank(fb.T.contiguous().numpy())
        self.desc = L.FrontendDesc(kind=1, n_fft=n_fft, win_length=n_fft, hop=hop, n_mels=self.n_mels, remove_dc=0,
                                   preemph=0.0, power=int(a['power']), use_log=0, log_floor=0.0)
        self.dct = None
        self.n_out = self.n_mels


class Spectrogram:
    """kwargs of torchaudio.transforms.Spectrogram (featurizer.py:43-44): the MelSpectrogram pipeline without the mel
    projection -- the "bank" is the identity over the n_fft/2+1 bins, so the same kernel emits |X|^power directly."""

    def __init__(self, **kwargs):
        for k in kwargs:
            if k not in _SPEC_DEFAULTS and k not in ('window_fn', 'wkwargs'):
                raise TypeError(f"Spectrogram.__init__() got an unexpected keyword argument '{k}'")
        a = dict(_SPEC_DEFAULTS)
        a.update(kwargs)
        self.kwargs = kwargs
        n_fft = a['n_fft']
        win_length = a['win_length'] if a['win_length'] is not None else n_fft
        hop = a['hop_length'] if a['hop_length'] is not None else win_length // 2
        bad = [k for k, b in (('window_fn', 'window_fn' in kwargs or 'wkwargs' in kwargs), ('pad', a['pad'] != 0),
                              ('normalized', bool(a['normalized'])), ('center', not a['center']),
                              ('pad_mode', a['pad_mode'] != 'reflect'), ('onesided', not a['onesided']),
                              ('power', a['power'] not in (1.0, 2.0, 1, 2)),
                              ('n_fft (2^a 3^b 5^c, multiple of 4, in [64, 2048])', not _fft_size_ok(n_fft)),
                              ('win_length', win_length > n_fft)) if b]
        if bad:
            raise NotImplementedError('Spectrogram options not lowered to the sm_100a front-end: ' + ', '.join(bad))
        nb = n_fft // 2 + 1
        self.n_fft, self.hop, self.n_mels = n_fft, hop, nb
        self.window = _stft_window(n_fft, win_length)
        self.win_length = n_fft
        idx = np.arange(nb, dtype=np.int32)
        self.bank = (idx, np.ones(nb, np.int32), idx.copy(), np.ones(nb, np.float32))
        self.desc = L.FrontendDesc(kind=1, n_fft=n_fft, win_length=n_fft, hop=hop, n_mels=nb, remove_dc=0, preemph=0.0,
                                   power=int(a['power']), use_log=0, log_floor=0.0)
        self.dct = None
        self.n_out = nb


class MFCC:
    """kwargs of torchaudio.transforms.MFCC (featurizer.py:45-46): MelSpectrogram(sample_rate, **melkwargs) ->
    AmplitudeToDB('power', top_db=80) (or log(mel + 1e-6) when log_mels) -> DCT-II (create_dct, functional.py:640-667)."""

    def __init__(self, **kwargs):
        for k in kwargs:
            if k not in _MFCC_DEFAULTS:
                raise TypeError(f"MFCC.__init__() got an unexpected keyword argument '{k}'")
        a = dict(_MFCC_DEFAULTS)
        a.update(kwargs)
        self.kwargs = kwargs
        if a['dct_type'] != 2:
            raise ValueError('DCT type not supported: {}'.format(a['dct_type']))
        mel = MelSpectrogram(sample_rate=a['sample_rate'], **(a['melkwargs'] or {}))
        n_mfcc, n_mels = a['n_mfcc'], mel.n_mels
        if n_mfcc > n_mels:
            raise ValueError('Cannot select more MFCC coefficients than # mel bins')
        self.n_fft, self.hop, self.n_mels, self.win_length = mel.n_fft, mel.hop, n_mels, mel.win_length
        self.window, self.bank = mel.window, mel.bank
        n = torch.arange(float(n_mels))
        k = torch.arange(float(n_mfcc)).unsqueeze(1)
        dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
        if a['norm'] is None:
            dct *= 2.0
        else:
            assert a['norm'] == 'ortho'
            dct[0] *= 1.0 / math.sqrt(2.0)
            dct *= math.sqrt(2.0 / float(n_mels))
        self.dct = np.ascontiguousarray(dct.t().numpy(), dtype=np.float32)          # [n_mels, n_mfcc]
        self.n_out = n_mfcc
        if a['log_mels']:
            self.desc = L.FrontendDesc(kind=1, n_fft=mel.n_fft, win_length=mel.n_fft, hop=mel.hop, n_mels=n_mels,
                                       remove_dc=0, preemph=0.0, power=mel.desc.power, use_log=3, log_floor=1e-6,
                                       post=1, n_out=n_mfcc, db_mult=0.0, top_db=-1.0)
        else:
            self.desc = L.FrontendDesc(kind=1, n_fft=mel.n_fft, win_length=mel.n_fft, hop=mel.hop, n_mels=n_mels,
                                       remove_dc=0, preemph=0.0, power=mel.desc.power, use_log=2, log_floor=1e-10,
                                       post=1, n_out=n_mfcc, db_mult=10.0, top_db=80.0)


class AudioFeaturizer:
    """音频特征器 (drop-in for mvector.data_utils.featurizer.AudioFeaturizer).

    :param feature_method: 'Fbank' | 'MelSpectrogram' | 'Spectrogram' | 'MFCC'  (HF models: not lowered, raise)
    :param method_args: forwarded as **kwargs exactly like the reference (unknown keys -> TypeError)
    """

    def __init__(self, feature_method='MelSpectrogram', use_hf_model=False, method_args={}, engine=None):
        self._method_args = method_args
        self._feature_method = feature_method
        self.use_hf_model = use_hf_model
        if use_hf_model:
            raise NotImplementedError('HF wav2vec-style feature models are outside the lowered path (SURVEY.md 2)')
        if feature_method == 'MelSpectrogram':
            self.feat_fun = MelSpectrogram(**method_args)
        elif feature_method == 'Fbank':
            self.feat_fun = KaldiFbank(**method_args)
        elif feature_method == 'Spectrogram':
            self.feat_fun = Spectrogram(**method_args)
        elif feature_method == 'MFCC':
            self.feat_fun = MFCC(**method_args)
        else:
            raise Exception(f'预处理方法 {self._feature_method} 不存在!')
        self._engine = engine
        self._configured = False
        logger.info(f'使用【{feature_method}】提取特征')

    # -- lazy device state so that constructing the object (config parsing) needs no GPU --
    def _ensure(self):
        if self._configured:
            return
        if self._engine is None:
            self._engine = Engine()
        f = self.feat_fun
        start, count, off, w = f.bank
        win = np.ascontiguousarray(f.window, dtype=np.float32)
        dct = getattr(f, 'dct', None)
        _check(self._engine.handle, L.lib().vp_frontend_set(
            self._engine.handle, C.byref(f.desc), win.ctypes.data_as(C.c_void_p), start.ctypes.data_as(C.c_void_p),
            count.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), int(w.size),
            dct.ctypes.data_as(C.c_void_p) if dct is not None else C.c_void_p()))
        self._configured = True

    @property
    def engine(self):
        self._ensure()
        return self._engine

    def num_frames(self, n_samples):
        f = self.feat_fun
        if f.desc.kind == 0:
            return 0 if n_samples < f.win_length else 1 + (n_samples - f.win_length) // f.hop
        return 1 + n_samples // f.hop

    @staticmethod
    def keep_frames(input_lens_ratio, T):
        """featurizer.py:82-84: mask_lens = round(ratio * T) in float32, half-to-even."""
        r = torch.as_tensor(input_lens_ratio, dtype=torch.float32).cpu()
        return torch.round(r * T).to(torch.int32)

    def forward(self, waveforms, input_lens_ratio=None, group=None):
        """waveforms [B, L] (or [L]) float32 (torch CPU/CUDA tensor or ndarray) -> CUDA tensor [B, T, F].

        ``group``: a torch.distributed process group over which ONE reference call is sharded by utterances (every rank
        passes its shard, padded to the global longest item): only MFCC has a cross-utterance term -- the top_db clamp
        against the maximum of the whole call -- which is then all-reduced (MAX) between the mel stage and the DCT."""
        self._ensure()
        dev = self._engine.device
        w = torch.as_tensor(waveforms)
        if w.dim() == 1:
            w = w.unsqueeze(0)
        w = w.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        B, Lp = w.shape
        T = self.num_frames(Lp)
        keep = None
        if input_lens_ratio is not None:
            keep = self.keep_frames(input_lens_ratio, T).to(dev, non_blocking=True)
        return self.forward_keep(w, keep, group=group)

    def forward_keep(self, w, keep=None, group=None):
        """Same with the mask lengths already on the device: ``w`` CUDA float32 [B, L] contiguous, ``keep`` CUDA int32 [B]
        (frames kept per utterance, featurizer.py:82-84) or None."""
        self._ensure()
        dev = self._engine.device
        B, Lp = w.shape
        f = self.feat_fun
        if f.desc.kind == 0:
            assert 2 <= f.win_length <= Lp, f'choose a window size {f.win_length} that is [2, {Lp}]'
        T = self.num_frames(Lp)
        feats = torch.empty(B, T, self.feature_dim, dtype=torch.float32, device=dev)
        lib = L.lib()
        scratch = torch.empty(max(int(lib.vp_frontend_scratch_floats(self._engine.handle, B, Lp)), 1),
                              dtype=torch.float32, device=dev)
        kp = C.c_void_p(keep.data_ptr()) if keep is not None else C.c_void_p()
        if group is not None and f.desc.post == 1 and f.desc.top_db >= 0:
            self.mfcc_sharded(w, B, Lp, kp, feats, scratch, torch.cuda.current_stream(dev), group)
            return feats
        fn = lib.vp_fbank if f.desc.kind == 0 else (lib.vp_mfcc if f.desc.post == 1 else lib.vp_melspec)
        _check(self._engine.handle, fn(self._engine.handle, C.c_void_p(w.data_ptr()), B, Lp, kp,
                                       C.c_void_p(feats.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                       self._engine.stream_ptr()))
        return feats

    def mfcc_sharded(self, w, B, Lp, kp, feats, scratch, stream, group):
        """MFCC of this rank's shard of ONE sharded call: vp_mfcc_mel -> all-reduce(MAX) of the clamp maximum over
        ``group`` -> vp_mfcc_finish, all enqueued on ``stream``.  Bit-identical to the single-process vp_mfcc of the whole
        batch (max is exact and order independent)."""
        import torch.distributed as dist
        lib = L.lib()
        h = self._engine.handle
        mx = torch.empty(1, dtype=torch.float32, device=feats.device)
        sp = C.c_void_p(stream.cuda_stream)
        with torch.cuda.stream(stream):
            _check(h, lib.vp_mfcc_mel(h, C.c_void_p(w.data_ptr()), B, Lp, C.c_void_p(scratch.data_ptr()),
                                      C.c_void_p(mx.data_ptr()), sp))
            if dist.is_initialized() and dist.get_world_size(group) > 1:
                from ..distributed import device_collectives
                if device_collectives(group):
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
                else:                                   # gloo: one float through the host
                    m = mx.cpu()
                    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
                    mx.copy_(m)
            _check(h, lib.vp_mfcc_finish(h, B, Lp, kp, C.c_void_p(feats.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                         C.c_void_p(mx.data_ptr()), sp))

    __call__ = forward

    @property
    def feature_dim(self):
        """featurizer.py:93-111."""
        if self._feature_method == 'MelSpectrogram':
            return self._method_args.get('n_mels', 128)
        elif self._feature_method == 'Spectrogram':
            return self._method_args.get('n_fft', 400) // 2 + 1
        elif self._feature_method == 'MFCC':
            return self._method_args.get('n_mfcc', 40)
        elif self._feature_method == 'Fbank':
            return self._method_args.get('num_mel_bins', 23)
        raise Exception('没有{}预处理方法'.format(self._feature_method))
