"""CPU emulation of frontend_kernel's arithmetic (numpy fp32, same op order) to localise the log-mel error term.
Compares against oracle.frontend.kaldi_fbank (== torchaudio, bit exact) and an fp64 evaluation of the same pipeline."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frontend as ofe

f32 = np.float32

def stockham(x, N, dtype):
    """x: [nf, N] complex (as two real arrays) -> FFT, radix-4 passes + radix-2, in `dtype`."""
    re, im = x[0].astype(dtype), x[1].astype(dtype)
    k = np.arange(N)
    ang = -2.0 * np.pi * k / N
    twr, twi = np.cos(ang).astype(dtype), np.sin(ang).astype(dtype)
    def cmul(ar, ai, br, bi):
        return (ar * br - ai * bi).astype(dtype), (ar * bi + ai * br).astype(dtype)
    Ns = 1
    while Ns < N:
        rem = N // Ns
        R = 4 if rem % 4 == 0 else 2
        step = N // (Ns * R)
        q = N // R
        j = np.arange(q)
        kk = j & (Ns - 1)
        base = (j - kk) * R + kk
        outr, outi = np.empty_like(re), np.empty_like(im)
        if R == 4:
            v = [(re[:, j + r * q], im[:, j + r * q]) for r in range(4)]
            if Ns > 1:
                for r in (1, 2, 3):
                    v[r] = cmul(v[r][0], v[r][1], twr[r * kk * step], twi[r * kk * step])
            a0 = (v[0][0] + v[2][0], v[0][1] + v[2][1]); a1 = (v[0][0] - v[2][0], v[0][1] - v[2][1])
            a2 = (v[1][0] + v[3][0], v[1][1] + v[3][1]); a3 = (v[1][1] - v[3][1], -(v[1][0] - v[3][0]))
            outr[:, base] = a0[0] + a2[0]; outi[:, base] = a0[1] + a2[1]
            outr[:, base + Ns] = a1[0] + a3[0]; outi[:, base + Ns] = a1[1] + a3[1]
            outr[:, base + 2 * Ns] = a0[0] - a2[0]; outi[:, base + 2 * Ns] = a0[1] - a2[1]
            outr[:, base + 3 * Ns] = a1[0] - a3[0]; outi[:, base + 3 * Ns] = a1[1] - a3[1]
        else:
            v0 = (re[:, j], im[:, j]); v1 = cmul(re[:, j + q], im[:, j + q], twr[kk * step], twi[kk * step])
            outr[:, base] = v0[0] + v1[0]; outi[:, base] = v0[1] + v1[1]
            outr[:, base + Ns] = v0[0] - v1[0]; outi[:, base + Ns] = v0[1] - v1[1]
        re, im = outr.astype(dtype), outi.astype(dtype)
        Ns *= R
    return re, im

def frames_fp32(w):
    """kaldi window pipeline in fp32 with torch ops (identical to the reference up to the rfft)."""
    w = torch.as_tensor(w, dtype=torch.float32)
    m = ofe.num_frames(w.numel(), 400, 160)
    fr = w.as_strided((m, 400), (160, 1))
    fr = fr - fr.mean(dim=1, keepdim=True)
    prev = torch.cat([fr[:, :1], fr[:, :-1]], dim=1)
    fr = fr - 0.97 * prev
    fr = fr * ofe.feature_window('povey', 400).unsqueeze(0)
    return torch.nn.functional.pad(fr, (0, 112)).numpy()

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    g = torch.Generator().manual_seed(1234)
    waves = torch.randn(B, 48000, generator=g) * 0.1
    banks = torch.nn.functional.pad(ofe.kaldi_mel_banks(80, 512, 16000.0, 20.0, 0.0).to(torch.float32), (0, 1)).numpy()
    eps = f32(np.finfo(np.float32).eps)
    errs = {}
    def upd(name, a, ref):
        errs[name] = max(errs.get(name, 0.0), float(np.abs(a.astype(np.float64) - ref.astype(np.float64)).max()))
    for b in range(B):
        ref = ofe.kaldi_fbank(waves[b], sample_frequency=16000, num_mel_bins=80).numpy()
        fr = frames_fp32(waves[b])
        nf = fr.shape[0]
        # fp64 truth on the same fp32 windowed frames
        X = np.fft.rfft(fr.astype(np.float64), axis=1)
        P64 = (X.real ** 2 + X.imag ** 2)
        truth = np.log(np.maximum(P64 @ banks.astype(np.float64).T, float(eps)))
        upd('reference(torchaudio) vs fp64', ref, truth)
        # (d) fp64 FFT, fp32 power/mel/log
        P = (X.real.astype(f32) ** 2 + X.imag.astype(f32) ** 2).astype(f32)
        mel = np.zeros((nf, 80), f32)
        for k in range(257):
            nz = np.nonzero(banks[:, k])[0]
            for m in nz:
                mel[:, m] = (P[:, k] * banks[m, k] + mel[:, m]).astype(f32)
        d = np.log(np.maximum(mel, eps)).astype(f32)
        upd('fp64 FFT + fp32 tail vs reference', d, ref)
        upd('fp64 FFT + fp32 tail vs fp64', d, truth)
        # (c) packed two-frame fp32 stockham (the kernel)
        for dtype, tag in ((f32, 'fp32'),):
            npair = (nf + 1) // 2
            fa = fr[0::2]; fb = np.zeros_like(fa); fb[:fr[1::2].shape[0]] = fr[1::2]
            zr, zi = stockham((fa, fb), 512, dtype)
            kidx = np.arange(257); nk = (512 - kidx) % 512
            ar = f32(0.5) * (zr[:, kidx] + zr[:, nk]); ai = f32(0.5) * (zi[:, kidx] - zi[:, nk])
            br = f32(0.5) * (zi[:, kidx] + zi[:, nk]); bi = f32(-0.5) * (zr[:, kidx] - zr[:, nk])
            Pa = (ar * ar + ai * ai).astype(f32); Pb = (br * br + bi * bi).astype(f32)
            Pk = np.empty((2 * npair, 257), f32); Pk[0::2] = Pa; Pk[1::2] = Pb; Pk = Pk[:nf]
            melk = (Pk.astype(np.float64) @ banks.astype(np.float64).T).astype(f32)
            c = np.log(np.maximum(melk, eps)).astype(f32)
            upd(f'packed {tag} stockham (kernel) vs reference', c, ref)
            upd(f'packed {tag} stockham (kernel) vs fp64', c, truth)
            # (e) unpacked: one frame per complex FFT (imag = 0)
            zr, zi = stockham((fr, np.zeros_like(fr)), 512, dtype)
            Pu = (zr[:, :257] ** 2 + zi[:, :257] ** 2).astype(f32)
            e = np.log(np.maximum((Pu.astype(np.float64) @ banks.astype(np.float64).T).astype(f32), eps)).astype(f32)
            upd(f'unpacked {tag} stockham vs reference', e, ref)
            upd(f'unpacked {tag} stockham vs fp64', e, truth)
    for k, v in errs.items():
        print(f'{k:50s} max-abs {v:.3e}')

main()
