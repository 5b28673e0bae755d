"""CPU study (no GPU): embedding error of operand-splitting schemes for the tensor-core GEMMs, emulated by rounding the
operands of every conv / linear of the oracle forward and accumulating the kept cross terms in fp64.

  tf32x3 : hi = tf32(x), lo = tf32(x - hi);  hi*hi + hi*lo + lo*hi          (what conv_tc.cu does today)
  fp16x2u: like fp16x2 but the ACTIVATIONS are split unscaled (weights still scaled per tensor)
  fp16x2 : per-tensor power-of-two scale s (max|x| * s <= 2^14); hi = fp16(x s), lo = fp16(x s - hi) (subnormals allowed);
           hi*hi + hi*lo + lo*hi, descaled                                      (round-2 candidate: kind::f16 runs at 2x tf32)
  bf16x3 : hi/mid/lo bf16, the six terms of order <= 2                           (for comparison)
  tf32x1, bf16x1, fp16x1(scaled) : single pass

Accumulation-order / accumulator-truncation effects are NOT modelled (they are the same for every scheme).
Usage: python tools/split_precision_study.py [model ...]"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle import models as om  # noqa: E402


def r_tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF                      # round to nearest (ties away), keep 10 mantissa bits
    return i.view(torch.float32)


def r_fp16(x):
    return x.to(torch.float16).to(torch.float32)


def r_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def pow2_scale(x, target=2.0 ** 14):
    m = float(x.abs().max())
    if m == 0.0:
        return 1.0
    return 2.0 ** int(torch.floor(torch.log2(torch.tensor(target / m))))


def parts(x, scheme):
    """-> list of (order, tensor) operand parts, and the descale factor."""
    if scheme == 'tf32x3':
        hi = r_tf32(x); lo = r_tf32(x - hi)
        return [(0, hi), (1, lo)], 1.0
    if scheme == 'exact':
        return [(0, x)], 1.0
    if scheme == 'tf32x1':
        return [(0, r_tf32(x))], 1.0
    if scheme == 'bf16x1':
        return [(0, r_bf16(x))], 1.0
    if scheme == 'bf16x3':
        a = r_bf16(x); b = r_bf16(x - a); c = r_bf16(x - a - b)
        return [(0, a), (1, b), (2, c)], 1.0
    if scheme == 'fp16x2u':                       # unscaled two-term fp16 split (activations): no max|x| needed
        hi = r_fp16(x)
        return [(0, hi), (1, r_fp16(x - hi))], 1.0
    if scheme in ('fp16x2', 'fp16x1'):
        s = pow2_scale(x)
        xs = x * s
        hi = r_fp16(xs)
        if scheme == 'fp16x1':
            return [(0, hi)], 1.0 / s
        return [(0, hi), (1, r_fp16(xs - hi))], 1.0 / s
    raise ValueError(scheme)


class Emulate:
    def __init__(self, scheme, min_rows=0):
        self.scheme, self.min_rows = scheme, min_rows
        self.orig = {}

    def _wrap(self, fn):
        scheme = self.scheme

        def wrapped(x, w, bias=None, *a, **k):
            xp, dx = parts(x.float(), scheme)
            wp, dw = parts(w.float(), 'fp16x2' if scheme == 'fp16x2u' else scheme)
            max_order = 0 if len(xp) == 1 else (1 if len(xp) == 2 else 2)
            acc = None
            for ox, xa in xp:
                for ow, wa in wp:
                    if ox + ow > max_order:
                        continue
                    t = fn(xa.double(), wa.double(), None, *a, **k)
                    acc = t if acc is None else acc + t
            out = (acc * (dx * dw)).float()
            if bias is not None:
                out = out + bias.view(1, -1, *([1] * (out.dim() - 2))) if out.dim() > 2 else out + bias
            return out
        return wrapped

    def __enter__(self):
        for name in ('conv1d', 'conv2d', 'linear'):
            self.orig[name] = getattr(F, name)
            setattr(F, name, self._wrap(self.orig[name]))
        return self

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(F, name, fn)


CASES = {
    'EcapaTdnn': (80, dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536]), 2, 298),
    'TDNN': (80, dict(embd_dim=192, channels=512, pooling_type='ASP'), 2, 218),
    'CAMPPlus': (80, dict(embd_dim=192), 2, 200),
    'ResNetSE': (64, dict(embd_dim=192, pooling_type='ASP'), 1, 120),
    'ERes2Net': (80, dict(embd_dim=192, m_channels=32), 1, 98),
}


def main(argv):
    names = argv or ['EcapaTdnn', 'TDNN']
    schemes = ['tf32x3', 'fp16x2', 'fp16x2u', 'bf16x3', 'tf32x1', 'fp16x1', 'bf16x1']
    print(f'{"model":10s} ' + ' '.join(f'{s:>9s}' for s in schemes) + '   (max rel-L2 of the embedding vs exact-contraction forward)')
    for name in names:
        fdim, margs, B, T = CASES[name]
        sd = om.random_state_dict(name, fdim, seed=0, gain=om.CONDITIONED_GAIN[name], **margs)
        x = torch.randn(B, T, fdim, generator=torch.Generator().manual_seed(7)) * 2.0
        with Emulate('exact'):                      # fp32 tensors between layers, exact (fp64) contractions
            ref = om.forward(name, sd, x, **margs).double()
        row = []
        for s in schemes:
            with Emulate(s):
                e = om.forward(name, sd, x, **margs).double()
            row.append(float(((e - ref).norm(dim=1) / ref.norm(dim=1)).max()))
        base = om.forward(name, sd, x, **margs).double()
        fp32 = float(((base - ref).norm(dim=1) / ref.norm(dim=1)).max())
        print(f'{name:10s} ' + ' '.join(f'{v:9.1e}' for v in row) + f'   plain fp32: {fp32:.1e}')


if __name__ == '__main__':
    main(sys.argv[1:])
