"""Attentive statistics pooling, lowered.  Mirrors mvector/models/pooling.py:68-127 (lengths=None, global_context).

The reference materialises attn = cat([x, mean.repeat(T), std.repeat(T)]) (a [B, 3C, T] copy, 1.4 GB at ECAPA B=256)
and runs a 3C -> A 1x1 conv over it.  Algebraically  W[:, :C] x_t + (W[:, C:2C] mean + W[:, 2C:] std + b)  -- the
second term is a per-utterance bias computed once by a [B, 2C] x [2C, A] product, so the lowered form is:
  stats  = colstats(mean, std)                       [B, 2C]
  ubias  = stats @ W[:, C:3C]^T + b                  [B, A]
  h      = tanh(bn(relu(x @ W[:, :C]^T + ubias)))    [B*T, A]      (TDNNBlock = conv -> ReLU -> BN, then tanh)
  logits = h @ Wc^T + bc                             [B*T, C]
  pooled = softmax_T(logits)-weighted mean / std     [B, 2C]
"""
import numpy as np

from .. import _lib as L
from .base import _np64, bn_affine


def asp_shapes(d, p, c, att=128):
    d[p + '.tdnn.conv.conv.weight'] = (att, c * 3, 1)
    d[p + '.tdnn.conv.conv.bias'] = (att,)
    for n in ('weight', 'bias', 'running_mean', 'running_var'):
        d[p + '.tdnn.norm.norm.' + n] = (att,)
    d[p + '.tdnn.norm.norm.num_batches_tracked'] = ()
    d[p + '.conv.conv.weight'] = (c, att, 1)
    d[p + '.conv.conv.bias'] = (c,)


def pack_asp(sd, p, arena, C, perm=None):
    """perm[j] = reference channel held by lowered column j (2-D backbones flatten (f, c) instead of (c, f))."""
    W = _np64(sd[p + '.tdnn.conv.conv.weight'])[:, :, 0]            # [A, 3C]
    A = W.shape[0]
    Wc = _np64(sd[p + '.conv.conv.weight'])[:, :, 0]                # [C, A]
    bc = _np64(sd[p + '.conv.conv.bias'])
    if perm is not None:
        W = np.concatenate([W[:, perm], W[:, C + perm], W[:, 2 * C + perm]], axis=1)
        Wc, bc = Wc[perm], bc[perm]
    s, h = bn_affine(sd, p + '.tdnn.norm.norm')
    return dict(A=A, C=C,
                wx=arena.add_conv(p + '.tdnn.wx', W[:, :C]), wms=arena.add(p + '.tdnn.wms', W[:, C:]),
                b=arena.add(p + '.tdnn.b', sd[p + '.tdnn.conv.conv.bias']),
                s=arena.add(p + '.tdnn.bn_s', s), h=arena.add(p + '.tdnn.bn_h', h),
                wc=arena.add_conv(p + '.conv.w', Wc), bc=arena.add(p + '.conv.b', bc))


def lower_asp(pb, o, x, B, T, pooled):
    """x: View [B*T, C]; pooled: View [B, 2C] (mean ; std)."""
    C, A = o['C'], o['A']
    stats = pb.alloc(B, 2 * C)
    pb.colstats(x, stats, T, L.STATS_MEAN_STD_CLAMP, eps=1e-12)
    ub = pb.alloc(B, A)
    pb.conv(stats, ub, o['wms'], 2 * C, 1, 1, bias=o['b'], engine=L.ENGINE_FFMA)
    h = pb.alloc(B * T, A)
    pb.conv(x, h, o['wx'], C, T, T, ubias=ub, act=L.ACT_RELU, post=(o['s'], o['h']), act2=L.ACT_TANH)
    logits = pb.alloc(B * T, C)
    pb.conv(h, logits, o['wc'], A, T, T, bias=o['bc'])
    pb.asp_pool(x, logits, pooled, T, eps=1e-12)
    for v in (logits, h, ub, stats):
        pb.free(v)
