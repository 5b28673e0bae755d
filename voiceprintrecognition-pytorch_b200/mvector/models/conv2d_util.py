"""Shared helpers for the 2-D (conv2d) backbones.  Activation maps are channel-last [B, T, F, C]: T (time) is the
conv2d W axis and F (frequency) the H axis of the reference's [B, C, F, T] tensors, so that the final
``reshape(B, C*F, T)`` (campplus.py:290-291, resnet_se.py:139) is free: row (b, t) already holds the F*C values,
in (f, c) order instead of the reference's (c, f) -- the permutation is folded into the next layer's weights."""
import numpy as np

from .base import _np64, bn_affine


def conv2d_weight(w):
    """[Cout, Cin, kF, kT] -> [Cout, (kt, kf, ci)]: the K order of the CONV op's gather."""
    w = _np64(w)
    return np.ascontiguousarray(w.transpose(0, 3, 2, 1)).reshape(w.shape[0], -1)


def fold_conv_bn(sd, conv_key, bn_prefix, bias_key=None):
    """conv -> BN (eval) == conv with W * s[n] and bias (b * s + h): returns (W2d fp64, bias fp64)."""
    W = conv2d_weight(sd[conv_key])
    s, h = bn_affine(sd, bn_prefix)
    b = _np64(sd[bias_key]) if bias_key is not None else 0.0
    return W * s[:, None], b * s + h


def out_len(n, k, s, pad, dil=1):
    return (n + 2 * pad - dil * (k - 1) - 1) // s + 1


def fc_perm(F8, C):
    """perm[j] for lowered column j = f*C + c  ->  reference flattened channel c*F8 + f."""
    f = np.arange(F8)[:, None]
    c = np.arange(C)[None, :]
    return (c * F8 + f).reshape(-1)


def bn_names(d, p, c, affine=True):
    if affine:
        d[p + '.weight'] = (c,)
        d[p + '.bias'] = (c,)
    d[p + '.running_mean'] = (c,)
    d[p + '.running_var'] = (c,)
    d[p + '.num_batches_tracked'] = ()
