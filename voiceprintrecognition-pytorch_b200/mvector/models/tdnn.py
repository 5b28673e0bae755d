"""TDNN (x-vector) mirror (reference: mvector/models/tdnn.py:9-68): five *valid* (unpadded) Conv1d layers, each
relu(conv) -> BN (tdnn.py:57-65; the 5th has no BN), ASP(512), bn5 -> linear -> bn6 folded into one product."""
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from .base import Backbone, _np64, bn_affine
from .ecapa_tdnn import conv1d_weight
from .pooling import check_pooling_type, lower_pool, pack_pool, pool_shapes, pool_width

_KS = (5, 3, 3, 1, 1)
_DIL = (1, 2, 3, 1, 1)


class TDNN(Backbone):
    def __init__(self, input_size, channels=512, embd_dim=192, pooling_type='ASP'):
        super().__init__()
        check_pooling_type(pooling_type)
        self.pooling_type = pooling_type
        self.input_size, self.channels, self.embd_dim = input_size, channels, embd_dim

    def param_shapes(self):
        d = OrderedDict()
        c = self.channels
        for i, k in enumerate(_KS, start=1):
            d[f'td_layer{i}.weight'] = (c, self.input_size if i == 1 else c, k)
            d[f'td_layer{i}.bias'] = (c,)
            if i < 5:
                for n in ('weight', 'bias', 'running_mean', 'running_var'):
                    d[f'bn{i}.{n}'] = (c,)
                d[f'bn{i}.num_batches_tracked'] = ()
        width = pool_shapes(d, 'pooling', self.pooling_type, c, 128)
        for nm, n_ in (('bn5', width),):
            for n in ('weight', 'bias', 'running_mean', 'running_var'):
                d[f'{nm}.{n}'] = (n_,)
            d[f'{nm}.num_batches_tracked'] = ()
        d['linear.weight'] = (self.embd_dim, width)
        d['linear.bias'] = (self.embd_dim,)
        for n in ('weight', 'bias', 'running_mean', 'running_var'):
            d[f'bn6.{n}'] = (self.embd_dim,)
        d['bn6.num_batches_tracked'] = ()
        return d

    def _pack(self, sd, arena):
        o = self._off
        for i in range(1, 6):
            e = dict(w=arena.add_conv(f'td{i}.w', conv1d_weight(sd[f'td_layer{i}.weight'])),
                     b=arena.add(f'td{i}.b', sd[f'td_layer{i}.bias']))
            if i < 5:
                s, h = bn_affine(sd, f'bn{i}')
                e['s'], e['h'] = arena.add(f'bn{i}.s', s), arena.add(f'bn{i}.h', h)
            o[f'td{i}'] = e
        o['asp'] = pack_pool(sd, 'pooling', self.pooling_type, arena, self.channels)
        s5, h5 = bn_affine(sd, 'bn5')
        s6, h6 = bn_affine(sd, 'bn6')
        W, b = _np64(sd['linear.weight']), _np64(sd['linear.bias'])
        o['fc_w'] = arena.add('fc.w', s6[:, None] * W * s5[None, :])
        o['fc_b'] = arena.add('fc.b', s6 * (W @ h5 + b) + h6)

    def _lower(self, pb, B, T):
        o, c = self._off, self.channels
        x = pb.input_view1d(self.input_size, B * T, T)
        t = T
        for i, (k, dil) in enumerate(zip(_KS, _DIL), start=1):
            tout = t - dil * (k - 1)
            if tout < 1:
                raise ValueError(f'{T} frames is too short for the TDNN receptive field')
            e = o[f'td{i}']
            y = pb.alloc(B * tout, c)
            pb.conv(x, y, e['w'], k * x.C, t, tout, KT=k, dT=dil, bias=e['b'], act=L.ACT_RELU,
                    post=(e['s'], e['h']) if i < 5 else None)
            if x.off != L.BUF_INPUT:
                pb.free(x)
            x, t = y, tout
        width = pool_width(self.pooling_type, c)
        pooled = pb.alloc(B, width)
        lower_pool(pb, o['asp'], self.pooling_type, x, B, t, pooled)
        pb.free(x)
        pb.conv(pooled, pb.output_view(self.embd_dim, B), o['fc_w'], width, 1, 1, bias=o['fc_b'], engine=L.ENGINE_FFMA)
