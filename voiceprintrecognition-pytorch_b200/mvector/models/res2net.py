"""Res2Net mirror (reference: mvector/models/res2net.py:89-174), lowered.

  stem      7x7 stride-3 conv2d on the one-channel feature map (CONV_C1, BN folded, ReLU) -> 3x3 stride-2 max pool (POOL2D)
  Bottle2neck (res2net.py:10-86) on channel-last [B, T, F, C] maps:
      conv1 1x1 (BN folded, ReLU) -> split into `scale` groups of `width` channels
      groups 0 .. scale-2: 3x3 conv (stride s) + BN + ReLU; in 'normal' blocks group j adds the previous group's output
                            to its input (CONV gather ADD), in 'stage' blocks (first of a layer) every group is independent
      last group: passed through ('normal': EW copy) or 3x3 average pooled with stride s ('stage': POOL2D)
      all written into their slots of the concat buffer; conv3 1x1 + BN + residual (+ 1x1 stride-s downsample) + ReLU
  head      free (f, c) flatten -> pooling (ASP/SAP/TAP/TSP) -> bn2 -> linear -> bn3 folded into one product
"""
import math
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from ..engine import View
from .base import Backbone, _np64, bn_affine
from .campplus import L_view1
from .conv2d_util import bn_names, fc_perm, fold_conv_bn, out_len
from .pooling import check_pooling_type, lower_pool, pack_pool, pool_perm, pool_shapes, pool_width


class Res2Net(Backbone):
    def __init__(self, input_size, m_channels=32, layers=[3, 4, 6, 3], base_width=32, scale=2, embd_dim=192,
                 pooling_type='ASP'):
        super().__init__()
        check_pooling_type(pooling_type)
        if scale < 2:
            raise NotImplementedError('Res2Net: scale == 1 is not lowered')
        self.pooling_type = pooling_type
        self.input_size, self.embd_dim = input_size, embd_dim
        self.m, self.layers, self.base_width, self.scale = m_channels, list(layers), base_width, scale
        self.cat = m_channels * 8 * 4 * (input_size // base_width)

    def _blocks(self):
        inpl = self.m
        for li, nb in enumerate(self.layers, start=1):
            planes = self.m * (2 ** (li - 1))
            stride = 1 if li == 1 else 2
            for b in range(nb):
                first = b == 0
                ds = first and (stride != 1 or inpl != planes * 4)
                width = int(math.floor(planes * (self.base_width / 64.0)))
                yield f'layer{li}.{b}', inpl, planes, width, (stride if first else 1), first, ds
                inpl = planes * 4

    def param_shapes(self):
        d = OrderedDict()
        d['conv1.weight'] = (self.m, 1, 7, 7)
        bn_names(d, 'bn1', self.m)
        nums = self.scale - 1
        for p, inpl, planes, w, stride, stage, ds in self._blocks():
            d[p + '.conv1.weight'] = (w * self.scale, inpl, 1, 1)
            bn_names(d, p + '.bn1', w * self.scale)
            for j in range(nums):
                d[f'{p}.convs.{j}.weight'] = (w, w, 3, 3)
            for j in range(nums):
                bn_names(d, f'{p}.bns.{j}', w)
            d[p + '.conv3.weight'] = (planes * 4, w * self.scale, 1, 1)
            bn_names(d, p + '.bn3', planes * 4)
            if ds:
                d[p + '.downsample.0.weight'] = (planes * 4, inpl, 1, 1)
                bn_names(d, p + '.downsample.1', planes * 4)
        width = pool_shapes(d, 'pooling', self.pooling_type, self.cat, 128)
        bn_names(d, 'bn2', width)
        d['linear.weight'] = (self.embd_dim, width)
        d['linear.bias'] = (self.embd_dim,)
        bn_names(d, 'bn3', self.embd_dim)
        return d

    def _pack(self, sd, arena):
        o = self._off

        def cb(name, conv_key, bn):
            W, b = fold_conv_bn(sd, conv_key, bn)
            o[name] = dict(w=arena.add_conv(name + '.w', W), b=arena.add(name + '.b', b))

        cb('stem', 'conv1.weight', 'bn1')
        for p, inpl, planes, w, stride, stage, ds in self._blocks():
            cb(p + '.c1', p + '.conv1.weight', p + '.bn1')
            for j in range(self.scale - 1):
                cb(f'{p}.k{j}', f'{p}.convs.{j}.weight', f'{p}.bns.{j}')
            cb(p + '.c3', p + '.conv3.weight', p + '.bn3')
            if ds:
                cb(p + '.ds', p + '.downsample.0.weight', p + '.downsample.1')
        C4 = self.m * 8 * 4
        f_last = self.cat // C4
        perm = fc_perm(f_last, C4)
        o['pool'] = pack_pool(sd, 'pooling', self.pooling_type, arena, self.cat, perm=perm)
        s2, h2 = bn_affine(sd, 'bn2')
        s3, h3 = bn_affine(sd, 'bn3')
        W, b = _np64(sd['linear.weight']), _np64(sd['linear.bias'])
        Wf = s3[:, None] * W * s2[None, :]
        o['fc_w'] = arena.add('fc.w', Wf[:, pool_perm(self.pooling_type, self.cat, perm)])
        o['fc_b'] = arena.add('fc.b', s3 * (W @ h2 + b) + h3)

    def _lower(self, pb, B, T):
        o, sc = self._off, self.scale
        F = self.input_size
        x_in = pb.input_view(F, B * T)
        t, f = out_len(T, 7, 3, 1), out_len(F, 7, 3, 1)
        if t < 1 or f < 1:
            raise ValueError(f'{T} frames x {F} bins is too small for the 7x7 stride-3 stem')
        s0 = pb.alloc(B * t * f, self.m)
        pb.conv(L_view1(x_in), s0, o['stem']['w'], 49, T, t, Fin=F, Fout=f, KT=7, KF=7, sT=3, sF=3, padT=1, padF=1,
                bias=o['stem']['b'], act=L.ACT_RELU, c1=True)
        tp, fp = out_len(t, 3, 2, 1), out_len(f, 3, 2, 1)
        x = pb.alloc(B * tp * fp, self.m)
        pb.pool2d(s0, x, L.POOL_MAX, t, f, tp, fp, k=3, stride=2, pad=1)
        pb.free(s0)
        t, f = tp, fp
        for p, inpl, planes, w, stride, stage, ds in self._blocks():
            to, fo = out_len(t, 3, stride, 1), out_len(f, 3, stride, 1)
            rows_in, rows_out = B * t * f, B * to * fo
            h = pb.alloc(rows_in, w * sc)
            pb.conv(x, h, o[p + '.c1']['w'], inpl, t, t, Fin=f, Fout=f, bias=o[p + '.c1']['b'], act=L.ACT_RELU)
            cat = pb.alloc(rows_out, w * sc)
            for j in range(sc - 1):
                e = o[f'{p}.k{j}']
                add_prev = (j > 0) and not stage
                pb.conv(h.cols(j * w, w), cat.cols(j * w, w), e['w'], 9 * w, t, to, Fin=f, Fout=fo, KT=3, KF=3, sT=stride,
                        sF=stride, padT=1, padF=1, bias=e['b'], act=L.ACT_RELU,
                        src2=cat.cols((j - 1) * w, w) if add_prev else None,
                        src2_mode=L.SRC2_ADD if add_prev else L.SRC2_NONE)
            last = h.cols((sc - 1) * w, w)
            if stage:
                pb.pool2d(last, cat.cols((sc - 1) * w, w), L.POOL_AVG, t, f, to, fo, k=3, stride=stride, pad=1)
            else:
                pb.ew(L.EW_COPY, last, cat.cols((sc - 1) * w, w), to * fo)
            pb.free(h)
            cout = planes * 4
            if ds:
                res = pb.alloc(rows_out, cout)
                pb.conv(x, res, o[p + '.ds']['w'], inpl, t, to, Fin=f, Fout=fo, sT=stride, sF=stride, bias=o[p + '.ds']['b'])
            else:
                res = x
            y = pb.alloc(rows_out, cout)
            pb.conv(cat, y, o[p + '.c3']['w'], w * sc, to, to, Fin=fo, Fout=fo, bias=o[p + '.c3']['b'], res=res,
                    act2=L.ACT_RELU)
            pb.free(cat)
            if ds:
                pb.free(res)
            pb.free(x)
            x, t, f = y, to, fo
        C4 = self.m * 8 * 4
        if f * C4 != self.cat:
            raise ValueError(f'input_size {F}: the flattened map has {f * C4} channels but the head was built for '
                             f'{self.cat} = m_channels*32*(input_size // base_width) (res2net.py:111)')
        flat = View(x.off, f * C4, 0, f * C4)
        width = pool_width(self.pooling_type, self.cat)
        pooled = pb.alloc(B, width)
        lower_pool(pb, o['pool'], self.pooling_type, flat, B, t, pooled)
        pb.free(x)
        pb.conv(pooled, pb.output_view(self.embd_dim, B), o['fc_w'], width, 1, 1, bias=o['fc_b'], engine=L.ENGINE_FFMA)
