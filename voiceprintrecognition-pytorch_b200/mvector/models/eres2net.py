"""ERes2Net mirror (reference: mvector/models/eres2net.py:173-263, two_emb_layer=False), lowered.

BasicBlockERes2Net / _diff_AFF (eres2net.py:55-170) on channel-last [B, T, F, C] maps; activation is Hardtanh(0, 20)
(eres2net.py:12-15) except the stem's plain ReLU (eres2net.py:243); every BN follows its conv and is folded.
  conv1 1x1 (stride s)              CONV (+hardtanh)
  split into `scale` groups; group j: input = prev + x_j (CONV gather ADD) or AFF(prev, x_j) (eres2net.py:32-52):
     att = BN(conv(SiLU(BN(conv(cat(prev, x_j))))))   two CONVs (gather CONCAT, BN folded)   then EW AFF blend
     3x3 conv + BN + hardtanh -> written into its slot of the concat buffer
  conv3 1x1 + BN + shortcut + hardtanh              one CONV (residual + act2 in the epilogue)
Bottom-up fusion (eres2net.py:246-253): 3x3 stride-2 downsample CONVs + AFF; TSTP pooling (pooling.py:140-148); seg_1.
"""
import math
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from ..engine import View
from .base import Backbone, _np64, bn_affine
from .campplus import L_view1
from .conv2d_util import bn_names, conv2d_weight, fc_perm, fold_conv_bn, out_len

HT = L.ACT_HARDTANH20


def _aff_names(d, p, channels, r=4):
    inter = int(channels // r)
    d[p + '.local_att.0.weight'] = (inter, channels * 2, 1, 1)
    d[p + '.local_att.0.bias'] = (inter,)
    bn_names(d, p + '.local_att.1', inter)
    d[p + '.local_att.3.weight'] = (channels, inter, 1, 1)
    d[p + '.local_att.3.bias'] = (channels,)
    bn_names(d, p + '.local_att.4', channels)


class ERes2Net(Backbone):
    def __init__(self, input_size, block=None, block_fuse=None, num_blocks=[3, 4, 6, 3], m_channels=32, mul_channel=1,
                 expansion=2, base_width=32, scale=2, embd_dim=192, two_emb_layer=False):
        super().__init__()
        if block is not None or block_fuse is not None or two_emb_layer:
            raise NotImplementedError('ERes2Net: custom blocks / two_emb_layer are not lowered')
        self.input_size, self.embd_dim = input_size, embd_dim
        self.num_blocks, self.m, self.mul, self.expansion = list(num_blocks), m_channels, mul_channel, expansion
        self.base_width, self.scale = base_width, scale
        self.F8 = int(input_size / 8)
        self.stats_dim = self.F8 * m_channels * 8
        if m_channels * 2 * mul_channel != m_channels * expansion:
            raise ValueError('mul_channel * 2 must equal expansion (layer1_downsample in-channels, eres2net.py:211)')

    def _blocks(self):
        inpl = self.m
        for li, nb in enumerate(self.num_blocks, start=1):
            planes = self.m * (2 ** (li - 1))
            width = int(math.floor(planes * (self.base_width / 64.0)))
            for b in range(nb):
                stride = 2 if (li > 1 and b == 0) else 1
                sc = stride != 1 or inpl != planes * self.expansion
                yield f'layer{li}.{b}', li, inpl, planes, width, stride, li >= 3, sc
                inpl = planes * self.expansion

    def param_shapes(self):
        d = OrderedDict()
        d['conv1.weight'] = (self.m, 1, 3, 3)
        bn_names(d, 'bn1', self.m)
        for p, li, inpl, planes, w, stride, fuse, sc in self._blocks():
            d[p + '.conv1.weight'] = (w * self.scale, inpl, 1, 1)
            bn_names(d, p + '.bn1', w * self.scale)
            for j in range(self.scale):
                d[f'{p}.convs.{j}.weight'] = (w, w, 3, 3)
            for j in range(self.scale):
                bn_names(d, f'{p}.bns.{j}', w)
            if fuse:
                for j in range(self.scale - 1):
                    _aff_names(d, f'{p}.fuse_models.{j}', w)
            d[p + '.conv3.weight'] = (planes * self.expansion, w * self.scale, 1, 1)
            bn_names(d, p + '.bn3', planes * self.expansion)
            if sc:
                d[p + '.shortcut.0.weight'] = (planes * self.expansion, inpl, 1, 1)
                bn_names(d, p + '.shortcut.1', planes * self.expansion)
        mc = self.m * self.mul
        d['layer1_downsample.weight'] = (mc * 4, mc * 2, 3, 3)
        d['layer2_downsample.weight'] = (mc * 8, mc * 4, 3, 3)
        d['layer3_downsample.weight'] = (mc * 16, mc * 8, 3, 3)
        _aff_names(d, 'fuse_mode12', mc * 4)
        _aff_names(d, 'fuse_mode123', mc * 8)
        _aff_names(d, 'fuse_mode1234', mc * 16)
        d['seg_1.weight'] = (self.embd_dim, self.stats_dim * self.expansion * 2)
        d['seg_1.bias'] = (self.embd_dim,)
        return d

    # ---- weights ----
    def _pack_aff(self, sd, p, arena):
        W0, b0 = fold_conv_bn(sd, p + '.local_att.0.weight', p + '.local_att.1', p + '.local_att.0.bias')
        W1, b1 = fold_conv_bn(sd, p + '.local_att.3.weight', p + '.local_att.4', p + '.local_att.3.bias')
        return dict(w0=arena.add_conv(p + '.w0', W0), b0=arena.add(p + '.b0', b0), w1=arena.add_conv(p + '.w1', W1),
                    b1=arena.add(p + '.b1', b1), inter=W0.shape[0], ch=W1.shape[0])

    def _pack(self, sd, arena):
        o = self._off

        def cb(name, conv_key, bn):
            W, b = fold_conv_bn(sd, conv_key, bn)
            o[name] = dict(w=arena.add_conv(name + '.w', W), b=arena.add(name + '.b', b))

        cb('stem', 'conv1.weight', 'bn1')
        for p, li, inpl, planes, w, stride, fuse, sc in self._blocks():
            cb(p + '.c1', p + '.conv1.weight', p + '.bn1')
            for j in range(self.scale):
                cb(f'{p}.k{j}', f'{p}.convs.{j}.weight', f'{p}.bns.{j}')
            if fuse:
                for j in range(self.scale - 1):
                    o[f'{p}.aff{j}'] = self._pack_aff(sd, f'{p}.fuse_models.{j}', arena)
            cb(p + '.c3', p + '.conv3.weight', p + '.bn3')
            if sc:
                cb(p + '.sc', p + '.shortcut.0.weight', p + '.shortcut.1')
        for nm in ('layer1_downsample', 'layer2_downsample', 'layer3_downsample'):
            o[nm] = arena.add_conv(nm + '.w', conv2d_weight(sd[nm + '.weight']))
        for nm in ('fuse_mode12', 'fuse_mode123', 'fuse_mode1234'):
            o[nm] = self._pack_aff(sd, nm, arena)
        C4 = self.m * 8 * self.expansion
        perm = fc_perm(self.F8, C4)
        n = self.F8 * C4
        perm2 = np.concatenate([perm, n + perm])
        o['fc_w'] = arena.add('fc.w', _np64(sd['seg_1.weight'])[:, perm2])
        o['fc_b'] = arena.add('fc.b', sd['seg_1.bias'])

    # ---- program ----
    def _aff(self, pb, e, x, y, rows_per_utt, rows, t, f):
        """xo = x * (1 + tanh(att)) + y * (1 - tanh(att)), att = local_att(cat(x, y)) (eres2net.py:44-50)."""
        a0 = pb.alloc(rows, e['inter'])
        pb.conv(x, a0, e['w0'], x.C + y.C, t, t, Fin=f, Fout=f, bias=e['b0'], act=L.ACT_SILU, src2=y,
                src2_mode=L.SRC2_CONCAT)
        a1 = pb.alloc(rows, e['ch'])
        pb.conv(a0, a1, e['w1'], e['inter'], t, t, Fin=f, Fout=f, bias=e['b1'])
        pb.free(a0)
        out = pb.alloc(rows, e['ch'])
        pb.ew(L.EW_AFF, x, out, rows_per_utt, y=y, att=a1)
        pb.free(a1)
        return out

    def _lower(self, pb, B, T):
        o = self._off
        F = self.input_size
        x_in = pb.input_view(F, B * T)
        x = pb.alloc(B * T * F, self.m)
        pb.conv(L_view1(x_in), x, o['stem']['w'], 9, T, T, Fin=F, Fout=F, KT=3, KF=3, padT=1, padF=1,
                bias=o['stem']['b'], act=L.ACT_RELU, c1=True)
        t, f = T, F
        layer_out = {}
        last_li = 1
        for p, li, inpl, planes, w, stride, fuse, sc in self._blocks():
            if li != last_li:
                layer_out[last_li] = (x, t, f)        # keep the stage output alive for the bottom-up fusion
                last_li = li
            keep_x = any(x is v[0] for v in layer_out.values())
            to, fo = out_len(t, 1, stride, 0), out_len(f, 1, stride, 0)
            rows = B * to * fo
            h = pb.alloc(rows, w * self.scale)
            pb.conv(x, h, o[p + '.c1']['w'], inpl, t, to, Fin=f, Fout=fo, sT=stride, sF=stride, bias=o[p + '.c1']['b'],
                    act=HT)
            cat = pb.alloc(rows, w * self.scale)
            for j in range(self.scale):
                e = o[f'{p}.k{j}']
                dst = cat.cols(j * w, w)
                kw = dict(Fin=fo, Fout=fo, KT=3, KF=3, padT=1, padF=1, bias=e['b'], act=HT)
                if j == 0:
                    pb.conv(h.cols(0, w), dst, e['w'], 9 * w, to, to, **kw)
                elif fuse:
                    fz = self._aff(pb, o[f'{p}.aff{j - 1}'], cat.cols((j - 1) * w, w), h.cols(j * w, w), to * fo, rows,
                                   to, fo)
                    pb.conv(fz, dst, e['w'], 9 * w, to, to, **kw)
                    pb.free(fz)
                else:
                    pb.conv(cat.cols((j - 1) * w, w), dst, e['w'], 9 * w, to, to, src2=h.cols(j * w, w),
                            src2_mode=L.SRC2_ADD, **kw)
            pb.free(h)
            cout = planes * self.expansion
            if sc:
                res = pb.alloc(rows, cout)
                pb.conv(x, res, o[p + '.sc']['w'], inpl, t, to, Fin=f, Fout=fo, sT=stride, sF=stride,
                        bias=o[p + '.sc']['b'])
            else:
                res = x
            y = pb.alloc(rows, cout)
            pb.conv(cat, y, o[p + '.c3']['w'], w * self.scale, to, to, Fin=fo, Fout=fo, bias=o[p + '.c3']['b'], res=res,
                    act2=HT)
            pb.free(cat)
            if sc:
                pb.free(res)
            if not keep_x:
                pb.free(x)
            x, t, f = y, to, fo
        layer_out[last_li] = (x, t, f)
        # ---- bottom-up fusion (eres2net.py:246-253) ----
        fused, ft, ff = layer_out[1]
        for li, (ds, aff) in enumerate((('layer1_downsample', 'fuse_mode12'), ('layer2_downsample', 'fuse_mode123'),
                                        ('layer3_downsample', 'fuse_mode1234')), start=2):
            xo, to, fo = layer_out[li]
            assert to == out_len(ft, 3, 2, 1) and fo == out_len(ff, 3, 2, 1)
            rows = B * to * fo
            d = pb.alloc(rows, xo.C)
            pb.conv(fused, d, o[ds], 9 * fused.C, ft, to, Fin=ff, Fout=fo, KT=3, KF=3, sT=2, sF=2, padT=1, padF=1)
            pb.free(fused)
            nf = self._aff(pb, o[aff], xo, d, to * fo, rows, to, fo)
            pb.free(d)
            if li < 4:
                pass                       # layer outputs 2, 3 feed the next stage only through `nf`
            pb.free(xo)
            fused, ft, ff = nf, to, fo
        assert ff == self.F8, 'input_size must be a multiple of 8'
        C4 = fused.C
        flat = View(fused.off, ff * C4, 0, ff * C4)
        stats = pb.alloc(B, 2 * ff * C4)
        pb.colstats(flat, stats, ft, L.STATS_MEAN_STD_TSTP, eps=1e-8)
        pb.free(fused)
        pb.conv(stats, pb.output_view(self.embd_dim, B), o['fc_w'], 2 * ff * C4, 1, 1, bias=o['fc_b'],
                engine=L.ENGINE_FFMA)
