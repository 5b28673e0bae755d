"""Oracle backbones: functional torch-fp32 CPU restatements of the reference model forwards.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each ``*_forward(sd, x, ...)`` takes a state dict with the
reference's parameter names (without the ``0.`` prefix that ``nn.Sequential(backbone)`` adds, predict.py:55)
and features ``x`` [B, T, F]; returns embeddings [B, embd_dim].  ``*_param_shapes`` enumerate the reference
state-dict entries (name -> shape) so weights can be generated where /root/reference is absent.

Restated from (paths relative to /root/reference/mvector/models):
  utils.py:28-138 (Conv1d reflect-same padding, TDNNBlock = conv -> ReLU -> BN), pooling.py:68-148,
  ecapa_tdnn.py:9-283, tdnn.py:9-68, campplus.py:27-357, resnet_se.py:7-145, eres2net.py:12-263.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def _bn(sd, p, x, eps=1e-5):
    """nn.BatchNorm{1,2}d in eval mode (running statistics)."""
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd.get(p + '.weight'),
                        sd.get(p + '.bias'), False, 0.0, eps)


def _bn_shapes(d, p, c, affine=True):
    if affine:
        d[p + '.weight'] = (c,)
        d[p + '.bias'] = (c,)
    d[p + '.running_mean'] = (c,)
    d[p + '.running_var'] = (c,)
    d[p + '.num_batches_tracked'] = ()


def _same_conv1d(sd, p, x, dilation=1):
    """utils.py:39-103: reflect 'same' padding then nn.Conv1d(bias=True)."""
    w = sd[p + '.conv.weight']
    k = w.shape[-1]
    pad = (dilation * (k - 1)) // 2                                            # utils.py:34-35 (stride 1)
    x = F.pad(x, (pad, pad), mode='reflect')
    return F.conv1d(x, w, sd[p + '.conv.bias'], dilation=dilation)


def _tdnn_block(sd, p, x, dilation=1):
    """utils.py:115-138: conv -> ReLU -> BatchNorm."""
    return _bn(sd, p + '.norm.norm', F.relu(_same_conv1d(sd, p + '.conv', x, dilation)))


def _tdnn_block_shapes(d, p, cin, cout, k):
    d[p + '.conv.conv.weight'] = (cout, cin, k)
    d[p + '.conv.conv.bias'] = (cout,)
    _bn_shapes(d, p + '.norm.norm', cout)


def asp_pool(sd, p, x, eps=1e-12):
    """pooling.py:86-127 with lengths=None (mask of ones) and global_context=True."""
    L = x.shape[-1]
    m = torch.ones(x.shape[0], 1, L) / L
    mean = (m * x).sum(2)
    std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
    attn = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
    attn = _same_conv1d(sd, p + '.conv', torch.tanh(_tdnn_block(sd, p + '.tdnn', attn)))
    attn = F.softmax(attn, dim=2)
    mean = (attn * x).sum(2)
    std = torch.sqrt((attn * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
    return torch.cat((mean, std), dim=1)


def pool_forward(sd, p, kind, x):
    """pooling.py: ASP (:68-127), SAP (:50-65), TAP (:8-26), TSP (:29-47, returns the VARIANCE)."""
    if kind == 'ASP':
        return asp_pool(sd, p, x)
    if kind == 'SAP':
        a = torch.tanh(F.conv1d(x, sd[p + '.linear1.weight'], sd[p + '.linear1.bias']))
        a = torch.softmax(F.conv1d(a, sd[p + '.linear2.weight'], sd[p + '.linear2.bias']), dim=2)
        return torch.sum(a * x, dim=2)
    if kind == 'TAP':
        return torch.mean(x, dim=2).flatten(start_dim=1)
    if kind == 'TSP':
        return torch.cat((torch.mean(x, dim=2), torch.var(x, dim=2)), dim=1)
    raise Exception(f'没有{kind}池化层！')


def _pool_shapes(d, p, kind, c, att=128):
    """Parameters of the pooling module and the width of its output."""
    if kind == 'ASP':
        _asp_shapes(d, p, c, att)
        return 2 * c
    if kind == 'SAP':
        d[p + '.linear1.weight'] = (128, c, 1)
        d[p + '.linear1.bias'] = (128,)
        d[p + '.linear2.weight'] = (c, 128, 1)
        d[p + '.linear2.bias'] = (c,)
        return c
    if kind == 'TAP':
        return c
    if kind == 'TSP':
        return 2 * c
    raise Exception(f'没有{kind}池化层！')


def _asp_shapes(d, p, c, att=128):
    _tdnn_block_shapes(d, p + '.tdnn', c * 3, att, 1)
    d[p + '.conv.conv.weight'] = (c, att, 1)
    d[p + '.conv.conv.bias'] = (c,)


# ---------------------------------------------------------------------------------------------
# EcapaTdnn (ecapa_tdnn.py:146-283), pooling_type 'ASP'
# ---------------------------------------------------------------------------------------------
def ecapa_param_shapes(input_size, embd_dim=192, pooling_type='ASP', channels=(512, 512, 512, 512, 1536),
                       kernel_sizes=(5, 3, 3, 3, 1), dilations=(1, 2, 3, 4, 1), attention_channels=128,
                       res2net_scale=8, se_channels=128, global_context=True):
    assert global_context
    d = OrderedDict()
    _tdnn_block_shapes(d, 'blocks.0', input_size, channels[0], kernel_sizes[0])
    for i in range(1, len(channels) - 1):
        p = f'blocks.{i}'
        cin, c = channels[i - 1], channels[i]
        _tdnn_block_shapes(d, p + '.tdnn1', cin, c, 1)
        w = c // res2net_scale
        for j in range(res2net_scale - 1):
            _tdnn_block_shapes(d, f'{p}.res2net_block.blocks.{j}', w, w, kernel_sizes[i])
        _tdnn_block_shapes(d, p + '.tdnn2', c, c, 1)
        d[p + '.se_block.conv1.conv.weight'] = (se_channels, c, 1)
        d[p + '.se_block.conv1.conv.bias'] = (se_channels,)
        d[p + '.se_block.conv2.conv.weight'] = (c, se_channels, 1)
        d[p + '.se_block.conv2.conv.bias'] = (c,)
        if cin != c:
            d[p + '.shortcut.conv.weight'] = (c, cin, 1)
            d[p + '.shortcut.conv.bias'] = (c,)
    _tdnn_block_shapes(d, 'mfa', channels[-1], channels[-1], kernel_sizes[-1])
    width = _pool_shapes(d, 'asp', pooling_type, channels[-1], attention_channels)
    _bn_shapes(d, 'asp_bn.norm' if pooling_type == 'ASP' else 'asp_bn', width)     # ecapa_tdnn.py:224 vs :232,239,246
    d['fc.conv.weight'] = (embd_dim, width, 1)
    d['fc.conv.bias'] = (embd_dim,)
    return d


def ecapa_forward(sd, x, embd_dim=192, pooling_type='ASP', channels=(512, 512, 512, 512, 1536),
                  kernel_sizes=(5, 3, 3, 3, 1), dilations=(1, 2, 3, 4, 1), attention_channels=128,
                  res2net_scale=8, se_channels=128, global_context=True):
    x = x.transpose(1, 2)                                                       # ecapa_tdnn.py:262
    x = _tdnn_block(sd, 'blocks.0', x, dilations[0])
    outs = []
    for i in range(1, len(channels) - 1):
        p = f'blocks.{i}'
        res = x
        if (p + '.shortcut.conv.weight') in sd:                                 # ecapa_tdnn.py:128-136
            res = _same_conv1d(sd, p + '.shortcut', x)
        h = _tdnn_block(sd, p + '.tdnn1', x)
        ys = []
        for j, hj in enumerate(torch.chunk(h, res2net_scale, dim=1)):           # ecapa_tdnn.py:40-51
            if j == 0:
                y = hj
            elif j == 1:
                y = _tdnn_block(sd, f'{p}.res2net_block.blocks.{j - 1}', hj, dilations[i])
            else:
                y = _tdnn_block(sd, f'{p}.res2net_block.blocks.{j - 1}', hj + y, dilations[i])
            ys.append(y)
        h = _tdnn_block(sd, p + '.tdnn2', torch.cat(ys, dim=1))
        s = h.mean(dim=2, keepdim=True)                                         # ecapa_tdnn.py:79
        s = F.relu(_same_conv1d(sd, p + '.se_block.conv1', s))
        s = torch.sigmoid(_same_conv1d(sd, p + '.se_block.conv2', s))
        x = s * h + res
        outs.append(x)
    x = _tdnn_block(sd, 'mfa', torch.cat(outs, dim=1), dilations[-1])          # ecapa_tdnn.py:273-274
    x = pool_forward(sd, 'asp', pooling_type, x)
    x = _bn(sd, 'asp_bn.norm' if pooling_type == 'ASP' else 'asp_bn', x)
    return _same_conv1d(sd, 'fc', x.unsqueeze(2)).squeeze(-1)                  # ecapa_tdnn.py:279-281


# ---------------------------------------------------------------------------------------------
# TDNN (tdnn.py:9-68), pooling_type 'ASP'
# ---------------------------------------------------------------------------------------------
def tdnn_param_shapes(input_size, channels=512, embd_dim=192, pooling_type='ASP'):
    d = OrderedDict()
    ks = (5, 3, 3, 1, 1)
    for i, k in enumerate(ks, start=1):
        d[f'td_layer{i}.weight'] = (channels, input_size if i == 1 else channels, k)
        d[f'td_layer{i}.bias'] = (channels,)
        if i < 5:
            _bn_shapes(d, f'bn{i}', channels)
    width = _pool_shapes(d, 'pooling', pooling_type, channels, 128)
    _bn_shapes(d, 'bn5', width)
    d['linear.weight'] = (embd_dim, width)
    d['linear.bias'] = (embd_dim,)
    _bn_shapes(d, 'bn6', embd_dim)
    return d


def tdnn_forward(sd, x, channels=512, embd_dim=192, pooling_type='ASP'):
    x = x.transpose(2, 1)
    for i, dil in enumerate((1, 2, 3, 1, 1), start=1):                          # valid (unpadded) convs
        x = F.relu(F.conv1d(x, sd[f'td_layer{i}.weight'], sd[f'td_layer{i}.bias'], dilation=dil))
        if i < 5:
            x = _bn(sd, f'bn{i}', x)
    x = _bn(sd, 'bn5', pool_forward(sd, 'pooling', pooling_type, x))
    return _bn(sd, 'bn6', F.linear(x, sd['linear.weight'], sd['linear.bias']))


# ---------------------------------------------------------------------------------------------
# CAMPPlus (campplus.py:295-357)
# ---------------------------------------------------------------------------------------------
_CAM_BLOCKS = ((12, 3, 1), (24, 3, 2), (16, 3, 2))


def campplus_param_shapes(input_size, embd_dim=512, growth_rate=32, bn_size=4, init_channels=128,
                          config_str='batchnorm-relu', memory_efficient=True):
    assert config_str == 'batchnorm-relu'
    d = OrderedDict()
    m = 32
    d['head.conv1.weight'] = (m, 1, 3, 3)
    _bn_shapes(d, 'head.bn1', m)
    for layer in ('layer1', 'layer2'):
        for b in range(2):
            p = f'head.{layer}.{b}'
            d[p + '.conv1.weight'] = (m, m, 3, 3)
            _bn_shapes(d, p + '.bn1', m)
            d[p + '.conv2.weight'] = (m, m, 3, 3)
            _bn_shapes(d, p + '.bn2', m)
            if b == 0:                                                          # stride 2 -> conv shortcut
                d[p + '.shortcut.0.weight'] = (m, m, 1, 1)
                _bn_shapes(d, p + '.shortcut.1', m)
    d['head.conv2.weight'] = (m, m, 3, 3)
    _bn_shapes(d, 'head.bn2', m)
    ch = m * math.ceil(input_size / 8)
    d['xvector.tdnn.linear.weight'] = (init_channels, ch, 5)
    _bn_shapes(d, 'xvector.tdnn.nonlinear.batchnorm', init_channels)
    ch = init_channels
    bn_ch = bn_size * growth_rate
    for bi, (nl, k, dil) in enumerate(_CAM_BLOCKS, start=1):
        for li in range(nl):
            p = f'xvector.block{bi}.tdnnd{li + 1}'
            cin = ch + li * growth_rate
            _bn_shapes(d, p + '.nonlinear1.batchnorm', cin)
            d[p + '.linear1.weight'] = (bn_ch, cin, 1)
            _bn_shapes(d, p + '.nonlinear2.batchnorm', bn_ch)
            d[p + '.cam_layer.linear_local.weight'] = (growth_rate, bn_ch, k)
            d[p + '.cam_layer.linear1.weight'] = (bn_ch // 2, bn_ch, 1)
            d[p + '.cam_layer.linear1.bias'] = (bn_ch // 2,)
            d[p + '.cam_layer.linear2.weight'] = (growth_rate, bn_ch // 2, 1)
            d[p + '.cam_layer.linear2.bias'] = (growth_rate,)
        ch = ch + nl * growth_rate
        _bn_shapes(d, f'xvector.transit{bi}.nonlinear.batchnorm', ch)
        d[f'xvector.transit{bi}.linear.weight'] = (ch // 2, ch, 1)
        ch //= 2
    _bn_shapes(d, 'xvector.out_nonlinear.batchnorm', ch)
    d['xvector.dense.linear.weight'] = (embd_dim, ch * 2, 1)
    _bn_shapes(d, 'xvector.dense.nonlinear.batchnorm', embd_dim, affine=False)
    return d


def _cam_seg_pool(x, seg_len=100):
    """campplus.py:101-111: avg_pool1d(ceil_mode=True) then expand back to T."""
    seg = F.avg_pool1d(x, kernel_size=seg_len, stride=seg_len, ceil_mode=True)
    shape = seg.shape
    seg = seg.unsqueeze(-1).expand(*shape, seg_len).reshape(*shape[:-1], -1)
    return seg[..., :x.shape[-1]]


def campplus_forward(sd, x, embd_dim=512, growth_rate=32, bn_size=4, init_channels=128,
                     config_str='batchnorm-relu', memory_efficient=True):
    x = x.permute(0, 2, 1).unsqueeze(1)                                         # [B,1,F,T]
    out = F.relu(_bn(sd, 'head.bn1', F.conv2d(x, sd['head.conv1.weight'], padding=1)))
    for layer in ('layer1', 'layer2'):
        for b in range(2):
            p = f'head.{layer}.{b}'
            stride = (2, 1) if b == 0 else (1, 1)
            h = F.relu(_bn(sd, p + '.bn1', F.conv2d(out, sd[p + '.conv1.weight'], stride=stride, padding=1)))
            h = _bn(sd, p + '.bn2', F.conv2d(h, sd[p + '.conv2.weight'], padding=1))
            sc = out
            if (p + '.shortcut.0.weight') in sd:
                sc = _bn(sd, p + '.shortcut.1', F.conv2d(out, sd[p + '.shortcut.0.weight'], stride=stride))
            out = F.relu(h + sc)
    out = F.relu(_bn(sd, 'head.bn2', F.conv2d(out, sd['head.conv2.weight'], stride=(2, 1), padding=1)))
    x = out.reshape(out.shape[0], out.shape[1] * out.shape[2], out.shape[3])    # campplus.py:290-291
    x = F.conv1d(x, sd['xvector.tdnn.linear.weight'], stride=2, padding=2)      # k5, padding=(5-1)//2
    x = F.relu(_bn(sd, 'xvector.tdnn.nonlinear.batchnorm', x))
    for bi, (nl, k, dil) in enumerate(_CAM_BLOCKS, start=1):
        for li in range(nl):
            p = f'xvector.block{bi}.tdnnd{li + 1}'
            h = F.relu(_bn(sd, p + '.nonlinear1.batchnorm', x))
            h = F.conv1d(h, sd[p + '.linear1.weight'])
            h = F.relu(_bn(sd, p + '.nonlinear2.batchnorm', h))
            y = F.conv1d(h, sd[p + '.cam_layer.linear_local.weight'], padding=(k - 1) // 2 * dil, dilation=dil)
            ctx = h.mean(-1, keepdim=True) + _cam_seg_pool(h)                   # campplus.py:96
            ctx = F.relu(F.conv1d(ctx, sd[p + '.cam_layer.linear1.weight'], sd[p + '.cam_layer.linear1.bias']))
            g = torch.sigmoid(F.conv1d(ctx, sd[p + '.cam_layer.linear2.weight'], sd[p + '.cam_layer.linear2.bias']))
            x = torch.cat([x, y * g], dim=1)
        p = f'xvector.transit{bi}'
        x = F.conv1d(F.relu(_bn(sd, p + '.nonlinear.batchnorm', x)), sd[p + '.linear.weight'])
    x = F.relu(_bn(sd, 'xvector.out_nonlinear.batchnorm', x))
    x = torch.cat([x.mean(dim=-1), x.std(dim=-1, unbiased=True)], dim=-1)       # campplus.py:27-33
    x = F.conv1d(x.unsqueeze(-1), sd['xvector.dense.linear.weight']).squeeze(-1)
    return _bn(sd, 'xvector.dense.nonlinear.batchnorm', x)


# ---------------------------------------------------------------------------------------------
# ResNetSE (resnet_se.py:65-145), pooling_type 'ASP'
# ---------------------------------------------------------------------------------------------
def resnetse_param_shapes(input_size, layers=(3, 4, 6, 3), num_filters=(32, 64, 128, 256), embd_dim=192,
                          pooling_type='ASP'):
    d = OrderedDict()
    d['conv1.weight'] = (num_filters[0], 1, 3, 3)
    _bn_shapes(d, 'bn1', num_filters[0])
    inpl = num_filters[0]
    for li, (nb, planes) in enumerate(zip(layers, num_filters), start=1):
        for b in range(nb):
            p = f'layer{li}.{b}'
            stride = 2 if (li > 1 and b == 0) else 1
            d[p + '.conv1.weight'] = (planes, inpl, 1, 1)
            _bn_shapes(d, p + '.bn1', planes)
            d[p + '.conv2.weight'] = (planes, planes, 3, 3)
            _bn_shapes(d, p + '.bn2', planes)
            d[p + '.conv3.weight'] = (planes * 2, planes, 1, 1)
            _bn_shapes(d, p + '.bn3', planes * 2)
            d[p + '.se.fc.0.weight'] = (planes * 2 // 8, planes * 2)
            d[p + '.se.fc.0.bias'] = (planes * 2 // 8,)
            d[p + '.se.fc.2.weight'] = (planes * 2, planes * 2 // 8)
            d[p + '.se.fc.2.bias'] = (planes * 2,)
            if b == 0 and (stride != 1 or inpl != planes * 2):
                d[p + '.downsample.0.weight'] = (planes * 2, inpl, 1, 1)
                _bn_shapes(d, p + '.downsample.1', planes * 2)
            inpl = planes * 2
    cat = num_filters[3] * 2 * (input_size // 8)
    width = _pool_shapes(d, 'pooling', pooling_type, cat, 128)
    _bn_shapes(d, 'bn2', width)
    d['linear.weight'] = (embd_dim, width)
    d['linear.bias'] = (embd_dim,)
    _bn_shapes(d, 'bn3', embd_dim)
    return d


def resnetse_forward(sd, x, layers=(3, 4, 6, 3), num_filters=(32, 64, 128, 256), embd_dim=192,
                     pooling_type='ASP'):
    x = x.transpose(2, 1).unsqueeze(1)
    x = F.relu(_bn(sd, 'bn1', F.conv2d(x, sd['conv1.weight'], padding=1)))
    for li, nb in enumerate(layers, start=1):
        for b in range(nb):
            p = f'layer{li}.{b}'
            stride = 2 if (li > 1 and b == 0) else 1
            out = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'])))
            out = F.relu(_bn(sd, p + '.bn2', F.conv2d(out, sd[p + '.conv2.weight'], stride=stride, padding=1)))
            out = _bn(sd, p + '.bn3', F.conv2d(out, sd[p + '.conv3.weight']))
            y = out.mean(dim=(2, 3))                                            # resnet_se.py:58-62
            y = F.relu(F.linear(y, sd[p + '.se.fc.0.weight'], sd[p + '.se.fc.0.bias']))
            y = torch.sigmoid(F.linear(y, sd[p + '.se.fc.2.weight'], sd[p + '.se.fc.2.bias']))
            out = out * y[:, :, None, None]
            res = x
            if (p + '.downsample.0.weight') in sd:
                res = _bn(sd, p + '.downsample.1', F.conv2d(x, sd[p + '.downsample.0.weight'], stride=stride))
            x = F.relu(out + res)
    x = x.reshape(x.shape[0], -1, x.shape[-1])                                  # resnet_se.py:139
    x = _bn(sd, 'bn2', pool_forward(sd, 'pooling', pooling_type, x))
    return _bn(sd, 'bn3', F.linear(x, sd['linear.weight'], sd['linear.bias']))


# ---------------------------------------------------------------------------------------------
# ERes2Net (eres2net.py:173-263), two_emb_layer False
# ---------------------------------------------------------------------------------------------
def _aff_shapes(d, p, channels, r=4):
    inter = int(channels // r)
    d[p + '.local_att.0.weight'] = (inter, channels * 2, 1, 1)
    d[p + '.local_att.0.bias'] = (inter,)
    _bn_shapes(d, p + '.local_att.1', inter)
    d[p + '.local_att.3.weight'] = (channels, inter, 1, 1)
    d[p + '.local_att.3.bias'] = (channels,)
    _bn_shapes(d, p + '.local_att.4', channels)


def _aff(sd, p, x, y):
    """eres2net.py:32-52."""
    a = F.conv2d(torch.cat((x, y), dim=1), sd[p + '.local_att.0.weight'], sd[p + '.local_att.0.bias'])
    a = F.silu(_bn(sd, p + '.local_att.1', a))
    a = _bn(sd, p + '.local_att.4', F.conv2d(a, sd[p + '.local_att.3.weight'], sd[p + '.local_att.3.bias']))
    a = 1.0 + torch.tanh(a)
    return x * a + y * (2.0 - a)


def _hardtanh20(x):
    return torch.clamp(x, 0.0, 20.0)                                            # eres2net.py:12-15


def eres2net_param_shapes(input_size, num_blocks=(3, 4, 6, 3), m_channels=32, mul_channel=1, expansion=2,
                          base_width=32, scale=2, embd_dim=192, two_emb_layer=False):
    assert not two_emb_layer
    d = OrderedDict()
    d['conv1.weight'] = (m_channels, 1, 3, 3)
    _bn_shapes(d, 'bn1', m_channels)
    inpl = m_channels
    for li, nb in enumerate(num_blocks, start=1):
        planes = m_channels * (2 ** (li - 1))
        fuse = li >= 3
        for b in range(nb):
            p = f'layer{li}.{b}'
            stride = 2 if (li > 1 and b == 0) else 1
            width = int(math.floor(planes * (base_width / 64.0)))
            d[p + '.conv1.weight'] = (width * scale, inpl, 1, 1)
            _bn_shapes(d, p + '.bn1', width * scale)
            for j in range(scale):
                d[f'{p}.convs.{j}.weight'] = (width, width, 3, 3)
            for j in range(scale):
                _bn_shapes(d, f'{p}.bns.{j}', width)
            if fuse:
                for j in range(scale - 1):
                    _aff_shapes(d, f'{p}.fuse_models.{j}', width)
            d[p + '.conv3.weight'] = (planes * expansion, width * scale, 1, 1)
            _bn_shapes(d, p + '.bn3', planes * expansion)
            if stride != 1 or inpl != planes * expansion:
                d[p + '.shortcut.0.weight'] = (planes * expansion, inpl, 1, 1)
                _bn_shapes(d, p + '.shortcut.1', planes * expansion)
            inpl = planes * expansion
    mc = m_channels * mul_channel
    d['layer1_downsample.weight'] = (mc * 4, mc * 2, 3, 3)
    d['layer2_downsample.weight'] = (mc * 8, mc * 4, 3, 3)
    d['layer3_downsample.weight'] = (mc * 16, mc * 8, 3, 3)
    _aff_shapes(d, 'fuse_mode12', mc * 4)
    _aff_shapes(d, 'fuse_mode123', mc * 8)
    _aff_shapes(d, 'fuse_mode1234', mc * 16)
    stats_dim = int(input_size / 8) * m_channels * 8
    d['seg_1.weight'] = (embd_dim, stats_dim * expansion * 2)
    d['seg_1.bias'] = (embd_dim,)
    return d


def _eres2net_stages(sd, x, num_blocks, m_channels, expansion, base_width, scale):
    """Stem + the four stages; returns [out1, out2, out3, out4] (eres2net.py:239-251)."""
    x = x.permute(0, 2, 1).unsqueeze(1)
    out = F.relu(_bn(sd, 'bn1', F.conv2d(x, sd['conv1.weight'], padding=1)))    # plain ReLU, eres2net.py:243
    outs = []
    for li, nb in enumerate(num_blocks, start=1):
        planes = m_channels * (2 ** (li - 1))
        width = int(math.floor(planes * (base_width / 64.0)))
        fuse = li >= 3
        for b in range(nb):
            p = f'layer{li}.{b}'
            stride = 2 if (li > 1 and b == 0) else 1
            h = _hardtanh20(_bn(sd, p + '.bn1', F.conv2d(out, sd[p + '.conv1.weight'], stride=stride)))
            spx = torch.split(h, width, 1)
            pieces = []
            for j in range(scale):
                if j == 0:
                    sp = spx[j]
                elif fuse:
                    sp = _aff(sd, f'{p}.fuse_models.{j - 1}', sp, spx[j])
                else:
                    sp = sp + spx[j]
                sp = _hardtanh20(_bn(sd, f'{p}.bns.{j}', F.conv2d(sp, sd[f'{p}.convs.{j}.weight'], padding=1)))
                pieces.append(sp)
            h = _bn(sd, p + '.bn3', F.conv2d(torch.cat(pieces, 1), sd[p + '.conv3.weight']))
            res = out
            if (p + '.shortcut.0.weight') in sd:
                res = _bn(sd, p + '.shortcut.1', F.conv2d(out, sd[p + '.shortcut.0.weight'], stride=stride))
            out = _hardtanh20(h + res)
        outs.append(out)
    return outs


def eres2net_forward(sd, x, num_blocks=(3, 4, 6, 3), m_channels=32, mul_channel=1, expansion=2,
                     base_width=32, scale=2, embd_dim=192, two_emb_layer=False):
    outs = _eres2net_stages(sd, x, num_blocks, m_channels, expansion, base_width, scale)
    o1, o2, o3, o4 = outs
    f12 = _aff(sd, 'fuse_mode12', o2, F.conv2d(o1, sd['layer1_downsample.weight'], stride=2, padding=1))
    f123 = _aff(sd, 'fuse_mode123', o3, F.conv2d(f12, sd['layer2_downsample.weight'], stride=2, padding=1))
    f1234 = _aff(sd, 'fuse_mode1234', o4, F.conv2d(f123, sd['layer3_downsample.weight'], stride=2, padding=1))
    mean = f1234.mean(dim=-1).flatten(start_dim=1)                              # pooling.py:140-148
    std = torch.sqrt(torch.var(f1234, dim=-1) + 1e-8).flatten(start_dim=1)
    return F.linear(torch.cat((mean, std), 1), sd['seg_1.weight'], sd['seg_1.bias'])


# ---------------------------------------------------------------------------------------------
# ERes2NetV2 (eres2net.py:383-456): same blocks as ERes2Net, only out3 -> layer3_ds -> fuse34 with out4
# ---------------------------------------------------------------------------------------------
def eres2netv2_param_shapes(input_size, num_blocks=(3, 4, 6, 3), m_channels=32, expansion=2, base_width=26, scale=2,
                            embd_dim=192, two_emb_layer=False):
    d = eres2net_param_shapes(input_size, num_blocks=num_blocks, m_channels=m_channels, mul_channel=1,
                              expansion=expansion, base_width=base_width, scale=scale, embd_dim=embd_dim)
    out = OrderedDict()
    for k, v in d.items():
        if k.startswith(('layer1_downsample', 'layer2_downsample', 'layer3_downsample', 'fuse_mode')):
            continue
        if k == 'seg_1.weight':
            out['layer3_ds.weight'] = (m_channels * 16, m_channels * 8, 3, 3)
            _aff_shapes(out, 'fuse34', m_channels * 16)
        out[k] = v
    return out


def eres2netv2_forward(sd, x, num_blocks=(3, 4, 6, 3), m_channels=32, expansion=2, base_width=26, scale=2, embd_dim=192,
                       two_emb_layer=False):
    outs = _eres2net_stages(sd, x, num_blocks, m_channels, expansion, base_width, scale)
    o3, o4 = outs[2], outs[3]
    f34 = _aff(sd, 'fuse34', o4, F.conv2d(o3, sd['layer3_ds.weight'], stride=2, padding=1))
    mean = f34.mean(dim=-1).flatten(start_dim=1)
    std = torch.sqrt(torch.var(f34, dim=-1) + 1e-8).flatten(start_dim=1)
    return F.linear(torch.cat((mean, std), 1), sd['seg_1.weight'], sd['seg_1.bias'])


# ---------------------------------------------------------------------------------------------
# Res2Net (res2net.py:89-174)
# ---------------------------------------------------------------------------------------------
def _res2net_blocks(m_channels, layers, base_width):
    inpl = m_channels
    for li, nb in enumerate(layers, start=1):
        planes = m_channels * (2 ** (li - 1))
        stride = 1 if li == 1 else 2
        for b in range(nb):
            first = b == 0
            ds = first and (stride != 1 or inpl != planes * 4)
            yield f'layer{li}.{b}', inpl, planes, int(math.floor(planes * (base_width / 64.0))), (stride if first else 1), \
                ('stage' if first else 'normal'), ds
            inpl = planes * 4


def res2net_param_shapes(input_size, m_channels=32, layers=(3, 4, 6, 3), base_width=32, scale=2, embd_dim=192,
                         pooling_type='ASP'):
    d = OrderedDict()
    d['conv1.weight'] = (m_channels, 1, 7, 7)
    _bn_shapes(d, 'bn1', m_channels)
    nums = 1 if scale == 1 else scale - 1
    for p, inpl, planes, width, stride, stype, ds in _res2net_blocks(m_channels, layers, base_width):
        d[p + '.conv1.weight'] = (width * scale, inpl, 1, 1)
        _bn_shapes(d, p + '.bn1', width * scale)
        for j in range(nums):
            d[f'{p}.convs.{j}.weight'] = (width, width, 3, 3)
        for j in range(nums):
            _bn_shapes(d, f'{p}.bns.{j}', width)
        d[p + '.conv3.weight'] = (planes * 4, width * scale, 1, 1)
        _bn_shapes(d, p + '.bn3', planes * 4)
        if ds:
            d[p + '.downsample.0.weight'] = (planes * 4, inpl, 1, 1)
            _bn_shapes(d, p + '.downsample.1', planes * 4)
    cat = m_channels * 8 * 4 * (input_size // base_width)
    width = _pool_shapes(d, 'pooling', pooling_type, cat, 128)
    _bn_shapes(d, 'bn2', width)
    d['linear.weight'] = (embd_dim, width)
    d['linear.bias'] = (embd_dim,)
    _bn_shapes(d, 'bn3', embd_dim)
    return d


def res2net_forward(sd, x, m_channels=32, layers=(3, 4, 6, 3), base_width=32, scale=2, embd_dim=192, pooling_type='ASP'):
    x = x.transpose(2, 1).unsqueeze(1)
    x = F.relu(_bn(sd, 'bn1', F.conv2d(x, sd['conv1.weight'], stride=3, padding=1)))        # res2net.py:100,144-146
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    nums = 1 if scale == 1 else scale - 1
    for p, inpl, planes, width, stride, stype, ds in _res2net_blocks(m_channels, layers, base_width):
        out = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'])))
        spx = torch.split(out, width, 1)
        pieces = []
        for j in range(nums):                                                                  # res2net.py:58-69
            sp = spx[j] if (j == 0 or stype == 'stage') else sp + spx[j]
            sp = F.relu(_bn(sd, f'{p}.bns.{j}', F.conv2d(sp, sd[f'{p}.convs.{j}.weight'], stride=stride, padding=1)))
            pieces.append(sp)
        if scale != 1:
            pieces.append(spx[nums] if stype == 'normal' else F.avg_pool2d(spx[nums], kernel_size=3, stride=stride, padding=1))
        out = _bn(sd, p + '.bn3', F.conv2d(torch.cat(pieces, 1), sd[p + '.conv3.weight']))
        res = x
        if ds:
            res = _bn(sd, p + '.downsample.1', F.conv2d(x, sd[p + '.downsample.0.weight'], stride=stride))
        x = F.relu(out + res)
    x = x.reshape(x.shape[0], -1, x.shape[-1])
    x = _bn(sd, 'bn2', pool_forward(sd, 'pooling', pooling_type, x))
    return _bn(sd, 'bn3', F.linear(x, sd['linear.weight'], sd['linear.bias']))


# ---------------------------------------------------------------------------------------------
# registry (mvector/models/__init__.py:15-21 builds by class name with **model_args)
# ---------------------------------------------------------------------------------------------
MODELS = {
    'EcapaTdnn': (ecapa_param_shapes, ecapa_forward),
    'TDNN': (tdnn_param_shapes, tdnn_forward),
    'CAMPPlus': (campplus_param_shapes, campplus_forward),
    'ResNetSE': (resnetse_param_shapes, resnetse_forward),
    'ERes2Net': (eres2net_param_shapes, eres2net_forward),
    'Res2Net': (res2net_param_shapes, res2net_forward),
    'ERes2NetV2': (eres2netv2_param_shapes, eres2netv2_forward),
}


def param_shapes(model, input_size, **model_args):
    return MODELS[model][0](input_size, **model_args)


def forward(model, sd, feats, **model_args):
    """Eval-mode backbone forward on CPU fp32: feats [B,T,F] -> [B, embd_dim]."""
    with torch.no_grad():
        return MODELS[model][1](sd, torch.as_tensor(feats, dtype=torch.float32), **model_args)


#: Conv/linear weight gain used for FULL-SIZE parity tests.  With plain He init the deep residual 2-D nets are
#: chaotic (a 1e-3 relative input perturbation moves the ERes2Net embedding by 7e-2..2e-1, and the fp32 CPU forward
#: itself differs from fp64 by 1e-4 (6.6M) / 4e-3 (55M) -- even changing the CPU thread count moves it by 2e-3), so a
#: 1e-4 parity gate would measure noise.  These gains bring the input->embedding amplification to O(1) (like a trained
#: net) and the fp32-vs-fp64 floor to <= 2e-6 while keeping every layer's contribution visible (measured in round 1).
CONDITIONED_GAIN = {'EcapaTdnn': 1.0, 'TDNN': 1.0, 'CAMPPlus': 1.0, 'ResNetSE': 0.8, 'ERes2Net': 0.7, 'Res2Net': 0.8, 'ERes2NetV2': 0.7}


def random_state_dict(model, input_size, seed=0, gain=1.0, **model_args):
    """Seeded weights with randomised BN statistics/affine (a fresh BN is near-identity
    and would hide BN bugs).  Conv/linear weights ~ N(0, gain^2 * 2/fan_in) (He), biases ~ N(0, 0.1^2), BN weight in
    [0.5, 1.5], BN bias ~ N(0, 0.2^2), running_mean ~ N(0, 0.2^2), running_var in [0.5, 1.5]."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(model, input_size, **model_args).items():
        if name.endswith('num_batches_tracked'):
            sd[name] = torch.tensor(100, dtype=torch.long)
        elif name.endswith('running_var'):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith('running_mean'):
            sd[name] = torch.randn(shape, generator=g) * 0.2
        elif len(shape) == 1 and name.endswith('.weight'):                      # BN affine weight
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:                                                    # biases
            sd[name] = torch.randn(shape, generator=g) * (0.2 if '.bias' in name and _is_bn_bias(name) else 0.1)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[name] = torch.randn(shape, generator=g) * (gain * math.sqrt(2.0 / fan_in))
    return sd


def _is_bn_bias(name):
    return any(t in name for t in ('norm.', 'bn', 'batchnorm', 'local_att.1', 'local_att.4', 'shortcut.1',
                                   'downsample.1'))
