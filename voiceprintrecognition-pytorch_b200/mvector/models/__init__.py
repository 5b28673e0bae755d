"""Backbone registry of the lowered path.

``build_model(input_size, configs)`` keeps the reference's contract (mvector/models/__init__.py:15-21): the class named by
``configs.model_conf.model`` (default CAMPPlus) is constructed with ``input_size`` plus ``**model_args``, so unknown yml
keys raise TypeError from the constructor exactly like the reference.  The lookup is an explicit table of the mirrors
that have a lowering (every backbone the shipped configs/*.yml name)."""
from loguru import logger

from . import campplus, ecapa_tdnn, eres2net, res2net, resnet_se, tdnn

_LOWERED = {
    'EcapaTdnn': ecapa_tdnn.EcapaTdnn,
    'TDNN': tdnn.TDNN,
    'CAMPPlus': campplus.CAMPPlus,
    'ResNetSE': resnet_se.ResNetSE,
    'ERes2Net': eres2net.ERes2Net,
    'ERes2NetV2': eres2net.ERes2NetV2,
    'Res2Net': res2net.Res2Net,
}
globals().update(_LOWERED)            # ``from mvector.models import EcapaTdnn`` keeps working
__all__ = ['build_model'] + sorted(_LOWERED)


def build_model(input_size, configs):
    name = configs.model_conf.get('model', 'CAMPPlus')
    kwargs = configs.model_conf.get('model_args', {})
    try:
        ctor = _LOWERED[name]
    except KeyError:
        raise AttributeError(f"module 'mvector.models' has no attribute '{name}' "
                             f"(lowered backbones: {', '.join(_LOWERED)})") from None
    backbone = ctor(input_size=input_size, **kwargs)
    logger.info(f'成功创建模型：{name}，参数为：{kwargs}')
    return backbone
