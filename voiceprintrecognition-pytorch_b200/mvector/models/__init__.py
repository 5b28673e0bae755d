"""Model registry.  Mirrors mvector/models/__init__.py:15-21: the class is looked up by name and constructed with
``input_size`` plus ``**model_args`` (so unknown yml keys raise TypeError exactly like the reference)."""
import importlib

from loguru import logger

from .campplus import CAMPPlus
from .ecapa_tdnn import EcapaTdnn
from .eres2net import ERes2Net, ERes2NetV2
from .res2net import Res2Net
from .resnet_se import ResNetSE
from .tdnn import TDNN

__all__ = ['build_model']


def build_model(input_size, configs):
    use_model = configs.model_conf.get('model', 'CAMPPlus')
    model_args = configs.model_conf.get('model_args', {})
    mod = importlib.import_module(__name__)
    if not hasattr(mod, use_model):
        raise AttributeError(f"module 'mvector.models' has no attribute '{use_model}' "
                             f"(lowered backbones: EcapaTdnn, TDNN, CAMPPlus, ResNetSE, ERes2Net, ERes2NetV2, Res2Net)")
    model = getattr(mod, use_model)(input_size=input_size, **model_args)
    logger.info(f'成功创建模型：{use_model}，参数为：{model_args}')
    return model
