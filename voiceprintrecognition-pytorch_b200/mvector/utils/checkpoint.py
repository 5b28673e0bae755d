"""Weight ingestion.  Mirrors the inference half of mvector/utils/checkpoint.py:11-51 (``load_pretrained``): a plain
``state_dict`` pickle whose backbone keys carry the ``0.`` prefix of ``nn.Sequential(backbone)`` (predict.py:55);
classifier keys (``1.*`` from training checkpoints) are reported as unexpected and ignored."""
import os

import torch
from loguru import logger


def load_pretrained(model, pretrained_model, use_gpu=True):
    if pretrained_model is None:
        return model
    if os.path.isdir(pretrained_model):
        pretrained_model = os.path.join(pretrained_model, 'model.pth')
    assert os.path.exists(pretrained_model), f"{pretrained_model} 模型不存在！"
    try:
        state = torch.load(pretrained_model, map_location='cpu', weights_only=True)
    except Exception as e:
        # Checkpoints pickled with non-tensor payloads need the unrestricted unpickler, which executes code from the file.
        # The reference always loads that way (checkpoint.py:29, torch.load default of its torch pin); here it is an
        # explicit opt-in so that a drop-in never runs a pickle payload silently.
        if os.environ.get('VPB_ALLOW_UNSAFE_PICKLE') != '1':
            raise RuntimeError(f'{pretrained_model} is not a plain tensor state_dict ({type(e).__name__}: {e}); set '
                               'VPB_ALLOW_UNSAFE_PICKLE=1 to load it with the unrestricted pickle loader') from e
        logger.warning(f'{pretrained_model}: falling back to torch.load(weights_only=False) (VPB_ALLOW_UNSAFE_PICKLE=1)')
        state = torch.load(pretrained_model, map_location='cpu', weights_only=False)
    shapes = model.param_shapes()
    for name in list(state.keys()):
        key = name[2:] if name.startswith('0.') else name
        if key in shapes and list(state[name].shape) != list(shapes[key]):
            logger.warning(f'{name} not used, shape {list(state[name].shape)} unmatched with {list(shapes[key])} in model.')
            state.pop(name)
    missing, unexpected = model.load_state_dict(state, strict=False)
    if len(unexpected) > 0:
        logger.warning('Unexpected key(s) in state_dict: {}. '.format(', '.join('"{}"'.format(k) for k in unexpected)))
    if len(missing) > 0:
        logger.warning('Missing key(s) in state_dict: {}. '.format(', '.join('"{}"'.format(k) for k in missing)))
    logger.info('成功加载预训练模型：{}'.format(pretrained_model))
    return model
