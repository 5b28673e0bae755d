"""Config helpers with the reference's names and behaviour (mvector/utils/utils.py:9-54): ``print_arguments`` logs an
argparse namespace and/or a (up to three levels deep) config dict, ``add_arguments`` registers one CLI flag,
``dict_to_object`` turns nested dicts into attribute-accessible ``Dict`` objects."""
from loguru import logger

_RULE = '-' * 48
_TRUE = frozenset(('y', 'yes', 't', 'true', 'on', '1'))
_FALSE = frozenset(('n', 'no', 'f', 'false', 'off', '0'))


def _log_tree(node, depth=0, max_depth=2):
    """Nested dict -> one log line per leaf, tab-indented; dicts below max_depth are printed inline like the reference."""
    for key, value in node.items():
        if isinstance(value, dict) and depth < max_depth:
            logger.info('\t' * depth + f'{key}:')
            _log_tree(value, depth + 1, max_depth)
        else:
            logger.info('\t' * depth + f'{key}: {value}')


def print_arguments(args=None, configs=None, title=None):
    if args:
        logger.info('----------- 额外配置参数 -----------')
        for name in sorted(vars(args)):
            logger.info('%s: %s' % (name, getattr(args, name)))
        logger.info(_RULE)
    if configs:
        logger.info(f'----------- {title or "配置文件参数"} -----------')
        _log_tree(configs)
        logger.info(_RULE)


def _to_bool(text):
    word = str(text).lower()
    if word in _TRUE:
        return 1
    if word in _FALSE:
        return 0
    raise ValueError(f'invalid truth value {text!r}')


def add_arguments(argname, type, default, help, argparser, **kwargs):
    argparser.add_argument('--' + argname, default=default, type=_to_bool if type == bool else type,
                           help=help + ' 默认: %(default)s.', **kwargs)


class Dict(dict):
    """dict whose keys are also attributes (missing attribute -> KeyError, as in the reference)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def dict_to_object(dict_obj):
    if isinstance(dict_obj, dict):
        return Dict((k, dict_to_object(v)) for k, v in dict_obj.items())
    return dict_obj
