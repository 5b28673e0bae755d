"""Config helpers.  Mirrors mvector/utils/utils.py:9-54 (print_arguments, add_arguments, Dict/dict_to_object)."""
from loguru import logger


def print_arguments(args=None, configs=None, title=None):
    if args:
        logger.info('----------- 额外配置参数 -----------')
        for arg, value in sorted(vars(args).items()):
            logger.info('%s: %s' % (arg, value))
        logger.info('------------------------------------------------')
    if configs:
        logger.info(f'----------- {title or "配置文件参数"} -----------')
        for a, v in configs.items():
            if isinstance(v, dict):
                logger.info(f'{a}:')
                for a1, v1 in v.items():
                    if isinstance(v1, dict):
                        logger.info(f'\t{a1}:')
                        for a2, v2 in v1.items():
                            logger.info(f'\t\t{a2}: {v2}')
                    else:
                        logger.info(f'\t{a1}: {v1}')
            else:
                logger.info(f'{a}: {v}')
        logger.info('------------------------------------------------')


def _strtobool(v):
    v = str(v).lower()
    if v in ('y', 'yes', 't', 'true', 'on', '1'):
        return 1
    if v in ('n', 'no', 'f', 'false', 'off', '0'):
        return 0
    raise ValueError(f'invalid truth value {v!r}')


def add_arguments(argname, type, default, help, argparser, **kwargs):
    type = _strtobool if type == bool else type
    argparser.add_argument('--' + argname, default=default, type=type, help=help + ' 默认: %(default)s.', **kwargs)


class Dict(dict):
    __setattr__ = dict.__setitem__
    __getattr__ = dict.__getitem__


def dict_to_object(dict_obj):
    if not isinstance(dict_obj, dict):
        return dict_obj
    inst = Dict()
    for k, v in dict_obj.items():
        inst[k] = dict_to_object(v)
    return inst
