"""Unpickle target for checkpoints that embed a ``parse_config.ConfigParser``
(E2VID+, FireNet+, HyperE2VID, ET-Net; reference parse_config.py:1-22).

``install()`` registers this module under the top-level name ``parse_config`` so
``torch.load(..., weights_only=False)`` finds the class the pickle refers to.
"""
import sys


class ConfigParser:
    def __init__(self, config):
        self._config = config

    def init_obj(self, name, module, *args, **kwargs):
        module_name = self[name]['type']
        module_args = dict(self[name]['args'])
        assert all([k not in module_args for k in kwargs]), 'Overwriting kwargs given in config file is not allowed'
        module_args.update(kwargs)
        return getattr(module, module_name)(*args, **module_args)

    def __getitem__(self, name):
        return self.config[name]

    @property
    def config(self):
        return self._config


def install():
    sys.modules.setdefault('parse_config', sys.modules[__name__])
