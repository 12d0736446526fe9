"""Unpickle target for checkpoints that embed a ``parse_config.ConfigParser`` instance (E2VID+, FireNet+, HyperE2VID,
ET-Net store one under checkpoint['config']; reference parse_config.py:1-22).

The pickle stream names the class ``parse_config.ConfigParser`` and carries the instance dictionary ``{'_config': {...}}``;
unpickling restores that dictionary without calling ``__init__``.  ``install()`` publishes this module under the top-level
name ``parse_config`` so that ``torch.load(..., weights_only=False)`` resolves the class here.  The loader then asks the
restored object for the model: ``checkpoint['config'].init_obj('arch', model_module)`` (eval.py:146-152).
"""
import sys


class ConfigParser:
    def __init__(self, config):
        self._config = config

    @property
    def config(self):
        return self._config

    def __getitem__(self, key):
        return self._config[key]

    def init_obj(self, name, module, *args, **kwargs):
        """Instantiate ``module.<config[name]['type']>`` with the stored keyword arguments plus ``kwargs``; a keyword that is
        already stored may not be overridden (AssertionError, like the reference)."""
        entry = self._config[name]
        stored = dict(entry['args'])
        clash = sorted(set(stored).intersection(kwargs))
        if clash:
            raise AssertionError('Overwriting kwargs given in config file is not allowed')
        factory = getattr(module, entry['type'])
        return factory(*args, **stored, **kwargs)


def install():
    sys.modules.setdefault('parse_config', sys.modules[__name__])
