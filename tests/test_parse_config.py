"""CPU: the ConfigParser unpickle shim (reference parse_config.py:1-22): a checkpoint pickled against the top-level module name
`parse_config` loads through it, and init_obj builds the class named in the stored config."""
import io
import pickle
import sys
import types

import pytest
import torch


def test_checkpoint_with_embedded_config_parser_round_trip(monkeypatch):
    from evreal_b200 import parse_config as shim
    # what the reference side pickles: an instance of parse_config.ConfigParser holding {'arch': {'type', 'args'}}
    fake = types.ModuleType('parse_config')
    exec("class ConfigParser:\n    def __init__(self, config):\n        self._config = config\n", fake.__dict__)
    fake.ConfigParser.__module__ = 'parse_config'
    monkeypatch.setitem(sys.modules, 'parse_config', fake)
    buf = io.BytesIO()
    torch.save({'config': fake.ConfigParser({'arch': {'type': 'Thing', 'args': {'a': 1, 'b': 2}}}), 'state_dict': {}}, buf)
    monkeypatch.delitem(sys.modules, 'parse_config')
    shim.install()
    assert sys.modules['parse_config'] is shim
    ck = torch.load(io.BytesIO(buf.getvalue()), map_location='cpu', weights_only=False)
    cfg = ck['config']
    assert isinstance(cfg, shim.ConfigParser) and cfg['arch']['type'] == 'Thing' and cfg.config['arch']['args'] == {'a': 1, 'b': 2}
    mod = types.SimpleNamespace(Thing=lambda *args, **kw: (args, kw))
    assert cfg.init_obj('arch', mod) == ((), {'a': 1, 'b': 2})
    assert cfg.init_obj('arch', mod, 7, c=3) == ((7,), {'a': 1, 'b': 2, 'c': 3})
    with pytest.raises(AssertionError):
        cfg.init_obj('arch', mod, a=5)
    assert pickle.loads(pickle.dumps(cfg)).config == cfg.config
