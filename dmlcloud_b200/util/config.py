"""Config container.  The reference passes an OmegaConf object around (pipeline.py:21-27, checkpoint.py:105-117);
omegaconf is an optional dependency here: when it is importable it is used unchanged, otherwise a YAML-backed
attribute dict provides the four calls the path needs (create / to_container / to_yaml / save / load)."""
import yaml

try:  # pragma: no cover - not installed in the build image
    from omegaconf import OmegaConf as _OmegaConf
except ImportError:
    _OmegaConf = None


class DictConfig(dict):
    """dict with attribute access; nested dicts are wrapped on the way in."""

    def __init__(self, data=None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, DictConfig):
            value = DictConfig(value)
        super().__setitem__(key, value)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value


def _plain(cfg):
    if isinstance(cfg, dict):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg


class _YamlConf:
    @staticmethod
    def create(obj=None):
        return DictConfig(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _plain(cfg)

    @staticmethod
    def to_yaml(cfg, resolve=True):
        return yaml.safe_dump(_plain(cfg)) if cfg else ''

    @staticmethod
    def save(config, f):
        yaml.safe_dump(_plain(config), f)

    @staticmethod
    def load(f):
        return DictConfig(yaml.safe_load(f) or {})

    @staticmethod
    def is_config(obj):
        return isinstance(obj, DictConfig)


if _OmegaConf is not None:  # pragma: no cover
    class Conf(_OmegaConf):
        @staticmethod
        def is_config(obj):
            return _OmegaConf.is_config(obj)
else:
    Conf = _YamlConf
