"""Test-infrastructure shim for `omegaconf` (absent from this image): a YAML-backed dict is all the reference needs
(pipeline.py:23-25,154,270; checkpoint.py:110,117). Used ONLY by oracle/gen_golden.py."""
import yaml


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def create(obj=None):
        return DictConfig(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return dict(cfg)

    @staticmethod
    def to_yaml(cfg, resolve=True):
        return yaml.safe_dump(dict(cfg)) if cfg else ''

    @staticmethod
    def save(config, f):
        yaml.safe_dump(dict(config), f)

    @staticmethod
    def load(f):
        return DictConfig(yaml.safe_load(f) or {})
