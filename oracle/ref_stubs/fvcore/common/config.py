"""Minimal stand-in for fvcore.common.config.CfgNode (attribute dict + YAML merge)."""
import yaml


class CfgNode(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(self, yaml.safe_load(f))

    @staticmethod
    def _merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                if k not in dst:
                    dst[k] = CfgNode()
                CfgNode._merge(dst[k], v)
            else:
                dst[k] = v
