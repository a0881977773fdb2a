"""Config node used by ``Detector.get_default_config``. The reference uses ``yacs.config.CfgNode``
(src/models.py:9, 406-431); when yacs is installed that class is used, otherwise this minimal stand-in with the
subset of behaviour the Detector needs (attribute access, ``in``, nested nodes, ``new_allowed``)."""
try:  # pragma: no cover - depends on the environment
    from yacs.config import CfgNode as CN
except ImportError:
    class CN(dict):
        def __init__(self, init_dict=None, key_list=None, new_allowed=False):
            super().__init__()
            self.__dict__["_new_allowed"] = new_allowed
            for k, v in (init_dict or {}).items():
                self[k] = CN(v, new_allowed=new_allowed) if isinstance(v, dict) and not isinstance(v, CN) else v

        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name)

        def __setattr__(self, name, value):
            self[name] = value

        def clone(self):
            out = CN(new_allowed=self.__dict__.get("_new_allowed", False))
            for k, v in self.items():
                out[k] = v.clone() if isinstance(v, CN) else (list(v) if isinstance(v, list) else v)
            return out

        def merge_from_other_cfg(self, other):
            for k, v in other.items():
                if isinstance(v, dict) and isinstance(self.get(k), CN):
                    self[k].merge_from_other_cfg(v)
                else:
                    self[k] = v

        def is_new_allowed(self):
            return self.__dict__.get("_new_allowed", False)
