"""Minimal reader of HandyRec's YAML feature-group schema (the subset of /root/reference/handyrec/config.py:61-168 that the
shipped configs use), on PyYAML instead of python-box."""
from __future__ import annotations

from typing import Dict, Union

import numpy as np
import yaml

from .features import DenseFeature, EmbdFeatureGroup, FeatureGroup, FeaturePool, SparseFeature, SparseSeqFeature


class ConfigLoader:
    def __init__(self, config: Union[str, Dict]):
        assert isinstance(config, (str, dict)), "config must be a string or a dict"
        self.config = yaml.safe_load(open(config)) if isinstance(config, str) else config

    @staticmethod
    def _sparse(name, cfg, feature_dim):
        cfg = dict(cfg or {})
        cfg.setdefault("name", name)
        cfg.setdefault("vocab_size", feature_dim[name] if "vocab_size" not in cfg else cfg["vocab_size"])
        return SparseFeature(**cfg)

    def get_feature_group(self, feature_group_name, feature_pool, feature_dim=None, value_dict=None):
        fg = self.config["FeatureGroups"][feature_group_name]
        feats = []
        for n, c in (fg.get("DenseFeatures") or {}).items():
            c = dict(c or {})
            c.setdefault("name", n)
            feats.append(DenseFeature(**c))
        for n, c in (fg.get("SparseFeatures") or {}).items():
            feats.append(self._sparse(n, c, feature_dim))
        for n, c in (fg.get("SparseSeqFeatures") or {}).items():
            unit_name, unit_cfg = next(iter(c["unit"].items()))
            feats.append(SparseSeqFeature(self._sparse(unit_name, unit_cfg, feature_dim), c.get("name", n), c["seq_len"]))
        kw = {"name": fg["name"], "features": feats, "feature_pool": feature_pool}
        if fg.get("l2_embd"):
            kw["l2_embd"] = float(fg["l2_embd"])
        if fg["type"] == "EmbdFeatureGroup":
            kw.update(id_name=fg["id_name"], embd_dim=fg.get("embd_dim"), value_dict=value_dict)
            if fg.get("pool_method"):
                kw["pool_method"] = fg["pool_method"]
            return EmbdFeatureGroup(**kw)
        return FeatureGroup(**kw)

    def prepare_features(self, feature_dim=None, data=None, pretrained_embd=None) -> Dict:
        pool = FeaturePool(pretrained_embd)
        result = {"feature_pool": pool}
        for name, fg in self.config["FeatureGroups"].items():
            if fg["type"] == "EmbdFeatureGroup":
                names = [x for k in ("DenseFeatures", "SparseFeatures", "SparseSeqFeatures") for x in (fg.get(k) or {})]
                value_dict = {f: np.array(list(data["item"][f])) for f in names}
                result["value_dict"] = value_dict
                result[name] = self.get_feature_group(name, pool, feature_dim, value_dict)
            else:
                result[name] = self.get_feature_group(name, pool, feature_dim)
        return result
