"""FeaturePool / FeatureGroup / EmbdFeatureGroup (mirror of /root/reference/handyrec/features/group.py:14-516)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple, Union

import numpy as np
import torch

from ..keras_lite import Input, KTensor, Lambda, Layer
from ..layers import CustomEmbedding, SequencePoolingLayer, ValueRows, ValueTable
from ..layers.core import Dense
from ..layers.utils import concat
from .type import DenseFeature, Feature, SparseFeature, SparseSeqFeature
from .utils import split_features


class FeaturePool:
    """Name-keyed registries of shared input / embedding / pooling layers (group.py:14-137)."""

    def __init__(self, pre_embd: Dict[str, np.ndarray] = None) -> None:
        self.input_layers = OrderedDict()
        self.embd_layers = OrderedDict()
        self.pool_layers = OrderedDict()
        self.pre_embd = pre_embd

    def init_input(self, name: str, params: Dict) -> KTensor:
        if name in self.input_layers:
            layer = self.input_layers[name]
            curr_dim, curr_dtype = layer.shape[-1], layer.dtype.name
            new_dim, new_dtype = params["shape"][0], getattr(params["dtype"], "name", params["dtype"])
            if curr_dim != new_dim or curr_dtype != new_dtype:  # group.py:69-74
                raise AttributeError(f"Params of {name} conflict with an existed input layer!\n\t existed shape:{curr_dim}, new shape:{new_dim}\n"
                                     f"\t existed dtype:{curr_dtype}, new dtype:{new_dtype}")
        else:
            layer = Input(**params)
            self.input_layers[name] = layer
        return layer

    def init_embd(self, name: str, params: Dict) -> CustomEmbedding:
        if name in self.embd_layers:
            layer = self.embd_layers[name]
            if layer.input_dim != params["input_dim"] or layer.output_dim != params["output_dim"]:  # group.py:107-112
                raise AttributeError(f"Params of {name} conflict with an existed embedding layer!\n\t existed input_dim:{layer.input_dim}, "
                                     f"new input_dim:{params['input_dim']}\n\t existed dtype:{layer.output_dim}, new dtype:{params['output_dim']}")
            if params["mask_zero"] and not layer.mask_zero:
                # group.py:113-121: a NEW, separately weighted table replaces the unmasked one (the earlier group keeps the old object)
                layer._name = layer._name + str(np.random.randint(1e5))
                layer = CustomEmbedding(**params)
        else:
            layer = CustomEmbedding(**params)
        self.embd_layers[name] = layer
        return layer

    def init_pool(self, name: str, params: Dict) -> Layer:
        if name in self.pool_layers:
            layer = self.pool_layers[name]
            if layer.method != params["method"]:  # group.py:129-133
                raise AttributeError(f"Params of {name} conflict with an existed pooling layer!\n\t existed pooling method:{layer.method}, "
                                     f"new method:{params['method']}")
        else:
            layer = SequencePoolingLayer(**params)
            self.pool_layers[name] = layer
        return layer


class _L2:
    """Stand-in for keras.regularizers.l2(l2_reg): `l2 * sum(w**2)`."""

    def __init__(self, l2):
        self.l2 = float(l2)


class FeatureGroup:
    """Inputs + embedding tables of a list of features; `embedding_lookup` runs lookup -> pool (group.py:164-336)."""

    def __init__(self, name: str, features: List[Feature], feature_pool: FeaturePool, l2_embd: float = 1e-6):
        self.name = name
        self.features = features
        self.feat_pool = feature_pool
        self.input_layers = self.construct_inputs(features, feature_pool)
        self.embd_layers = self.construct_embds(features, feature_pool, l2_embd)

    @classmethod
    def construct_inputs(cls, features, feature_pool) -> "OrderedDict[str, KTensor]":
        input_layers = OrderedDict()
        for feat in features:
            if isinstance(feat, SparseSeqFeature):
                dim = feat.seq_len
            elif isinstance(feat, DenseFeature):
                dim = feat.dim
            else:
                dim = 1
            input_layers[feat.name] = feature_pool.init_input(feat.name, {"name": feat.name, "shape": (dim,), "dtype": feat.dtype})
        return input_layers

    @classmethod
    def construct_embds(cls, features, feature_pool, l2_reg) -> "OrderedDict[str, CustomEmbedding]":
        embd_layers = OrderedDict()
        _, sparse, sparse_seq = split_features(features)
        mask_embds = [x.unit.name for x in sparse_seq.values()]  # units of sequence features need mask_zero (group.py:272-273)
        todo = list(sparse.values())
        for feat in sparse_seq.values():
            if feat.unit.name not in sparse:
                todo.append(feat.unit)
        for feat in todo:
            weights = None
            if feature_pool.pre_embd and feat.name in feature_pool.pre_embd:
                weights = [feature_pool.pre_embd[feat.name]]
            params = {"name": "embd_" + feat.name, "input_dim": feat.vocab_size, "output_dim": feat.embdding_dim,
                      "embeddings_regularizer": _L2(l2_reg), "trainable": feat.trainable, "weights": weights, "mask_zero": feat.name in mask_embds}
            embd_layers[feat.name] = feature_pool.init_embd(feat.name, params)
        return embd_layers

    def embedding_lookup(self, pool_method: str = "mean") -> Tuple[List[KTensor], List[KTensor]]:
        """-> (dense inputs, embeddings): sparse features first, then pooled sequence features, each (B,1,D) (group.py:299-336)."""
        dense, sparse, sparse_seq = split_features(self.features)
        dense_output = [self.input_layers[k] for k in dense]
        embd_outputs = OrderedDict()
        for name in sparse:
            embd_outputs[name] = self.embd_layers[name](self.input_layers[name])
        for feat in sparse_seq.values():
            sparse_embd = self.embd_layers[feat.unit.name]
            seq_input = self.input_layers[feat.name]
            pool_layer = self.feat_pool.init_pool(feat.name + "_POOL", {"name": feat.name + "_POOL", "method": pool_method})
            embd_outputs[feat.name] = pool_layer(sparse_embd(seq_input))
        return dense_output, list(embd_outputs.values())


class EmbdFeatureGroup:
    """Concatenated embeddings of a whole item catalogue (group.py:339-516)."""

    def __init__(self, name: str, id_name: str, features: List[Feature], feature_pool: FeaturePool, value_dict: Dict[str, np.ndarray],
                 embd_dim: int = None, l2_embd: float = 1e-6, pool_method: str = "mean"):
        if id_name not in [x.name for x in features]:
            raise ValueError("`id_name` should be the name of a feature in `features`")
        for feat in features:
            if isinstance(feat, SparseSeqFeature) and not isinstance(feat.unit, SparseFeature) and not isinstance(feat.unit, EmbdFeatureGroup):
                raise ValueError("Only an `EmbdFeatureGroup` or a `SparseFeature` can be the unit of a `SparseSeqFeature`")
        self.name, self.id_name, self.feat_pool, self.embd_dim = name, id_name, feature_pool, embd_dim
        id_feat = {x.name: x for x in features}[id_name]
        self.id_input = feature_pool.init_input(id_name, {"name": id_name, "shape": (1,), "dtype": id_feat.dtype})
        self.embd_layers = FeatureGroup.construct_embds(features, feature_pool, l2_embd)
        self.features = features
        self._value_dict = value_dict
        self._layers = {}
        for feat in features:
            self._layers[feat.name] = ValueTable(value_dict[feat.name], name=feat.name + "_list", dtype=feat.dtype)
            if isinstance(feat, SparseSeqFeature):
                self._layers[feat.name + "_pool"] = SequencePoolingLayer(pool_method, name=feat.name + "_" + pool_method)
        if self.embd_dim is not None:
            self._output_layer = Dense(self.embd_dim, name="reduce_dim")

    def get_embd(self, index, compress: bool = False):
        """(n_items, p*d+q) matrix of every item's features (group.py:439-484); `index` only anchors the graph."""
        return self._features_to_vector(index, self._layers, compress)

    def _features_to_vector(self, index, source, compress: bool):
        """The arithmetic of `get_embd` (group.py:464-483) on whatever `source[name](index)` yields per feature: the whole value
        list (ValueTable -> every catalogue row) or the rows picked by `index` (ValueRows -> lazy evaluation)."""
        embd_outputs = OrderedDict()
        dense, sparse, sparse_seq = split_features(self.features)
        for name in dense:  # a dense feature is treated as a 1-d embedding (group.py:464-468)
            embd = source[name](index)
            if len(embd.shape) == 1:
                embd = Lambda(lambda t: t.unsqueeze(-1).to(torch.float32), lambda s: tuple(s) + (1,), name=name + "_expand")(embd)
            else:
                embd = Lambda(lambda t: t.to(torch.float32), lambda s: tuple(s), name=name + "_cast")(embd)
            embd_outputs[name] = embd
        for name in sparse:
            embd_outputs[name] = self.embd_layers[name](source[name](index))  # (n, d)
        for name, feat in sparse_seq.items():
            embd_seq = self.embd_layers[feat.unit.name](source[name](index))
            pooled = self._layers[name + "_pool"](embd_seq)
            embd_outputs[name] = Lambda(lambda t: t.squeeze(1), lambda s: (s[0], s[2]), name=name + "_squeeze")(pooled)  # (n, d)
        output = concat([], list(embd_outputs.values()))
        if compress:
            output = self._output_layer(output)
        return output

    def get_embd_rows(self, index, compress: bool = False):
        """Feature vectors of the items in `index` (m, 1) ONLY: gather each feature's values for those ids, then embed / pool /
        concatenate exactly like `get_embd`.  `get_embd` evaluates all n catalogue rows every step although a sampled-softmax step
        reads B + num_sampled of them (10 M rows -> 1.6 GB written per step); row-wise the two are the same function."""
        if not hasattr(self, "_row_layers"):
            self._row_layers = {f.name: ValueRows(self._value_dict[f.name], name=f.name + "_rows", dtype=f.dtype) for f in self.features}
        return self._features_to_vector(index, self._row_layers, compress)

    def row_tower(self, compress: bool = False, head=None):
        """A callable `ids (m,1) -> (m, width)` over `get_embd_rows` (+ an optional layer `head`, e.g. DSSM's item DNN) that
        `SampledSoftmaxLayer(item_rows=...)` evaluates for the rows a step needs."""
        from ..keras_lite import Model

        feat = {x.name: x for x in self.features}[self.id_name]
        idx = Input(shape=(1,), name=self.id_name + "_rows", dtype=feat.dtype)
        out = self.get_embd_rows(idx, compress)
        if head is not None:
            out = head(out)
        tower = Model(inputs=[idx], outputs=out, name=self.name + "_row_tower")

        def rows(ids, training=False):
            return tower({idx.name: ids}, training=training)

        rows.layers = tower.layers + ([self._output_layer] if compress and self.embd_dim is not None else [])
        rows.width = out.shape[-1]
        return rows

    def lookup(self, index, compress: bool = False):
        """Feature vectors of the given ids (group.py:486-506)."""
        embedding = self.get_embd(index, compress)
        last = index.shape[-1]
        width = embedding.shape[-1]

        def gather(emb, idx):
            out = emb[idx.long()]
            return out.squeeze(1) if last == 1 else out

        shape_fn = (lambda se, si: (si[0], width)) if last == 1 else (lambda se, si: tuple(si) + (width,))
        return Lambda(gather, shape_fn, name=self.name + "_lookup")([embedding, index])

    def __call__(self, seq_input):
        output = self.lookup(seq_input, compress=True)
        width = output.shape[-1]
        mask = Lambda(lambda idx: (idx != 0).unsqueeze(-1).expand(*idx.shape, width), lambda s: tuple(s) + (width,), dtype_fn=lambda i: "bool",
                      name=self.name + "_mask")(seq_input)
        return output, mask
