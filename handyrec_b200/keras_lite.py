"""A very small Keras-functional-API look-alike, just enough for HandyRec's feature groups, layers and model
constructors to be written the way the reference writes them (Input -> layers -> Model -> compile/fit/predict).

* Graph construction is symbolic (`KTensor`s with shapes/dtypes, no GPU needed); execution binds `Input`s to torch
  CUDA tensors and runs each layer's `call`, which launches libhrb200 kernels (handyrec_b200.autograd_ops).
* Masks travel with tensors like in Keras: `compute_mask` of the producing layer, handed to consumers whose `call`
  takes a `mask` argument.
* There is no CPU execution path: building a graph works anywhere, running one needs the CUDA library.
"""
from __future__ import annotations

import inspect
import itertools
import math
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

_uid = itertools.count()


def device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class ShardedTables:
    """Creation scope for multi-GPU models (the role of `tf.distribute.Strategy.scope()`), one process per GPU under torch.distributed:

        with ShardedTables():                      # or ShardedTables(comm, min_rows=..., seed=...)
            model = DeepFM(fm_group, dnn_group, ...)
            model.compile(Adam(1e-3), binary_crossentropy)
        model.fit(x_of_this_rank, y_of_this_rank, batch_size=per_rank_batch)

    Embedding tables with more than `min_rows` rows are created row-sharded: this rank allocates and initialises only rows
    r with r % N == rank (the counter-based initialiser gives every row the value it would have in the full table, so N ranks
    together hold exactly the table one GPU would have built from the same seed).  Such a table is read by the fused engine
    only (`compile` lowers the graph onto `ShardedDeepFMEngine`); a graph that does not lower raises instead of running the
    layer path on a shard.  Smaller tables and dense weights are replicated and broadcast from rank 0 at compile time.

    Collective calls: `fit` / `train_on_batch` / the first `predict` (they build or step the engine) must be made by every rank, with
    the same `batch_size` and the same number of batches -- like any data-parallel training loop.  `predict` after the first `fit`
    and `save_weights` / `load_weights` are rank-local."""

    active: Optional["ShardedTables"] = None

    def __init__(self, comm=None, min_rows: int = 131072, seed: int = 2022):
        if comm is None:
            from .sharded import TorchDistComm

            comm = TorchDistComm()
        self.comm, self.min_rows, self.seed, self.n_built = comm, int(min_rows), int(seed), 0

    def __enter__(self):
        ShardedTables.active = self
        return self

    def __exit__(self, *exc):
        ShardedTables.active = None
        return False

    def next_seed(self) -> int:
        self.n_built += 1
        return self.seed + 7919 * self.n_built


_DTYPES = {"int32": torch.int32, "int64": torch.int64, "float32": torch.float32, "bool": torch.bool, "uint8": torch.uint8}


def _tdtype(d):
    if isinstance(d, torch.dtype):
        return d
    return _DTYPES[getattr(d, "name", d)]


class _DType:
    """Minimal stand-in for tf.DType (`.name`, `.is_integer`)."""

    def __init__(self, name):
        self.name = name

    @property
    def is_integer(self):
        return self.name.startswith("int") or self.name.startswith("uint")

    def __eq__(self, o):
        return self.name == getattr(o, "name", o)

    def __repr__(self):
        return self.name


class KTensor:
    """Symbolic tensor: shape with None for the batch axis, dtype, producing node."""

    def __init__(self, shape, dtype, node=None, index=0, name=None):
        self.shape = tuple(shape)
        self.dtype = _DType(getattr(dtype, "name", dtype) if not isinstance(dtype, torch.dtype) else str(dtype).split(".")[-1])
        self.node, self.index, self.name = node, index, name
        self._keras_mask: Optional["KTensor"] = None

    def __add__(self, other):
        return Lambda(lambda a, b: a + b, lambda sa, sb: sa, name="add")([self, other])

    def __mul__(self, other):
        if isinstance(other, KTensor):
            return Lambda(lambda a, b: a * b, lambda sa, sb: sa, name="multiply")([self, other])
        c = float(other)
        return Lambda(lambda a: a * c, lambda sa: sa, name="scale")(self)

    __rmul__ = __mul__

    def __repr__(self):
        return f"<KTensor {self.name or ''} shape={self.shape} dtype={self.dtype}>"


class Node:
    def __init__(self, layer, inputs, kwargs):
        self.layer, self.inputs, self.kwargs = layer, inputs, kwargs
        self.outputs: List[KTensor] = []
        self.id = next(_uid)


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for e in x:
            out += _flatten(e)
        return out
    return [x]


def _map_structure(fn, x):
    if isinstance(x, (list, tuple)):
        return [_map_structure(fn, e) for e in x]
    return fn(x)


class Layer:
    def __init__(self, name: Optional[str] = None, trainable: bool = True, dtype=None, **kwargs):
        self.name = name or f"{type(self).__name__.lower()}_{next(_uid)}"
        self._name = self.name
        self.trainable = trainable
        self.built = False
        self._weights: "OrderedDict[str, torch.nn.Parameter]" = OrderedDict()
        self._weight_l2: Dict[str, float] = {}
        self._sublayers: List["Layer"] = []
        sig = inspect.signature(self.call)
        self._call_takes_mask = "mask" in sig.parameters
        self._call_takes_training = "training" in sig.parameters or any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())

    # ---- weights -------------------------------------------------------------------------
    def add_weight(self, name=None, shape=None, initializer="glorot_uniform", dtype=None, trainable=True, l2: float = 0.0, value=None):
        shape = tuple(int(s) for s in shape)
        if value is not None:
            t = torch.as_tensor(np.asarray(value), dtype=torch.float32).reshape(shape).clone()
        else:
            t = _init(initializer, shape)
        p = torch.nn.Parameter(t.to(device()), requires_grad=bool(trainable and self.trainable))
        key = name or f"w{len(self._weights)}"
        self._weights[key] = p
        if l2:
            self._weight_l2[key] = float(l2)
        return p

    def _track(self, layer: "Layer") -> "Layer":
        self._sublayers.append(layer)
        return layer

    @property
    def weights(self) -> List[torch.nn.Parameter]:
        out = list(self._weights.values())
        for l in self._sublayers:
            out += l.weights
        return out

    def weights_with_l2(self) -> List[Tuple[torch.nn.Parameter, float]]:
        out = [(p, self._weight_l2.get(k, 0.0)) for k, p in self._weights.items()]
        for l in self._sublayers:
            out += l.weights_with_l2()
        return out

    def get_weights(self):
        return [w.detach().cpu().numpy() for w in self.weights]

    def set_weights(self, values):
        ws = self.weights
        assert len(ws) == len(values), f"{self.name}: expected {len(ws)} arrays, got {len(values)}"
        for w, v in zip(ws, values):
            w.data.copy_(torch.as_tensor(np.asarray(v), dtype=w.dtype).reshape(w.shape))

    # ---- protocol -------------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def compute_mask(self, inputs, mask=None):
        return None

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    def output_dtype(self, inputs):
        return "float32"

    def __call__(self, inputs, **kwargs):
        flat = _flatten(inputs)
        symbolic = any(isinstance(t, KTensor) for t in flat)
        shapes = _map_structure(lambda t: tuple(t.shape), inputs)
        if not self.built:
            self.build(shapes)
            self.built = True
        if symbolic:
            node = Node(self, inputs, kwargs)
            out_shape = self.compute_output_shape(shapes)
            multi = isinstance(out_shape, list)
            outs = []
            for i, s in enumerate(out_shape if multi else [out_shape]):
                t = KTensor(s, self.output_dtype(inputs), node, i, name=f"{self.name}:{i}")
                outs.append(t)
            node.outputs = outs
            in_masks = _map_structure(lambda t: t._keras_mask if isinstance(t, KTensor) else None, inputs)
            has_mask = self.symbolic_has_mask(inputs, in_masks)
            for t in outs:
                t._keras_mask = KTensor(t.shape, "bool", node, -1, name=f"{self.name}:mask") if has_mask else None
            return outs if multi else outs[0]
        # eager: inputs are torch tensors; masks ride on the `_keras_mask` attribute
        in_masks = _map_structure(lambda t: getattr(t, "_keras_mask", None), inputs)
        return self._run(inputs, in_masks, kwargs.get("training", False))

    def symbolic_has_mask(self, inputs, in_masks) -> bool:
        """Whether compute_mask will return a mask (decidable at graph-build time for every layer used here)."""
        return False

    def _run(self, inputs, in_masks, training):
        kw = {}
        single_mask = in_masks  # same nesting as `inputs` (a list for multi-input layers)
        if self._call_takes_mask:
            kw["mask"] = single_mask
        if self._call_takes_training:
            kw["training"] = training
        out = self.call(inputs, **kw)
        mask = self.compute_mask(inputs, single_mask)
        if isinstance(out, torch.Tensor):
            out._keras_mask = mask
        return out


def _init(initializer, shape):
    name = initializer if isinstance(initializer, str) else getattr(initializer, "__name__", "zeros")
    name = name.lower()
    if name in ("zeros", "zero"):
        return torch.zeros(shape)
    if name in ("ones", "one"):
        return torch.ones(shape)
    if name in ("uniform", "random_uniform", "randomuniform"):
        n = 1
        for d in shape:
            n *= d
        if n >= (1 << 22) and device().type == "cuda":  # large tables are drawn on the device (no multi-GB host tensor)
            return torch.rand(shape, device=device()).sub_(0.5).mul_(0.1)
        return (torch.rand(shape) - 0.5) * 0.1  # Keras RandomUniform(-0.05, 0.05)
    if name in ("glorot_uniform", "glorotuniform"):
        fan_in, fan_out = (shape[0], shape[-1]) if len(shape) > 1 else (shape[0], shape[0])
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape) * 2 - 1) * lim
    raise ValueError(f"unknown initializer {initializer!r}")


class Zeros:  # tensorflow.keras.initializers.Zeros look-alike
    __name__ = "zeros"


class InputLayer(Layer):
    def call(self, inputs, **kwargs):
        return inputs


def Input(shape=None, name=None, dtype="float32", **kwargs) -> KTensor:
    t = KTensor((None,) + tuple(shape), dtype, node=None, name=name or f"input_{next(_uid)}")
    t.is_input = True
    return t


class Lambda(Layer):
    """fn over torch tensors + a shape function (used for `dnn_output + fm_output`, squeeze etc.)."""

    def __init__(self, fn: Callable, shape_fn: Callable, dtype_fn: Optional[Callable] = None, **kw):
        super().__init__(**kw)
        self.fn, self.shape_fn, self.dtype_fn = fn, shape_fn, dtype_fn

    def call(self, inputs):
        return self.fn(*inputs) if isinstance(inputs, (list, tuple)) else self.fn(inputs)

    def compute_output_shape(self, input_shape):
        return self.shape_fn(*input_shape) if isinstance(input_shape, list) else self.shape_fn(input_shape)

    def output_dtype(self, inputs):
        return self.dtype_fn(inputs) if self.dtype_fn else "float32"


class Activation(Layer):
    """keras.layers.Activation for the names HandyRec's configs use; sigmoid outputs remember their logits."""

    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.activation = activation

    def call(self, inputs):
        from .autograd_ops import ActivationFn

        if self.activation in (None, "linear"):
            return inputs
        if self.activation not in ("relu", "sigmoid", "tanh"):
            raise ValueError(f"unsupported activation {self.activation!r}")
        out = ActivationFn.apply(inputs, self.activation)
        if self.activation == "sigmoid":
            out._keras_logits = inputs  # like tf.keras.activations.sigmoid: binary_crossentropy uses the logits
        return out


# ---------------------------------------------------------------------------------------------
# optimisers / losses (Keras formulas, kernels from libhrb200)
# ---------------------------------------------------------------------------------------------
class Adam:
    def __init__(self, learning_rate=1e-3, lr=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr = float(lr if lr is not None else learning_rate)
        self.b1, self.b2, self.eps, self.t = beta_1, beta_2, epsilon, 0
        self.state: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def step(self, params_l2: Sequence[Tuple[torch.nn.Parameter, float]]):
        from . import kernels as K

        self.t += 1
        ps, gs, ms, vs, l2s = [], [], [], [], []
        for p, l2 in params_l2:
            if p.grad is None or not p.requires_grad:
                continue
            if id(p) not in self.state:
                self.state[id(p)] = (torch.zeros_like(p.data), torch.zeros_like(p.data))
            m, v = self.state[id(p)]
            ps.append(p.data)
            gs.append(p.grad.contiguous())
            ms.append(m)
            vs.append(v)
            l2s.append(2.0 * l2)
        K.adam_step_multi(ps, gs, ms, vs, l2s, self.lr, self.b1, self.b2, self.eps, self.t)  # all variables in one launch


class SGD:
    def __init__(self, learning_rate=0.01, lr=None):
        self.lr = float(lr if lr is not None else learning_rate)

    def step(self, params_l2):
        from . import kernels as K

        live = [(p, l2) for p, l2 in params_l2 if p.grad is not None and p.requires_grad]
        K.sgd_step_multi([p.data for p, _ in live], [p.grad.contiguous() for p, _ in live], [2.0 * l2 for _, l2 in live], self.lr)


def binary_crossentropy(y_true, y_pred):
    from .autograd_ops import SigmoidBCEFn

    logits = getattr(y_pred, "_keras_logits", None)
    if logits is None:
        # Keras' fallback when the prediction is not the direct output of a sigmoid (e.g. BatchNormalization / Dropout behind the
        # output activation, layers/core.py:66-73): probabilities clipped to [eps, 1-eps], eps = 1e-7
        from .autograd_ops import ClippedBCEFn

        return ClippedBCEFn.apply(y_pred, y_true)
    return SigmoidBCEFn.apply(logits, y_true)


class Model:
    """keras.Model look-alike over the symbolic graph: __call__/predict/compile/fit/train_on_batch."""

    def __init__(self, inputs, outputs, name=None):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs
        self.name = name or "model"
        self._order = self._toposort()
        self.optimizer = None
        self.loss = None

    def _toposort(self) -> List[Node]:
        order, seen = [], set()

        def visit(t):
            if not isinstance(t, KTensor) or t.node is None or t.node.id in seen:
                return
            seen.add(t.node.id)
            for i in _flatten(t.node.inputs):
                visit(i)
            order.append(t.node)

        for o in _flatten(self.outputs):
            visit(o)
        return order

    @property
    def layers(self) -> List[Layer]:
        self.sync()
        out, seen = [], set()
        for n in self._order:
            if id(n.layer) not in seen:
                seen.add(id(n.layer))
                out.append(n.layer)
        return out

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"no layer named {name}")

    def weights_with_l2(self):
        out, seen = [], set()
        for l in self.layers:
            for p, l2 in l.weights_with_l2():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append((p, l2))
        return out

    @property
    def trainable_weights(self):
        return [p for p, _ in self.weights_with_l2() if p.requires_grad]

    # ---- execution --------------------------------------------------------------------------
    def _bind(self, x) -> Dict[str, torch.Tensor]:
        dev = device()
        if dev.type != "cuda":
            raise RuntimeError("handyrec_b200 executes on CUDA only (no CPU fallback); graph construction alone works on CPU")
        if isinstance(x, dict):
            feed = x
        else:
            xs = x if isinstance(x, (list, tuple)) else [x]
            feed = {t.name: v for t, v in zip(self.inputs, xs)}
        bound = {}
        for t in self.inputs:
            if t.name not in feed:
                raise KeyError(f"missing input {t.name!r}")
            v = feed[t.name]
            v = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v
            v = v.to(dev, _tdtype(t.dtype.name))
            if v.dim() == 1:
                v = v.reshape(-1, 1)
            bound[t.name] = v.contiguous()
        return bound

    def __call__(self, x, training=False):
        self.sync()
        feed = self._bind(x)
        vals: Dict[Tuple[int, int], torch.Tensor] = {}

        def get(t):
            if getattr(t, "is_input", False):
                return feed[t.name]
            return vals[(t.node.id, t.index)]

        for n in self._order:
            ins = _map_structure(get, n.inputs)
            masks = _map_structure(lambda v: getattr(v, "_keras_mask", None), ins)
            out = n.layer._run(ins, masks, training)
            for i, o in enumerate(out if isinstance(out, (list, tuple)) else [out]):
                vals[(n.id, i)] = o
        return _map_structure(get, self.outputs)

    # ---- fused execution (handyrec_b200.lowering) ----------------------------------------------------
    _fused = None
    fuse = True  # set to False before compile() to keep the layer-by-layer path (tests compare the two)

    def sync(self) -> None:
        """Make the layers' weights current after fused training steps (train_on_batch leaves them in the engine)."""
        if self._fused is not None:
            self._fused.sync_to_layers()

    def predict(self, x, batch_size=None):
        if (self._fused is not None and self._fused.engine is None and isinstance(x, dict) and self._comm is not None and self._comm.N > 1
                and self.optimizer is not None):
            # a multi-GPU model before its first fit: only the engine can read row-sharded tables, so it is built here (collective:
            # every rank calls predict, as with fit)
            n = len(np.asarray(next(iter(x.values()))))
            self._fused.build(min(batch_size or n, n), self.optimizer)
        if self._fused is not None and self._fused.engine is not None and isinstance(x, dict):
            return self._fused.predict(x, batch_size)
        self.sync()
        n = len(next(iter(x.values()))) if isinstance(x, dict) else len(x[0] if isinstance(x, (list, tuple)) else x)
        bs = batch_size or n
        outs = []
        from .kernels import deferred_id_checks

        with torch.no_grad():
            for s in range(0, n, bs):
                xb = {k: v[s : s + bs] for k, v in x.items()} if isinstance(x, dict) else [v[s : s + bs] for v in x]
                with deferred_id_checks():
                    outs.append(self(xb, training=False).detach().cpu())
        return torch.cat(outs).numpy()

    _comm = None

    def compile(self, optimizer=None, loss=None, distribute=None, **kwargs):
        """keras.Model.compile.  `distribute` (extra): True under an initialised torch.distributed process group (one process per
        GPU), or a `handyrec_b200.sharded.TorchDistComm` -- a DeepFM-shaped graph then trains on the row-sharded multi-GPU engine."""
        self.optimizer = optimizer if optimizer is not None else Adam()
        self.loss = loss if loss is not None else binary_crossentropy
        self._fused = None
        if distribute is None and ShardedTables.active is not None:
            distribute = ShardedTables.active.comm
            self.dense_table_max_rows = ShardedTables.active.min_rows
        if distribute:
            from .sharded import TorchDistComm

            self._comm = distribute if not isinstance(distribute, bool) else TorchDistComm()
        if self.fuse and self.loss is binary_crossentropy and isinstance(self.optimizer, (Adam, SGD)):
            from .lowering import lower  # a DeepFM-shaped graph runs on the fused engine (one lookup, tcgen05 towers)

            self._fused = lower(self)
        if self._comm is not None and self._comm.N > 1 and self._fused is None:
            raise ValueError("distribute: only graphs the fused DeepFM engine covers train multi-GPU (lowering.py lists the conditions)")

    def train_on_batch(self, x, y) -> float:
        if self._fused is not None and isinstance(x, dict):
            n = len(np.asarray(next(iter(x.values()))))
            self._fused.build(n, self.optimizer)
            return self._fused.train_on_batch(x, y) + self._fused.reg_loss()
        self.sync()
        from .kernels import deferred_id_checks

        with deferred_id_checks():  # the lookups' out-of-range flags are read once, with the loss, instead of one host sync per lookup
            return self._train_on_batch_layers(x, y)

    def _train_on_batch_layers(self, x, y) -> float:
        out = self(x, training=True)
        yt = torch.as_tensor(np.asarray(y), dtype=torch.float32).to(device()) if not isinstance(y, torch.Tensor) else y.to(device())
        loss = self.loss(yt, out)
        params = self.weights_with_l2()
        for p, _ in params:
            p.grad = None
        loss.backward()
        sparse = [l for l in self._all_layers() if getattr(l, "_sparse_grads", None)]
        skip = {id(l.embeddings) for l in sparse}
        # Keras adds the regularisation losses to the reported loss; their gradient 2*l2*W is applied inside the update kernels
        # (tables under the sparse row update are left out of the reported term: a full pass over them is what that update avoids)
        reg = sum(float(l2) * float((p.detach() * p.detach()).sum()) for p, l2 in params if l2 and id(p) not in skip)
        self.optimizer.step(params)
        for l in sparse:  # large tables: IndexedSlices-style gradient -> sorted-segment row update
            l.apply_sparse(self.optimizer)
        self._update_moving_stats()
        return float(loss.detach()) + reg

    def save_weights(self, filepath, **kwargs):
        """keras.Model.save_weights: a directory of `.npy` files (tables row-sharded), see handyrec_b200.checkpoint."""
        from .checkpoint import save_weights

        save_weights(self, filepath)

    def load_weights(self, filepath, **kwargs):
        from .checkpoint import load_weights

        load_weights(self, filepath)

    def _all_layers(self):
        out, seen = [], set()
        for l in self.layers:
            for sub in [l] + _all_sublayers(l):
                if id(sub) not in seen:
                    seen.add(id(sub))
                    out.append(sub)
        return out

    def _update_moving_stats(self):
        for sub in self._all_layers():
            fn = getattr(sub, "_commit_moving_stats", None)
            if fn:
                fn()

    def fit(self, x=None, y=None, batch_size=32, epochs=1, validation_data=None, verbose=0, **kwargs):
        history = {"loss": []}
        if isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], dict) and y is None:
            x, y = x
        batches = _batches(x, y, batch_size)
        for _ in range(epochs):
            if self._fused is not None and isinstance(x, dict):
                # dict-of-arrays input: batches are packed into pinned memory one step ahead of the GPU and the whole epoch
                # runs on the fused engine with asynchronous loss read-back (lowering.FusedDeepFM.fit_epoch)
                n = len(np.asarray(next(iter(x.values()))))
                self._fused.build(min(batch_size, n), self.optimizer)
                losses = self._fused.fit_epoch(x, y, batch_size)
                sizes = [min(batch_size, n - s) for s in range(0, n, batch_size)]
                history["loss"].append(float(np.average(losses, weights=sizes)) + self._fused.reg_loss())
                self.last_losses = losses
            else:
                tot, cnt = 0.0, 0
                for xb, yb in batches():  # Keras reports the sample-weighted running mean of the batch losses
                    nb = len(yb) if hasattr(yb, "__len__") else 1
                    tot += self.train_on_batch(xb, yb) * nb
                    cnt += nb
                history["loss"].append(tot / max(cnt, 1))
            # fused training leaves the dense / FM weights in the engine; they are handed back to the layers lazily, the first time
            # anything reads the layers (`layers`, `get_layer`, `predict` on the layer path, `save_weights`, `sync()`)
            fused_val = (validation_data is not None and self._fused is not None and self._fused.engine is not None
                         and isinstance(validation_data, tuple) and len(validation_data) == 2 and isinstance(validation_data[0], dict))
            if fused_val:
                # the fused engine's forward (the only reader a row-sharded table has); Keras' BCE on clipped probabilities
                xv, yv = validation_data
                p = np.clip(self._fused.predict(xv, batch_size).reshape(-1).astype(np.float64), 1e-7, 1.0 - 1e-7)
                yt = np.asarray(yv, dtype=np.float64).reshape(-1)
                history.setdefault("val_loss", []).append(float(-(yt * np.log(p) + (1.0 - yt) * np.log1p(-p)).mean()))
            elif validation_data is not None:
                vl, vc = 0.0, 0
                with torch.no_grad():
                    for xb, yb in _batches(validation_data, None, batch_size)():
                        out = self(xb, training=False)
                        vl += float(self.loss(torch.as_tensor(np.asarray(yb), dtype=torch.float32).to(device()), out))
                        vc += 1
                history.setdefault("val_loss", []).append(vl / max(vc, 1))

        class H:
            pass

        h = H()
        h.history = history
        return h


def _all_sublayers(l: Layer) -> List[Layer]:
    out = []
    for s in l._sublayers:
        out += [s] + _all_sublayers(s)
    return out


def _batches(x, y, batch_size):
    """x: dict of arrays (+ y), or an iterable of (dict, label) batches (what tf.data yields in the reference tests)."""

    def gen():
        if isinstance(x, dict):
            n = len(next(iter(x.values())))
            for s in range(0, n, batch_size):
                yield {k: v[s : s + batch_size] for k, v in x.items()}, y[s : s + batch_size]
        elif isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], dict):
            xd, yd = x
            n = len(next(iter(xd.values())))
            for s in range(0, n, batch_size):
                yield {k: v[s : s + batch_size] for k, v in xd.items()}, yd[s : s + batch_size]
        else:
            for xb, yb in x:
                yield xb, yb

    return gen
