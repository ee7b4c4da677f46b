"""DNN (mirror of /root/reference/handyrec/layers/core.py:11-99)."""
from __future__ import annotations

from typing import Tuple

from ..autograd_ops import BatchNormFn, DenseFn, DropoutFn
from ..keras_lite import Activation, Layer


class Dense(Layer):
    """keras.layers.Dense on the last axis; element-wise activations are fused into the GEMM epilogue."""

    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.units, self.activation, self.use_bias = int(units), activation, use_bias
        self.l2 = float(getattr(kernel_regularizer, "l2", kernel_regularizer or 0.0))

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), initializer="glorot_uniform", l2=self.l2)
        self.bias = self.add_weight("bias", (self.units,), initializer="zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        return DenseFn.apply(inputs, self.kernel, self.bias, self.activation)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)


class BatchNormalization(Layer):
    def __init__(self, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        u = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (u,), initializer="ones")
        self.beta = self.add_weight("beta", (u,), initializer="zeros")
        self.moving_mean = self.add_weight("moving_mean", (u,), initializer="zeros", trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (u,), initializer="ones", trainable=False)
        self._pending = None
        self.built = True

    def call(self, inputs, training=False):
        out = BatchNormFn.apply(inputs, self.gamma, self.beta, self.moving_mean.data, self.moving_variance.data, bool(training), self.epsilon)
        if training and out.grad_fn is not None:
            self._pending = out.grad_fn.batch_stats
        return out

    def _commit_moving_stats(self):
        if self._pending is not None:
            bm, bv = self._pending
            self.moving_mean.data.mul_(self.momentum).add_(bm, alpha=1 - self.momentum)
            self.moving_variance.data.mul_(self.momentum).add_(bv, alpha=1 - self.momentum)
            self._pending = None


class Dropout(Layer):
    def __init__(self, rate, seed=0, **kw):
        super().__init__(**kw)
        self.rate, self.seed, self._calls = float(rate), int(seed), 0

    def call(self, inputs, training=False):
        if not training or self.rate == 0.0:
            return inputs
        self._calls += 1
        return DropoutFn.apply(inputs, self.rate, self.seed * 1000003 + self._calls)


def get_activation_layer(activation: str) -> Layer:
    """layers/utils.py:117-133."""
    from .activation import Dice

    if activation == "dice":
        return Dice()
    return Activation(activation)


class DNN(Layer):
    """Dense chain `[in] + hidden_units` (core.py:57 prepends Dense(in)); per layer: Dense -> activation (hidden) or
    output_activation (last; also on hidden layers when `activation` is falsy, core.py:66-69) -> [BN] -> Dropout."""

    def __init__(self, hidden_units: Tuple[int], activation: str = "relu", l2_reg: float = 0, dropout_rate: float = 0, use_bn: bool = False,
                 output_activation: str = None, seed: int = 2022, **kwargs):
        self.hidden_units, self.activation, self.l2_reg = hidden_units, activation, l2_reg
        self.dropout_rate, self.use_bn, self.output_activation, self.seed = dropout_rate, use_bn, output_activation, seed
        self.layers = None
        super().__init__(**kwargs)

    def build(self, input_shape):
        input_size = input_shape[-1]
        hidden_units = [int(input_size)] + list(self.hidden_units)
        self.layers = []
        shape = tuple(input_shape)
        for i, unit in enumerate(hidden_units):
            act = None
            if i + 1 != len(hidden_units) and self.activation:
                act = self.activation
            elif self.output_activation:
                act = self.output_activation
            fused = act in ("relu", "sigmoid", "tanh", "linear", None)
            dense = self._track(Dense(unit, activation=act if fused and act != "sigmoid" else None, kernel_regularizer=self.l2_reg))
            dense.build(shape)
            shape = dense.compute_output_shape(shape)
            self.layers.append(dense)
            if act is not None and not (fused and act != "sigmoid"):
                a = self._track(get_activation_layer(act))  # dice, or sigmoid (kept separate: it remembers its logits)
                a.build(shape)
                a.built = True
                self.layers.append(a)
            if self.use_bn:
                bn = self._track(BatchNormalization())
                bn.build(shape)
                self.layers.append(bn)
            self.layers.append(self._track(Dropout(self.dropout_rate, seed=self.seed + i)))
        self.built = True

    def call(self, inputs, **kwargs):
        training = kwargs.get("training", False)
        x = inputs
        for l in self.layers:
            x = l.call(x, training=training) if l._call_takes_training else l.call(x)
        return x

    def compute_output_shape(self, input_shape):
        if len(self.hidden_units) > 0:
            return tuple(input_shape[:-1]) + (self.hidden_units[-1],)
        return tuple(input_shape)

    def get_config(self):
        return {"hidden_units": self.hidden_units, "activation": self.activation, "l2_reg": self.l2_reg, "dropout_rate": self.dropout_rate,
                "use_bn": self.use_bn, "output_activation": self.output_activation, "seed": self.seed, **super().get_config()}
