"""Dice (mirror of /root/reference/handyrec/layers/activation.py:6-50)."""
import torch

from ..autograd_ops import DiceFn
from ..keras_lite import Layer


class Dice(Layer):
    """p = sigmoid(BN_{center=False, scale=False, eps}(x)); out = p*x + (1-p)*alpha*x."""

    def __init__(self, axis=-1, epsilon=1e-9, momentum=0.99, **kwargs):
        self.axis, self.epsilon, self.momentum = axis, epsilon, momentum
        self.alphas = None
        super().__init__(**kwargs)

    def build(self, input_shape):
        u = int(input_shape[-1])
        self.alphas = self.add_weight("dice_alpha", (u,), initializer="zeros")
        self.moving_mean = self.add_weight("moving_mean", (u,), initializer="zeros", trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (u,), initializer="ones", trainable=False)
        self._pending = None
        self.built = True

    def call(self, inputs, **kwargs):
        training = bool(kwargs.get("training", False))  # activation.py:40
        out = DiceFn.apply(inputs, self.alphas, self.moving_mean.data, self.moving_variance.data, training, self.epsilon)
        if training and out.grad_fn is not None:
            self._pending = out.grad_fn.batch_stats
        return out

    def _commit_moving_stats(self):
        if self._pending is not None:  # Keras BN: moving = moving*momentum + batch*(1-momentum)
            bm, bv = self._pending
            self.moving_mean.data.mul_(self.momentum).add_(bm, alpha=1 - self.momentum)
            self.moving_variance.data.mul_(self.momentum).add_(bv, alpha=1 - self.momentum)
            self._pending = None

    def get_config(self):
        return {"axis": self.axis, "epsilon": self.epsilon, **super().get_config()}
