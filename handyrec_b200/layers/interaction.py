"""FM (mirror of /root/reference/handyrec/layers/interaction.py:9-42)."""
from ..autograd_ops import FMFn
from ..keras_lite import Layer


class FM(Layer):
    """Factorization Machine: w0 + sum_f Dense1(x_f) + 0.5*sum_d[(sum_f x)^2 - sum_f x^2] -> (B,1)."""

    def __init__(self, **kwargs):
        self.linear = None
        self.w_0 = None
        super().__init__(**kwargs)

    def build(self, input_shape):
        self.linear = self.add_weight("linear_kernel", (input_shape[-1], 1), initializer="glorot_uniform")  # Dense(1, use_bias=False)
        self.w_0 = self.add_weight("W_0", (1,), initializer="zeros")
        self.built = True

    def call(self, inputs, mask=None, *args, **kwargs):
        return FMFn.apply(inputs, self.linear, self.w_0)

    def compute_output_shape(self, input_shape):
        return (None, 1)
