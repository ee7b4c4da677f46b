"""Mirror of handyrec.layers (same exported names, /root/reference/handyrec/layers/__init__.py:14-26) on libhrb200 kernels.
AUGRUCell / PositionEmbedding belong to DIEN / FMLPRec, which are out of scope."""
from .core import DNN
from .interaction import FM
from .sequence import SequencePoolingLayer, LocalActivationUnit
from .activation import Dice
from .tools import ValueTable, ValueRows, CustomEmbedding, SqueezeMask, AttentionPooling, SampledSoftmaxLayer

__all__ = ["DNN", "FM", "SequencePoolingLayer", "Dice", "ValueTable", "CustomEmbedding", "SqueezeMask", "LocalActivationUnit", "AttentionPooling", "SampledSoftmaxLayer", "ValueRows"]
