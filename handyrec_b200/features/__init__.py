"""Mirror of handyrec.features (/root/reference/handyrec/features/__init__.py): same names, arguments and exceptions."""
from .type import DenseFeature, SparseFeature, SparseSeqFeature
from .group import FeaturePool, FeatureGroup, EmbdFeatureGroup

__all__ = ["DenseFeature", "SparseFeature", "SparseSeqFeature", "FeaturePool", "FeatureGroup", "EmbdFeatureGroup"]
