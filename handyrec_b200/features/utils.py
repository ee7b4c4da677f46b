"""Feature utilities (mirror of /root/reference/handyrec/features/utils.py:7-38)."""
from collections import OrderedDict

from .type import DenseFeature, SparseFeature, SparseSeqFeature


def split_features(features):
    """-> (dense, sparse, sparse_seq) ordered dicts keyed by feature name, in list order."""
    buckets = (OrderedDict(), OrderedDict(), OrderedDict())
    for feat in features:
        for kind, bucket in zip((DenseFeature, SparseFeature, SparseSeqFeature), buckets):
            if isinstance(feat, kind):
                bucket[feat.name] = feat
                break
    return buckets
