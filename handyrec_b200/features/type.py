"""Feature schema objects (mirror of /root/reference/handyrec/features/type.py:1-132)."""


class Feature:
    """Base: a distinct `name` and a `dtype` string."""

    def __init__(self, name: str, dtype: str):
        self.name, self.dtype = name, dtype


class DenseFeature(Feature):
    """type.py:26-50: `dim` numbers per sample, dtype "float32" unless stated."""

    def __init__(self, name: str, dim: int = 1, dtype: str = "float32"):
        super().__init__(name, dtype)
        self.dim = dim


class SparseFeature(Feature):
    """type.py:53-96.  The embedding width is stored under the reference's attribute name `embdding_dim` (sic, type.py:95),
    which features/group.py:288 reads."""

    def __init__(self, name: str, vocab_size: int, embedding_dim: int, trainable: bool = True, dtype: str = "int32"):
        super().__init__(name, dtype)
        self.vocab_size = vocab_size
        self.embdding_dim = embedding_dim
        self.trainable = trainable

    @property
    def embedding_dim(self):
        return self.embdding_dim


class SparseSeqFeature(Feature):
    """type.py:99-132: a fixed-length (pre-padded with 0) sequence of `unit` ids; dtype is always "int32"."""

    def __init__(self, unit, name: str, seq_len: int):
        super().__init__(name, "int32")
        self.unit = unit
        self.seq_len = seq_len
        self.is_group = not isinstance(unit, SparseFeature)  # an EmbdFeatureGroup otherwise
