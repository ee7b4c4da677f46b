"""Model constructors that consume the hot path (mirror of handyrec.models.ranking: DeepFM, DIN)."""
from .deepfm import DeepFM
from .din import DIN

__all__ = ["DeepFM", "DIN"]
