"""Model constructors over the hot path: the reference's public names and signatures (handyrec.models.ranking /
handyrec.models.retrieval), assembled from shared building blocks (`_blocks.py`) instead of one file per model.

`handyrec_b200.compat.install()` additionally makes the reference's OWN model files importable unchanged on top of this
package's features / layers."""
from .ranking import DIN, DeepFM, YouTubeRankDNN
from .retrieval import DSSM, YouTubeMatchDNN

__all__ = ["DeepFM", "DIN", "YouTubeRankDNN", "DSSM", "YouTubeMatchDNN"]
