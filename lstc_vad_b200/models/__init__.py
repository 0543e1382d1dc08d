"""Host-side mirror of the reference `models/` package (same class names, constructor and forward signatures,
parameter / buffer names, shapes and registration order), running on the lstc_vad_b200 CUDA kernels."""
from .Encoder import Encoder
from .EncoderLayer import EncoderLayer
from .MultiHeadAttention import MultiHeadAttention, ScaledDotProductAttention
from .FFN import PositionwiseFeedForward
from .PatchEmbedding import PatchEmbedding
from .Classifier import Classifier
from .Regressor import Regressor

__all__ = ["Encoder", "EncoderLayer", "MultiHeadAttention", "ScaledDotProductAttention", "PositionwiseFeedForward",
           "PatchEmbedding", "Classifier", "Regressor"]
