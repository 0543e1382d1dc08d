from lstc_vad_b200.models.MultiHeadAttention import MultiHeadAttention, ScaledDotProductAttention  # noqa: F401
