from lstc_vad_b200.models.PatchEmbedding import PatchEmbedding  # noqa: F401
