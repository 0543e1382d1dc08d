from lstc_vad_b200.models.FFN import PositionwiseFeedForward  # noqa: F401
