from lstc_vad_b200.models.Encoder import Encoder  # noqa: F401
