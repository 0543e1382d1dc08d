from lstc_vad_b200.models.EncoderLayer import EncoderLayer  # noqa: F401
