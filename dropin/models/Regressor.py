from lstc_vad_b200.models.Regressor import Regressor  # noqa: F401
