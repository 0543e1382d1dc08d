from lstc_vad_b200.models.Classifier import Classifier  # noqa: F401
