"""Drop-in `models` package: put this directory's parent in front of the reference checkout on sys.path and the
reference's `from models.X import X` statements resolve to the lstc_vad_b200 CUDA implementation."""
