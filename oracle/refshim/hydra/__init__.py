"""Import shim (test infrastructure only): the `hydra` package is not installed in
this image and there is no network.  Only `hydra.utils.instantiate` is used by the
reference (evaluations/infer_arvc.py:54,69,89,98,111)."""
from . import utils  # noqa: F401
