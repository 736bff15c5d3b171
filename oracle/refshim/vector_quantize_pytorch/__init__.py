"""Import shim (test infrastructure only).  `vector-quantize-pytorch==1.14.24`
(requirements.txt:26 of the reference) is not installed; the reference vendors a
near-verbatim twin of the one class it needs, so the shim re-exports that twin
(modules/bicodec_speaker_encoder/fsq/residual_fsq.py:269-336).  Requires
/root/reference on sys.path."""
from modules.bicodec_speaker_encoder.fsq.residual_fsq import GroupedResidualFSQ  # noqa: F401
