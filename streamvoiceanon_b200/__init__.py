"""B200-native engine for the streaming voice-conversion hot path of Plachtaa/StreamVoiceAnon.

Three reference-facing surfaces (same names / arguments as the reference's classes) over one C-ABI CUDA
library (include/svanon.h, streamvoiceanon_b200/libsvanon_b200.so):

    ARVCWrapper        <-> modules.arvc_wrapper.ARVCWrapper
    ContentTokenizer   <-> modules.vqgan.modules.firefly_encoder.FireflyArchitecture   (encode)
    Vocoder            <-> modules.vqgan.modules.firefly.FireflyArchitecture           (quantizer.decode, head)
    EncoderStream      <-> incremental `encode` (new samples in, new ids out; offline-encode semantics)       [stateful entries,
    VocoderStream      <-> incremental `head(quantizer.decode(.))` (new code frames in, new samples out)       SURVEY 8b]
    StreamSession      <-> InferenceWrapper.process_one_chunk as one library call per chunk
    BatchSession       <-> the same loop for N concurrent streams in lock-step (the reference is batch-1)
    StreamPool         <-> streams that join / leave at any chunk boundary, grouped into lock-step cohorts
    InferenceWrapper   <-> evaluations.infer_arvc.InferenceWrapper (prompt from waves, stream_infer, infer)
    speaker.CAMPPlus / speaker.SpeakerEncoder <-> the two speaker encoders of the prompt path
"""
from . import synth  # noqa: F401

__all__ = ["ARVCWrapper", "ContentTokenizer", "Vocoder", "EncoderStream", "VocoderStream", "StreamSession", "BatchSession", "StreamPool", "InferenceWrapper",
           "synth"]


def __getattr__(name):
    if name == "ARVCWrapper":
        from .arvc_wrapper import ARVCWrapper
        return ARVCWrapper
    if name in ("ContentTokenizer", "Vocoder", "EncoderStream", "VocoderStream"):
        from . import firefly
        return getattr(firefly, name)
    if name in ("StreamSession", "BatchSession"):
        from . import streaming
        return getattr(streaming, name)
    if name == "StreamPool":
        from .server import StreamPool
        return StreamPool
    if name == "InferenceWrapper":
        from .inference import InferenceWrapper
        return InferenceWrapper
    raise AttributeError(name)
