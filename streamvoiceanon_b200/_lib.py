"""ctypes binding of libsvanon_b200.so (include/svanon.h).

The library is the product: there is NO fallback.  If the shared object is missing or a
call fails, the error is raised."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ["SVANON_LIB"]) if os.environ.get("SVANON_LIB") else PKG / "libsvanon_b200.so"   # tuning builds
HEADER = PKG.parent / "include" / "svanon.h"

MODEL_AR, MODEL_TOKENIZER, MODEL_VOCODER, MODEL_STYLE, MODEL_TIMBRE = 0, 1, 2, 3, 4

_lib = None

_p = C.c_void_p
_SIGS = {
    "svanon_last_error": (C.c_char_p, []),
    "svanon_kernel_launches": (C.c_int64, []),
    "svanon_engine_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "svanon_engine_destroy": (None, [_p]),
    "svanon_load_tensor": (C.c_int, [_p, C.c_int, C.c_char_p, _p, C.c_int, C.POINTER(C.c_int64)]),
    "svanon_finalize_weights": (C.c_int, [_p, C.c_int]),
    "svanon_enc_num_ids": (C.c_int, [C.c_int64]),
    "svanon_enc_encode": (C.c_int, [_p, _p, C.c_int64, _p, _p]),
    "svanon_voc_quantizer_decode": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "svanon_voc_head": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "svanon_voc_decode": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "svanon_stream_create": (C.c_int, [_p, C.c_int, C.POINTER(_p)]),
    "svanon_stream_destroy": (None, [_p]),
    "svanon_ar_set_delay": (C.c_int, [_p, C.c_int]),
    "svanon_ar_set_sampling": (C.c_int, [_p, C.c_float, C.c_float, C.c_uint64]),
    "svanon_ar_set_generate_sampling": (C.c_int, [_p, C.c_float, C.c_float]),
    "svanon_ar_generate_many": (C.c_int, [C.POINTER(_p), C.c_int, C.POINTER(_p), C.POINTER(_p), C.POINTER(C.c_int), C.POINTER(_p),
                                         C.POINTER(C.c_int), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), _p]),
    "svanon_ar_prefill_prompt": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, _p]),
    "svanon_ar_prefill_delay": (C.c_int, [_p, _p, C.c_int, _p]),
    "svanon_ar_decode_one": (C.c_int, [_p, _p, _p, _p, C.POINTER(C.c_int32), _p]),
    "svanon_ar_decode_batch": (C.c_int, [C.POINTER(_p), C.c_int, _p, _p, _p, _p]),
    "svanon_ar_generate": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int, _p, _p, _p, _p, _p]),
    "svanon_ar_position": (C.c_int, [_p]),
    "svanon_ar_debug_logits": (C.c_int, [_p, C.c_int]),
    "svanon_set_gemm_mode": (C.c_int, [C.c_int]),
    "svanon_set_pdl": (C.c_int, [C.c_int]),
    "svanon_set_chain_mode": (C.c_int, [C.c_int]),
    "svanon_debug_chain_gemm": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "svanon_debug_enc_transformer": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    "svanon_set_precision": (C.c_int, [C.c_int]),
    "svanon_set_gemm_pair": (C.c_int, [C.c_int]),
    "svanon_gemm_pair_launches": (C.c_longlong, []),
    "svanon_debug_gemm_alo": (C.c_int, [_p]),
    "svanon_debug_gemm_fused": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p]),
    "svanon_debug_gemm_weights_static": (C.c_int, [C.c_int]),
    "svanon_debug_gemm": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "svanon_debug_gemm_taps": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int, C.POINTER(C.c_int), _p, _p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "svanon_ar_set_kernel_variant": (C.c_int, [_p, C.c_int]),
    "svanon_gemm_timing": (C.c_int, [_p, C.c_int]),
    "svanon_gemm_timing_read": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "svanon_ar_profile": (C.c_int, [_p, C.c_int, C.POINTER(C.c_uint64)]),
    "svanon_debug_grid_barrier": (C.c_int, [_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "svanon_ar_read_debug": (C.c_int, [_p, _p, _p, _p]),
    "svanon_stream_set_prompt": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, C.c_int, C.c_int, _p]),
    "svanon_stream_setup": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "svanon_stream_process_chunk": (C.c_int, [_p, _p, C.c_int, _p, _p, _p]),
    "svanon_stream_set_vocoder_mode": (C.c_int, [_p, C.c_int]),
    "svanon_stream_set_encoder_mode": (C.c_int, [_p, C.c_int]),
    "svanon_batch_set_encoder_mode": (C.c_int, [_p, C.c_int]),
    "svanon_stream_set_timing": (C.c_int, [_p, C.c_int]),
    "svanon_stream_last_timing": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "svanon_stream_history": (C.c_int, [_p, _p, C.POINTER(C.c_int), _p, C.POINTER(C.c_int), C.c_int]),
    "svanon_resample": (C.c_int, [_p, _p, C.c_int64, _p, C.c_int, C.c_int, C.c_int, _p, C.c_int64, _p]),
    "svanon_noise_mix": (C.c_int, [_p, _p, _p, C.c_int64, C.c_float, _p, _p]),
    "svanon_voc_encode": (C.c_int, [_p, _p, C.c_int, C.c_int64, _p, _p]),
    "svanon_kaldi_fbank": (C.c_int, [_p, _p, C.c_int64, _p, _p]),
    "svanon_campplus_forward": (C.c_int, [_p, _p, C.c_int64, C.c_int, _p, _p]),
    "svanon_style_vector": (C.c_int, [_p, _p, C.c_int64, _p, _p]),
    "svanon_timbre_latent": (C.c_int, [_p, _p, C.c_int64, C.c_int64, _p, _p, _p]),
    "svanon_enc_encode_batch": (C.c_int, [_p, _p, C.c_int, C.c_int64, _p, _p]),
    "svanon_ar_decode_many": (C.c_int, [C.POINTER(_p), C.c_int, _p, _p, _p, _p]),
    "svanon_batch_create": (C.c_int, [_p, C.POINTER(_p), C.c_int, C.POINTER(_p)]),
    "svanon_batch_destroy": (None, [_p]),
    "svanon_batch_setup": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "svanon_batch_process_chunk": (C.c_int, [_p, _p, C.c_int, _p, _p, _p]),
    "svanon_batch_set_ar_path": (C.c_int, [_p, C.c_int]),
    "svanon_batch_merge": (C.c_int, [_p, _p, C.POINTER(_p), _p]),
    "svanon_batch_select": (C.c_int, [_p, C.POINTER(C.c_int), C.c_int, C.POINTER(_p), _p]),
    "svanon_batch_set_timing": (C.c_int, [_p, C.c_int]),
    "svanon_batch_last_timing": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "svanon_enc_stream_create": (C.c_int, [_p, C.c_int, C.POINTER(_p)]),
    "svanon_enc_stream_destroy": (None, [_p]),
    "svanon_enc_stream_reset": (C.c_int, [_p, _p]),
    "svanon_enc_stream_position": (C.c_int64, [_p]),
    "svanon_enc_push_chunk": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "svanon_voc_stream_create": (C.c_int, [_p, C.c_int, C.c_int, C.POINTER(_p)]),
    "svanon_voc_stream_destroy": (None, [_p]),
    "svanon_voc_stream_reset": (C.c_int, [_p, _p]),
    "svanon_voc_push_frames": (C.c_int, [_p, _p, _p, _p]),
}


def header_symbols():
    """Every function name declared in include/svanon.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svanon_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m streamvoiceanon_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("svanon_b200: " + load().svanon_last_error().decode(errors="replace"))


def kernel_launches() -> int:
    return int(load().svanon_kernel_launches())
