"""Builds the UNMODIFIED reference models in this container.  TEST INFRASTRUCTURE ONLY.

Usable where /root/reference exists (the build container) or where `oracle/_ref/` -- the same modules byte-compiled by
oracle/build_ref.py, which travels to the GPU box -- is present.  Adds
the import shims (oracle/refshim) and the reference root to sys.path, instantiates the
three hot-path models from the reference's own YAMLs through its own
`hydra.utils.instantiate` call pattern (evaluations/infer_arvc.py:53-54,68-69,88-89),
loads the synthetic checkpoints, and applies the three external patches of SURVEY.md
section 8c-3 (no reference file is edited):

  (i)   fp32 KV caches        -- the hard-coded fp16 cache crashes on CPU
  (ii)  torch.cuda.Event / synchronize no-ops on CPU
  (iii) `multinomial_sample_one_no_sync` reads Exp(1) noise from the shared tape
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch
import yaml

SHIMS = Path(__file__).resolve().parent / "refshim"
BUILT = Path(__file__).resolve().parent / "_ref"


def _root() -> Path:
    env = os.environ.get("SVANON_REFERENCE")
    if env:
        return Path(env)
    src = Path("/root/reference")
    return src if (src / "modules" / "arvc_wrapper.py").exists() else BUILT


REF_ROOT = _root()


BYTECODE_SUFFIX = ".pycode"          # oracle/build_ref.py


def available() -> bool:
    return (REF_ROOT / "modules" / "arvc_wrapper.py").exists() or (REF_ROOT / "modules" / ("arvc_wrapper" + BYTECODE_SUFFIX)).exists()


def kind() -> str:
    """'source' (the reference tree itself) or 'bytecode' (oracle/_ref)."""
    return "source" if (REF_ROOT / "modules" / "arvc_wrapper.py").exists() else "bytecode"


def _bytecode_hook(path):
    """sys.path_hooks entry: directories under the byte-compiled reference are searched for `<module>.pycode` files
    (sourceless byte code, oracle/build_ref.py); every other path is left to the standard hooks."""
    from importlib.machinery import FileFinder, SourcelessFileLoader
    if not os.path.abspath(path).startswith(str(REF_ROOT)):
        raise ImportError
    return FileFinder(path, (SourcelessFileLoader, [BYTECODE_SUFFIX]))


def _paths():
    if kind() == "bytecode" and _bytecode_hook not in sys.path_hooks:
        sys.path_hooks.insert(0, _bytecode_hook)
        sys.path_importer_cache.clear()
    for p in (str(SHIMS), str(REF_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)


class _Event:
    def __init__(self, enable_timing=False):
        import time
        self._t = 0.0
        self._time = time

    def record(self):
        self._t = self._time.perf_counter()

    def elapsed_time(self, other):
        return (other._t - self._t) * 1e3


class Tape:
    """State for patch (iii): counts decode_one_token_ar calls (`step`) and sampler
    calls within the current step (`slot`)."""

    def __init__(self, noise_fn):
        self.noise_fn = noise_fn
        self.step = -1
        self.slot = 0

    def new_step(self):
        self.step += 1
        self.slot = 0

    def draw(self, probs):
        q = self.noise_fn(self.step, self.slot, probs.shape[-1])
        self.slot += 1
        return torch.argmax(probs / q[: probs.shape[-1]], dim=-1, keepdim=True).to(dtype=torch.int)


def build(ar_sd, tok_sd, voc_sd, noise_fn):
    """Returns (ar_model, tokenizer, vocoder, tape) -- reference modules, eval mode, CPU."""
    _paths()
    import hydra
    import modules.dual_ar_stream as das

    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        top = yaml.safe_load(open("configs/config_firefly_arvcasr_8192_delay0_8.yaml"))
        model = hydra.utils.instantiate(yaml.safe_load(open(top["model_params"]["config_path"])))
        tok = hydra.utils.instantiate(yaml.safe_load(open(top["speech_tokenizer"]["config_path"])))
        voc = hydra.utils.instantiate(yaml.safe_load(open(top["firefly"]["config_path"])))
    finally:
        os.chdir(cwd)

    # patch (i): fp32 caches, same call as infer_arvc.py:55-59 otherwise
    model.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float32)
    model.eval()
    missing, unexpected = model.load_state_dict(ar_sd, strict=False)
    assert not unexpected, unexpected
    assert set(missing) <= {"decoder.model.embeddings.weight"}, missing

    missing, unexpected = tok.load_state_dict(tok_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("head.", "quantizer.post_module.", "quantizer.pre_module.freqs_cis",
                             "quantizer.pre_module.causal_mask", "quantizer.residual_bsq.rvqs.0.project_out",
                             "quantizer.residual_bsq.rvqs.0.mask")) for k in missing), missing
    tok.eval()

    missing, unexpected = voc.load_state_dict(voc_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("backbone.", "quantizer.downsample.")) or "project_in" in k for k in missing), missing
    voc.remove_parametrizations()          # infer_arvc.py:94
    voc.eval()

    # patch (iii)
    tape = Tape(noise_fn)
    das.multinomial_sample_one_no_sync = tape.draw
    orig = das.decode_one_token_ar.__wrapped__ if hasattr(das.decode_one_token_ar, "__wrapped__") else None
    real = das.decode_one_token_ar

    def counted(*a, **k):
        tape.new_step()
        return real(*a, **k)

    das.decode_one_token_ar = counted
    del orig
    return model, tok, voc, tape


def make_inference_wrapper(model, tok, voc, style_vectors, timbre_latents, ref_audio_codes):
    """An `InferenceWrapper` (evaluations/infer_arvc.py:26) around already-built models.
    `calculate_prompt`'s setup-path calls (two speaker encoders, and the vocoder's own
    encoder that turns the reference wave into codec ids) are replaced by supplied tensors;
    the content tokenizer still runs on the reference wave.  `prefill_prompt`,
    `setup_stream_caches` and `process_one_chunk` run unmodified."""
    _paths()
    # patch (ii)
    if not torch.cuda.is_available():
        torch.cuda.Event = _Event
        torch.cuda.synchronize = lambda *a, **k: None
    from evaluations.infer_arvc import InferenceWrapper

    w = object.__new__(InferenceWrapper)
    w.device = torch.device("cpu")
    w.model, w.speech_tokenizer, w.firefly = model, tok, voc
    w.compiled_speech_tokenizer_encode = tok.encode
    w.sr = 44100

    def calculate_prompt(ref_wav_tensors, alpha=1.0, spk_emb_collate_type="concat_mel"):
        lst = ref_wav_tensors if isinstance(ref_wav_tensors, list) else [ref_wav_tensors]
        ref = torch.cat(lst, dim=-1) if len(lst) > 1 else lst[0]
        lens = w.create_wave_lens_tensor(ref)
        ref_content_codes, _ = w.speech_tokenizer.encode(ref, lens)
        return ref_audio_codes, ref_content_codes.squeeze(0), style_vectors, timbre_latents, ref

    w.calculate_prompt = calculate_prompt
    return w
