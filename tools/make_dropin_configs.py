"""Writes the drop-in YAML set of INTEGRATION.md section 2: the reference's own configs with the five `_target_`s pointed
at the engine's shims, so that the reference's UNMODIFIED `InferenceWrapper` (evaluations/infer_arvc.py:33-144),
`infer_arvc.py` CLI and `real-time-gui.py:52-57` `load_models` construct the engine instead of the torch modules.

    python tools/make_dropin_configs.py <reference root> <output dir> [--checkpoints DIR]

reads  <reference root>/configs/config_firefly_arvcasr_8192_delay0_8.yaml and the five YAMLs it names
writes <output dir>/configs/... (same relative paths; run the reference with <output dir> as its working directory or pass
       the new top-level YAML as --config)

Edits, and nothing else:
  * top-level `_target_` of the five model YAMLs -> streamvoiceanon_b200.{arvc_wrapper.ARVCWrapper, firefly.ContentTokenizer,
    firefly.Vocoder, speaker.CAMPPlus, speaker.SpeakerEncoder};
  * `decoder` / `decoder.model` / `decoder.model.config` -> the shim's light config carriers (they are read for `delay`);
  * every other nested `_target_` (torch.nn.Embedding, ConvNeXtEncoder, HiFiGANGenerator, LogMelSpectrogram,
    torchaudio MelSpectrogram, ...) -> `builtins.dict`: the shims accept and ignore those arguments, and 150 M unused torch
    parameters are not instantiated;
  * with --checkpoints: the five `checkpoint_path`s are re-rooted there (file names kept).
"""
from __future__ import annotations

import sys
from pathlib import Path

import yaml

TOP = "configs/config_firefly_arvcasr_8192_delay0_8.yaml"
SHIMS = {
    "model_params": "streamvoiceanon_b200.arvc_wrapper.ARVCWrapper",
    "speech_tokenizer": "streamvoiceanon_b200.firefly.ContentTokenizer",
    "firefly": "streamvoiceanon_b200.firefly.Vocoder",
    "style_encoder": "streamvoiceanon_b200.speaker.CAMPPlus",
    "timbre_encoder": "streamvoiceanon_b200.speaker.SpeakerEncoder",
}
DECODER = {
    "modules.dual_ar_stream.DualARWrapper": "streamvoiceanon_b200.arvc_wrapper.DualARWrapper",
    "modules.dual_ar_stream.DualARTransformer": "streamvoiceanon_b200.arvc_wrapper.DualARTransformer",
    "modules.dual_ar_stream.DualARModelArgs": "streamvoiceanon_b200.arvc_wrapper.DualARModelArgs",
}


def _neutralise(node):
    """Nested `_target_`s: decoder carriers -> shim carriers, everything else -> builtins.dict."""
    if isinstance(node, dict):
        if "_target_" in node:
            node["_target_"] = DECODER.get(node["_target_"], "builtins.dict")
        for v in node.values():
            _neutralise(v)
    elif isinstance(node, list):
        for v in node:
            _neutralise(v)


def make(reference_root, out_root, checkpoints=None) -> Path:
    """Returns the path of the new top-level YAML."""
    ref, out = Path(reference_root), Path(out_root)
    top = yaml.safe_load(open(ref / TOP))
    for section, target in SHIMS.items():
        rel = top[section]["config_path"]
        cfg = yaml.safe_load(open(ref / rel))
        for v in cfg.values():
            _neutralise(v)
        cfg["_target_"] = target
        dst = out / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        yaml.safe_dump(cfg, open(dst, "w"), sort_keys=False)
        if checkpoints is not None and "checkpoint_path" in top[section]:
            top[section]["checkpoint_path"] = str(Path(checkpoints) / Path(top[section]["checkpoint_path"]).name)
    dst = out / TOP
    dst.parent.mkdir(parents=True, exist_ok=True)
    yaml.safe_dump(top, open(dst, "w"), sort_keys=False)
    return dst


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    ck = sys.argv[sys.argv.index("--checkpoints") + 1] if "--checkpoints" in sys.argv else None
    print(make(sys.argv[1], sys.argv[2], ck))
