import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def gold():
    def load(name):
        return np.load(GOLD / f"{name}.npz")
    return load


@pytest.fixture(scope="session")
def weights():
    """Synthetic checkpoints (reference key names), generated once per session."""
    from streamvoiceanon_b200 import synth
    seed = int(np.load(GOLD / "weights_digest.npz")["seed"])
    ar = synth.make_ar_state_dict(seed)
    tok = synth.make_tokenizer_state_dict(seed)
    voc = synth.make_vocoder_state_dict(seed)
    voc_enc = synth.make_vocoder_encoder_state_dict(seed)      # the vocoder's encode path (prompt: wave -> codec ids)
    return dict(ar=ar, tok=tok, voc=voc, voc_enc=voc_enc, voc_folded=synth.fold_weight_norm(voc))


@pytest.fixture(scope="session")
def models(weights):
    """The three reference-facing model objects, loaded ONCE per process (one engine per GPU)."""
    from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(weights["ar"], strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(weights["tok"], strict=False)
    voc = Vocoder()
    voc.load_state_dict({**weights["voc"], **weights["voc_enc"]}, strict=False)   # weight-norm form, folded by the library
    voc.remove_parametrizations()
    return ar, tok, voc


@pytest.fixture(scope="session")
def tape():
    from streamvoiceanon_b200 import synth

    def make(seed):
        cache = {}

        def noise_fn(step, slot, V):
            if step not in cache:
                cache.clear()
                cache[step] = synth.noise_tape(seed, step)
            return cache[step][slot]
        return noise_fn
    return make
