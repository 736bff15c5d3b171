"""`generate(..., temperature=, top_p=)` on the GPU against the UNMODIFIED reference (tests/golden/ar_generate_kwargs.npz,
oracle/make_golden_generate_kwargs.py): the first frame is sampled with the default arguments, the later frames with
the caller's (modules/dual_ar_stream.py:723,745-752).  Codec ids bit-exact.
"""
import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]


def test_generate_sampling_kwargs_vs_reference(models, gold, tape):
    ar, _, _ = models
    g, s = gold("ar_generate_kwargs"), gold("ar_stream")
    style, timbre = synth.synth_speaker(int(s["spk_seed"]))
    args = (torch.from_numpy(s["ref_content"]).cuda(), torch.from_numpy(s["ref_audio"]).cuda(),
            torch.from_numpy(s["src_content"])[:, : int(g["n_src"])].cuda(), style.cuda(), timbre.cuda())
    ar.set_delay(delay=2)
    ar.set_noise_fn(tape(int(g["tape_seed"])), 0)
    out = ar.generate(*args, temperature=float(g["temperature"]), top_p=float(g["top_p"]), repetition_penalty=1.5)
    assert np.array_equal(out.cpu().numpy(), g["codes"])
    # the per-call arguments do not stick: the next call without them reproduces the default-argument fixture
    ar.set_noise_fn(tape(int(g["tape_seed"])), 0)
    again = ar.generate(args[0], args[1], args[2][:, :10], args[3], args[4])
    assert np.array_equal(again.cpu().numpy(), gold("ar_generate")["codes"])
    with pytest.raises(TypeError):
        ar.generate(*args, top_k=5)
