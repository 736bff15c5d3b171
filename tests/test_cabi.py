"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every symbol
include/svanon.h declares, and argument errors surface as error codes + messages (no GPU needed)."""
import ctypes as C

import pytest

from streamvoiceanon_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_header_symbol(lib):
    names = _lib.header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svanon.h but not exported"
    assert set(names) == set(_lib._SIGS), set(names) ^ set(_lib._SIGS)


def test_enc_num_ids(lib):
    # one content id per 2048 samples (hop 512, two stride-2 down-samplers)
    assert lib.svanon_enc_num_ids(128 * 2048) == 128
    assert lib.svanon_enc_num_ids(2047) == 0
    assert lib.svanon_enc_num_ids(343483) == 167          # test_waves/trump_0.wav, SURVEY section 8d


def test_errors_are_reported_not_crashed(lib):
    h = C.c_void_p()
    rc = lib.svanon_engine_create(0, C.byref(h))
    import torch
    if not torch.cuda.is_available():
        assert rc != 0 and lib.svanon_last_error()          # no device here: must fail loudly
    assert lib.svanon_ar_set_delay(None, 2) != 0
    assert b"null" in lib.svanon_last_error()


def test_speaker_entry_points_reject_null_arguments(lib):
    """The prompt-path entry points fail with a message, not a crash, on null handles / buffers."""
    buf = (C.c_float * 8)()
    for rc in (lib.svanon_kaldi_fbank(None, buf, 1000, buf, None),
               lib.svanon_campplus_forward(None, buf, 10, 5, buf, None),
               lib.svanon_style_vector(None, buf, 1000, buf, None),
               lib.svanon_timbre_latent(None, buf, 2000, 2000, buf, None, None)):
        assert rc != 0 and lib.svanon_last_error()


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from streamvoiceanon_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine.get(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under streamvoiceanon_b200/ may reference it."""
    from pathlib import Path
    pkg = Path(_lib.__file__).resolve().parent
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p
        # ... and the host build of the speaker-encoder source (tests/hostemu) is test infrastructure too
        assert "libspeaker_host" not in text and "hostemu_" not in text, p


def test_custom_ops_registered_with_fake_impls():
    """torch.compile(fullgraph=True) in the reference's caller code needs the engine entries as custom ops whose output
    shapes can be derived without running them (streamvoiceanon_b200/ops.py)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from streamvoiceanon_b200 import ops  # noqa: F401
    with FakeTensorMode():
        assert tuple(torch.ops.svanon_b200.enc_encode(torch.empty(2, 40960)).shape) == (1, 2, 20)
        assert tuple(torch.ops.svanon_b200.voc_head(torch.empty(1, 512, 8)).shape) == (1, 1, 4096)
        assert tuple(torch.ops.svanon_b200.voc_quantizer_decode(torch.empty(1, 8, 3, dtype=torch.int64)).shape) == (1, 12, 512)
        assert tuple(torch.ops.svanon_b200.voc_decode(torch.empty(1, 8, 3, dtype=torch.int64)).shape) == (1, 1, 6144)
