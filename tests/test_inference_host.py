"""Host logic of streamvoiceanon_b200.inference.InferenceWrapper that needs no GPU: the reference's chunking rule of
`stream_infer` (left padding to whole chunks, a full extra chunk when the source is already aligned,
evaluations/infer_arvc.py:648-652), reference-list handling and crop lengths (:234-259), and the order of calls."""
import numpy as np
import pytest
import torch

from streamvoiceanon_b200.inference import InferenceWrapper


class _FakeSession:
    def __init__(self, log):
        self.log = log

    def process_chunk(self, chunk, out=None):
        self.log.append(("chunk", chunk.clone()))
        if out is None:
            out = torch.empty_like(chunk)
        out.copy_(chunk * 2)
        return out


def _wrapper():
    iw = InferenceWrapper(None, None, None, None, None, device="cpu")
    log = []

    def prefill_prompt(refs, **kw):
        log.append(("prefill", [tuple(r.shape) for r in refs], kw))
        iw._session = _FakeSession(log)

    def setup_stream_caches(*a):
        log.append(("setup", a))
        iw.decode_chunk_frames = a[4]

    iw.prefill_prompt, iw.setup_stream_caches = prefill_prompt, setup_stream_caches
    return iw, log


@pytest.mark.parametrize("n,chunk,n_chunks", [(2048 * 3, 1, 4), (2048 * 3 + 1, 1, 4), (2048 * 4 - 1, 2, 2), (4096 * 2, 2, 3), (10, 1, 1)])
def test_stream_infer_chunking_rule(n, chunk, n_chunks):
    iw, log = _wrapper()
    src = torch.arange(1, n + 1, dtype=torch.float32)
    out = iw.stream_infer(src, torch.zeros(5000), decode_chunk_frames=chunk, delay=2, alpha=0.7)
    step = 2048 * chunk
    assert out.shape == (n_chunks * step,)
    pad = n_chunks * step - n
    assert 1 <= pad <= step                                     # never zero: an aligned source gets a whole silent chunk
    assert np.array_equal(out[:pad], np.zeros(pad, np.float32)) and np.array_equal(out[pad:], 2 * src.numpy())
    kinds = [e[0] for e in log]
    assert kinds == ["prefill", "setup"] + ["chunk"] * n_chunks
    assert log[0][2] == dict(max_prompt_frames=256, delay=2, alpha=0.7, spk_emb_collate_type="concat_mel")
    assert log[1][1] == (128, 64, 768, 32, chunk)               # the CLI defaults of stream_infer
    assert all(e[1].numel() == step for e in log[2:])


def test_reference_lists_and_crops():
    iw, log = _wrapper()
    refs = [torch.zeros(44100 * 2), np.zeros(44100, np.float32)]
    iw.stream_infer(torch.zeros(100), refs, ref_crop_lengths=[0.5, None])
    assert log[0][1] == [(1, 22050), (1, 44100)]
    iw.stream_infer(torch.zeros(100), refs, ref_crop_lengths=0.25)
    assert [e for e in log if e[0] == "prefill"][1][1] == [(1, 11025), (1, 11025)]
    assert iw.process_ref_paths("a.wav", None) == (["a.wav"], [None])
    with pytest.raises(ValueError):
        iw.process_ref_paths(refs, [1.0])
    with pytest.raises(NotImplementedError):
        iw.stream_infer(torch.zeros(100), refs, save_result=True)


def test_wav_paths_are_read_like_librosa(tmp_path):
    """16-bit PCM -> float / 32768, stereo -> channel mean, at the model rate no resampling."""
    from scipy.io import wavfile
    iw, _ = _wrapper()
    pcm = (np.arange(-500, 500, dtype=np.int16) * 30)
    wavfile.write(tmp_path / "mono.wav", 44100, pcm)
    wavfile.write(tmp_path / "stereo.wav", 44100, np.stack([pcm, -pcm // 2], axis=1).astype(np.int16))
    mono = iw._load(tmp_path / "mono.wav")
    assert tuple(mono.shape) == (1, 1000) and torch.equal(mono[0], torch.from_numpy(pcm).float() / 32768.0)
    stereo = iw._load(str(tmp_path / "stereo.wav"), crop_seconds=0.01)
    assert tuple(stereo.shape) == (1, 441)
    want = (torch.from_numpy(pcm).float() + torch.from_numpy((-pcm // 2).astype(np.int16)).float()) / 2 / 32768.0
    assert torch.allclose(stereo[0], want[:441])


def test_process_one_chunk_checks_the_chunk_length():
    iw, _ = _wrapper()
    iw._session, iw.decode_chunk_frames = _FakeSession([]), 2
    assert tuple(iw.process_one_chunk(torch.ones(1, 4096)).shape) == (1, 4096)
    with pytest.raises(ValueError):
        iw.process_one_chunk(torch.ones(1, 2048))


def test_from_config_reads_the_reference_layout(tmp_path, monkeypatch):
    """`from_config` follows the reference constructor (infer_arvc.py:33-126): five checkpoints named by the top-level
    YAML, tokenizer unwrapped from {'net': ...} / 'module.', relative paths against the root, sr from preprocess_params."""
    import yaml
    (tmp_path / "ckpt").mkdir()
    names = {"speech_tokenizer": "tok.pth", "firefly": "voc.pth", "style_encoder": "style.bin", "timbre_encoder": "timbre.pth"}
    cfg = {"preprocess_params": {"sr": 44100}, "model_params": {"config_path": "unused.yaml"}}
    for section, f in names.items():
        cfg[section] = {"config_path": "unused.yaml", "checkpoint_path": f"ckpt/{f}"}
        sd = {"w": torch.full((2,), float(len(section)))}
        torch.save({"net": {"module.w": sd["w"]}} if section == "speech_tokenizer" else sd, tmp_path / "ckpt" / f)
    torch.save({"embedding.weight": torch.ones(3)}, tmp_path / "ckpt" / "ar.pth")
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
    seen = {}

    def fake(cls, ar_sd, tok_sd, voc_sd, style_sd, timbre_sd, **kw):
        seen.update(ar=ar_sd, tok=tok_sd, voc=voc_sd, style=style_sd, timbre=timbre_sd)
        return InferenceWrapper(None, None, None, None, None, device="cpu")

    monkeypatch.setattr(InferenceWrapper, "from_state_dicts", classmethod(fake))
    iw = InferenceWrapper.from_config("config.yaml", "ckpt/ar.pth", compile_ar=True, compile_decoder=True, compile_encoder=True,
                                      root=tmp_path)
    assert iw.sr == 44100 and iw.config["firefly"]["checkpoint_path"] == "ckpt/voc.pth"
    assert list(seen["ar"]) == ["embedding.weight"]
    assert list(seen["tok"]) == ["w"] and float(seen["tok"]["w"][0]) == len("speech_tokenizer")
    assert float(seen["voc"]["w"][0]) == len("firefly") and float(seen["style"]["w"][0]) == len("style_encoder")
    assert float(seen["timbre"]["w"][0]) == len("timbre_encoder")
