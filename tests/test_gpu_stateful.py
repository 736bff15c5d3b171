"""GPU parity of the stateful stage entries (`enc_push_chunk`, `voc_push_frames`; section 7 step 6b),
through the C ABI (svanon_enc_push_chunk / svanon_voc_push_frames).

Parity targets, both reference functions:
  * stateful encoder == the reference's OFFLINE `FireflyArchitecture.encode()` (modules/vqgan/modules/firefly_encoder.py:
    553-566) on the same prefix (causal-prefix equality): against the fixture the unmodified reference wrote
    (tests/golden/encoder_40f.npz) and against the engine's own offline encode -- which test_gpu_parity.py pins to the
    reference -- over 300 frames, ragged push sizes, several streams side by side.  Ids bit-exact.
  * stateful vocoder == `head(quantizer.decode(codes))` of the whole utterance (firefly.py:280-293, fsq.py:112-116): against
    the reference fixture (tests/golden/vocoder_20f.npz) and the engine's window decode.  Waveform MSE < 1e-8.
The loop's encoder mode 3 runs the stateful encoder inside process_one_chunk: its content ids are the offline encode of
the source so far (NOT the reference's window re-encode; documented opt-in, include/svanon.h)."""
import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

WAVE_MSE_TOL = 1e-8


def test_enc_push_vs_reference_fixture(models, gold):
    from streamvoiceanon_b200 import EncoderStream
    _, tok, _ = models
    g = gold("encoder_40f")
    wav = synth.synth_audio_44k(int(g["audio_seed"]), 2.0)[: int(g["n_samples"])]
    es = EncoderStream(tok, 1)
    ids = torch.cat([es.push(wav[i * 2048:(i + 1) * 2048][None].cuda()).cpu() for i in range(40)], dim=1)
    assert es.position == 40
    assert np.array_equal(ids.numpy()[None], g["ids"])
    es.reset()                                                       # a new utterance: same ids again, host buffers, 8-frame pushes
    ids2 = torch.cat([es.push(wav[i * 16384:(i + 1) * 16384][None]) for i in range(5)], dim=1)
    assert np.array_equal(ids2.numpy()[None], g["ids"])
    es.close()


def test_enc_push_equals_offline_encode_300_frames(models):
    """Causal-prefix equality over 300 frames (past the 128-frame streaming window and the 39-frame conv receptive field),
    three streams side by side, ragged push sizes."""
    from streamvoiceanon_b200 import EncoderStream
    _, tok, _ = models
    T = 300
    wavs = torch.stack([synth.synth_audio_44k(3300 + i, 14.2)[: T * 2048] for i in range(3)])
    want, _ = tok.encode(wavs.cuda(), torch.LongTensor([T * 2048] * 3).cuda())          # [1, 3, 300]
    es = EncoderStream(tok, 3)
    sizes, got, t = [1, 2, 1, 3, 8, 1, 5], [], 0
    while t < T:
        k = min(sizes[len(got) % len(sizes)], T - t)
        got.append(es.push(wavs[:, t * 2048:(t + k) * 2048].cuda()))
        t += k
    got = torch.cat(got, dim=1)
    assert tuple(got.shape) == (3, T)
    mism = int((got != want[0]).sum())
    assert mism == 0, f"{mism} of {3 * T} ids differ from the offline encode"
    es.close()
    with pytest.raises(ValueError):
        EncoderStream(tok, 1).push(torch.zeros(1, 1000))


def test_voc_push_vs_reference_fixture_and_window_decode(models, gold):
    from streamvoiceanon_b200 import VocoderStream
    _, _, voc = models
    g = gold("vocoder_20f")
    codes = torch.from_numpy(g["codes"]).cuda()                      # [1, 8, 20]
    vs = VocoderStream(voc, 1, 1)
    wave = torch.cat([vs.push(codes[:, :, t:t + 1]) for t in range(20)], dim=1)
    assert float(((wave[0].cpu().numpy() - g["wave"]) ** 2).mean()) < WAVE_MSE_TOL
    vs.close()
    # two streams, two frames per push, against the engine's decode of each whole utterance
    gen = torch.Generator().manual_seed(4300)
    codes2 = torch.randint(0, 1000, (2, 8, 40), generator=gen)
    want = torch.cat([voc.decode_codes(codes2[b:b + 1].cuda())[0] for b in range(2)])    # [2, 40 * 2048]
    vs = VocoderStream(voc, 2, 2)
    got = torch.cat([vs.push(codes2[:, :, t:t + 2].cuda()) for t in range(0, 40, 2)], dim=1)
    assert float(((got - want) ** 2).mean()) < WAVE_MSE_TOL
    vs.reset()
    again = torch.cat([vs.push(codes2[:, :, t:t + 2]) for t in range(0, 8, 2)], dim=1)   # host buffers after a reset
    assert float(((again - want[:, : 8 * 2048].cpu()) ** 2).mean()) < WAVE_MSE_TOL
    vs.close()


def _session(tok, b, n_ref, tape_fn, delay=2):
    from streamvoiceanon_b200 import StreamSession
    style, timbre = synth.synth_speaker(5100 + b)
    ref_wave = synth.synth_audio_44k(5100 + b, 3.5)[: n_ref * 2048][None]
    gen = torch.Generator().manual_seed(300 + b)
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=gen).int()
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    sess = StreamSession()
    sess.set_noise_fn(tape_fn, 0)
    sess.set_prompt(ref_content[0].cuda(), ref_audio.cuda(), style.cuda(), timbre.cuda(), max_prompt_frames=256, delay=delay)
    return sess


def test_loop_with_stateful_encoder(models, tape):
    """Encoder mode 3 inside the loop: the content ids of the stream are the OFFLINE encode of the source so far; a
    3-stream lock-step batch in the same mode reproduces each stream alone (ids bit-exact, waveform to fp32 rounding)."""
    from streamvoiceanon_b200 import BatchSession
    _, tok, _ = models
    n_chunks = 40
    cfg = dict(encode_window_frames=32, decode_window_frames=24, max_seq_frames=200, buffer_frames=8, decode_chunk_frames=1)
    srcs = [synth.synth_audio_44k(1800 + b, 2.0)[: n_chunks * 2048] for b in range(3)]
    singles = []
    for b in range(3):
        sess = _session(tok, b, 30 + 4 * b, tape(7900 + b))
        sess.set_encoder_mode(3)
        sess.setup(**cfg)
        waves = torch.cat([sess.process_chunk(srcs[b][i * 2048:(i + 1) * 2048].cuda()).cpu() for i in range(n_chunks)])
        src_hist, pred_hist = sess.history()
        want, _ = tok.encode(srcs[b][None].cuda(), torch.LongTensor([n_chunks * 2048]).cuda())
        assert torch.equal(src_hist, want[0, 0].cpu()), b
        assert pred_hist.shape[1] == n_chunks - 2 and float(waves.abs().max()) > 0
        singles.append((src_hist, pred_hist, waves))
        sess.close()
    sessions = [_session(tok, b, 30 + 4 * b, tape(7900 + b)) for b in range(3)]
    batch = BatchSession(sessions)
    batch.set_encoder_mode(3)
    batch.setup(**cfg)
    batch.set_ar_path(1)
    outs = torch.cat([batch.process_chunk(torch.stack([s[i * 2048:(i + 1) * 2048] for s in srcs]).cuda()).cpu()
                      for i in range(n_chunks)], dim=1)
    for b, sess in enumerate(sessions):
        src_hist, pred_hist = sess.history()
        assert torch.equal(src_hist, singles[b][0]), b
        assert torch.equal(pred_hist, singles[b][1]), b
        assert float(((outs[b] - singles[b][2]) ** 2).mean()) < 1e-10, b
    batch.close()
    for s in sessions:
        s.close()
