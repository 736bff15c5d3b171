"""BASELINE config 5 (anonymisation: three references, alpha 0.7, two-frame chunks, truncated prompt, re-prompt) on the
GPU against the fixtures written by the UNMODIFIED reference (`InferenceWrapper.calculate_prompt` / `prefill_prompt` /
`process_one_chunk`; tests/golden/prompt_config5.npz, stream_config5.npz, oracle/make_golden_prompt.py).

Everything the engine computes is checked bit-exact (codec ids and content ids of the concatenated references, the
loop's content ids and codec ids) or to the fp32 tolerance (noise mix, waveform).  The two speaker embeddings BEFORE
the anonymisation mix come from the CPU oracle here -- their CUDA ports are the open item of SURVEY section 8f-3 -- and are
mixed on the GPU with the reference's recorded draws."""
import numpy as np
import pytest
import torch

from streamvoiceanon_b200 import synth

pytestmark = pytest.mark.gpu

WAVE_MSE_TOL = 1e-8


def _reference_wave(gp):
    refs = [synth.synth_audio_44k(int(s), float(gp["ref_seconds"]))[None] for s in gp["ref_seeds"]]
    ref = torch.cat(refs, dim=-1)
    assert ref.shape[-1] == int(gp["n_samples"])
    return ref


def test_config5_prompt_ids_vs_reference(models, gold):
    """`calculate_prompt` rows the engine owns (infer_arvc.py:423-441): codec ids of the 4.8 s concatenation through
    svanon_voc_encode and its content ids through svanon_enc_encode, both bit-exact."""
    _, tok, voc = models
    gp = gold("prompt_config5")
    ref = _reference_wave(gp)
    lens = torch.LongTensor([ref.shape[-1]]).cuda()
    (codes, _), n = voc.encode(ref.cuda(), lens)
    assert int(n[0]) == gp["ref_audio_codes"].shape[-1]
    assert np.array_equal(codes.cpu().numpy(), gp["ref_audio_codes"])
    content, _ = tok.encode(ref.cuda(), lens)
    assert np.array_equal(content[0].cpu().numpy(), gp["ref_content_codes"])


def test_config5_stream_loop_vs_reference(models, weights, gold, tape):
    """The loop of config 5 from the reference WAVES: prompt ids from the engine, speaker embeddings mixed on the GPU
    (alpha 0.7, the reference's draws), prompt truncated to 48 frames, chunk = 2, re-prompt firing -- content ids and codec
    ids bit-exact vs the unmodified reference, waveform MSE < 1e-8."""
    from oracle import prompt as P
    from oracle import speaker as S
    from streamvoiceanon_b200 import StreamSession
    from streamvoiceanon_b200.prompt import apply_noise_mixing
    _, tok, voc = models
    g, gp = gold("stream_config5"), gold("prompt_config5")
    ws = int(g["weight_seed"])
    ref = _reference_wave(gp)
    lens = torch.LongTensor([ref.shape[-1]]).cuda()
    (codes, _), _ = voc.encode(ref.cuda(), lens)
    content, _ = tok.encode(ref.cuda(), lens)
    with torch.no_grad():                                    # pre-mix embeddings: CPU oracle (see the module docstring)
        ref16 = P.resample(ref, 44100, 16000)
        lens16 = torch.LongTensor([ref16.shape[-1]])
        style0 = S.calculate_style_vec(ref16, lens16, synth.make_campplus_state_dict(ws))
        timbre0 = S.calculate_timbre_latent(ref16, lens16, synth.make_timbre_encoder_state_dict(ws))[0]
    alpha = float(g["alpha"])
    style = apply_noise_mixing(style0.cuda(), alpha, torch.from_numpy(gp["noise_style"]).cuda())
    timbre = apply_noise_mixing(timbre0.cuda(), alpha, torch.from_numpy(gp["noise_timbre"]).cuda())
    assert np.abs(style.cpu().numpy() - gp["style_vectors"]).max() < 1e-5
    assert np.abs(timbre.cpu().numpy() - gp["timbre_latents"]).max() < 1e-5

    chunk, n = int(g["decode_chunk_frames"]), int(g["n_chunks"])
    sess = StreamSession()
    sess.set_noise_fn(tape(int(g["tape_seed"])), 0)
    sess.set_prompt(content[0], codes, style, timbre, max_prompt_frames=int(g["max_prompt_frames"]), delay=int(g["delay"]))
    sess.setup(int(g["encode_window_frames"]), int(g["decode_window_frames"]), int(g["max_seq_frames"]),
               int(g["buffer_frames"]), chunk)
    src = synth.synth_audio_44k(int(g["src_seed"]), 1.5)[: n * chunk * 2048].view(n, chunk * 2048)
    wave = torch.cat([sess.process_chunk(src[i].cuda()).cpu() for i in range(n)])
    src_hist, pred_hist = sess.history()
    sess.close()
    assert np.array_equal(src_hist.numpy()[None], g["src_content"])
    assert np.array_equal(pred_hist.numpy()[None], g["pred_codes"])
    mse = float(((wave.numpy() - g["wave"]) ** 2).mean())
    assert mse < WAVE_MSE_TOL, mse
