"""CPU oracle for the streaming voice-conversion hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain torch-fp32 (CPU) restatement of the reference's algorithm for
the three stages of `InferenceWrapper.process_one_chunk`
(/root/reference/evaluations/infer_arvc.py:492-596):

  E  content encoder   oracle/content_encoder.py
  A  dual-AR decode    oracle/dual_ar.py
  V  vocoder           oracle/vocoder.py
  loop                 oracle/streaming.py

and of the prompt (setup) path that feeds it (`InferenceWrapper.calculate_prompt`, :382-441):

  speaker encoders     oracle/speaker.py   (kaldi fbank + CAMPPlus; slaney mel + ECAPA trunk + Perceiver + FSQ)
  resample, noise mix, calculate_prompt    oracle/prompt.py
  reference wave -> codec ids              oracle/vocoder.py (wav2codes)

Every function cites the reference file:line it follows.  Nothing in the product
package (`streamvoiceanon_b200/`) imports this package: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs do,
and only as the checker / the timed CPU baseline.

Pinning: the reference ships no tests and no golden vectors, so the
oracle is pinned against outputs of the reference's own modules executed in the build
container: `oracle/make_golden.py` imports /root/reference (through the import shims in
`oracle/refshim/`), loads the same synthetic weights, runs the same inputs and writes
`tests/golden/*.npz` (`make_golden_vocenc.py`, `make_golden_noise_mix.py`, `make_golden_style.py`,
`make_golden_prompt.py` do the same for the prompt path, up to the unmodified
`prefill_prompt` + `process_one_chunk` of BASELINE config 5; `make_golden_infer.py` runs the
unmodified `infer()` and `stream_infer()` from .wav files: configs 1 and 2; `make_golden_speaker_full.py` the speaker
encoders on 15 s; `make_golden_generate_kwargs.py` `generate` with sampling arguments); `tests/test_oracle_golden.py`
checks the oracle against those fixtures on every CPU run.  The third-party FSQ arithmetic
(`vector-quantize-pytorch==1.14.24`, reference requirements.txt:26) is not vendored as a
package; it is pinned through the reference's own vendored twin
(modules/bicodec_speaker_encoder/fsq/residual_fsq.py:269-336).
"""
