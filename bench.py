#!/usr/bin/env python
"""Benchmark of the per-chunk streaming loop (BASELINE.json configs[1]: --simulate_streaming,
decode_chunk_frames=1, delay=2, single stream per GPU) and of BASELINE config 4 (many concurrent streams per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of InferenceWrapper.process_one_chunk (evaluations/infer_arvc.py:492-596) over one
2048-sample chunk of synthetic audio: wave-ring update, content-encoder window re-encode (128 frames), one
dual-AR decode step, vocoder on the last 64 code frames, tail select.  Prints ONE JSON line (rank 0).

  value  frames/s with the chunk already resident in HBM, device-timed (CUDA events on the launching stream)
  e2e    the same loop through the C ABI with HOST buffers: pinned-host chunk in, host waveform out, the
         host<->device copies and the per-chunk synchronisation inside the timed region
  N > 1  independent replicas, one process per GPU, no collective ("replicas only"); value is the
         sum over ranks / max-over-ranks time
  --impl reference   the reference loop on the host cores: the UNMODIFIED reference (byte-compiled into oracle/_ref by
                     oracle/build_ref.py, kind "reference") when present, else the CPU oracle port (oracle/streaming.py)
  roofline           the kernel with the largest share of the step -- the tcgen05 3xTF32 GEMM (gemm_tc.cu; stage E's window
                     encode and stage V's wide levels): executed flops of the step's ACTUAL GEMM shapes / their event-timed
                     launch durations (svanon_gemm_timing) against the TF32 tensor peak
  roofline_ar        the stage-A decode kernel against measured HBM bandwidth
  roofline_gemm_many_streams   the wide-GEMM kernel of the many-stream batches (gemm_pair.cu) on the encoder MLP shape
  stage_compute      executed GFLOP / stage time of the compute-bound stages E and V against the 3xTF32 ceiling
  concurrent_streams BASELINE config 4 on every GPU: B = 128 streams in lock-step (svanon_batch_process_chunk) for >= 700
                     chunks with HOST buffers, so that every stream's re-prompt falls inside the window: mean / p50 / p99 /
                     max step ms, aggregate frames/s; at N = 1 also the largest B of a short ladder whose p99 stays under
                     the 46.44 ms frame period ("concurrent streams/GPU at RTF < 1")
  prompt_path        (N = 1) the setup path beside the headline: calculate_prompt on 5 s of reference audio, per step, timed by
                     tools/bench_prompt.py in a child process with a time limit; {"unavailable": why} if that fails
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAME_S = 2048 / 44100.0
WORKLOAD = dict(workload="streaming chunk=1 delay=2, single stream per GPU (BASELINE configs[1])",
                encode_window_frames=128, decode_window_frames=64, max_prompt_frames=256, max_seq_frames=768,
                buffer_frames=32, decode_chunk_frames=1, delay=2, prompt_frames=107, source_seconds=10.0,
                weights="seeded random fp32 (streamvoiceanon_b200.synth, seed 1234)",
                parallelism="replicas: one process and one stream per GPU, no collective on the data path",
                l2="per-step working set (~0.8 GB of fp32 weights streamed once) exceeds the 126 MB L2; no explicit flush")
# algorithmic bytes of one AR decode launch (fp32 weights): every slow/fast layer, norms,
# fast_output and the touched embedding rows once; the discarded 768->8192 head is skipped
AR_WEIGHT_PARAMS = 129_782_016 - 6_291_456 - 768


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the `ncu --set full` capture recorded in
    profiles/ncu_traffic.json (written by tools/ncu_traffic.py from the capture of the commit it names); None when no
    capture of this round exists -- never a constant carried over from an older build."""
    f = ROOT / "profiles" / "ncu_traffic.json"
    if not f.exists():
        return None, None
    try:
        row = json.loads(f.read_text()).get(kernel)
        return (row["dram_bytes_per_launch"], row["source"]) if row else (None, None)
    except Exception:
        return None, None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=10, help="chunks of the CPU baseline sample (main arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--concurrent", default="128,192,240,256", help="stream-count ladder of the config-4 leg: the first entry runs on "
                    "every GPU, the rest (N=1 only) are tried in order while p99 stays under the frame period ('' = skip)")
    ap.add_argument("--concurrent-chunks", type=int, default=720, help="chunks per stream of the config-4 leg")
    ap.add_argument("--stateful", default="384,512,640,768", help="stream-count ladder of the stateful-encoder leg (N=1 only; '' = skip)")
    ap.add_argument("--stateful-chunks", type=int, default=150, help="chunks per stream of the stateful-encoder leg")
    ap.add_argument("--config5", type=int, default=128, help="streams per GPU of the BASELINE config-5 leg (0 = skip)")
    ap.add_argument("--config5-steps", type=int, default=240, help="two-frame chunks per stream of the config-5 leg")
    ap.add_argument("--perf", default="128,256,320", help="stream-count ladder of the perf-mode leg, window encoder (N=1 only; '' = skip)")
    ap.add_argument("--perf-stateful", default="512,640,768", help="the same with the stateful encoder")
    ap.add_argument("--no-prompt-path", action="store_true", help="skip the setup-path leg (child process, N=1 only)")
    return ap.parse_args()


def cpu_prompt_path(threads: int):
    """CPU side of the prompt_path leg: the oracle's `calculate_prompt` (oracle/prompt.py: resample, both speaker encoders,
    noise mix, codec ids, content ids; torch fp32) on the same 5 s of synthetic reference audio, host threads as given."""
    import numpy as np
    from oracle import prompt as P
    from streamvoiceanon_b200 import synth
    torch.set_num_threads(threads)
    ref = synth.synth_audio_44k(5000, 5.0)[None]
    sds = (synth.make_campplus_state_dict(1234), synth.make_timbre_encoder_state_dict(1234), synth.make_tokenizer_state_dict(1234),
           synth.make_vocoder_encoder_state_dict(1234))
    z1, z2 = np.zeros((1, 192), np.float32), np.zeros((1, 32, 128), np.float32)
    ms = []
    with torch.no_grad():
        for _ in range(3):
            t0 = time.perf_counter()
            P.calculate_prompt([ref], 0.7, z1, z2, *sds)
            ms.append((time.perf_counter() - t0) * 1e3)
    return {"ms": sorted(ms)[1], "cores": threads, "kind": "port",
            "sample": "oracle/prompt.py calculate_prompt, 5 s of reference audio, median of 3 calls, torch fp32"}


def prompt_path_leg(timeout_s: int = 150):
    """Setup path beside the headline (never inside it): `PromptBuilder.calculate_prompt` on 5 s of reference audio --
    resample, both speaker encoders, noise mix, codec ids, content ids -- per step, CUDA-event ms and kernel launches,
    measured by tools/bench_prompt.py in a CHILD process with a hard time limit, so that nothing this leg does (its
    speaker-encoder kernels were written after round 1's GPU minutes ran out and first execute here and in
    tests/test_zz_gpu_speaker.py) can cost the bench line.  Returns the tool's JSON row or {"unavailable": why}."""
    what = ("setup path, once per stream: streamvoiceanon_b200.prompt.PromptBuilder.calculate_prompt on 5 s of synthetic "
            "reference audio (tools/bench_prompt.py, child process): per step CUDA-event ms after warm-up and kernel launches")
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_prompt.py"), "5"], capture_output=True, text=True,
                           timeout=timeout_s)
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not rows:
            tail = (r.stderr or r.stdout).strip().splitlines()[-1:] or ["no output"]
            return {"what": what, "unavailable": f"exit {r.returncode}: {tail[0][:300]}"}
        return dict(json.loads(rows[-1]), what=what)
    except Exception as exc:                                                   # timeout included
        return {"what": what, "unavailable": repr(exc)[:300]}


def make_inputs(rank: int):
    from streamvoiceanon_b200 import synth
    ref_wave = synth.synth_audio_44k(5000 + rank, 5.0)
    ref_wave = ref_wave[: (ref_wave.numel() // 2048) * 2048]
    src = synth.synth_audio_44k(1000 + rank, WORKLOAD["source_seconds"])
    src = src[: (src.numel() // 2048) * 2048].view(-1, 2048)
    n_ref = ref_wave.numel() // 2048
    g = torch.Generator().manual_seed(99 + rank)
    ref_audio = torch.randint(0, 1000, (1, 8, n_ref), generator=g).int()       # stand-in for the setup-path codec encoder
    style, timbre = synth.synth_speaker(5000 + rank)
    return ref_wave[None], ref_audio, style, timbre, src


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_loop(rank: int, n_chunks: int, warm: int, threads: int):
    """The reference loop's CPU port (oracle/streaming.py), timed on the host cores: median ms per chunk."""
    from oracle import content_encoder as E
    from oracle.streaming import StreamOracle
    from streamvoiceanon_b200 import synth
    torch.set_num_threads(threads)
    ar_sd = synth.make_ar_state_dict(1234)
    tok_sd = synth.make_tokenizer_state_dict(1234)
    voc_sd = synth.fold_weight_norm(synth.make_vocoder_state_dict(1234))
    ref_wave, ref_audio, style, timbre, src = make_inputs(rank)
    noise = {}

    def noise_fn(step, slot, V):
        if step not in noise:
            noise.clear()
            noise[step] = synth.noise_tape(7000 + rank, step)
        return noise[step][slot]
    so = StreamOracle(ar_sd, tok_sd, voc_sd, noise_fn)
    with torch.no_grad():
        ref_content = E.encode(ref_wave, tok_sd)[0].squeeze(0)
        so.prefill_prompt(ref_audio, ref_content, style, timbre, WORKLOAD["max_prompt_frames"], WORKLOAD["delay"])
        so.setup_stream_caches(WORKLOAD["encode_window_frames"], WORKLOAD["decode_window_frames"],
                               WORKLOAD["max_seq_frames"], WORKLOAD["buffer_frames"], 1)
        times = []
        for i in range(warm + n_chunks):
            t0 = time.perf_counter()
            so.process_one_chunk(src[i % src.shape[0]][None])
            if i >= warm:
                times.append(time.perf_counter() - t0)
    times.sort()
    stage = {k: (sorted(v)[len(v) // 2] * 1e3 if v else None) for k, v in so.timings.items()}
    return sum(times) / len(times) * 1e3, times[len(times) // 2] * 1e3, stage


def cpu_reference_loop(rank: int, n_chunks: int, warm: int, threads: int):
    """The UNMODIFIED reference's `InferenceWrapper.process_one_chunk` (evaluations/infer_arvc.py:492-596) with the
    reference's own three models, from oracle/_ref (byte-compiled from /root/reference by oracle/build_ref.py; the source
    tree itself in the build container), on the host cores: same inputs, windows and weights as the engine arm.  The
    setup-path encoders are replaced by supplied tensors (oracle/ref_harness.make_inference_wrapper), as in the fixtures;
    external patches of SURVEY section 8c-3 only (fp32 KV cache, torch.cuda.Event no-ops, noise tape)."""
    import contextlib
    import io
    import re
    from oracle import ref_harness
    from streamvoiceanon_b200 import synth
    torch.set_num_threads(threads)
    ref_wave, ref_audio, style, timbre, src = make_inputs(rank)
    noise = {}

    def noise_fn(step, slot, V):
        if step not in noise:
            noise.clear()
            noise[step] = synth.noise_tape(7000 + rank, step)
        return noise[step][slot]
    model, tok, voc, tape = ref_harness.build(synth.make_ar_state_dict(1234), synth.make_tokenizer_state_dict(1234),
                                              synth.make_vocoder_state_dict(1234), noise_fn)
    w = ref_harness.make_inference_wrapper(model, tok, voc, style, timbre, ref_audio)
    times, stage = [], {"E": [], "A": [], "V": []}
    pat = re.compile(r"Time taken for (content encoder|AR|vocoder): ([0-9.eE+-]+)ms")
    names = {"content encoder": "E", "AR": "A", "vocoder": "V"}
    with torch.no_grad():
        tape.step = -1
        w.prefill_prompt([ref_wave], max_prompt_frames=WORKLOAD["max_prompt_frames"], delay=WORKLOAD["delay"])
        w.setup_stream_caches(WORKLOAD["encode_window_frames"], WORKLOAD["decode_window_frames"], WORKLOAD["max_seq_frames"],
                              WORKLOAD["buffer_frames"], 1)
        for i in range(warm + n_chunks):
            buf = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(buf):                    # the reference prints its three stage timings
                w.process_one_chunk(src[i % src.shape[0]][None])
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
                for k, v in pat.findall(buf.getvalue()):
                    stage[names[k]].append(float(v))
    times.sort()
    med = {k: (sorted(v)[len(v) // 2] if v else None) for k, v in stage.items()}
    return sum(times) / len(times) * 1e3, times[len(times) // 2] * 1e3, med


def cpu_baseline(n_chunks: int, warm: int):
    """CPU arm on this box's host cores: the unmodified reference when oracle/_ref (or /root/reference) is there
    (kind "reference"), else the oracle port (kind "port")."""
    cores = os.cpu_count() or 1
    from oracle import ref_harness
    if ref_harness.available():
        mean_ms, med_ms, stage = cpu_reference_loop(0, n_chunks, warm, cores)
        kind = "reference"
        what = (f"the UNMODIFIED reference's InferenceWrapper.process_one_chunk ({ref_harness.kind()} of the reference's own "
                f"modules, oracle/build_ref.py) with its own tokenizer / dual-AR / vocoder modules")
    else:
        mean_ms, med_ms, stage = cpu_oracle_loop(0, n_chunks, warm, cores)
        kind = "port"
        what = "oracle/streaming.py (CPU port of process_one_chunk)"
    return {"value": 1e3 / mean_ms, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{n_chunks} chunks of the same workload (same stream, windows, weights) after {warm} warm-up chunks; {what}; "
                      f"torch fp32, {cores} threads",
            "ms_per_step": mean_ms, "ms_per_step_median": med_ms, "stage_ms_median": stage}


def _pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def concurrent_leg(tok, B, chunks, rank=0, warm=5, enc_mode=None):
    """BASELINE config 4 on this GPU: B streams advanced in lock-step by ONE library call per chunk
    (svanon_batch_process_chunk), CLI-default windows, delay 2, pinned HOST buffers in and out (each call returns after
    its device->host copy), `chunks` chunks per stream.  Reference audio of 2.8-5 s per stream (60..107 prompt frames), so
    the streams' re-prompts (evaluations/infer_arvc.py:547-564) fall on DIFFERENT chunks near the end of the window instead
    of all on one.  Per-step wall time (host clock around the call, which synchronises) -> mean / p50 / p99 / max."""
    from streamvoiceanon_b200 import BatchSession, StreamSession, _lib, synth
    base_wave = synth.synth_audio_44k(5000 + rank, 5.0)
    style, timbre = synth.synth_speaker(5000 + rank)
    contents = {}
    sessions = []
    t_setup = time.perf_counter()
    for b in range(B):
        n_ref = 60 + (b * 48) // B
        if n_ref not in contents:
            w = base_wave[: n_ref * 2048][None].cuda()
            contents[n_ref] = tok.encode(w, torch.LongTensor([w.shape[1]]).cuda())[0][0]
        g = torch.Generator().manual_seed(99 + b)
        s = StreamSession()
        s.set_sampling(0.7, 0.7, seed=7000 + 1000 * rank + b)
        s.set_prompt(contents[n_ref], torch.randint(0, 1000, (1, 8, n_ref), generator=g).int().cuda(), style.cuda(),
                     timbre.cuda(), WORKLOAD["max_prompt_frames"], WORKLOAD["delay"])
        sessions.append(s)
    batch = BatchSession(sessions)
    if enc_mode is not None:
        batch.set_encoder_mode(enc_mode)
    batch.setup(WORKLOAD["encode_window_frames"], WORKLOAD["decode_window_frames"], WORKLOAD["max_seq_frames"],
                WORKLOAD["buffer_frames"], 1)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    n_src = 40
    src = torch.stack([synth.synth_audio_44k(1000 + (b % 8), 2.0)[: n_src * 2048] for b in range(B)])
    pin_in = src.view(B, n_src, 2048).transpose(0, 1).contiguous().pin_memory()          # [chunk][B][2048]
    pin_out = torch.empty(B, 2048).pin_memory()
    lib = _lib.load()
    pos = lambda: [int(lib.svanon_ar_position(s._h)) for s in sessions]                 # noqa: E731
    it = 0
    for _ in range(warm + WORKLOAD["delay"]):
        batch.process_chunk(pin_in[it % n_src], pin_out); it += 1
    torch.cuda.synchronize()
    last, reprompts, steps_with = pos(), 0, 0
    wall = []
    l0 = _lib.kernel_launches()
    t_all = time.perf_counter()
    for _ in range(chunks):
        t0 = time.perf_counter()
        batch.process_chunk(pin_in[it % n_src], pin_out); it += 1
        wall.append((time.perf_counter() - t0) * 1e3)
        now = pos()
        k = sum(1 for a, b_ in zip(last, now) if b_ < a)
        reprompts += k
        steps_with += 1 if k else 0
        last = now
    t_all = time.perf_counter() - t_all
    launches = _lib.kernel_launches() - l0
    # host issue time vs device time with DEVICE buffers (the call returns without synchronising): is the host the limiter?
    dev_in, dev_out = pin_in.cuda(), torch.empty(B, 2048, device="cuda")
    for _ in range(3):
        batch.process_chunk(dev_in[it % n_src], dev_out); it += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_dev = 30
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n_dev):
        batch.process_chunk(dev_in[it % n_src], dev_out); it += 1
    host_ms = (time.perf_counter() - t0) * 1e3 / n_dev
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / n_dev
    batch.set_timing(True)
    batch.process_chunk(dev_in[0], dev_out)
    st = batch.last_timing()
    batch.close()
    for s in sessions:
        s.close()
    mean = sum(wall) / len(wall)
    return {"streams": B, "chunks": chunks, "ms_per_step_mean": mean, "ms_per_step_p50": _pct(wall, 0.5),
            "ms_per_step_p99": _pct(wall, 0.99), "ms_per_step_max": max(wall), "rtf_mean": mean / 1e3 / FRAME_S,
            "rtf_p99": _pct(wall, 0.99) / 1e3 / FRAME_S, "frames_per_s": B * chunks / t_all,
            "reprompts_in_window": reprompts, "steps_with_a_reprompt": steps_with,
            "steps_over_frame_period": sum(1 for w in wall if w > FRAME_S * 1e3),
            "h2d_bytes_per_step": B * 2048 * 4, "d2h_bytes_per_step": B * 2048 * 4, "gpu_launches_per_step": launches / chunks,
            "device_resident": {"ms_per_step": dev_ms, "host_issue_ms_per_step": host_ms,
                                "note": "device buffers, no synchronisation inside the call: host time to ISSUE one step "
                                        "beside the device time of the step; host-bound when the two are close"},
            "stage_ms": {"E": st[0], "A": st[1], "V": st[2]}, "setup_s": t_setup}


def config5_leg(tok, voc, B, steps, rank=0, warm=4):
    """BASELINE config 5 on this GPU: anonymisation alpha = 0.7 with THREE mixed references per stream, streaming with
    two-frame chunks, delay 2.  The prompts come from the real prompt path -- `PromptBuilder.calculate_prompt` on the
    concatenation of three synthetic references (both speaker encoders, noise mix with torch's generator, codec and content
    ids; evaluations/infer_arvc.py:382-441), eight distinct prompts of 3 x 4.3..5.0 s shared round-robin by the B streams
    (prefilled untruncated, kept truncated to 256 frames: :468-489) -- then B streams in lock-step, HOST buffers,
    `steps` chunks of 2 frames each, every stream re-prompting once inside the window."""
    from streamvoiceanon_b200 import BatchSession, StreamSession, synth
    from streamvoiceanon_b200.prompt import PromptBuilder
    from streamvoiceanon_b200.speaker import CAMPPlus, SpeakerEncoder
    style_enc, timbre_enc = CAMPPlus(), SpeakerEncoder()
    style_enc.load_state_dict(synth.make_campplus_state_dict(1234))
    timbre_enc.load_state_dict(synth.make_timbre_encoder_state_dict(1234))
    pb = PromptBuilder(tok, voc, style_enc, timbre_enc)
    torch.manual_seed(77 + rank)
    t0 = time.perf_counter()
    prompts = []
    for k in range(8):
        sec = 4.3 + 0.1 * k
        refs = [synth.synth_audio_44k(5200 + 10 * k + j + 100 * rank, sec)[None].cuda() for j in range(3)]
        prompts.append(pb.calculate_prompt(refs, 0.7, "concat_mel"))
    torch.cuda.synchronize()
    t_prompt = (time.perf_counter() - t0) / 8
    sessions = []
    for b in range(B):
        codes, content, style, timbre, _ = prompts[b % 8]
        s = StreamSession()
        s.set_sampling(0.7, 0.7, seed=9000 + 1000 * rank + b)
        s.set_prompt(content[0], codes, style, timbre, WORKLOAD["max_prompt_frames"], 2)
        sessions.append(s)
    batch = BatchSession(sessions)
    batch.setup(WORKLOAD["encode_window_frames"], WORKLOAD["decode_window_frames"], WORKLOAD["max_seq_frames"],
                WORKLOAD["buffer_frames"], 2)
    n_src = 20
    src = torch.stack([synth.synth_audio_44k(1000 + (b % 8), 2.0)[: n_src * 4096] for b in range(B)])
    pin_in = src.view(B, n_src, 4096).transpose(0, 1).contiguous().pin_memory()
    pin_out = torch.empty(B, 4096).pin_memory()
    lib = _lib_mod().load()
    pos = lambda: [int(lib.svanon_ar_position(s._h)) for s in sessions]                 # noqa: E731
    it = 0
    for _ in range(warm):
        batch.process_chunk(pin_in[it % n_src], pin_out); it += 1
    last, reprompts, wall = pos(), 0, []
    t_all = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        batch.process_chunk(pin_in[it % n_src], pin_out); it += 1
        wall.append((time.perf_counter() - t1) * 1e3)
        now = pos()
        reprompts += sum(1 for a, b_ in zip(last, now) if b_ < a)
        last = now
    t_all = time.perf_counter() - t_all
    batch.close()
    for s in sessions:
        s.close()
    period = 2 * FRAME_S * 1e3
    mean = sum(wall) / len(wall)
    return {"what": "BASELINE config 5: alpha 0.7, three mixed references (concat_mel, prompt path on the GPU), chunk = 2 frames, delay 2",
            "streams": B, "steps": steps, "chunk_frames": 2, "chunk_period_ms": period, "ms_per_step_mean": mean,
            "ms_per_step_p50": _pct(wall, 0.5), "ms_per_step_p99": _pct(wall, 0.99), "ms_per_step_max": max(wall),
            "rtf_mean": mean / period, "rtf_p99": _pct(wall, 0.99) / period, "frames_per_s": 2 * B * steps / t_all,
            "reprompts_in_window": reprompts, "steps_over_chunk_period": sum(1 for w in wall if w > period),
            "prompt_frames": [int(p[1].shape[-1]) for p in prompts], "calculate_prompt_ms": t_prompt * 1e3,
            "h2d_bytes_per_step": B * 4096 * 4, "d2h_bytes_per_step": B * 4096 * 4}


def _lib_mod():
    from streamvoiceanon_b200 import _lib
    return _lib


def perf_mode_leg(tok, args, rank):
    """PERF mode beside the parity-mode line, never instead of it: the many-stream loop with fp16 single-pass tensor-core
    GEMMs (svanon_set_precision 1 -- the reference's own GPU precision: fp16 autocast, evaluations/infer_arvc.py:493), with
    the window encoder and with the stateful encoder, plus what the mode costs in fidelity (tools/eval_perf_mode.py in a child
    process: content-id agreement, teacher-forced codec-id agreement, fast-head logit error, vocoder SNR vs parity mode)."""
    from streamvoiceanon_b200.engine import Engine
    eng = Engine.get(torch.cuda.current_device())
    frame_ms = FRAME_S * 1e3
    out = {"what": "svanon_set_precision(1): one kind::f16 tcgen05 pass per GEMM (activations rounded to fp16 on the way into the "
                   "tensor core, fp16 weight copies, fp32 accumulation) instead of the 3xTF32 split; ids are NOT bit-exact in this "
                   "mode (see `fidelity`); the headline `value` / `e2e` / `concurrent_streams` above are parity mode",
           "dtype": "f16 operands, f32 accumulate", "window_encoder": [], "stateful_encoder": []}
    eng.set_precision(1)
    try:
        for key, ladder, mode in (("window_encoder", args.perf, None), ("stateful_encoder", args.perf_stateful, 3)):
            for B in [int(x) for x in ladder.split(",") if x.strip()]:
                try:
                    out[key].append(concurrent_leg(tok, B, args.stateful_chunks, rank, enc_mode=mode))
                except Exception as exc:
                    out[key].append({"streams": B, "error": repr(exc)[:300]})
                    break
                if out[key][-1]["ms_per_step_p99"] >= frame_ms:
                    break
            ok = [r["streams"] for r in out[key] if "error" not in r and r["ms_per_step_p99"] < frame_ms]
            out[f"max_streams_per_gpu_p99_lt_frame_period_{key}"] = max(ok) if ok else None
    finally:
        eng.set_precision(0)
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "eval_perf_mode.py"), "32", "40"], capture_output=True, text=True,
                           timeout=240)
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        out["fidelity"] = json.loads(rows[-1]) if rows else {"unavailable": (r.stderr or r.stdout).strip().splitlines()[-1:][:1]}
    except Exception as exc:
        out["fidelity"] = {"unavailable": repr(exc)[:300]}
    return out


def gemm_roofline(peaks):
    """Tensor-pipe roofline of the wide-GEMM kernel of the many-stream batches (gemm_pair.cu: CTA pairs, operands by tensor-map
    TMA) on the encoder MLP shape (M = 16384 rows = 128 streams x 128 window tokens, N = 2048, K = 512, +bias, GELU), timed
    live with CUDA events the way the engine runs it: engine-owned (static) weights, the lo term of A already written by
    the producer of A.  `with_split_pass_us` is the same launch when the kernel has to split A itself; `single_cta_us` is
    the round-1/2 kernel (gemm_tc.cu, 128 x 256 tile) on the same problem.  Every fp32-grade product costs three TF32 MMAs, so
    `achieved` counts 3 x 2MNK TF32 flops; `peak` = half the measured dense bf16 throughput of MEASURED_PEAKS.json (TF32 runs at
    half the bf16 rate on B200), else half the nominal 2250."""
    from streamvoiceanon_b200 import _lib
    from streamvoiceanon_b200.engine import Engine, ptr
    eng, lib = Engine.get(torch.cuda.current_device()), _lib.load()
    M, N, K = 16384, 2048, 512
    A = torch.randn(M, K, device="cuda")
    A_lo = A - (A.view(torch.int32) & -8192).view(torch.float32)          # x - trunc_tf32(x), exact in fp32
    Ws = [torch.randn(N, K, device="cuda") for _ in range(8)]
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")

    def timed(pair_mode, a_lo):
        _lib.check(lib.svanon_set_gemm_pair(pair_mode))
        _lib.check(lib.svanon_debug_gemm_alo(ptr(a_lo) if a_lo is not None else None))
        for i in range(8):
            _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(Ws[i % 8]), ptr(b), ptr(out), M, N, K, 1, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40
        n0 = lib.svanon_gemm_pair_launches()
        e0.record()
        for i in range(n):
            _lib.check(lib.svanon_debug_gemm(eng.handle, ptr(A), ptr(Ws[i % 8]), ptr(b), ptr(out), M, N, K, 1, None))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3, lib.svanon_gemm_pair_launches() - n0

    _lib.check(lib.svanon_debug_gemm_weights_static(2))       # like engine weights: converted copies cached per weight pointer
    try:
        us, took = timed(1, A_lo)
        us_split, _ = timed(1, None)
        us_single, _ = timed(0, None)
    finally:
        _lib.check(lib.svanon_set_gemm_pair(-1))
        _lib.check(lib.svanon_debug_gemm_alo(None))
        _lib.check(lib.svanon_debug_gemm_weights_static(0))
    fp32_tflops = 2.0 * M * N * K / us / 1e6
    bf16 = peaks.get("bf16_tflops") or peaks.get("bf16_tflops_sustained")        # kernel timed alone: the burst figure
    peak, src = (bf16 / 2, "0.5 x measured dense bf16 burst (MEASURED_PEAKS.json)") if bf16 else (1125.0, "0.5 x nominal 2250 bf16")
    return {"kernel": "gemm_pair_kernel<256> (tcgen05.mma.cta_group::2 3xTF32, 256x256 tile per CTA pair, A and B by tensor-map "
                      "TMA, persistent with two TMEM accumulators)", "bound": "tensor", "shape": [M, N, K],
            "where": "encoder transformer MLP of 128 lock-step streams (M = 128 x 128 window tokens); does not occur in the "
                     "single-stream step", "pair_kernel_launches_timed": int(took),
            "launch_us": us, "with_split_pass_us": us_split, "single_cta_us": us_single,
            "fp32_equivalent_tflops": fp32_tflops, "achieved": 3 * fp32_tflops, "peak": peak,
            "unit": "TFLOP/s", "frac": 3 * fp32_tflops / peak, "peak_source": src,
            "traffic": None, "note": "profiles/r2x_gemm_pair_ncu_summary.txt: tensor pipe active 79 % of elapsed cycles (the single-CTA "
            "kernel: 34 %); what is left is the first tile's fill, the last tile's epilogue and 512 tiles on 74 pairs (6.9 rounds)"}


# Executed floating-point work per stream and chunk (chunk = 1), DESIGN.md section 4.  Stage E: conv stack incl. the
# DFT-as-GEMM spectrogram 26.3 GFLOP per 128 frames = 0.2055 per content frame, window transformer 6.98 GFLOP of which the
# last layer runs for the kept token only (-0.67); a single stream sends two 41-frame spans through the conv stack, >= 8
# lock-step streams one span + 0.2 GFLOP for the newest frames from conv history.  Stage V: 2.647 GFLOP per frame
# (incremental vocoder).  The reference computes 29.0 (E) and 169.4 (V) GFLOP for the same chunk.
# From 8 lock-step streams the window-start span recomputes only the rows the zero padding reaches ("rings of
# steady-state layer inputs"): 3 log-mel rows of DFT (0.026), stem on 9 rows, ConvNeXt block j on 9 + 6 j rows (3.13 over the 18
# blocks), stage transitions 0.05, the two down-sampled levels 0.53 = 3.74 GFLOP instead of the span's 41 * 0.2055 = 8.43
# (SVANON_ENC_HEAD_TRI=0 restores the whole span, and this accounting).
E_GFLOP_SINGLE = 82 * 0.2055 + 6.98 - 0.67
E_HEAD_SPAN_GFLOP = 41 * 0.2055 if os.environ.get("SVANON_ENC_HEAD_TRI", "1") == "0" else 3.74
E_GFLOP_MANY = E_HEAD_SPAN_GFLOP + 0.2 + 6.98 - 0.67
V_GFLOP = 2.647


def stage_compute(e_ms, v_ms, streams, peaks):
    """Achieved fp32-equivalent TFLOP/s of the two compute-bound stages against the 3xTF32 ceiling (one third of the
    TF32 tensor peak = one sixth of the measured dense bf16 peak): the compute roofline SURVEY.md section 8d asks for."""
    bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")     # stages are long steps: sustained figure
    ceiling = (bf16 / 6.0) if bf16 else 2250.0 / 6.0
    e_g = (E_GFLOP_SINGLE if streams < 8 else E_GFLOP_MANY) * streams
    v_g = V_GFLOP * streams
    out = {"streams": streams, "ceiling_fp32_equivalent_tflops": ceiling,
           "ceiling_source": ("measured dense bf16 (MEASURED_PEAKS.json)" if bf16 else "nominal 2250 bf16") + " / 6: TF32 at half the "
                             "bf16 rate, three MMAs per fp32-grade product"}
    for name, g, ms in (("E", e_g, e_ms), ("V", v_g, v_ms)):
        t = g / ms if ms > 0 else 0.0                                         # GFLOP / ms = TFLOP/s
        out[name] = {"executed_gflop": g, "ms": ms, "achieved_tflops": t, "frac": t / ceiling}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 40)
    warm = min(max(args.warmup, 3), 5)
    base = cpu_baseline(steps, warm)
    fps, mean_ms = base["value"], base["ms_per_step"]
    line = {"impl": "reference", "metric": "streaming_frames_per_sec_chunk1_delay2", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": mean_ms, "rtf": mean_ms / 1e3 / FRAME_S,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": WORKLOAD, "cpu_baseline": base,
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_engine(args):
    import torch.distributed as dist
    from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, StreamSession, Vocoder, _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(synth.make_tokenizer_state_dict(1234), strict=False)
    voc = Vocoder()
    voc.load_state_dict({**synth.make_vocoder_state_dict(1234), **synth.make_vocoder_encoder_state_dict(1234)}, strict=False)

    ref_wave, ref_audio, style, timbre, src = make_inputs(rank)
    ref_content, _ = tok.encode(ref_wave.cuda(), torch.LongTensor([ref_wave.shape[1]]).cuda())
    sess = StreamSession()
    sess.set_sampling(0.7, 0.7, seed=7000 + rank)
    sess.set_prompt(ref_content[0], ref_audio.cuda(), style.cuda(), timbre.cuda(), WORKLOAD["max_prompt_frames"], WORKLOAD["delay"])
    sess.setup(WORKLOAD["encode_window_frames"], WORKLOAD["decode_window_frames"], WORKLOAD["max_seq_frames"],
               WORKLOAD["buffer_frames"], 1)
    src_dev = src.cuda()
    n_src = src.shape[0]
    out_dev = torch.empty(2048, device="cuda")
    W, K = max(args.warmup, 3), args.steps
    it = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ("value")
    for _ in range(W):
        sess.process_chunk(src_dev[it % n_src], out_dev); it += 1
    sess.set_timing(True)
    stage = [[], [], []]
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.kernel_launches()
    pos0 = int(_lib.load().svanon_ar_position(sess._h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sess.process_chunk(src_dev[it % n_src], out_dev); it += 1
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    launches = _lib.kernel_launches() - l0
    pos1 = int(_lib.load().svanon_ar_position(sess._h))
    barrier()
    # per-stage device time (events inside the library, same stream), a short extra pass
    for _ in range(min(K, 20)):
        sess.process_chunk(src_dev[it % n_src], out_dev); it += 1
        t = sess.last_timing()
        for j in range(3):
            stage[j].append(t[j])
    sess.set_timing(False)

    # ---------------- end-to-end arm: host buffers through the C ABI
    pin_in = src.clone().pin_memory()
    pin_out = torch.empty(2048).pin_memory()
    for _ in range(3):
        sess.process_chunk(pin_in[it % n_src], pin_out); it += 1
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        sess.process_chunk(pin_in[it % n_src], pin_out); it += 1
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    barrier()

    # ---------------- GEMM timing pass: every GEMM launch bracketed by events inside the library (its own pass: the events
    # break the programmatic-dependent-launch overlap, so this pass is slower than the `value` pass)
    lib = _lib.load()
    eng_h = ar._engine.handle
    for _ in range(2):
        sess.process_chunk(src_dev[it % n_src], out_dev); it += 1
    torch.cuda.synchronize()
    _lib.check(lib.svanon_gemm_timing(eng_h, 1))
    n_gt = min(K, 20)
    e0.record()
    for _ in range(n_gt):
        sess.process_chunk(src_dev[it % n_src], out_dev); it += 1
    e1.record()
    torch.cuda.synchronize()
    gt_pass_ms = e0.elapsed_time(e1)
    import ctypes as C
    g_ms, g_gflop, g_n = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_int64 * 4)()
    _lib.check(lib.svanon_gemm_timing_read(eng_h, g_ms, g_gflop, g_n))
    _lib.check(lib.svanon_gemm_timing(eng_h, 0))
    gemm_t = {name: {"ms_per_step": g_ms[i] / n_gt, "gflop_per_step": g_gflop[i] / n_gt, "launches_per_step": g_n[i] / n_gt}
              for i, name in enumerate(("gemm_tc_kernel", "gemm_pipe_kernel", "gemm_kernel", "conv_small_kernel"))}

    # ---------------- BASELINE config 4 on this GPU (every rank): 128 streams in lock-step, host buffers
    counts = [int(x) for x in args.concurrent.split(",") if x.strip()]
    conc = []
    if counts:
        sess.close()
        sess = None
        barrier()
        sampler4 = ClockSampler(local)
        sampler4.start()
        conc.append(concurrent_leg(tok, counts[0], args.concurrent_chunks, rank))
        conc[0]["clocks"] = sampler4.stop()
        barrier()
        if world == 1:
            frame_ms = FRAME_S * 1e3
            for B in counts[1:]:
                if conc[-1]["ms_per_step_p99"] >= frame_ms:
                    break
                conc.append(concurrent_leg(tok, B, args.concurrent_chunks, rank))
    # ---------------- BASELINE config 5 on this GPU (every rank): alpha 0.7, three references, two-frame chunks
    cfg5 = None
    if counts and args.config5 > 0:
        barrier()
        try:
            cfg5 = config5_leg(tok, voc, args.config5, args.config5_steps, rank)
        except Exception as exc:
            cfg5 = {"error": repr(exc)[:400]}
        barrier()
    # opt-in mode: stateful content encoder (offline-encode semantics, svanon_stream_set_encoder_mode 3), N = 1 only
    stateful = []
    if world == 1 and counts:
        frame_ms = FRAME_S * 1e3
        for B in [int(x) for x in args.stateful.split(",") if x.strip()]:
            try:
                stateful.append(concurrent_leg(tok, B, args.stateful_chunks, rank, enc_mode=3))
            except Exception as exc:                                            # e.g. out of memory at a large stream count
                stateful.append({"streams": B, "error": repr(exc)[:300]})
                break
            if stateful[-1]["ms_per_step_p99"] >= frame_ms:
                break

    # opt-in PERF mode (svanon_set_precision 1: fp16 single-pass tensor-core GEMMs, the reference's own GPU precision), N = 1
    perf_mode = None
    if world == 1 and counts and (args.perf or args.perf_stateful):
        perf_mode = perf_mode_leg(tok, args, rank)

    t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, conc[0] if conc else None)
        per_rank5 = [None] * world
        dist.all_gather_object(per_rank5, cfg5)
    else:
        per_rank = [conc[0] if conc else None]
        per_rank5 = [cfg5]
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = dev_ms / K
    fps = world * K / (dev_ms / 1e3)
    med = [sorted(s_)[len(s_) // 2] for s_ in stage]
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    s_valid = (pos0 + pos1) / 2 + 2
    ar_bytes = AR_WEIGHT_PARAMS * 4 + (2 * 12 * 12 * 64 * 4) * s_valid + 2 * 2 * 12 * 12 * 64 * 4
    ar_ms = med[1]
    achieved = ar_bytes / (ar_ms / 1e3) / 1e9
    # dominant kernel of the step: the tcgen05 GEMM.  Three TF32 MMAs per fp32-grade product -> 3 x executed flops of TF32
    # work; peak = half the measured dense bf16 rate (TF32 runs at half the bf16 rate), sustained figure (timed inside a step)
    tc = gemm_t["gemm_tc_kernel"]
    bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
    tc_peak, tc_peak_src = (bf16 / 2, "0.5 x measured sustained dense bf16 (MEASURED_PEAKS.json)") if bf16 else (1125.0, "0.5 x nominal 2250 bf16 (fallback)")
    tc_tf32 = 3 * tc["gflop_per_step"] / tc["ms_per_step"] if tc["ms_per_step"] > 0 else 0.0      # GFLOP / ms = TFLOP/s
    tc_traffic, tc_traffic_src = ncu_traffic("gemm_tc_kernel")
    ar_traffic, ar_traffic_src = ncu_traffic("ar_decode_staged_kernel")
    line = {
        "metric": "streaming_frames_per_sec_chunk1_delay2", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_step, "rtf": ms_step / 1e3 / FRAME_S, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": WORKLOAD,
        "stage_ms_median": {"E_window_encode": med[0], "A_decode": med[1], "V_vocoder": med[2]},
        "streams_per_gpu_rtf_lt_1_sequential": int(FRAME_S * 1e3 / ms_step),
        "gpu_launches": int(launches),
        "e2e": {"value": world * K / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": 2048 * 4,
                "d2h_bytes_per_step": 2048 * 4, "ms_per_step": e2e_ms / K},
        "roofline": {"kernel": "gemm_tc_kernel + chain_kernel (tcgen05 3xTF32; every tile configuration the single-stream step "
                               "launches: stage E conv stack, stage V levels with >= 64 channels; the encoder's transformer half "
                               "is ONE chain_kernel launch whose 41 GEMM phases, row phases and grid barriers are all inside "
                               "the timed duration)",
                     "bound": "tensor", "achieved": tc_tf32, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": tc_tf32 / tc_peak if tc_peak else None, "traffic": tc_traffic, "traffic_source": tc_traffic_src,
                     "peak_source": tc_peak_src, "fp32_equivalent_tflops": tc_tf32 / 3,
                     "algorithmic_gflop_per_step": tc["gflop_per_step"], "launches_per_step": tc["launches_per_step"],
                     "avg_launch_us": tc["ms_per_step"] / tc["launches_per_step"] * 1e3 if tc["launches_per_step"] else None,
                     "share_of_step": tc["ms_per_step"] / (gt_pass_ms / n_gt),
                     "how": "svanon_gemm_timing: CUDA events around every GEMM launch on the launching stream, over "
                            f"{n_gt} steady-state chunks; achieved = 3 x sum(2MNK taps) / sum(launch durations); share = that "
                            "sum / the event-timed duration of the same pass",
                     "other_gemm_backends": {k: v for k, v in gemm_t.items() if k != "gemm_tc_kernel"}},
        "roofline_ar": {"kernel": "ar_decode_staged_kernel (one launch per frame: 12 slow + 8x4 fast layers + 8 samplers)",
                        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": ar_traffic, "traffic_source": ar_traffic_src, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": ar_bytes, "launch_ms": ar_ms, "s_valid": s_valid,
                        "share_of_step": ar_ms / ms_step},
        "clocks": clocks,
    }
    if world == 1:
        line["roofline_gemm_many_streams"] = gemm_roofline(peaks)
    if conc:
        frame_ms = FRAME_S * 1e3
        ranks = [r for r in per_rank if r]
        total_fps = sum(r["frames_per_s"] for r in ranks)
        line["concurrent_streams"] = {
            "what": "BASELINE config 4: B streams per GPU in lock-step through svanon_batch_process_chunk (one pass over the "
                    "weights per chunk for all streams), HOST buffers, every stream re-prompts once inside the window; the "
                    "reference is batch-1 (1 stream x its RTF)",
            "streams_per_gpu": counts[0], "n_gpus": world, "total_streams": counts[0] * world,
            "frames_per_s_all_gpus": total_fps,
            "ms_per_step_mean_max_over_ranks": max(r["ms_per_step_mean"] for r in ranks),
            "ms_per_step_p99_max_over_ranks": max(r["ms_per_step_p99"] for r in ranks),
            "host_issue_ms_per_step_max_over_ranks": max(r["device_resident"]["host_issue_ms_per_step"] for r in ranks),
            "device_ms_per_step_max_over_ranks": max(r["device_resident"]["ms_per_step"] for r in ranks),
            "p99_under_frame_period": all(r["ms_per_step_p99"] < frame_ms for r in ranks),
            "per_rank": ranks if world > 1 else None,
            "ladder": conc,
            "max_streams_per_gpu_p99_lt_frame_period": max([r["streams"] for r in conc if r["ms_per_step_p99"] < frame_ms], default=None),
            "max_streams_per_gpu_mean_rtf_lt_1": max([r["streams"] for r in conc if r["rtf_mean"] < 1.0], default=None)}
    if cfg5:
        ranks5 = [r for r in per_rank5 if r and "error" not in r]
        line["config5"] = dict(per_rank5[0]) if per_rank5[0] else {}
        if ranks5:
            line["config5"].update({"n_gpus": world, "total_streams": sum(r["streams"] for r in ranks5),
                                    "frames_per_s_all_gpus": sum(r["frames_per_s"] for r in ranks5),
                                    "ms_per_step_mean_max_over_ranks": max(r["ms_per_step_mean"] for r in ranks5),
                                    "ms_per_step_p99_max_over_ranks": max(r["ms_per_step_p99"] for r in ranks5)})
    if stateful:
        ok = [r["streams"] for r in stateful if "error" not in r and r["ms_per_step_p99"] < FRAME_S * 1e3]
        line["concurrent_streams_stateful_encoder"] = {
            "what": "the same lock-step loop with the STATEFUL content encoder (encoder mode 3: ids of the reference's offline "
                    "encode() of the stream so far, 0.23 GFLOP per frame, instead of the reference's 128-frame window re-encode; "
                    "opt-in, not the reference's streaming semantics -- include/svanon.h); host buffers; short window "
                    f"({args.stateful_chunks} chunks: no re-prompt inside)",
            "max_streams_per_gpu_p99_lt_frame_period": max(ok) if ok else None, "ladder": stateful}
    if perf_mode:
        line["perf_mode"] = perf_mode
    try:
        line["stage_compute"] = [stage_compute(med[0], med[2], 1, peaks)]
        best = [r for r in conc if r["rtf_mean"] < 1.0]
        if best:
            r = max(best, key=lambda r: r["streams"])
            line["stage_compute"].append(stage_compute(r["stage_ms"]["E"], r["stage_ms"]["V"], r["streams"], peaks))
    except Exception as exc:                                                   # never lose the bench line over a derived figure
        line["stage_compute"] = {"error": repr(exc)}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample, 2)
    if world == 1 and not args.no_prompt_path:
        line["prompt_path"] = prompt_path_leg()
        if not args.no_cpu_baseline:
            try:
                line["prompt_path"]["cpu_baseline"] = cpu_prompt_path(os.cpu_count() or 1)
            except Exception as exc:
                line["prompt_path"]["cpu_baseline"] = {"unavailable": repr(exc)[:300]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print to fd 1 (e.g. "NCCL version ...") was
    re-routed to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
