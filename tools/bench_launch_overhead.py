"""Is the single-stream loop bound by the host's launch rate or by the GPU's dependent kernel chain?

Runs the steady-state chunk loop of one stream (device buffers, CLI-default windows) and prints one JSON line:
  host_issue_ms   wall time the host spends inside `process_chunk` per chunk when nothing waits for the GPU
                  (calls are issued back to back, one synchronize at the very end).  NOTE: over 100 chunks the driver's
                  launch queue (~1000 entries) fills and the host is throttled to the device's pace, so this figure can
                  never come out far below device_ms;
  host_burst_ms   the same for bursts of TWO chunks issued into an EMPTY queue (synchronize before every burst): the
                  host's true cost of issuing one chunk;
  device_ms       CUDA-event time per chunk over the back-to-back calls;
  launches        kernel launches per chunk.
host_issue_ms close to device_ms means the host is the limiter (CUDA graphs of the E and V stages would pay);
host_issue_ms well below device_ms means the kernels' own latency chain is (persistent phase kernels would).

    python tools/bench_launch_overhead.py [chunks]"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, StreamSession, Vocoder, _lib, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = 1234
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(seed), strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(synth.make_tokenizer_state_dict(seed), strict=False)
    voc = Vocoder()
    voc.load_state_dict(synth.make_vocoder_state_dict(seed), strict=False)
    ref_wave = synth.synth_audio_44k(5000, 5.0)[None].cuda()
    ref_content, _ = tok.encode(ref_wave, torch.LongTensor([ref_wave.shape[1]]))
    T = ref_content.shape[-1]
    g = torch.Generator().manual_seed(1)
    ref_audio = torch.randint(0, 1000, (1, 8, T), generator=g).int().cuda()
    style, timbre = synth.synth_speaker(5000)
    sess = StreamSession()
    sess.set_prompt(ref_content[0], ref_audio, style.cuda(), timbre.cuda(), max_prompt_frames=256, delay=2)
    sess.setup(128, 64, 768, 32, 1)
    warm = 8
    src = synth.synth_audio_44k(1000, (n + warm + 2) * 2048 / 44100 + 0.1)[: (n + warm) * 2048].view(n + warm, 2048).cuda()
    out = torch.empty(2048, device="cuda")
    for i in range(warm):
        sess.process_chunk(src[i], out)
    torch.cuda.synchronize()
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for i in range(warm, warm + n):
        sess.process_chunk(src[i], out)
    host = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    launches = (_lib.kernel_launches() - l0) // n
    dev_ms = e0.elapsed_time(e1) / n
    burst = []
    for r in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sess.process_chunk(src[warm + (2 * r) % n], out)
        sess.process_chunk(src[warm + (2 * r + 1) % n], out)
        burst.append((time.perf_counter() - t0) / 2)
    burst.sort()
    print(json.dumps({"chunks": n, "host_issue_ms": round(host / n * 1e3, 4), "host_burst_ms": round(burst[len(burst) // 2] * 1e3, 4),
                      "host_burst_us_per_launch": round(burst[len(burst) // 2] * 1e6 / launches, 2),
                      "device_ms": round(dev_ms, 4), "launches": launches}))
    sess.close()


if __name__ == "__main__":
    main()
