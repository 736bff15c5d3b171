"""Steady-state chunks of the single-stream loop for ncu, inside a cudaProfilerStart/Stop window:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \\
        python tools/profile_single.py [chunks in the window, default 2]
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 6 -o prof \\
        python tools/profile_single.py 1

Only the window's launches are profiled (warm-up, prompt prefill and weight upload are not)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, StreamSession, Vocoder, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    seed = 1234
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(seed), strict=False)
    tok = ContentTokenizer()
    tok.load_state_dict(synth.make_tokenizer_state_dict(seed), strict=False)
    Vocoder().load_state_dict(synth.make_vocoder_state_dict(seed), strict=False)
    ref_wave = synth.synth_audio_44k(5000, 5.0)[None].cuda()
    ref_content, _ = tok.encode(ref_wave, torch.LongTensor([ref_wave.shape[1]]))
    T = ref_content.shape[-1]
    g = torch.Generator().manual_seed(1)
    ref_audio = torch.randint(0, 1000, (1, 8, T), generator=g).int().cuda()
    style, timbre = synth.synth_speaker(5000)
    sess = StreamSession()
    sess.set_sampling(0.7, 0.7, seed=7000)
    sess.set_prompt(ref_content[0], ref_audio, style.cuda(), timbre.cuda(), max_prompt_frames=256, delay=2)
    sess.setup(128, 64, 768, 32, 1)
    warm = 8
    src = synth.synth_audio_44k(1000, 2.0)[: (n + warm) * 2048].view(n + warm, 2048).cuda()
    out = torch.empty(2048, device="cuda")
    for i in range(warm):
        sess.process_chunk(src[i], out)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(warm, warm + n):
        sess.process_chunk(src[i], out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sess.close()


if __name__ == "__main__":
    main()
