"""One steady-state lock-step batch step inside a cudaProfilerStart/Stop window (many-stream path):

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file l.csv \
        python tools/profile_batch.py 128
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:arb_attn_slow -c 2 -o prof \
        python tools/profile_batch.py 128 [prompt seconds, default 5] [steps to advance before the window, default 0]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["48"])
import tools.bench_batch as bb  # noqa: E402

if __name__ == "__main__":
    from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder, synth
    ar = ARVCWrapper()
    ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
    ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
    ContentTokenizer().load_state_dict(synth.make_tokenizer_state_dict(1234), strict=False)
    Vocoder().load_state_dict(synth.make_vocoder_state_dict(1234), strict=False)
    prompt_s = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
    advance = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    print(bb.run(int(sys.argv[1]), steps=2, warm=4, profile_steps=1, prompt_s=prompt_s, advance=advance))
