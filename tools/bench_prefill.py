"""Prompt prefill / re-prompt cost: ARVCWrapper.prefill_prompt for T prompt frames = 33 + 2T tokens
through the multi-token path (tensor-core GEMMs + tiled attention), CUDA-event timed."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import ARVCWrapper, synth  # noqa: E402

ar = ARVCWrapper()
ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
ar.set_delay(delay=2)
style, timbre = synth.synth_speaker(5000)
g = torch.Generator().manual_seed(1)
for T in (107, 256, 288):
    rc = torch.randint(0, 8192, (1, T), generator=g).cuda()
    ra = torch.randint(0, 1000, (1, 8, T), generator=g).int().cuda()
    for _ in range(3):
        ar.prefill_prompt(rc, ra, style.cuda(), timbre.cuda())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ar.prefill_prompt(rc, ra, style.cuda(), timbre.cuda())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tok = 33 + 2 * T
    gflop = tok * 2 * 92_030_976 / 1e9
    print(f"prefill_prompt T={T} frames ({tok} tokens): {ms:.2f} ms  ({gflop / ms:.1f} TFLOP/s fp32-grade on the projections)")
