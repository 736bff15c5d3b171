"""In-kernel timeline of the batch-1 decode kernel (svanon_ar_profile): where thread 0 of CTA 0 spends its cycles."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import ARVCWrapper, _lib, synth  # noqa: E402

ar = ARVCWrapper()
ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
ar.set_delay(delay=2)
lib = _lib.load()
g = torch.Generator().manual_seed(1)
T = 107
style, timbre = synth.synth_speaker(5000)
ar.prefill_prompt(torch.randint(0, 8192, (1, T), generator=g).cuda(), torch.randint(0, 1000, (1, 8, T), generator=g).int().cuda(),
                  style.cuda(), timbre.cuda())
ar.prefill_src_condition4delay(torch.randint(0, 8192, (1, 2), generator=g).cuda())
ids = torch.randint(0, 8192, (200, 1, 1), generator=g).cuda()
for i in range(20):
    ar.decode_one(ids[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20, 70):
    ar.decode_one(ids[i])
e1.record()
torch.cuda.synchronize()
base = e0.elapsed_time(e1) / 50
_lib.check(lib.svanon_ar_profile(ar._engine.handle, 1, None))
n = 50
e0.record()
for i in range(70, 70 + n):
    ar.decode_one(ids[i])
e1.record()
torch.cuda.synchronize()
prof_ms = e0.elapsed_time(e1) / n
out = (C.c_uint64 * 8)()
_lib.check(lib.svanon_ar_profile(ar._engine.handle, 0, out))
names = ["activation load + norm", "wait for staged weights", "dot products + stores", "grid barrier", "attention", "sampler", "other", "-"]
tot = sum(out)
print(f"decode_one: {base * 1e3:.1f} us per frame unprofiled (incl. host), {prof_ms * 1e3:.1f} us with counters armed")
for nm, v in zip(names, out):
    if v:
        print(f"  {nm:26s} {v / n / 1.965e3:8.1f} us per frame  {100 * v / tot:5.1f} %   (at 1965 MHz)")
print(f"  total {tot / n / 1.965e3:.1f} us per frame")
