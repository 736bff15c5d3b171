"""Pure cost of the grid barrier of the persistent decode kernels (148 CTAs x 512 threads, cooperative launch)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from streamvoiceanon_b200 import ARVCWrapper, _lib, synth  # noqa: E402

ar = ARVCWrapper()
ar.setup_caches(max_batch_size=1, max_seq_len=2048, dtype=torch.float16)
ar.load_state_dict(synth.make_ar_state_dict(1234), strict=False)
lib = _lib.load()
for exchange in (0, 1):
    ms = C.c_float()
    iters = 2000
    _lib.check(lib.svanon_debug_grid_barrier(ar._engine.handle, iters, exchange, C.byref(ms)))
    print(f"grid barrier (arrival counter), exchange {exchange}: {ms.value / iters * 1e3:.2f} us per barrier")
