#!/bin/bash
# Round 2, second GPU call: new attention kernel (A/B + ncu of both), stateful encoder / vocoder entries, the unmodified caller
# over the shims, the host's true launch cost.
set -u
O=gpurun_out/r2b
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stateful.py tests/test_gpu_batch.py tests/test_gpu_zz_unmodified_caller.py -x -q -rs 2>&1 | tail -40 ) > $O/pytest_new.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q -rs --deselect tests/test_gpu_stateful.py --deselect tests/test_gpu_batch.py --deselect tests/test_gpu_zz_unmodified_caller.py 2>&1 | tail -15 ) > $O/pytest_rest.txt 2>&1
timeout 120 python tools/bench_launch_overhead.py 100 > $O/launch_overhead.json 2>&1
# attention A/B at S ~ 950 (256-frame prompts, 200 steps in): step and stage-A time with the old and the new kernel
SVANON_ATTN_TMA=0 timeout 300 python -c "
import sys, json; sys.argv=['x']; sys.path.insert(0,'.')
import tools.bench_batch as bb, torch
from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder, synth
ar=ARVCWrapper(); ar.setup_caches(max_batch_size=1,max_seq_len=2048); ar.load_state_dict(synth.make_ar_state_dict(1234),strict=False)
ContentTokenizer().load_state_dict(synth.make_tokenizer_state_dict(1234),strict=False); Vocoder().load_state_dict(synth.make_vocoder_state_dict(1234),strict=False)
print(json.dumps(bb.run(128, steps=20, warm=4, prompt_s=11.9, advance=200)))" > $O/attn_ab_old.json 2>&1
SVANON_ATTN_TMA=1 timeout 300 python -c "
import sys, json; sys.argv=['x']; sys.path.insert(0,'.')
import tools.bench_batch as bb, torch
from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder, synth
ar=ARVCWrapper(); ar.setup_caches(max_batch_size=1,max_seq_len=2048); ar.load_state_dict(synth.make_ar_state_dict(1234),strict=False)
ContentTokenizer().load_state_dict(synth.make_tokenizer_state_dict(1234),strict=False); Vocoder().load_state_dict(synth.make_vocoder_state_dict(1234),strict=False)
print(json.dumps(bb.run(128, steps=20, warm=4, prompt_s=11.9, advance=200)))" > $O/attn_ab_new.json 2>&1
SVANON_ATTN_TMA=0 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:arb_attn_slow -c 2 \
    -o $O/attn_old_s950 python tools/profile_batch.py 128 11.9 200 > $O/ncu_attn_old.log 2>&1
SVANON_ATTN_TMA=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:arb_attn_slow -c 2 \
    -o $O/attn_tma_s950 python tools/profile_batch.py 128 11.9 200 > $O/ncu_attn_new.log 2>&1
# stateful-encoder mode: step time at 128 / 256 streams
timeout 300 python -c "
import sys, json, time; sys.path.insert(0,'.')
import torch, bench
from streamvoiceanon_b200 import ARVCWrapper, ContentTokenizer, Vocoder, synth
ar=ARVCWrapper(); ar.setup_caches(max_batch_size=1,max_seq_len=2048); ar.load_state_dict(synth.make_ar_state_dict(1234),strict=False)
tok=ContentTokenizer(); tok.load_state_dict(synth.make_tokenizer_state_dict(1234),strict=False); Vocoder().load_state_dict(synth.make_vocoder_state_dict(1234),strict=False)
for B in (128, 256):
    r=bench.concurrent_leg(tok, B, 100, 0, enc_mode=3); print(json.dumps(r))" > $O/stateful_legs.json 2>&1
ls -la $O; tail -6 $O/pytest_new.txt; tail -4 $O/pytest_rest.txt; cat $O/launch_overhead.json $O/attn_ab_old.json $O/attn_ab_new.json; tail -3 $O/stateful_legs.json
