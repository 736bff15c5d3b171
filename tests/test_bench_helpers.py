"""bench.py's derived figures (no GPU): the per-stage compute roofline arithmetic and that the script imports cleanly."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("svanon_bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_stage_compute_arithmetic():
    b = _bench()
    r = b.stage_compute(2.0, 1.0, 1, {"bf16_tflops_sustained": 1200.0})
    assert abs(r["ceiling_fp32_equivalent_tflops"] - 200.0) < 1e-9               # bf16 / 2 (TF32) / 3 (split)
    assert abs(r["E"]["achieved_tflops"] - b.E_GFLOP_SINGLE / 2.0) < 1e-9 and abs(r["V"]["achieved_tflops"] - 2.647) < 1e-9
    assert abs(r["E"]["frac"] - r["E"]["achieved_tflops"] / 200.0) < 1e-12
    many = b.stage_compute(20.0, 5.0, 128, {})
    assert abs(many["E"]["executed_gflop"] - 128 * b.E_GFLOP_MANY) < 1e-6         # conv-history encoder from 8 streams
    assert abs(many["ceiling_fp32_equivalent_tflops"] - 375.0) < 1e-9             # nominal fallback 2250 / 6
    assert b.stage_compute(0.0, 0.0, 1, {})["E"]["achieved_tflops"] == 0.0
    # executed work is below what the reference computes for the same chunk (29.0 GFLOP window re-encode)
    assert b.E_GFLOP_MANY < b.E_GFLOP_SINGLE < 29.0


def test_default_sweep_and_workload_are_the_baseline_config():
    b = _bench()
    assert b.WORKLOAD["decode_chunk_frames"] == 1 and b.WORKLOAD["delay"] == 2
    assert b.WORKLOAD["encode_window_frames"] == 128 and b.WORKLOAD["decode_window_frames"] == 64
