"""Builds `oracle/_ref/`: the UNMODIFIED reference, byte-compiled from the sources where they lie.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref            # build container only (needs /root/reference)

The reference is a Python program, so "compiling it from its own source files" means `py_compile`: every module of the
hot path and of its caller (`modules/**/*.py`, `evaluations/infer_arvc.py`) becomes a sourceless byte-code file under
`oracle/_ref/` (same directory layout; suffix `.pycode` -- plain `.pyc` files are stripped from the snapshot that travels to
the GPU box -- which `oracle/ref_harness.py` makes importable with a path hook restricted to that directory, so
`import modules.dual_ar_stream` / `import evaluations.infer_arvc` resolve), and
the YAML configs the reference's constructor reads (`configs/`, data files) are placed beside them.  No reference source
enters the repository: `oracle/_ref/` is git-ignored (it travels to the GPU box with the snapshot, like the built
`libsvanon_b200.so`), and nothing under `streamvoiceanon_b200/` imports it (tests/test_cabi.py checks).

Used by
  * `bench.py --impl reference` and the `cpu_baseline` leg: the reference's own `process_one_chunk` timed on the host
    cores (`cpu_baseline.kind == "reference"`) instead of the oracle port;
  * `tests/test_gpu_zz_unmodified_caller.py`: the reference's own `InferenceWrapper` (constructor, `infer`,
    `stream_infer`) driven over the engine's shims on the GPU box, where /root/reference does not exist;
  * `oracle/ref_harness.py` falls back to it when /root/reference is absent.
The byte code is tied to this image's CPython (3.12; same image on the GPU box): `is_current()` checks the magic number.
"""
from __future__ import annotations

import importlib.util
import py_compile
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
PY_TREES = ("modules",)
PY_FILES = ("evaluations/infer_arvc.py",)
DATA_TREES = ("configs",)
STAMP = "BUILD_INFO"
SUFFIX = ".pycode"


def is_current() -> bool:
    """True when oracle/_ref holds byte code this interpreter can import."""
    probe = OUT / "modules" / ("arvc_wrapper" + SUFFIX)
    if not probe.exists():
        return False
    with open(probe, "rb") as f:
        return f.read(4) == importlib.util.MAGIC_NUMBER


def build(force: bool = False) -> Path:
    if not (SRC / "modules" / "arvc_wrapper.py").exists():
        raise RuntimeError(f"{SRC} is not present: oracle/_ref can only be built in the build container")
    if is_current() and not force:
        return OUT
    if OUT.exists():
        shutil.rmtree(OUT)
    files = [SRC / f for f in PY_FILES]
    for tree in PY_TREES:
        files += sorted((SRC / tree).rglob("*.py"))
    for src in files:
        rel = src.relative_to(SRC)
        dst = (OUT / rel).with_suffix(SUFFIX)
        dst.parent.mkdir(parents=True, exist_ok=True)
        # dfile: the path tracebacks show (the reference's own file name; the source itself is not shipped)
        py_compile.compile(str(src), cfile=str(dst), dfile=str(rel), doraise=True, optimize=0)
    for tree in DATA_TREES:
        for src in sorted((SRC / tree).rglob("*")):
            if src.is_file() and src.suffix in (".yaml", ".yml", ".json"):
                dst = OUT / src.relative_to(SRC)
                dst.parent.mkdir(parents=True, exist_ok=True)
                shutil.copyfile(src, dst)
    (OUT / STAMP).write_text(f"byte-compiled from {SRC} by oracle/build_ref.py with CPython {sys.version.split()[0]}; "
                             f"{len(files)} modules\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
