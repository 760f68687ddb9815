"""Stage the UNMODIFIED reference modules of the path into ``oracle/_ref/`` (build container only).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so the files the reference's
own forward needs are copied -- byte for byte, never edited -- into ``oracle/_ref/src/...``, which is git-ignored (it never
enters the history) but travels with the ``gpurun`` snapshot like the built ``.so`` files.  ``oracle/run_reference.py``
imports them from there when ``/root/reference`` is absent, so that ``bench.py --impl reference`` and the ``cpu_baseline``
leg time the reference's own ``SequentialDualStreamModel.forward`` (src/model/sequential_dual_stream_model.py:63-167)
rather than a port of it.

    python -m oracle.stage_reference          # idempotent; prints the manifest

``oracle/_ref/MANIFEST.json`` records the sha256 of every staged file next to the sha256 of its source, so "unmodified"
is checkable.  The two third-party packages the reference imports but this image lacks (librosa, opensmile) are NOT
staged code: they are the stand-ins defined in ``oracle/run_reference.py`` (the librosa calls resolve to the restatement
in ``oracle/koemorph_oracle.py``).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = os.environ.get("KOEMORPH_REFERENCE", "/root/reference")
STAGE_ROOT = os.path.join(HERE, "_ref")

# what `import src.model.sequential_dual_stream_model` + one forward touch (probed with sys.modules)
FILES = [
    "src/__init__.py",
    "src/model/__init__.py",
    "src/model/dual_stream_attention.py",
    "src/model/simplified_dual_stream_model.py",
    "src/model/sequential_dual_stream_model.py",
    "src/features/__init__.py",
    "src/features/emotion_extractor.py",
    "src/features/opensmile_extractor.py",
    "src/features/mel_sliding_window.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available() -> bool:
    return all(os.path.isfile(os.path.join(SOURCE_ROOT, f)) for f in FILES)


def staged_available() -> bool:
    return all(os.path.isfile(os.path.join(STAGE_ROOT, f)) for f in FILES)


def stage(verbose: bool = False) -> dict:
    """Copy the reference files (when the source tree is present) and return the manifest."""
    if not source_available():
        if staged_available():
            with open(os.path.join(STAGE_ROOT, "MANIFEST.json")) as f:
                return json.load(f)
        raise RuntimeError(f"reference tree not found at {SOURCE_ROOT} and nothing staged under {STAGE_ROOT}")
    manifest = {"source_root": SOURCE_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(SOURCE_ROOT, rel), os.path.join(STAGE_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest["files"][rel] = {"sha256_source": _sha(src), "sha256_staged": _sha(dst)}
        assert manifest["files"][rel]["sha256_source"] == manifest["files"][rel]["sha256_staged"]
    with open(os.path.join(STAGE_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return manifest


if __name__ == "__main__":
    stage(verbose=True)
