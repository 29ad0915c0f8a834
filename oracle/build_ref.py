"""TEST / BASELINE INFRASTRUCTURE ONLY — stages the reference's own hot-path files for the GPU box.

The reference (RockingMat/Prot2Text-V2-esm3) is pure Python, so "building" it is packing the few files that hold the
Stage-1 hot path, from where they lie under /root/reference, into ONE archive:

    oracle/_ref/ref_hotpath.tgz        (git-ignored; NOT gpurun-ignored, so it travels to the GPU box like a built .so)

        models/modality_config.py, models/configuration_esm2llama_instruct.py, models/modeling_esm2llama_instruct.py
        scripts/__init__.py, scripts/train_contrast.py, scripts/utils_argparse.py

No reference source enters the repository's history: the archive is produced by `__graft_entry__.build()` in the
authoring container (where /root/reference is mounted) and unpacked at run time by `oracle/reference_loader.py` into a
temporary directory, which then serves as the reference root when /root/reference itself is absent.  `bench.py`'s
`cpu_baseline` leg and `--impl reference` then time the reference's OWN ModalityAdapter / readout_embeddings /
SegmentedBatchInfoNCELoss (`kind: "reference"`), and tests/test_oracle.py's live-reference check runs on the GPU box
too.  Without the archive they fall back to the restatement (`kind: "port"`).

    python oracle/build_ref.py
"""
from __future__ import annotations

import io
import os
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "ref_hotpath.tgz")
SOURCE_ROOT = os.environ.get("P2T_REFERENCE_SOURCE", "/root/reference")
FILES = [
    "models/modality_config.py", "models/configuration_esm2llama_instruct.py", "models/modeling_esm2llama_instruct.py",
    "scripts/__init__.py", "scripts/train_contrast.py", "scripts/utils_argparse.py",
]


def build(force: bool = False) -> str | None:
    """Pack the hot-path files; returns the archive path, or None when the reference tree is not mounted."""
    if not os.path.isfile(os.path.join(SOURCE_ROOT, "scripts", "train_contrast.py")):
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    newest = max(os.path.getmtime(os.path.join(SOURCE_ROOT, f)) for f in FILES)
    if not force and os.path.isfile(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        return ARCHIVE
    os.makedirs(OUT_DIR, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for f in FILES:
            tar.add(os.path.join(SOURCE_ROOT, f), arcname=f)
    with open(ARCHIVE, "wb") as fh:
        fh.write(buf.getvalue())
    return ARCHIVE


if __name__ == "__main__":
    print(build(force=True))
