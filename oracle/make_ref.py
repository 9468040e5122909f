#!/usr/bin/env python
"""Recipe that stages the UNMODIFIED reference files of the scoring path into oracle/_ref/ (git-ignored, but shipped to
the GPU box with the working-tree snapshot exactly like the built .so).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python: there is nothing to compile, so "building" it is
staging the few source files the path lives in, byte for byte, where `oracle/ref_import.py` can import them with
MADRIGAL_REFERENCE_ROOT=oracle/_ref — `bench.py --impl reference` and `cpu_baseline` then time the reference's own
code (`kind: "reference"`), and only fall back to the numpy port (`kind: "port"`) when oracle/_ref is absent.
Nothing here is tracked by git; no reference source enters the repository's history.

    python oracle/make_ref.py [--reference-root /root/reference]

Files staged (paths relative to the reference root):
    madrigal/__init__.py, madrigal/utils.py, madrigal/models/__init__.py, madrigal/models/models.py
        BilinearDDIScorer / Symmetric / TransformerFusion / NovelDDIMultilabel  (models.py:352-455, 521-547, 914-953)
    notebooks/normalize_scores.py
        classwise_normalized_rank_3d_numpy / run_slice                          (normalize_scores.py:36-74)
The chemCPA sub-package the two modules import at top level is NOT staged (it pulls a large optional stack and is off
the timed path); ref_import stubs it, as it does for torch_geometric / torchdrug / torch_scatter.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "madrigal/__init__.py",
    "madrigal/utils.py",
    "madrigal/models/__init__.py",
    "madrigal/models/models.py",
    "notebooks/normalize_scores.py",
]


def stage(reference_root: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(os.path.join(reference_root, "madrigal")):
        raise FileNotFoundError(f"reference not found at {reference_root}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference_root, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    return manifest


def staged(dest: str = DEST) -> bool:
    return all(os.path.exists(os.path.join(dest, rel)) for rel in FILES)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference-root", default=os.environ.get("MADRIGAL_REFERENCE_ROOT", "/root/reference"))
    args = ap.parse_args()
    m = stage(args.reference_root)
    print(f"staged {len(m)} reference files into {DEST}")
    sys.exit(0)
