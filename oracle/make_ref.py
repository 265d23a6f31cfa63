#!/usr/bin/env python
"""Stage the UNMODIFIED reference (`/root/reference/blvm`, pure Python/PyTorch) under git-ignored `oracle/_ref/` so that
it travels to the GPU box with the snapshot (it is not in the history: `.gitignore` lists `oracle/_ref/`).

Test infrastructure only.  What may import `oracle/_ref`: `tests/` (the `-m gpu` model tests run every reference audio
model unpatched and under `patch_blvm()` on the B200), and `bench.py --impl reference` / the `reference_eager_cuda`
sub-record (the reference's own CPU / eager-CUDA path as the timed baseline).  The product (`benchmarking-lvms_b200/`)
never imports it.

    python oracle/make_ref.py            # copy (idempotent; skipped when /root/reference is absent, e.g. on the GPU box)

Layout written:
    oracle/_ref/blvm/...                 the reference package, byte-identical .py files (sha256 manifest alongside)
    oracle/_ref/_shims/...               import-only stand-ins for packages absent from this image (torchtyping, ...),
                                         copied from tests/golden/_ref_shims
    oracle/_ref/MANIFEST.json            {relative path: sha256} of every copied reference file + the source commit note
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("BLVM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "_ref_shims")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True) -> bool:
    """Returns True if oracle/_ref is present and complete afterwards."""
    src_pkg = os.path.join(SRC, "blvm")
    if not os.path.isdir(src_pkg):
        ok = os.path.isfile(os.path.join(DST, "MANIFEST.json"))
        if verbose:
            print(f"[make_ref] {src_pkg} not present; " + ("keeping the staged copy" if ok else "nothing staged"))
        return ok
    manifest = {}
    for dirpath, dirnames, filenames in os.walk(src_pkg):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        for fn in filenames:
            if not fn.endswith((".py", ".txt", ".json", ".yaml", ".yml")):
                continue
            s = os.path.join(dirpath, fn)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            digest = _sha(s)
            if not (os.path.exists(d) and _sha(d) == digest):
                shutil.copyfile(s, d)
            manifest[rel] = digest
    os.makedirs(os.path.join(DST, "_shims"), exist_ok=True)
    for fn in sorted(os.listdir(SHIMS)):
        if fn.endswith(".py"):
            shutil.copyfile(os.path.join(SHIMS, fn), os.path.join(DST, "_shims", fn))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=0, sort_keys=True)
    if verbose:
        print(f"[make_ref] staged {len(manifest)} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
