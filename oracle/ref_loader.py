"""Import the staged, unmodified reference (`oracle/_ref/blvm`, written by oracle/make_ref.py) — test infrastructure.

Only `tests/` and `bench.py`'s reference legs call this; the product never does.  The staged copy is preferred over
`/root/reference` even in the build container so that what is exercised is exactly what travels to the GPU box.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "blvm", "__init__.py"))


def why_unavailable() -> str:
    return (f"{REF}/blvm is not staged: run `python oracle/make_ref.py` (or `__graft_entry__.build()`) in the build "
            "container, where /root/reference exists; the git-ignored copy then travels with the snapshot")


def load():
    """Put the staged reference (and the import shims for packages this image lacks) on sys.path and return the `blvm`
    package.  `blvm.settings` prompts for a data directory unless BLVM_DATA_ROOT_DIRECTORY is set (settings.py:33-37)."""
    if not available():
        raise ImportError(why_unavailable())
    os.environ.setdefault("BLVM_DATA_ROOT_DIRECTORY", "/tmp/blvmdata")
    os.makedirs(os.environ["BLVM_DATA_ROOT_DIRECTORY"], exist_ok=True)
    os.environ.setdefault("WANDB_MODE", "disabled")
    for p in (os.path.join(REF, "_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    blvm = importlib.import_module("blvm")
    importlib.import_module("blvm.models")
    return blvm
