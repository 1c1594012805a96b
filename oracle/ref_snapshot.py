"""ORACLE / reference-arm infrastructure (never imported by open_muse_b200).

``snapshot()`` copies the reference's pure-Python ``muse/`` and ``training/`` trees from /root/reference into the
git-ignored ``oracle/_ref/`` (listed in .gitignore, NOT in .gpurunignore, so it travels to the GPU box like a built
``.so``).  It is the recipe VERDICT r1 asked for: the UNMODIFIED reference can then be timed as the CPU baseline
(bench.py, ``cpu_baseline.kind == "reference"``) and its training script can be executed against the drop-in package on
the GPU box (tests/test_train_script_gpu.py).  Nothing under oracle/_ref is tracked by git or imported by the product.

``import_reference()`` imports that copy as the package ``muse`` with the two-symbol ``accelerate`` stub of SURVEY 8c."""
import contextlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MUSE_REFERENCE", "/root/reference")


def snapshot():
    """Returns DEST if a snapshot exists (refreshing it when /root/reference is present), else None."""
    if os.path.isdir(os.path.join(SRC, "muse")):
        for sub in ("muse", "training"):
            dst = os.path.join(DEST, sub)
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        with open(os.path.join(DEST, "SOURCE"), "w") as f:
            f.write("unmodified copy of huggingface/open-muse muse/ and training/ taken from %s by oracle/ref_snapshot.py\n" % SRC)
    return DEST if os.path.isdir(os.path.join(DEST, "muse")) else None


def available():
    return os.path.isdir(os.path.join(DEST, "muse"))


def import_reference():
    """The unmodified reference package (from oracle/_ref) as module ``muse``."""
    if not available():
        raise ImportError("oracle/_ref/muse is absent: run __graft_entry__.build() where /root/reference exists")
    import transformers  # noqa: F401  (before the accelerate stub: its own accelerate probing breaks on a spec-less stub)

    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.init_empty_weights = contextlib.nullcontext
        accu = types.ModuleType("accelerate.utils")
        accu.set_module_tensor_to_device = lambda *a, **k: None
        acc.utils = accu
        acc.__spec__ = None
        sys.modules["accelerate"] = acc
        sys.modules["accelerate.utils"] = accu
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import muse

    assert os.path.realpath(muse.__file__).startswith(os.path.realpath(DEST)), muse.__file__
    return muse
