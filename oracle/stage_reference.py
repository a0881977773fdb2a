"""Stage the UNMODIFIED reference files of the hot path under ``baseline/_ref/`` (git-ignored, NOT gpurun-ignored) so
that the reference's own ``Detector.predict`` can be timed on the GPU box's host cores (``bench.py --impl reference``
and the ``cpu_baseline`` leg, ``kind: "reference"``), where ``/root/reference`` does not exist.

ORACLE-side infrastructure: files are copied byte for byte at build time (``__graft_entry__.build()`` calls this when
``/root/reference`` is present), never committed, never imported by the product package. ``python
oracle/stage_reference.py`` re-stages by hand.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("DFD_REFERENCE", "/root/reference")
STAGED = os.path.join(ROOT, "baseline", "_ref")
# the files `Detector.predict` needs: src/models.py:394-780 and the CLIP package it loads (src/clip/*)
FILES = ["src/models.py", "src/clip/__init__.py", "src/clip/clip.py", "src/clip/model.py",
         "src/clip/simple_tokenizer.py", "src/clip/bpe_simple_vocab_16e6.txt.gz"]


def staged_root():
    """Path of the staged reference tree, or None when it has not been staged."""
    return STAGED if all(os.path.isfile(os.path.join(STAGED, f)) for f in FILES) else None


def stage(verbose=True):
    if not os.path.isdir(REFERENCE):
        return staged_root()
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
    open(os.path.join(STAGED, "src", "__init__.py"), "a").close()  # `from src.models import Detector`
    if verbose:
        print("staged %d reference files under %s" % (len(FILES), os.path.relpath(STAGED, ROOT)))
    return STAGED


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
