#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference package, copied from the read-only checkout.

    python oracle/make_ref.py          # copies /root/reference/nn_fac -> oracle/_ref/nn_fac

The reference (ax-le/nn-fac v0.3.4) is pure Python, so "building" it is a copy of its package directory.  oracle/_ref/ is
git-ignored (no reference source ever enters the history) but not gpurun-ignored, so the copy travels to the GPU box, where
/root/reference does not exist.  Its one missing dependency, tensorly == 0.6.0 (setup.py:30, not installable here), is the
numpy stand-in in oracle/ref_shim/tensorly.  bench.py --impl reference times this copy (oracle/ref_timing.py); nothing in
nn-fac_b200/ imports it.  Called by __graft_entry__.build() when /root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NNFAC_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(verbose=True):
    src = os.path.join(SRC, "nn_fac")
    if not os.path.isdir(src):
        if verbose:
            print(f"make_ref: {src} not found; keeping {'the existing' if os.path.isdir(DST) else 'no'} oracle/_ref")
        return os.path.isdir(os.path.join(DST, "nn_fac"))
    dst = os.path.join(DST, "nn_fac")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        n = sum(len(files) for _, _, files in os.walk(dst))
        print(f"make_ref: copied {n} files {src} -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
