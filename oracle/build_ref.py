"""Builds the UNMODIFIED reference (softwaredoug/searcharray) into oracle/_ref/ -- test infrastructure.

    python oracle/build_ref.py            # idempotent; no-op when /root/reference is absent

Recipe: the reference is Python + Cython.  /root/reference is read-only and the
cythonize step writes generated C next to the .pyx files, so the tree is copied to a scratch
directory under /tmp first, then installed with the image's own pip / setuptools / Cython /
numpy (no index access):

    cp -r /root/reference /tmp/sa_ref_build_<pid>
    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref /tmp/sa_ref_build_<pid>

Outputs go ONLY into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like
the repo's own built .so files; nothing of the reference enters the history).  It is used
  * by tests/ to validate the restatement in oracle/ against the real thing, and
  * by `bench.py --impl reference` / `cpu_baseline` as the CPU arm (`kind: "reference"`): the
    reference's own `SearchArray.score` on the host cores.
The product path (searcharray_b200/) never imports it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SA_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
STAMP = os.path.join(REF_DST, ".built_from")


def have_ref():
    """True when oracle/_ref holds an importable build of the reference."""
    return os.path.exists(os.path.join(REF_DST, "searcharray", "__init__.py")) and os.path.exists(STAMP)


def build(force=False, verbose=False):
    if have_ref() and not force:
        return REF_DST
    if not os.path.isdir(os.path.join(REF_SRC, "searcharray")):
        return None                      # GPU box: only the prebuilt files exist
    scratch = f"/tmp/sa_ref_build_{os.getpid()}"
    shutil.rmtree(scratch, ignore_errors=True)
    shutil.copytree(REF_SRC, scratch, ignore=shutil.ignore_patterns("fixtures", "test", ".git", "scripts"))
    shutil.rmtree(REF_DST, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--target", REF_DST, scratch]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or p.returncode:
        sys.stderr.write(p.stdout)
    shutil.rmtree(scratch, ignore_errors=True)
    if p.returncode:
        raise RuntimeError("building the reference into oracle/_ref failed")
    with open(STAMP, "w") as f:
        f.write(f"pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of {REF_SRC}>\n")
    return REF_DST


def import_reference():
    """Imports the reference package from oracle/_ref (raises ImportError if it was never built)."""
    if not have_ref():
        raise ImportError("oracle/_ref is not built (run `python oracle/build_ref.py` where /root/reference exists)")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import searcharray                     # noqa: F401  (the reference, not this repo's package)
    return searcharray


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
