"""Recipe for ``oracle/_ref/``: the three reference modules of the hot path (GPnet.py, GPnetRegressor.py,
GPnetClassifier.py) compiled -- unmodified, from the sources where they lie under the read-only mount /root/reference --
to CPython bytecode (``py_compile``), the Python counterpart of compiling a C reference into ``oracle/_ref/*.so``.

    python -m oracle.make_ref

Run by ``__graft_entry__.build()`` in the build container.  No reference SOURCE is copied anywhere; ``oracle/_ref/`` holds
build outputs only (``*.bytecode`` = CPython bytecode under a neutral extension, because snapshot tools drop ``*.pyc``, + a manifest with the sha256 of each source), is listed in .gitignore (never enters this
repository) but not in .gpurunignore (travels to the GPU box like the built ``.so``; same image, same interpreter, so the
bytecode loads there).  ``bench.py --impl reference`` imports it through the shims of oracle/ref_shim.py and drives the
literal ``GPnetRegressor.logPosterior``.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import hashlib
import os
import py_compile
import sys

from oracle.ref_shim import REFERENCE_MODULES, REFERENCE_PATH, VENDORED_EXT, VENDORED_PATH


def make_ref(src=REFERENCE_PATH, dst=VENDORED_PATH):
    """Returns [(module, sha256 of its source)] compiled into ``dst``; empty when the mount is absent (GPU box: prebuilt)."""
    if not all(os.path.exists(os.path.join(src, f)) for f in REFERENCE_MODULES):
        return []
    os.makedirs(dst, exist_ok=True)
    out = []
    for f in REFERENCE_MODULES:
        path = os.path.join(src, f)
        py_compile.compile(path, cfile=os.path.join(dst, f[:-3] + VENDORED_EXT), dfile=f, doraise=True)      # sourceless import
        with open(path, "rb") as fh:
            out.append((f, hashlib.sha256(fh.read()).hexdigest()))
    with open(os.path.join(dst, "MANIFEST.txt"), "w") as fh:
        fh.write("CPython %d.%d bytecode of the unmodified reference modules (filippocastelli/GPnet), built by oracle/make_ref.py; "
                 "sha256 of each source:\n" % sys.version_info[:2])
        for f, h in out:
            fh.write(f"{h}  {f}\n")
    return out


if __name__ == "__main__":
    for f, h in make_ref():
        print(h, f)
