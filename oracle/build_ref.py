#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY -- builds ``oracle/_ref/``: the reference's own mapping modules, byte-compiled.

The reference is pure Python, so "compiling it from the sources where they lie" means ``py_compile``: this recipe
imports the UNMODIFIED reference through ``oracle/ref_shim.py`` (which stubs the absent third-party packages), notes
every module that import pulled in from ``/root/reference``, and byte-compiles each of those files into a sourceless
``<same relative path>.pyc`` inside ONE archive, ``oracle/_ref/reference_bytecode.zip`` (loose ``*.pyc`` files do not
survive the snapshot that ``gpurun`` sends to the GPU box).  No reference source text is copied; the output is a
binary, ``oracle/_ref/`` is git-ignored (it stays out of history) but not gpurun-ignored, so it travels to the GPU box
like the built ``.so`` files.  There ``ref_shim`` imports the same modules from the archive (zipimport; same
interpreter: the box runs this image), which is what lets ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference ITSELF
(``cpu_baseline.kind = "reference"``) on the box's host cores.

    python -m oracle.build_ref          # in the build container (needs /root/reference); __graft_entry__.build() calls it
"""
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "oracle", "_ref")
ARCHIVE = "reference_bytecode.zip"
SRC = "/root/reference"


def build(force=False):
    """Returns the list of compiled modules, or None when the reference tree is absent (GPU box: uses what was built)."""
    if not os.path.isfile(os.path.join(SRC, "src", "mapping_replay.py")):
        return None
    stamp = os.path.join(OUT, "BUILT_FROM")
    if not force and os.path.exists(stamp):
        return open(stamp).read().split("\n")[1:]
    os.environ["SMAP_REFERENCE_DIR"] = SRC
    from oracle import ref_shim
    ref_shim.REF = SRC
    ref_shim._loaded = None
    ref_shim.load_reference()
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if isinstance(m.__dict__.get("__file__"), str) and m.__file__.endswith(".py")
                    and os.path.realpath(m.__file__).startswith(SRC + os.sep)})
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    done = []
    tmp = tempfile.mkdtemp()
    with zipfile.ZipFile(os.path.join(OUT, ARCHIVE), "w", zipfile.ZIP_STORED) as z:
        for path in files:
            rel = os.path.relpath(path, SRC)
            dst = os.path.join(tmp, rel + "c")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            # dfile: the path tracebacks show; it names the reference file, it does not have to exist on the box
            py_compile.compile(path, cfile=dst, dfile=os.path.join("reference", rel), doraise=True)
            z.write(dst, rel + "c")
            done.append(rel)
    shutil.rmtree(tmp)
    with open(stamp, "w") as f:
        f.write("byte-compiled by oracle/build_ref.py from %s (python %s)\n" % (SRC, sys.version.split()[0]))
        f.write("\n".join(done))
    return done


if __name__ == "__main__":
    mods = build(force=True)
    print("reference tree absent: nothing built" if mods is None else "oracle/_ref: %d modules\n  " % len(mods) + "\n  ".join(mods))
