"""TEST INFRASTRUCTURE ONLY.  Byte-compile the UNMODIFIED reference model code into ``oracle/_ref/``.

    python oracle/build_ref.py          # needs /root/reference (the build container)

The reference is pure Python, so "building" it means ``py_compile``: every module the hot path
imports (``model/**``, ``utils.py``, ``configs/**``) is compiled from the sources where they lie
under ``/root/reference`` into a sourceless tree of compiled code objects (``*.refbin``: the bytes
``py_compile`` produces; not named ``.pyc`` because snapshot tools drop those like ``__pycache__``).
No reference source is copied; the output directory is git-ignored (it travels to the GPU box with
``gpurun`` like the built ``.so``) and is importable, through the finder in ``oracle/ref_shim.py``, by
the same interpreter version that built it.  It lets the GPU box run

  * ``bench.py --impl reference`` on the reference's OWN modules (``cpu_baseline.kind = "reference"``),
  * ``tests/test_dropin_reference_gpu.py``: the reference's ``model/DrugLAMP*.py:forward`` on top of
    ``druglamp_b200.patch_reference()``.

``oracle/ref_shim.py`` falls back to this tree when ``/root/reference`` is absent.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.environ.get("DRUGLAMP_REFERENCE_ROOT", "/root/reference")
TREES = ("model", "configs")
FILES = ("utils.py",)


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "model")):
        return OUT if os.path.isdir(os.path.join(OUT, "model")) else None
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    todo = [f for f in FILES]
    for tree in TREES:
        for d, _, files in os.walk(os.path.join(SRC, tree)):
            for f in files:
                if f.endswith(".py"):
                    todo.append(os.path.relpath(os.path.join(d, f), SRC))
    for rel in sorted(todo):
        dst = os.path.join(OUT, rel[:-3] + ".refbin")      # foo.py -> foo.refbin beside where foo.py would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(SRC, rel), cfile=dst, dfile=rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", rel)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
