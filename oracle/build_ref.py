#!/usr/bin/env python
"""Build the UNMODIFIED pyDVS Cython modules into ``oracle/_ref/`` (test infrastructure only).

This is oracle tooling, not product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import what it produces.

Recipe (our own; the reference's ``setup.py`` is not run): for each of the three ``.pyx`` files the
reference's ``setup.py:13-29`` lists, run the ``cython`` compiler on the source *where it lies* under
``/root/reference/pydvs`` with the generated C going to ``oracle/_ref/pydvs/``, then ``gcc -shared`` it
into a CPython extension next to it.  Nothing is copied from the reference; ``oracle/_ref/`` is
git-ignored (build output) but travels to the GPU box with the snapshot.

Importing the result (see ``oracle/ref_loader.py``) needs ``numpy.float = float`` set first for the
uncoded/decode modules (``generate_spikes_uncoded.pyx:32``, ``decode_spikes.pyx:34``).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PYDVS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "pydvs")
MODULES = ("generate_spikes", "generate_spikes_uncoded", "decode_spikes")
# the reference's only native compute, DVSOperator (pydvs/cpp/dvs_op.cpp:4-33), compiled UNCHANGED against the
# stand-in OpenCV headers of oracle/opencv_stub/ plus our C-ABI driver oracle/dvs_op_ref_driver.cpp
DVS_OP_LIB = os.path.join(HERE, "_ref", "libdvs_op_ref.so")


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def built_paths():
    return [os.path.join(OUT, m + ext_suffix()) for m in MODULES]


def is_built():
    return all(os.path.exists(p) for p in built_paths())


# the reference's CALLERS of the path, byte-compiled from the sources where they lie (bytecode is build output, like the
# .so files above): tests run them unchanged on top of the drop-in module to show that callers need no edits
CALLERS_DIR = os.path.join(HERE, "_ref", "callers")
CALLERS = {"stand_alone": "stand_alone.py", "virtual_cam": os.path.join("pydvs", "virtual_cam.py")}


def caller_path(name):
    return os.path.join(CALLERS_DIR, name + ".code")   # CPython bytecode (py_compile layout); not named .pyc so that
    #                                                      snapshot tools that skip *.pyc still carry it to the GPU box


def callers_are_built():
    return all(os.path.exists(caller_path(n)) for n in CALLERS)


def build_callers(force=False, verbose=True):
    """py_compile the reference's stand_alone.py and pydvs/virtual_cam.py into oracle/_ref/callers/*.code.  `dfile`
    makes the code objects report a path inside that directory, which is where virtual_cam.py looks for its
    fading_mask.png (virtual_cam.py:113-115; tests write that fixture there from their golden arrays)."""
    if callers_are_built() and not force:
        return True
    if not all(os.path.exists(os.path.join(REF_SRC, rel)) for rel in CALLERS.values()):
        return callers_are_built()
    import py_compile
    os.makedirs(CALLERS_DIR, exist_ok=True)
    for name, rel in CALLERS.items():
        py_compile.compile(os.path.join(REF_SRC, rel), cfile=caller_path(name), dfile=os.path.join(CALLERS_DIR, name + ".py"),
                           doraise=True)
        if verbose:
            print("[oracle/_ref] byte-compiled", caller_path(name))
    return True


def load_caller_code(name):
    """The code object of a byte-compiled reference caller (or None)."""
    import marshal
    path = caller_path(name)
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        data = f.read()
    return marshal.loads(data[16:])


def dvs_op_is_built():
    return os.path.exists(DVS_OP_LIB)


def build_dvs_op(force=False, verbose=True):
    """g++ on the reference's dvs_op.cpp where it lies (no OpenCV needed: see oracle/opencv_stub).  x86-64 gcc does
    not contract `relax * ref + diff` into an FMA; -ffp-contract=off states it."""
    if dvs_op_is_built() and not force:
        return True
    cpp_dir = os.path.join(REF_SRC, "pydvs", "cpp")
    if not os.path.exists(os.path.join(cpp_dir, "dvs_op.cpp")):
        return dvs_op_is_built()
    os.makedirs(os.path.dirname(DVS_OP_LIB), exist_ok=True)
    cc = ["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-w", "-I", os.path.join(HERE, "opencv_stub"),
          "-I", cpp_dir, os.path.join(HERE, "dvs_op_ref_driver.cpp"), os.path.join(cpp_dir, "dvs_op.cpp"), "-o", DVS_OP_LIB]
    r = subprocess.run(cc, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for dvs_op.cpp:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("[oracle/_ref] built", DVS_OP_LIB)
    return True


def build(force=False, verbose=True):
    """Returns True if oracle/_ref is available after the call."""
    try:
        build_dvs_op(force=force, verbose=verbose)
    except Exception as e:   # the Cython modules below do not depend on it
        if verbose:
            print("[oracle/_ref] DVSOperator not built: %s" % e)
    try:
        build_callers(force=force, verbose=verbose)
    except Exception as e:
        if verbose:
            print("[oracle/_ref] callers not byte-compiled: %s" % e)
    if is_built() and not force:
        return True
    src_dir = os.path.join(REF_SRC, "pydvs")
    if not os.path.isdir(src_dir):
        if verbose:
            print("[oracle/_ref] reference sources not present at %s; skipping" % src_dir)
        return is_built()
    import numpy

    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "__init__.py"), "w") as f:
        f.write("# build output of oracle/build_ref.py (compiled pyDVS reference modules)\n")
    py_inc = sysconfig.get_paths()["include"]
    np_inc = numpy.get_include()
    for mod in MODULES:
        pyx = os.path.join(src_dir, mod + ".pyx")
        c_out = os.path.join(OUT, mod + ".c")
        so_out = os.path.join(OUT, mod + ext_suffix())
        # language_level 3str is Cython 3's default; stated explicitly so the recipe is reproducible.
        cy = [sys.executable, "-m", "cython", "--3str", "--module-name", "pydvs." + mod,
              "-o", c_out, pyx]
        r = subprocess.run(cy, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("cython failed for %s:\n%s\n%s" % (mod, r.stdout, r.stderr))
        cc = ["gcc", "-shared", "-fPIC", "-O2", "-fwrapv", "-fno-strict-aliasing", "-w",
              "-DNPY_NO_DEPRECATED_API=0", "-I", py_inc, "-I", np_inc, c_out, "-o", so_out]
        r = subprocess.run(cc, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for %s:\n%s\n%s" % (mod, r.stdout, r.stderr))
        os.remove(c_out)  # generated C is derived from reference text: never keep it around
        if verbose:
            print("[oracle/_ref] built", so_out)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
