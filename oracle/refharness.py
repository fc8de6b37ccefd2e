"""Import the REAL reference package from /root/reference.  TEST INFRASTRUCTURE, build container only.

/root/reference does not exist on the GPU box; nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call this module.  It is used
  * by tests/test_oracle_ref.py (CPU, skipped when the reference is absent) to check the numpy
    restatement in oracle/cluster.py and the C libdist restatement against the reference's own
    Python loops and its own Cython ``libdist``;
  * by scripts/make_golden.py to generate tests/golden/*.npz.

Recipe: the reference imports ``mdtraj``, ``mdtraj.io`` and ``tables`` at
module top although the clustering path only needs ``md.rmsd`` and a Trajectory duck type, so
those three modules are stubbed in ``sys.modules``; ``mdtraj.rmsd`` is the restated RMSD of
oracle/distances.py.  ``enspara.geometry.libdist`` is compiled from the .pyx WHERE IT LIES
(cython -> oracle/_ref/build/libdist.c -> oracle/_ref/enspara_geometry/libdist.*.so); no
reference source is copied into the repository.
"""
import importlib
import os
import subprocess
import sys
import sysconfig
import types

import numpy as np

REFERENCE = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(_HERE, "_ref")
_GEOM_DIR = os.path.join(REF_OUT, "enspara_geometry")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "enspara"))


def build_libdist(force=False):
    """cython + gcc on /root/reference/enspara/geometry/libdist.pyx -> oracle/_ref/."""
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(_GEOM_DIR, "libdist" + ext)
    if os.path.exists(so) and not force:
        return so
    if not available():
        raise RuntimeError("reference tree not present; cannot build oracle/_ref")
    build_dir = os.path.join(REF_OUT, "build")
    os.makedirs(build_dir, exist_ok=True)
    os.makedirs(_GEOM_DIR, exist_ok=True)
    c_file = os.path.join(build_dir, "libdist.c")
    pyx = os.path.join(REFERENCE, "enspara", "geometry", "libdist.pyx")
    subprocess.run([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file], check=True,
                   capture_output=True)
    inc = sysconfig.get_paths()["include"]
    # same flags as /root/reference/setup.py:10-20,31-37 (-fopenmp, numpy include dir)
    cmd = ["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-Wno-unused-function",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", "-I", inc, "-I", np.get_include(),
           c_file, "-o", so]
    subprocess.run(cmd, check=True, capture_output=True)
    return so


_loaded = None


def load():
    """Return the reference ``enspara`` package (cluster + mpi + libdist usable)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present")
    from . import distances as od

    so = build_libdist()
    # load_compiled_libdist() may have registered a stub package: the real one replaces it
    for name in [m for m in sys.modules if m == "enspara" or m.startswith("enspara.")]:
        if getattr(sys.modules[name], "_eb_stub", False) or \
                getattr(sys.modules.get("enspara"), "_eb_stub", False):
            del sys.modules[name]

    md = types.ModuleType("mdtraj")
    md.rmsd = od.rmsd
    md.Trajectory = od.Trajectory
    md_io = types.ModuleType("mdtraj.io")
    md.io = md_io
    for name in ("load", "load_frame", "open", "join"):
        def _missing(*a, _n=name, **k):
            raise RuntimeError("mdtraj.%s is not available in the oracle harness" % _n)
        setattr(md, name, _missing)
    sys.modules.setdefault("mdtraj", md)
    sys.modules.setdefault("mdtraj.io", md_io)
    sys.modules.setdefault("tables", types.ModuleType("tables"))

    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import enspara  # noqa: F401  (the reference package)
        import enspara.geometry as geom
        if _GEOM_DIR not in geom.__path__:
            geom.__path__.append(_GEOM_DIR)
        importlib.import_module("enspara.geometry.libdist")
        importlib.import_module("enspara.cluster")
    assert os.path.samefile(sys.modules["enspara.geometry.libdist"].__file__, so)
    _loaded = sys.modules["enspara"]
    return _loaded


def modules():
    """(kcenters, kmedoids, hybrid, util, libdist, mpi) of the reference."""
    load()
    m = sys.modules
    return (m["enspara.cluster.kcenters"], m["enspara.cluster.kmedoids"],
            m["enspara.cluster.hybrid"], m["enspara.cluster.util"],
            m["enspara.geometry.libdist"], m["enspara.mpi"])


def load_compiled_libdist():
    """The reference's OWN compiled Cython ``libdist`` (oracle/_ref/, built by build_libdist()
    from /root/reference/enspara/geometry/libdist.pyx) WITHOUT the reference's Python package:
    the only thing the extension imports from it is ``enspara.exception`` (libdist.pyx:4), which
    is stubbed here.  This is what travels to the GPU box, where /root/reference does not
    exist: bench CPU baselines time the reference's real euclidean / manhattan code there."""
    import glob
    import importlib.util
    name = "enspara.geometry.libdist"
    if name in sys.modules and hasattr(sys.modules[name], "euclidean"):
        return sys.modules[name]
    if available():
        # build container: the real package is importable, no stub needed
        load()
        return sys.modules[name]
    hits = glob.glob(os.path.join(_GEOM_DIR, "libdist*.so"))
    if not hits:
        if not available():
            raise RuntimeError("oracle/_ref/enspara_geometry/libdist*.so is missing and the "
                               "reference tree is not present to build it")
        hits = [build_libdist()]
    if "enspara" not in sys.modules:
        pkg = types.ModuleType("enspara")
        pkg.__path__ = []
        pkg._eb_stub = True
        exc = types.ModuleType("enspara.exception")

        class ImproperlyConfigured(Exception):
            pass

        class DataInvalid(Exception):
            pass
        exc.ImproperlyConfigured, exc.DataInvalid = ImproperlyConfigured, DataInvalid
        pkg.exception = exc
        sys.modules["enspara"] = pkg
        sys.modules["enspara.exception"] = exc
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod
