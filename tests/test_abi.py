"""CPU: the C-ABI library loads without a GPU and exports exactly what include/*.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "enspara_b200.h")


def _split_planned():
    raw = open(HEADER).read()
    if "#ifdef EB_PLANNED" not in raw:
        return raw, ""
    head, rest = raw.split("#ifdef EB_PLANNED", 1)
    block, tail = rest.split("#endif /* EB_PLANNED */", 1)
    return head + tail, block


def _names(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(eb_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    from enspara_b200 import _lib, build
    build.build()
    live, planned = _split_planned()
    declared = _names(live)
    planned_names = _names(planned)
    assert declared, "no declarations parsed from the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r"\bT (eb_[a-z0-9_]+)", out))
    assert declared <= exported, "declared but not exported: %s" % sorted(declared - exported)
    assert exported <= declared | planned_names, \
        "exported but not declared: %s" % sorted(exported - declared)
    # the ctypes table covers the header one to one
    assert set(_lib.SIGNATURES) == declared | planned_names
    assert set(_lib.PLANNED) == planned_names


def test_library_loads_and_layout_helpers_work_without_gpu():
    from enspara_b200 import _lib
    L = _lib.load()
    assert L.eb_version() >= 100
    assert L.eb_rmsd_apad(500) == 512 and L.eb_rmsd_apad(264) == 264 and L.eb_rmsd_apad(22) == 24
    assert L.eb_rmsd_apad(1000) == 1024 and L.eb_rmsd_apad(33) == 40
    assert L.eb_rmsd_record_bytes(500) == 32 + 12 * 512
    assert L.eb_feat_record_bytes(64, _lib.DT_F32) == 32 + 256
    assert L.eb_feat_record_bytes(3, _lib.DT_I8) == 32 + 16
    assert L.eb_kc_partials_bytes() >= 16 * 1024


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under enspara_b200/ may reference it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "enspara_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or \
                        "libenspara_oracle" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from enspara_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()


def test_struct_mirrors_match_the_compiled_layout():
    """ctypes mirrors of eb_kc_state / eb_pam_ctx have the size the library was compiled with
    (host-only call: no GPU needed)."""
    import ctypes
    from enspara_b200 import _lib
    lib = _lib.load()
    assert lib.eb_struct_bytes(0) == ctypes.sizeof(_lib.KcState) == 64
    assert lib.eb_struct_bytes(1) == ctypes.sizeof(_lib.PamCtx)
