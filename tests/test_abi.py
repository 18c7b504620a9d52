"""The C-ABI library loads and exports every symbol include/csn_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "csn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:int|const char\*|unsigned long long)\s+(csn_\w+)\s*\(", src, flags=re.M)
    assert len(names) >= 20
    return names


def test_library_exports_every_declared_symbol():
    from cerebralsignalnetworks_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m cerebralsignalnetworks_b200.build` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "library does not export %s" % name


def test_binding_table_matches_header():
    from cerebralsignalnetworks_b200 import _lib
    declared = set(declared_functions())
    bound = set(_lib.SIGNATURES) | set(_lib.EXTRA_SYMBOLS)
    assert declared == bound, (declared - bound, bound - declared)
    # argument counts agree with the header
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        n_args = 0 if m.group(1).strip() in ("", "void") else m.group(1).count(",") + 1
        assert n_args == len(argtypes), "%s: header has %d args, binding %d" % (name, n_args, len(argtypes))


def test_debug_hooks_stay_out_of_the_product_header():
    """The bring-up hooks are declared in include/csn_b200_debug.h only (and are exported for the tests that use them)."""
    from cerebralsignalnetworks_b200 import _lib
    assert not [n for n in declared_functions() if n.startswith("csn_dbg_")]
    dbg = open(os.path.join(ROOT, "include", "csn_b200_debug.h")).read()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.DEBUG_SIGNATURES:
        assert name in dbg and hasattr(lib, name)


def test_no_float_atomics_in_the_library():
    """Every cross-CTA sum is a fixed-order fold: the CUDA sources hold no float atomicAdd / red.global.add.f32."""
    csrc = os.path.join(ROOT, "cerebralsignalnetworks_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh")):
            continue
        code = re.sub(r"//[^\n]*", "", open(os.path.join(csrc, f)).read())
        assert "red.global.add" not in code, f
        for m in re.finditer(r"atomicAdd\(([^;]*);", code):
            assert "ticket" in m.group(1), "%s: atomicAdd on something other than an arrival ticket: %s" % (f, m.group(0))


def test_version_and_error_string():
    from cerebralsignalnetworks_b200 import _lib
    lib = _lib.load()
    assert lib.csn_version() == 200
    assert isinstance(lib.csn_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Pure host-side validation paths return an error code + message and never touch the device."""
    from cerebralsignalnetworks_b200 import _lib
    lib = _lib.load()
    rc = lib.csn_sosfilt_f32(None, None, None, 4, 1, 1, 1, 0, 0, 0, None)
    assert rc == -1 and b"null" in lib.csn_last_error()
    with pytest.raises(_lib.CsnError):
        _lib.call("csn_adam_step", None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 1, 1.0, None)


def test_sass_contains_blackwell_tensor_path():
    """cuobjdump of the built library shows tcgen05 (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG)."""
    import shutil
    import subprocess
    from cerebralsignalnetworks_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
