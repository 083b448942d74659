"""The C-ABI library loads and exports every symbol include/birda_b200.h declares (CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "birda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import birda_b200
    syms = declared_symbols()
    assert len(syms) >= 30
    dll = ctypes.CDLL(birda_b200.lib_path)
    missing = [s for s in syms if not hasattr(dll, s)]
    assert not missing, f"not exported: {missing}"


def test_binding_table_matches_header():
    from birda_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_device_fails_loudly():
    """No CPU fallback: without a GPU the context refuses to exist (and says why)."""
    import birda_b200 as b
    from birda_b200.api import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(b.BirdaError) as e:
        b.Context(0)
    assert e.value.code == -8 and "no CPU fallback" in e.value.message


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under birda_b200/ may reference it."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "birda_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src.replace("NOT the oracle", ""):
                    bad.append(fn)
    assert not bad, bad


def test_pool_without_device_fails_loudly():
    """The multi-GPU file pool refuses to exist without a GPU, like the context it is built on."""
    import ctypes as C

    import birda_b200 as b
    from birda_b200 import _lib
    from birda_b200.api import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    cb = _lib.CLASSIFY_FN(lambda *a: 1)
    cfg = (_lib.PipelineCfg * 1)(_lib.PipelineCfg(48_000, 3.0, 0.0, 8, 0, _lib.PostCfg(1, 0.1, 5, 0.01, 1, 0), None, None))
    h = C.c_void_p()
    rc = _lib.lib.bb_pool_create((C.c_int32 * 1)(0), 1, cfg, cb, None, C.byref(h))
    assert rc == -8 and not h.value
    assert b"no CPU fallback" in _lib.lib.bb_last_error(None)
