"""The C-ABI library loads on a CPU-only box and exports every symbol include/bump.h declares.
No compute calls here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bump.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bump_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from bumpcosmology_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from bumpcosmology_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bump.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"


def test_header_constants_match_binding():
    from bumpcosmology_b200 import _lib
    src = open(os.path.join(ROOT, "include", "bump.h")).read()
    macros = dict(re.findall(r"#define\s+(BUMP_[A-Z0-9_]+)\s+(\d+)u?", src))
    assert int(macros["BUMP_NTHETA"]) == _lib.NTHETA == 14
    assert int(macros["BUMP_OUT_HEADER"]) == _lib.OUT_HEADER
    assert int(macros["BUMP_OUT_DLOGLIKE"]) == _lib.OUT_DLOGLIKE
    assert int(macros["BUMP_OUT_DLOG_MU"]) == _lib.OUT_DLOG_MU
    assert int(macros["BUMP_PARTIAL_LEN"]) == _lib.PARTIAL_LEN
    assert int(macros["BUMP_OUT_NOBS"]) == _lib.OUT_NOBS


def test_no_cpu_fallback(lib):
    """Without a GPU the product must fail loudly (BUMP_E_NOGPU), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.bump_device_count() == 0
    ctx = C.c_void_p()
    rc = lib.bump_ctx_create(C.byref(ctx), 0, 0)
    assert rc == 3 and not ctx
    assert b"no CPU fallback" in lib.bump_last_error()
    from bumpcosmology_b200._lib import BumpError
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    one = np.ones((1, 2))
    with pytest.raises(BumpError):
        Hyperlikelihood(one, one, one, one, one[0], one[0], one[0], one[0], 10.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bumpcosmology_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_jax_ffi_binding_is_gated_on_jax():
    """The XLA typed-FFI adaptor is source-only where JAX is absent: importing the binding must fail loudly (never
    fall back to a host path), and the adaptor must stay a valid translation unit without the XLA headers."""
    import importlib.util
    import subprocess
    if importlib.util.find_spec("jax") is not None:
        pytest.skip("jax present: the binding is importable")
    with pytest.raises(ImportError, match="needs jax"):
        importlib.import_module("bumpcosmology_b200.jax_ffi")
    src = os.path.join(ROOT, "bumpcosmology_b200", "csrc", "bump_xla_ffi.cc")
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", src], check=True)
    txt = open(src).read()
    assert "XLA_FFI_DEFINE_HANDLER_SYMBOL(BumpLoglikeFfi" in txt and "bump_eval_device" in txt
