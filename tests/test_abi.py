"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares.
No compute call is made (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "acvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acvae_[a-z_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as g
    g.build()
    from acvae_b200 import _lib
    l = _lib.lib()
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    assert set(declared) == set(_lib.SYMBOLS), (set(declared) ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(l, name), name
    assert l.acvae_abi_version() == _lib.ABI_VERSION


def test_workspace_queries_and_errors():
    from acvae_b200 import _lib, functional as F
    import ctypes as C
    l = _lib.lib()
    d = F.make_dims(32, 62, 19, 256, 256, 4400, 512, 20)
    n = l.acvae_train_workspace_bytes(C.byref(d))
    assert 10e6 < n < 2e9
    bad = F.make_dims(32, 62, 19, 255, 256, 4400, 512, 20)     # E not a multiple of 4
    assert l.acvae_train_workspace_bytes(C.byref(bad)) == 0
    rc = l.acvae_train_fwd(C.byref(bad), None, None, None, 0, None)
    assert rc != 0 and b"multiples of 4" in l.acvae_last_error()
    assert l.acvae_vocab_workspace_bytes(608, 4400, 256) > 608 * 4400 * 4


def test_sass_is_sm100():
    """The shipped library carries sm_100a code only."""
    import subprocess
    from acvae_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
