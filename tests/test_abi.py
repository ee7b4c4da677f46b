"""CPU: the C-ABI library loads and exports every symbol include/hrb200.h declares (no compute)."""
import ctypes
import os
import subprocess

import pytest

import __graft_entry__ as entry
from handyrec_b200 import _lib


@pytest.fixture(scope="module")
def built():
    entry.build()
    return _lib.LIB_PATH


def test_header_is_parsed():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for must in ("hrb_lookup_fwd", "hrb_lookup_fm_fwd", "hrb_lookup_bwd_update", "hrb_fm_fwd", "hrb_dense_fwd", "hrb_lau_fwd"):
        assert must in protos


def test_library_exports_every_declared_symbol(built):
    assert os.path.exists(built)
    L = ctypes.CDLL(built)
    missing = [n for n in _lib.parse_header() if not hasattr(L, n)]
    assert not missing, missing


def test_only_c_abi_symbols_are_exported(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("hrb_") for s in exported), [s for s in exported if not s.startswith("hrb_")]


def test_status_strings_and_bad_args_without_gpu(built):
    L = _lib.lib()
    assert L.hrb_abi_version() == 1
    assert L.hrb_status_str(0) == b"HRB_OK" and L.hrb_status_str(2) == b"HRB_UNSUPPORTED"
    # argument validation happens before any CUDA call, so it is testable on a CPU-only box
    rc = L.hrb_embedding_fwd(None, 10, 8, None, 4, None, None, None, None)
    assert rc == _lib.HRB_BAD_ARG and b"hrb_embedding_fwd" in L.hrb_last_error()
    with pytest.raises(_lib.HrbError):
        _lib.call("hrb_seq_pool_fwd", None, None, 1, 1, 1, 1, None, None)


def test_sm100a_cubin_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if "sm_" in l}
    assert archs == {"sm_100a"}, out
