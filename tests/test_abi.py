"""The C-ABI shared library loads and exports every symbol include/pmnet_b200.h declares (no GPU compute)."""

import ctypes
import os
import re

from pharmaconet_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pmnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmnet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_all_declared_symbols():
    _lib.build()
    L = ctypes.CDLL(_lib.SO_PATH)
    names = _declared_functions()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n


def test_abi_version_and_host_only_calls():
    L = _lib.lib()
    assert L.pmnet_abi_version() == _abi.ABI_VERSION
    cfg = _abi.PmScoreConfig(8, 296, 8192, 0)
    a = L.pmnet_score_workspace_bytes(35, 26, ctypes.byref(cfg))
    cfg2 = _abi.PmScoreConfig(8, 296, 16384, 0)
    b = L.pmnet_score_workspace_bytes(35, 26, ctypes.byref(cfg2))
    assert 0 < a < b
    assert L.pmnet_topk_workspace_bytes(0, 10) > 0


def test_struct_layouts_match_header():
    # 4 x int32 + 8 pointers; int32 + pad + 8 pointers + int64 + 4 x int32; 8 x int32 (ABI 3: rescore_status + 3 reserved)
    assert ctypes.sizeof(_abi.PmModel) == 16 + 8 * 8
    assert ctypes.sizeof(_abi.PmLigandBatch) == 8 + 8 * 8 + 24 + 8  # + the optional `order` pointer (ABI 2)
    assert ctypes.sizeof(_abi.PmScoreConfig) == 32


def test_null_arguments_are_rejected_without_a_gpu():
    L = _lib.lib()
    rc = L.pmnet_score_batch(None, None, None, None, None, None, None, None, 0, None, None)
    assert rc == _abi.PMNET_EINVAL
    assert b"null" in L.pmnet_last_error_string()
