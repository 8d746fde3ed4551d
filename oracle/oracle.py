"""TEST INFRASTRUCTURE - ctypes wrapper of the CPU oracle (oracle/pmnet_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
The product package (pharmaconet_b200/) never does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pharmaconet_b200 import _abi
from pharmaconet_b200.constants import weights_vector

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpmnet_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pmnet_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "pmnet_b200.h")
    stale = not os.path.exists(_SO) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.pmnet_oracle_score.restype = C.c_int
        _lib.pmnet_oracle_score.argtypes = [
            C.POINTER(_abi.PmModel), C.POINTER(_abi.PmLigandBatch), C.c_void_p, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
        ]  # fmt: skip
        _lib.pmnet_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().pmnet_oracle_max_threads())


def score(model, batch, weights=None, threads: int = 0, begin: int = 0, end: int | None = None, with_conf=False):
    """model: PackedModel, batch: LigandBatch (host numpy). Returns dict(scores f64[n], status i32[n],
    stats u64[n,4] = tree nodes, leaves, entries, pair entries[, conf f64[n,128]])."""
    L = lib()
    end = batch.num_ligands if end is None else end
    n = end - begin
    marr = {k: np.ascontiguousarray(v) for k, v in model.arrays().items()}
    barr = {k: np.ascontiguousarray(v) for k, v in batch.arrays().items()}
    ms = _abi.model_struct(
        model.num_nodes, model.num_clusters, int(model.cluster_node_off[-1]), {k: v.ctypes.data for k, v in marr.items()}
    )
    bs = _abi.batch_struct(batch.num_ligands, {k: v.ctypes.data for k, v in barr.items()})
    w = np.asarray(weights_vector(weights) if not isinstance(weights, np.ndarray) else weights, dtype=np.float32)
    scores = np.zeros(n, dtype=np.float64)
    status = np.zeros(n, dtype=np.int32)
    stats = np.zeros((n, 4), dtype=np.uint64)
    conf = np.zeros((n, 128), dtype=np.float64) if with_conf else None
    rc = L.pmnet_oracle_score(
        C.byref(ms), C.byref(bs), w.ctypes.data, begin, end, scores.ctypes.data,
        conf.ctypes.data if with_conf else None, 128, status.ctypes.data, stats.ctypes.data, int(threads),
    )  # fmt: skip
    if rc != 0:
        raise RuntimeError(f"pmnet_oracle_score failed with code {rc}")
    out = dict(scores=scores, status=status, stats=stats)
    if with_conf:
        out["conf"] = conf
    return out
