"""Loader of the C-ABI shared library (include/pmnet_b200.h). The CUDA extension is the product: if it cannot be
loaded this module raises - there is no CPU or PyTorch fallback for the scoring path."""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

from . import _abi

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.environ.get("PMNET_B200_SO") or os.path.join(_PKG, "libpmnet_b200.so")  # env override: developer A/B builds
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("scoring.cu", "conv3d.cu", "pointwise.cu", "swin_ops.cu", "gemm.cu")]
HEADERS = [os.path.join(_ROOT, "include", "pmnet_b200.h"), os.path.join(_PKG, "csrc", "scoring_fast.cuh")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]  # fmt: skip

_lib = None


def _nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libpmnet_b200.so next to this file (in-tree, so it travels with the repo)."""
    if not force and not is_stale():
        return SO_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libpmnet_b200.so")
    import fcntl

    # one builder at a time (several ranks of a fresh checkout may get here together); every builder writes its own
    # temporary file and the finished library is moved into place atomically
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while this one waited
                return SO_PATH
            tmp = f"{SO_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, *SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
            os.replace(tmp, SO_PATH)
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    # missing, or older than a source / header (an edited .cu must not run stale). A developer override
    # (PMNET_B200_SO = an A/B variant built with other flags) is loaded as it is.
    if not os.environ.get("PMNET_B200_SO") and is_stale() and _nvcc() is not None:
        build()
    L = C.CDLL(SO_PATH)
    L.pmnet_abi_version.restype = C.c_int
    if L.pmnet_abi_version() != _abi.ABI_VERSION:
        raise RuntimeError("libpmnet_b200.so ABI version mismatch; rebuild with pharmaconet_b200._lib.build(force=True)")
    L.pmnet_last_error_string.restype = C.c_char_p
    L.pmnet_score_workspace_bytes.restype = C.c_size_t
    L.pmnet_score_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.POINTER(_abi.PmScoreConfig)]
    L.pmnet_score_batch.restype = C.c_int
    L.pmnet_score_batch.argtypes = [
        C.POINTER(_abi.PmModel), C.POINTER(_abi.PmLigandBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_abi.PmScoreConfig), C.c_void_p,
    ]  # fmt: skip
    L.pmnet_order_workspace_bytes.restype = C.c_size_t
    L.pmnet_order_workspace_bytes.argtypes = [C.c_int32]
    L.pmnet_cost_order.restype = C.c_int
    L.pmnet_cost_order.argtypes = [
        C.POINTER(_abi.PmModel), C.POINTER(_abi.PmLigandBatch), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_topk_workspace_bytes.restype = C.c_size_t
    L.pmnet_topk_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    L.pmnet_topk.restype = C.c_int
    L.pmnet_topk.argtypes = [
        C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_conv3d_k3_c96.restype = C.c_int
    L.pmnet_conv3d_k3_c96.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_conv3d_k3_c96_pass.restype = C.c_int
    L.pmnet_conv3d_k3_c96_pass.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_lateral_c96_split.restype = C.c_int
    L.pmnet_lateral_c96_split.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_box_combine_c96_split.restype = C.c_int
    L.pmnet_box_combine_c96_split.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_lateral_c96.restype = C.c_int
    L.pmnet_lateral_c96.argtypes = [
        C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_box_combine_c96.restype = C.c_int
    L.pmnet_box_combine_c96.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_density_post.restype = C.c_int
    L.pmnet_density_post.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32, C.c_int32,
        C.c_void_p,
    ]  # fmt: skip
    L.pmnet_window_attention.restype = C.c_int
    L.pmnet_window_attention.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
        C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_ln_residual.restype = C.c_int
    L.pmnet_ln_residual.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
        C.c_void_p,
    ]  # fmt: skip
    L.pmnet_window_attention_split.restype = C.c_int
    L.pmnet_window_attention_split.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
        C.c_int32, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_ln_residual_split.restype = C.c_int
    L.pmnet_ln_residual_split.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
        C.c_int32, C.c_float, C.c_void_p,
    ]  # fmt: skip
    L.pmnet_gemm_bf16.restype = C.c_int
    L.pmnet_gemm_bf16.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
        C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]  # fmt: skip
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pmnet_last_error_string()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


EXPORTS = (
    "pmnet_abi_version",
    "pmnet_last_error_string",
    "pmnet_score_workspace_bytes",
    "pmnet_score_batch",
    "pmnet_order_workspace_bytes",
    "pmnet_cost_order",
    "pmnet_topk_workspace_bytes",
    "pmnet_topk",
    "pmnet_conv3d_k3_c96",
    "pmnet_conv3d_k3_c96_pass",
    "pmnet_lateral_c96",
    "pmnet_lateral_c96_split",
    "pmnet_box_combine_c96",
    "pmnet_box_combine_c96_split",
    "pmnet_density_post",
    "pmnet_window_attention",
    "pmnet_window_attention_split",
    "pmnet_ln_residual",
    "pmnet_ln_residual_split",
    "pmnet_gemm_bf16",
)
