"""ctypes mirror of include/pmnet_b200.h (struct layouts and constants). No compute here."""

from __future__ import annotations

import ctypes as C

ABI_VERSION = 4

PMNET_OK, PMNET_EINVAL, PMNET_EWORKSPACE, PMNET_ELIMIT, PMNET_ECUDA = range(5)
LIG_OK, LIG_EMPTY, LIG_OVERFLOW, LIG_UNSUPPORTED, LIG_DEFERRED, LIG_HEAVY = range(6)
MAX_CONFORMERS = 128

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)


class PmModel(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_clusters", C.c_int32),
        ("n_cluster_nodes", C.c_int32),
        ("reserved", C.c_int32),
        ("node_type", C.c_void_p),
        ("edge_mu", C.c_void_p),
        ("edge_sigma", C.c_void_p),
        ("cluster_mask", C.c_void_p),
        ("cluster_node_off", C.c_void_p),
        ("cluster_nodes", C.c_void_p),
        ("cluster_dist", C.c_void_p),
        ("cluster_size_sum", C.c_void_p),
    ]


class PmLigandBatch(C.Structure):
    _fields_ = [
        ("n_ligands", C.c_int32),
        ("lig_node_off", C.c_void_p),
        ("lig_cluster_off", C.c_void_p),
        ("cluster_node_off", C.c_void_p),
        ("cluster_nodes", C.c_void_p),
        ("node_type_mask", C.c_void_p),
        ("n_conf", C.c_void_p),
        ("coord_off", C.c_void_p),
        ("coords", C.c_void_p),
        ("coord_base", C.c_int64),
        ("node_base", C.c_int32),
        ("cluster_base", C.c_int32),
        ("cnode_base", C.c_int32),
        ("reserved", C.c_int32),
        ("order", C.c_void_p),
    ]


class PmScoreConfig(C.Structure):
    _fields_ = [
        ("warps_per_block", C.c_int32),
        ("blocks", C.c_int32),
        ("scratch_rows", C.c_int32),
        ("max_conformers", C.c_int32),
        ("rescore_status", C.c_int32),
        ("heavy_budget", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


MODEL_FIELDS = (
    "node_type",
    "edge_mu",
    "edge_sigma",
    "cluster_mask",
    "cluster_node_off",
    "cluster_nodes",
    "cluster_dist",
    "cluster_size_sum",
)
BATCH_FIELDS = (
    "lig_node_off",
    "lig_cluster_off",
    "cluster_node_off",
    "cluster_nodes",
    "node_type_mask",
    "n_conf",
    "coord_off",
    "coords",
)


def model_struct(n_nodes: int, n_clusters: int, n_cluster_nodes: int, ptrs: dict[str, int]) -> PmModel:
    """ptrs: field name -> raw address (host or device, the callee decides what it expects)."""
    m = PmModel()
    m.n_nodes, m.n_clusters, m.n_cluster_nodes = int(n_nodes), int(n_clusters), int(n_cluster_nodes)
    for f in MODEL_FIELDS:
        setattr(m, f, int(ptrs[f]) or None)
    return m


def batch_struct(n_ligands: int, ptrs: dict[str, int], bases: dict[str, int] | None = None) -> PmLigandBatch:
    """bases: optional {coord_base, node_base, cluster_base, cnode_base} for a chunk view of a larger library."""
    b = PmLigandBatch()
    b.n_ligands = int(n_ligands)
    for f in BATCH_FIELDS:
        setattr(b, f, int(ptrs[f]) or None)
    for k, v in (bases or {}).items():
        setattr(b, k, int(v))
    return b
