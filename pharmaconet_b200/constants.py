"""Shared constants of the scoring path.

Reference: src/pmnet/scoring/graph_match.py:32-40 (DEFAULT_WEIGHTS), :43-60 (priority_fn ranks),
src/pmnet/pharmacophore_model.py:22-33 (INTERACTION_TO_PHARMACOPHORE),
src/pmnet/data/constant.py:3-14 (INTERACTION_LIST), src/pmnet/utils/density_map.py:46-48 (cluster kinds).
"""

from __future__ import annotations

# The 7 ligand-side pharmacophore types; the index is the bit position used in every packed type mask.
PHARMACOPHORE_TYPES: tuple[str, ...] = (
    "Hydrophobic",
    "Aromatic",
    "Cation",
    "Anion",
    "HBond_donor",
    "HBond_acceptor",
    "Halogen",
)
TYPE_INDEX: dict[str, int] = {t: i for i, t in enumerate(PHARMACOPHORE_TYPES)}
NUM_TYPES = len(PHARMACOPHORE_TYPES)

DEFAULT_WEIGHTS: dict[str, float] = dict(
    Cation=8,
    Anion=8,
    Aromatic=4,
    HBond_donor=4,
    HBond_acceptor=4,
    Halogen=4,
    Hydrophobic=1,
)

# 10 protein-side interaction (NCI) types and the ligand pharmacophore type each one matches.
INTERACTION_LIST: tuple[str, ...] = (
    "Hydrophobic",
    "PiStacking_P",
    "PiStacking_T",
    "PiCation_lring",
    "PiCation_pring",
    "HBond_ldon",
    "HBond_pdon",
    "SaltBridge_lneg",
    "SaltBridge_pneg",
    "XBond",
)
INTERACTION_TO_PHARMACOPHORE: dict[str, str] = {
    "Hydrophobic": "Hydrophobic",
    "PiStacking_P": "Aromatic",
    "PiStacking_T": "Aromatic",
    "PiCation_lring": "Aromatic",
    "PiCation_pring": "Cation",
    "HBond_pdon": "HBond_acceptor",
    "HBond_ldon": "HBond_donor",
    "SaltBridge_pneg": "Cation",
    "SaltBridge_lneg": "Anion",
    "XBond": "Halogen",
}
INTERACTION_TO_HOTSPOT: dict[str, str] = {
    "Hydrophobic": "Hydrophobic",
    "PiStacking_P": "Aromatic",
    "PiStacking_T": "Aromatic",
    "PiCation_lring": "Cation",
    "PiCation_pring": "Aromatic",
    "HBond_pdon": "HBond_donor",
    "HBond_ldon": "HBond_acceptor",
    "SaltBridge_pneg": "Anion",
    "SaltBridge_lneg": "Cation",
    "XBond": "Halogen",
}

# Cluster kinds shared by ligand graphs and pharmacophore models (dict insertion order of the reference).
CLUSTER_KINDS: tuple[str, ...] = ("Cation", "Anion", "HBond", "Aromatic", "Hydrophobic", "Halogen")

# priority_fn (graph_match.py:43-60): (group, -size, rank, min atom index of first node)
CLUSTER_PRIORITY: dict[str, tuple[int, int]] = {
    "Aromatic": (0, 0),
    "Cation": (0, 1),
    "Anion": (0, 2),
    "HBond": (1, 0),
    "Halogen": (1, 1),
    "Hydrophobic": (1, 2),
}

MAX_TREE_DEPTH = 20  # graph_match.py:88
MIN_MATCHES_NO_SKIP = 5  # tree.py:98


def weights_vector(weights: dict[str, float] | None = None) -> list[float]:
    """7 floats in PHARMACOPHORE_TYPES order (graph_match.py:82-84: defaults updated by the user's dict)."""
    w = dict(DEFAULT_WEIGHTS)
    if weights is not None:
        for k in weights:
            if k not in TYPE_INDEX:
                raise KeyError(f"unknown pharmacophore type {k!r}")
        w.update(weights)
    return [float(w[t]) for t in PHARMACOPHORE_TYPES]
