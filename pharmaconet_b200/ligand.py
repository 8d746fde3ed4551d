"""Host-side ligand graph construction for the scoring path (array-based, no object graph).

What the reference does with OpenBabel objects in `src/pmnet/scoring/ligand.py` is done here on plain
integer arrays so that the result can be packed for the CUDA kernel:

* pharmacophore nodes keyed by their atom set, several types per node (ligand.py:134-156),
* "dependence" of Hydrophobic-on-Aromatic and HBond-on-Cation/Anion nodes (ligand.py:303-329),
* functional-group grouping of HBond / Hydrophobic nodes (ligand.py:158-213),
* clusters with one optional high-priority node followed by low-priority nodes (ligand.py:215-259, :387-395),
* the matcher's level order `priority_fn` (graph_match.py:43-60), which does not depend on the model and is
  therefore applied here once per ligand.

Chemistry perception (which atoms carry which pharmacophore type) needs OpenBabel and is NOT part of this
path: a `TypedLigand` already carries the typed atoms, exactly like the
reference's `Ligand.pharmacophore_list`.
"""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass, field

import numpy as np

from .constants import CLUSTER_PRIORITY, TYPE_INDEX

AtomKey = int | tuple[int, ...]


@dataclass
class TypedLigand:
    """A ligand after pharmacophore typing: heavy atoms, bonds, typed atom groups and conformer coordinates.

    pharmacophores: (type, atom_indices, center_indices) in the order of the reference's
        `Ligand.pharmacophore_list` (ligand.py:56-59): all Hydrophobic, then Aromatic, Cation, Anion,
        HBond_donor, HBond_acceptor, Halogen. `atom_indices`/`center_indices` are an int or a tuple of ints.
    atom_positions: float32 [num_atoms, num_conformers, 3] (the reference's internal layout, ligand.py:44).
    """

    atomic_nums: Sequence[int]
    neighbors: Sequence[Sequence[int]]
    pharmacophores: Sequence[tuple[str, AtomKey, AtomKey]]
    atom_positions: np.ndarray | None = None
    name: str = ""

    @property
    def num_atoms(self) -> int:
        return len(self.atomic_nums)

    @property
    def num_conformers(self) -> int:
        assert self.atom_positions is not None
        return int(self.atom_positions.shape[1])


@dataclass
class LigandCluster:
    kind: str  # one of constants.CLUSTER_KINDS
    high: int | None = None
    low: list[int] = field(default_factory=list)

    @property
    def nodes(self) -> list[int]:
        # ligand.py:387-395: the high-priority node first, then the low-priority ones in insertion order
        return ([self.high] if self.high is not None else []) + self.low


@dataclass
class LigandTopology:
    """Conformer-independent part of a ligand graph, clusters already in matcher priority order."""

    num_atoms: int
    node_type_mask: np.ndarray  # uint8 [Nn]
    node_center_atoms: list[tuple[int, ...]]  # atoms averaged for the node position (ligand.py:293-301)
    node_atoms: list[tuple[int, ...]]
    clusters: list[LigandCluster]  # graph order (ligand.py:256)
    cluster_order: list[int]  # indices into `clusters`, stable-sorted by priority_fn

    @property
    def num_nodes(self) -> int:
        return len(self.node_type_mask)

    def ordered_cluster_nodes(self) -> list[list[int]]:
        return [self.clusters[i].nodes for i in self.cluster_order]


def _as_tuple(key: AtomKey) -> tuple[int, ...]:
    return (int(key),) if isinstance(key, (int, np.integer)) else tuple(int(k) for k in key)


def _has(types: list[str], *prefixes: str) -> bool:
    return any(t.startswith(prefixes) for t in types)


def build_topology(lig: TypedLigand) -> LigandTopology:
    # ---- nodes (ligand.py:134-156): one node per distinct atom key, types appended in list order ----
    key_to_node: dict[AtomKey, int] = {}
    types: list[list[str]] = []
    atoms: list[frozenset[int]] = []
    centers: list[tuple[int, ...]] = []
    center_is_scalar: list[bool] = []
    depends: list[set[int]] = []
    by_type: dict[str, list[int]] = {}
    for typ, atom_key, center_key in lig.pharmacophores:
        hkey = int(atom_key) if isinstance(atom_key, (int, np.integer)) else tuple(int(a) for a in atom_key)
        n = key_to_node.get(hkey)
        if n is not None:
            types[n].append(typ)
            by_type.setdefault(typ, []).append(n)
            continue
        new = len(types)
        key_to_node[hkey] = new
        types.append([typ])
        atoms.append(frozenset(_as_tuple(atom_key)))
        centers.append(_as_tuple(center_key))
        center_is_scalar.append(isinstance(center_key, (int, np.integer)))
        depends.append(set())
        by_type.setdefault(typ, []).append(new)
        # dependence rules, evaluated old-node vs new-node with the types known at this moment
        # (ligand.py:303-329: an if/elif chain, the first type match wins even when the subset test fails)
        for old in range(new):
            t_old, t_new = types[old], types[new]
            if _has(t_old, "Hydrophobic") and _has(t_new, "Aromatic"):
                if atoms[old] <= atoms[new]:
                    depends[old].add(new)
            elif _has(t_old, "Aromatic") and _has(t_new, "Hydrophobic"):
                if atoms[new] <= atoms[old]:
                    depends[new].add(old)
            elif _has(t_old, "HBond") and _has(t_new, "Cation", "Anion"):
                if atoms[old] <= atoms[new]:
                    depends[old].add(new)
            elif _has(t_old, "Cation", "Anion") and _has(t_new, "HBond"):
                if atoms[new] <= atoms[old]:
                    depends[new].add(old)
    n_nodes = len(types)

    # ---- functional-group grouping (ligand.py:158-213) ----
    group: list[set[int]] = [set() for _ in range(n_nodes)]

    def heavy_neighbors(atom: int) -> list[int]:
        return [int(a) for a in lig.neighbors[atom] if lig.atomic_nums[a] != 1]

    def link(members: list[int], node: int) -> None:
        for m in members:
            group[m].add(node)
            group[node].add(m)

    hbond_groups: dict[int, list[int]] = {}
    hydrop_groups: dict[int, list[int]] = {}
    for n in range(n_nodes):
        if "HBond_acceptor" in types[n] or "HBond_donor" in types[n]:
            table = hbond_groups
        elif "Hydrophobic" in types[n]:
            table = hydrop_groups
        else:
            continue
        assert len(atoms[n]) == 1
        nbrs = heavy_neighbors(next(iter(atoms[n])))
        if len(nbrs) == 1:
            members = table.setdefault(nbrs[0], [])
            link(members, n)
            members.append(n)

    # connected hydrophobic carbons become one group; start nodes are taken last-inserted-first
    pending: dict[int, int] = {next(iter(atoms[n])): n for n in by_type.get("Hydrophobic", [])}
    while pending:
        _, start = pending.popitem()
        members = [start] + sorted(group[start])
        frontier = [next(iter(atoms[m])) for m in members]
        for atom in frontier:  # grows while iterating (breadth-first)
            for nbr in lig.neighbors[atom]:
                if lig.atomic_nums[nbr] != 6:
                    continue
                other = pending.pop(int(nbr), None)
                if other is None:
                    continue
                frontier.append(int(nbr))
                link(members, other)
                members.append(other)

    # ---- clusters (ligand.py:215-259) ----
    clusters: list[LigandCluster] = []
    owner: dict[int, int] = {}  # node that opened a cluster -> cluster index
    placed: set[int] = set()
    for typ in ("Aromatic", "Cation", "Anion", "Halogen"):
        for n in by_type.get(typ, []):
            if n in placed:
                continue
            placed.add(n)
            owner[n] = len(clusters)
            clusters.append(LigandCluster(kind=typ, high=n))
    for typ in ("Hydrophobic", "HBond_donor", "HBond_acceptor"):
        for n in by_type.get(typ, []):
            if n in placed:
                continue
            placed.add(n)
            target: int | None = None
            if depends[n]:
                target = owner[min(depends[n])]
            elif group[n]:
                # reference iterates a set of objects (address order); ties are resolved here by lowest index
                for g in sorted(group[n]):
                    if g in owner:
                        target = owner[g]
                        break
            if target is None:
                owner[n] = len(clusters)
                clusters.append(LigandCluster(kind="HBond" if typ.startswith("HBond") else "Hydrophobic", low=[n]))
            else:
                clusters[target].low.append(n)

    # ---- matcher level order (graph_match.py:43-60, :87): stable sort by priority ----
    def priority(ci: int):
        c = clusters[ci]
        grp, rank = CLUSTER_PRIORITY[c.kind]
        return (grp, -len(c.nodes), rank, min(atoms[c.nodes[0]]))

    order = sorted(range(len(clusters)), key=priority)

    mask = np.zeros(n_nodes, dtype=np.uint8)
    for n, ts in enumerate(types):
        for t in ts:
            mask[n] |= 1 << TYPE_INDEX[t]
    return LigandTopology(
        num_atoms=lig.num_atoms,
        node_type_mask=mask,
        node_center_atoms=[c for c in centers],
        node_atoms=[tuple(sorted(a)) for a in atoms],
        clusters=clusters,
        cluster_order=order,
    )


def node_positions(topology: LigandTopology, atom_positions: np.ndarray) -> np.ndarray:
    """float32 [Nn, C, 3] node coordinates (ligand.py:293-301: the atom itself or the fp32 mean of atoms)."""
    atom_positions = np.asarray(atom_positions, dtype=np.float32)
    out = np.empty((topology.num_nodes,) + atom_positions.shape[1:], dtype=np.float32)
    for n, c in enumerate(topology.node_center_atoms):
        if len(c) == 1:
            out[n] = atom_positions[c[0]]
        else:
            out[n] = np.mean(atom_positions[list(c), :], axis=0, dtype=np.float32)
    return out
