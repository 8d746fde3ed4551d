"""Ligand pharmacophore typing: molecule -> `TypedLigand` (the reference's `Ligand.pharmacophore_list`).

Two layers:
* `type_atoms(table)` - the functional-group rules of src/pmnet/scoring/ligand_utils.py:25-184 on a plain
  `AtomTable` (atomic numbers, neighbour lists, degrees, hybridisation, donor / acceptor flags, aromatic rings).
  Pure Python, unit-tested on hand-made tables.
* `table_from_pbmol` / `typed_ligand_from_file|smiles|pbmol` - fills the table from OpenBabel, exactly the
  queries the reference makes (`IsHbondAcceptor`, `IsHbondDonor` on a polar-hydrogen clone, `sssr` aromatic rings,
  `GetHyb`, `GetExplicitDegree`, `GetHvyDegree`). OpenBabel is not available in the build / benchmark environment
  (SURVEY section 0.6), so this layer is UNVERIFIED offline and raises ImportError when the toolkit is missing.
  Perception itself (aromaticity, donors / acceptors) stays OpenBabel's: it is third-party arithmetic outside the
  reference repository.
"""

from __future__ import annotations

import itertools
import os
from dataclasses import dataclass, field

import numpy as np

from .ligand import TypedLigand

HALOGENS = (9, 17, 35, 53)


@dataclass
class AtomTable:
    """Heavy atoms of a hydrogen-stripped molecule (indices 0-based). `neighbors[i]` lists ALL bonded atoms present
    in the stripped molecule; `neighbor_z[i]` their atomic numbers."""

    atomic_nums: list[int]
    neighbors: list[list[int]]
    explicit_degree: list[int]
    heavy_degree: list[int]
    hyb: list[int]
    is_acceptor: list[bool]
    is_donor: list[bool]
    aromatic_rings: list[tuple[int, ...]] = field(default_factory=list)

    def nz(self, i: int) -> list[int]:
        return [self.atomic_nums[j] for j in self.neighbors[i]]


def _is_quartamine_n(t: AtomTable, i: int) -> bool:  # ligand_utils.py:94-103
    return t.atomic_nums[i] == 7 and t.explicit_degree[i] == 4 and all(z != 1 for z in t.nz(i))


def _is_tertamine_n(t: AtomTable, i: int) -> bool:  # :106-107
    return t.atomic_nums[i] == 7 and t.hyb[i] == 3 and t.heavy_degree[i] == 3


def _is_sulfonium_s(t: AtomTable, i: int) -> bool:  # :110-118
    return t.atomic_nums[i] == 16 and t.explicit_degree[i] == 3 and all(z != 1 for z in t.nz(i))


def _is_guanidine_c(t: AtomTable, i: int) -> bool:  # :121-133
    if t.atomic_nums[i] != 6:
        return False
    n_terminal = 0
    for j in t.neighbors[i]:
        if t.atomic_nums[j] != 7:
            return False
        n_terminal += t.heavy_degree[j] == 1
    return len(t.neighbors[i]) == 3 and n_terminal > 0


def _count(t: AtomTable, i: int, z: int) -> int:
    return sum(1 for v in t.nz(i) if v == z)


def type_atoms(t: AtomTable) -> list[tuple[str, int | tuple[int, ...], int | tuple[int, ...]]]:
    """(type, atom key, centre key) in the reference's list order: Hydrophobic, Aromatic, Cation, Anion,
    HBond_donor, HBond_acceptor, Halogen (ligand_utils.py:80-88, ligand.py:56-59)."""
    n = len(t.atomic_nums)
    z = t.atomic_nums
    hydrophobic = [i for i in range(n) if z[i] == 6 and all(v in (1, 6) for v in t.nz(i))]
    acceptors = [i for i in range(n) if z[i] not in HALOGENS and t.is_acceptor[i]]
    donors = [i for i in range(n) if t.is_donor[i]]
    rings = sorted(tuple(sorted(r)) for r in t.aromatic_rings)
    cations: list[tuple] = [(i, i) for i in range(n) if _is_quartamine_n(t, i) or _is_tertamine_n(t, i) or _is_sulfonium_s(t, i)]
    anions: list[tuple] = []
    for i in range(n):
        if _is_guanidine_c(t, i):
            ns = tuple(j for j in t.neighbors[i] if z[j] == 7)
            cations.append(((i,) + ns, i))
        elif (z[i] == 15 and all(v == 8 for v in t.nz(i))) or (z[i] == 16 and _count(t, i, 8) == 4):  # phosphate / sulfate
            anions.append(((i,) + tuple(t.neighbors[i]), i))
        elif z[i] == 16 and _count(t, i, 8) == 3:  # sulfonic acid
            anions.append(((i,) + tuple(j for j in t.neighbors[i] if z[j] == 8), i))
        elif z[i] == 6 and _count(t, i, 8) == 2 and _count(t, i, 6) == 1:  # carboxylate: centre = the two oxygens
            os_ = tuple(j for j in t.neighbors[i] if z[j] == 8)
            anions.append(((i,) + os_, os_))
    halogens = [i for i in range(n) if z[i] in HALOGENS and 6 in t.nz(i)]
    out: list = [("Hydrophobic", i, i) for i in hydrophobic]
    out += [("Aromatic", r, r) for r in rings]
    out += [("Cation", a, c) for a, c in cations]
    out += [("Anion", a, c) for a, c in anions]
    out += [("HBond_donor", i, i) for i in donors]
    out += [("HBond_acceptor", i, i) for i in acceptors]
    out += [("Halogen", i, i) for i in halogens]
    return out


# ---------------------------------------------------------------------------------------------- OpenBabel layer
def _openbabel():
    try:
        from openbabel import pybel
        from openbabel.pybel import ob
    except Exception as e:  # noqa: BLE001
        raise ImportError(
            "ligand typing from files / SMILES needs OpenBabel (openbabel-wheel>=3.1.1.20, as in the reference); "
            "score pre-typed ligands with PharmacophoreModel.scoring_batch(LigandBatch) instead"
        ) from e
    return pybel, ob


def table_from_pbmol(pbmol) -> AtomTable:
    """pbmol: hydrogen-stripped pybel.Molecule (ligand.py:37-40)."""
    _, ob = _openbabel()
    obmol = pbmol.OBMol
    atoms = list(ob.OBMolAtomIter(obmol))
    hyd = pbmol.clone
    hyd.OBMol.AddPolarHydrogens()
    atoms_h = list(ob.OBMolAtomIter(hyd.OBMol))[: len(atoms)]
    return AtomTable(
        atomic_nums=[a.GetAtomicNum() for a in atoms],
        neighbors=[[nb.GetIdx() - 1 for nb in ob.OBAtomAtomIter(a)] for a in atoms],
        explicit_degree=[a.GetExplicitDegree() for a in atoms],
        heavy_degree=[a.GetHvyDegree() for a in atoms],
        hyb=[a.GetHyb() for a in atoms],
        is_acceptor=[bool(a.IsHbondAcceptor()) for a in atoms],
        is_donor=[bool(a.IsHbondDonor()) for a in atoms_h],
        aromatic_rings=[tuple(i - 1 for i in ring._path) for ring in pbmol.sssr if ring.IsAromatic()],
    )


def typed_ligand_from_pbmol(pbmol, atom_positions, conformer_axis: int | None = None, _unsafe: bool = False) -> TypedLigand:
    """ligand.py:17-61: positions as (C, N, 3) (axis 0 / None), (N, C, 3) (axis 1) or a list of (N, 3) arrays."""
    mol = pbmol if _unsafe else pbmol.clone
    mol.removeh()
    if isinstance(atom_positions, list):
        pos = np.stack(atom_positions, axis=1, dtype=np.float32)
    else:
        pos = np.asarray(atom_positions, dtype=np.float32)
        if conformer_axis in (0, None):
            pos = np.ascontiguousarray(np.moveaxis(pos, 0, 1))
    table = table_from_pbmol(mol)
    assert len(table.atomic_nums) == pos.shape[0]
    return TypedLigand(table.atomic_nums, table.neighbors, type_atoms(table), pos)


_warned_builtin = False


def typed_ligand_from_file(filename, num_conformers: int | None = None, perception: str = "auto") -> TypedLigand:
    """ligand.py:63-84: every molecule record of the file is one conformer of the same ligand.

    perception: "openbabel" (the reference's), "builtin" (`pharmaconet_b200.sdf`: toolkit-free, approximate,
    `.sdf` only) or "auto" (OpenBabel when importable, else the built-in reader for `.sdf`)."""
    ext = os.path.splitext(str(filename))[1]
    assert ext in [".sdf", ".pdb", ".mol2"]
    if perception not in ("auto", "openbabel", "builtin"):
        raise ValueError(f"unknown perception mode {perception!r}")
    if perception != "openbabel" and ext == ".sdf":
        use_builtin = perception == "builtin"
        if not use_builtin:
            try:
                _openbabel()
            except ImportError:
                use_builtin = True
        if use_builtin:
            global _warned_builtin
            if perception == "auto" and not _warned_builtin:
                import warnings

                warnings.warn(
                    "OpenBabel is not installed: typing .sdf ligands with the built-in approximate perception "
                    "(pharmaconet_b200.sdf); pharmacophore types may differ from the reference's",
                    stacklevel=2,
                )
                _warned_builtin = True
            from .sdf import typed_ligand_from_sdf

            return typed_ligand_from_sdf(str(filename), num_conformers)
    pybel, _ = _openbabel()
    it = pybel.readfile(ext[1:], str(filename))
    mols = list(it if num_conformers is None else itertools.islice(it, num_conformers))
    base = mols[0]
    base.removeh()
    coords = []
    for m in mols:
        m.removeh()
        assert len(m.atoms) == len(base.atoms)
        coords.append(np.asarray([a.coords for a in m.atoms], dtype=np.float32))
    lig = typed_ligand_from_pbmol(base, coords, _unsafe=True)
    lig.name = str(filename)
    return lig


def typed_ligand_from_smiles(smiles: str, num_conformers: int) -> TypedLigand:
    """ligand.py:86-107: RDKit srETKDGv3 embedding written to a temporary SDF, then read like a file."""
    import tempfile

    from rdkit import Chem
    from rdkit.Chem import rdDistGeom

    mol = Chem.AddHs(Chem.MolFromSmiles(smiles))
    rdDistGeom.EmbedMultipleConfs(mol, num_conformers, params=rdDistGeom.srETKDGv3())
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "lig.sdf")
        with Chem.SDWriter(path) as w:
            for i in range(mol.GetNumConformers()):
                w.write(mol, confId=i)
        return typed_ligand_from_file(path)
