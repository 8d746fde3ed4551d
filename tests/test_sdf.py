"""Toolkit-free SDF reader (pharmaconet_b200/sdf.py): parsing and the approximate perception rules on hand-made
molecules. Typing parity with OpenBabel is unpinned (the toolkit is absent); these tests pin the stated rules."""

import numpy as np

from sdf_util import molblock, ring

from pharmaconet_b200 import sdf
from pharmaconet_b200.ligand_typing import type_atoms, typed_ligand_from_file


def types_of(block):
    table, heavy = sdf.perceive(sdf.parse_sdf(block)[0])
    out = {}
    for t, atoms, _ in type_atoms(table):
        out.setdefault(t, []).append(atoms)
    return table, out


def test_benzene_is_one_aromatic_ring_of_hydrophobic_carbons():
    atoms, bonds = ring(6, "CCCCCC", [2, 1, 2, 1, 2, 1])
    atoms += [("H", 2.4 * np.cos(2 * np.pi * i / 6), 2.4 * np.sin(2 * np.pi * i / 6), 0.0) for i in range(6)]
    bonds += [(i + 1, 7 + i, 1) for i in range(6)]
    table, t = types_of(molblock(atoms, bonds))
    assert len(table.atomic_nums) == 6  # hydrogens stripped
    assert t["Aromatic"] == [(0, 1, 2, 3, 4, 5)]
    assert sorted(t["Hydrophobic"]) == list(range(6))
    assert "HBond_donor" not in t and "HBond_acceptor" not in t


def test_cyclohexane_and_cyclohexene_are_not_aromatic():
    for orders in ([1] * 6, [2, 1, 1, 1, 1, 1], [2, 1, 2, 1, 1, 1]):
        atoms, bonds = ring(6, "CCCCCC", orders)
        _, t = types_of(molblock(atoms, bonds))
        assert "Aromatic" not in t


def test_pyridine_n_accepts_pyrrole_nh_donates_furan_o_is_silent():
    atoms, bonds = ring(6, "NCCCCC", [2, 1, 2, 1, 2, 1])
    _, t = types_of(molblock(atoms, bonds))
    assert t["Aromatic"] == [(0, 1, 2, 3, 4, 5)] and t["HBond_acceptor"] == [0] and "HBond_donor" not in t
    atoms, bonds = ring(5, "NCCCC", [1, 2, 1, 2, 1], [("H", 2.4, 0.0, 0.0)], [(1, 6, 1)])
    _, t = types_of(molblock(atoms, bonds))
    assert t["Aromatic"] == [(0, 1, 2, 3, 4)] and t["HBond_donor"] == [0] and "HBond_acceptor" not in t
    atoms, bonds = ring(5, "OCCCC", [1, 2, 1, 2, 1])
    _, t = types_of(molblock(atoms, bonds))
    assert t["Aromatic"] == [(0, 1, 2, 3, 4)] and "HBond_acceptor" not in t


def test_naphthalene_has_two_aromatic_rings_in_either_kekule_form():
    # atoms 1-10, fusion bond 5-10 ... ring A: 1 2 3 4 5 10, ring B: 5 6 7 8 9 10
    atoms = [("C", float(i), 0.0, 0.0) for i in range(10)]
    ring_bonds = [(1, 2), (2, 3), (3, 4), (4, 5), (5, 10), (10, 1), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10)]
    for doubles in ({(1, 2), (3, 4), (5, 10), (6, 7), (8, 9)}, {(10, 1), (2, 3), (4, 5), (6, 7), (8, 9)}):
        bonds = [(i, j, 2 if (i, j) in doubles else 1) for i, j in ring_bonds]
        _, t = types_of(molblock(atoms, bonds))
        assert sorted(t["Aromatic"]) == [(0, 1, 2, 3, 4, 9), (4, 5, 6, 7, 8, 9)]


def test_functional_groups():
    # acetic acid CH3-C(=O)-OH: carboxylate anion rule, carbonyl O accepts, hydroxyl donates and accepts
    atoms = [("C", 0, 0, 0), ("C", 1.5, 0, 0), ("O", 2.2, 1.0, 0), ("O", 2.2, -1.0, 0), ("H", 3.1, -1.0, 0)]
    bonds = [(1, 2, 1), (2, 3, 2), (2, 4, 1), (4, 5, 1)]
    _, t = types_of(molblock(atoms, bonds))
    assert t["Anion"] == [(1, 2, 3)] and t["HBond_donor"] == [3] and sorted(t["HBond_acceptor"]) == [2, 3]
    assert t["Hydrophobic"] == [0]
    # trimethylamine: tertiary amine cation + acceptor; acetamide N: neither cation nor acceptor, but a donor
    atoms = [("N", 0, 0, 0), ("C", 1.4, 0, 0), ("C", -0.7, 1.2, 0), ("C", -0.7, -1.2, 0)]
    _, t = types_of(molblock(atoms, [(1, 2, 1), (1, 3, 1), (1, 4, 1)]))
    assert t["Cation"] == [0] and t["HBond_acceptor"] == [0]
    atoms = [("C", 0, 0, 0), ("C", 1.5, 0, 0), ("O", 2.2, 1.0, 0), ("N", 2.2, -1.0, 0), ("H", 3.1, -1.0, 0), ("H", 1.8, -1.9, 0)]
    _, t = types_of(molblock(atoms, [(1, 2, 1), (2, 3, 2), (2, 4, 1), (4, 5, 1), (4, 6, 1)]))
    assert "Cation" not in t and t["HBond_acceptor"] == [2] and t["HBond_donor"] == [3]
    # chlorobenzene-like C-Cl: halogen on carbon; the halogen never accepts
    _, t = types_of(molblock([("C", 0, 0, 0), ("Cl", 1.7, 0, 0), ("C", -1.5, 0, 0)], [(1, 2, 1), (1, 3, 1)]))
    assert t["Halogen"] == [1] and "HBond_acceptor" not in t


def test_hydrogen_suppressed_records_get_implicit_polar_hydrogens():
    """Vendor libraries usually ship without hydrogens; the reference adds the polar ones (OpenBabel) before typing.
    Implicit H = standard valence - bond orders: donors are typed without any H atom in the connection table."""
    # ethanol C-C-O: the hydroxyl donates and accepts
    _, t = types_of(molblock([("C", 0, 0, 0), ("C", 1.5, 0, 0), ("O", 2.2, 1.0, 0)], [(1, 2, 1), (2, 3, 1)]))
    assert t["HBond_donor"] == [2] and t["HBond_acceptor"] == [2]
    # acetamide C-C(=O)-N: the N donates, does not accept; the carbonyl O accepts, does not donate
    atoms = [("C", 0, 0, 0), ("C", 1.5, 0, 0), ("O", 2.2, 1.0, 0), ("N", 2.2, -1.0, 0)]
    _, t = types_of(molblock(atoms, [(1, 2, 1), (2, 3, 2), (2, 4, 1)]))
    assert t["HBond_donor"] == [3] and t["HBond_acceptor"] == [2]
    # Kekule pyrrole (N with two single ring bonds): aromatic, N-H donates, no acceptor; pyridine: no hydrogen on N
    atoms, bonds = ring(5, "NCCCC", [1, 2, 1, 2, 1])
    table, t = types_of(molblock(atoms, bonds))
    assert t["Aromatic"] == [(0, 1, 2, 3, 4)] and t["HBond_donor"] == [0] and "HBond_acceptor" not in t
    atoms, bonds = ring(6, "NCCCCC", [2, 1, 2, 1, 2, 1])
    _, t = types_of(molblock(atoms, bonds))
    assert t["HBond_acceptor"] == [0] and "HBond_donor" not in t
    # trimethylammonium cation written without hydrogens and with its charge: one implicit H on N+
    atoms = [("N", 0, 0, 0), ("C", 1.4, 0, 0), ("C", -0.7, 1.2, 0), ("C", -0.7, -1.2, 0)]
    _, t = types_of(molblock(atoms, [(1, 2, 1), (1, 3, 1), (1, 4, 1)], charges=[(1, 1)]))
    assert t["HBond_donor"] == [0] and "HBond_acceptor" not in t
    # a record WITH explicit hydrogens is taken as complete: ether oxygen next to explicit C-H stays a non-donor
    atoms = [("C", 0, 0, 0), ("O", 1.4, 0, 0), ("C", 2.8, 0, 0), ("H", -0.5, 0.9, 0)]
    _, t = types_of(molblock(atoms, [(1, 2, 1), (2, 3, 1), (1, 4, 1)]))
    assert "HBond_donor" not in t


def test_charges_and_conformers(tmp_path):
    # tetramethylammonium (quaternary N, M  CHG line) written as three conformers
    atoms = [("N", 0, 0, 0), ("C", 1, 1, 1), ("C", -1, -1, 1), ("C", -1, 1, -1), ("C", 1, -1, -1)]
    bonds = [(1, k, 1) for k in range(2, 6)]
    text = ""
    for shift in (0.0, 0.5, 1.0):
        text += molblock([(s, x + shift, y, z) for s, x, y, z in atoms], bonds, charges=[(1, 1)])
    path = tmp_path / "tma.sdf"
    path.write_text(text)
    recs = sdf.parse_sdf(text)
    assert len(recs) == 3 and recs[0].charges[0] == 1
    lig = typed_ligand_from_file(str(path), perception="builtin")
    assert lig.num_atoms == 5 and lig.num_conformers == 3
    assert lig.atom_positions.shape == (5, 3, 3) and lig.atom_positions.dtype == np.float32
    assert np.allclose(lig.atom_positions[0, :, 0], [0.0, 0.5, 1.0])
    assert ("Cation", 0, 0) in [tuple(p) for p in lig.pharmacophores]
    assert not any(p[0] == "HBond_acceptor" for p in lig.pharmacophores)  # N+ with four bonds
    assert typed_ligand_from_file(str(path), num_conformers=2, perception="builtin").num_conformers == 2


def test_pack_library_tool_roundtrip(tmp_path):
    """tools/pack_library.py: directory of .sdf files -> packed .npz that load_library reads back identically."""
    import os
    import subprocess
    import sys

    from pharmaconet_b200.packing import LigandBatch, load_library

    lib = tmp_path / "lib"
    lib.mkdir()
    atoms, bonds = ring(6, "CCCCCC", [2, 1, 2, 1, 2, 1], [("O", 2.8, 0.0, 0.0), ("H", 3.4, 0.7, 0.0)], [(1, 7, 1), (7, 8, 1)])
    (lib / "b_phenol.sdf").write_text(molblock(atoms, bonds) * 3)
    (lib / "a_amine.sdf").write_text(
        molblock([("N", 0, 0, 0), ("C", 1.4, 0, 0), ("C", -0.7, 1.2, 0), ("C", -0.7, -1.2, 0)], [(1, 2, 1), (1, 3, 1), (1, 4, 1)]) * 2
    )
    (lib / "c_broken.sdf").write_text("not a molfile\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "lib.npz"
    r = subprocess.run(
        [sys.executable, os.path.join(root, "tools", "pack_library.py"), "-d", str(lib), "-o", str(out), "--perception", "builtin"],
        check=True, capture_output=True, text=True, timeout=120,
    )  # fmt: skip
    assert "2 ligands, 5 conformers" in r.stdout and "c_broken.sdf" in r.stderr
    batch, names = load_library(out)
    assert [os.path.basename(n) for n in names] == ["a_amine.sdf", "b_phenol.sdf"]
    assert list(batch.n_conf) == [2, 3]
    ref = LigandBatch.from_typed([typed_ligand_from_file(str(lib / n), perception="builtin") for n in ("a_amine.sdf", "b_phenol.sdf")])
    for k, v in ref.arrays().items():
        assert np.array_equal(v, batch.arrays()[k]), k
