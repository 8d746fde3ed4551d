"""TEST INFRASTRUCTURE - pin the functional-group typing rules on the reference's own code (build container only).

    python oracle/make_golden_typing.py      # writes tests/golden/typing_examples.json

The reference types a ligand in two steps: OpenBabel perception (donor / acceptor flags, SSSR aromatic rings,
hybridisation, degrees - third-party, absent here) and its own functional-group rules on top of those flags
(`get_pharmacophore_nodes`, src/pmnet/scoring/ligand_utils.py:25-88, `is_*` :91-184). This script runs the SECOND step
of the unmodified reference on duck-typed atoms that answer the OpenBabel queries from a `pharmaconet_b200` AtomTable
(built by the toolkit-free reader from the reference's examples/library.tar plus hand-made molecules with the rarer
groups), and stores tables + the reference's pharmacophore list. `tests/test_ligand_typing.py` checks that
`ligand_typing.type_atoms` gives the same list for the same table: the rules layer is pinned, the perception is not.
"""

from __future__ import annotations

import json
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_harness  # noqa: E402
from sdf_util import molblock, ring  # noqa: E402

from pharmaconet_b200 import sdf  # noqa: E402
from pharmaconet_b200.ligand_typing import AtomTable, type_atoms  # noqa: E402

LIBRARY = "/root/reference/examples/library.tar"
N_FILES = 400


class FakeAtom:
    def __init__(self, table: AtomTable, i: int, donor_view: bool = False):
        self.t, self.i, self.donor_view = table, i, donor_view
        self.nb = []

    def GetIdx(self):
        return self.i + 1

    def GetAtomicNum(self):
        return self.t.atomic_nums[self.i]

    def GetExplicitDegree(self):
        return self.t.explicit_degree[self.i]

    def GetHvyDegree(self):
        return self.t.heavy_degree[self.i]

    def GetHyb(self):
        return self.t.hyb[self.i]

    def IsHbondAcceptor(self):
        return self.t.is_acceptor[self.i]

    def IsHbondDonor(self):
        return self.t.is_donor[self.i]


class FakeRing:
    def __init__(self, path0):
        self._path = [i + 1 for i in path0]

    def IsAromatic(self):
        return True


class FakeOBMol:
    def __init__(self, atoms):
        self.atoms = atoms

    def AddPolarHydrogens(self):
        pass


class FakeMol:
    def __init__(self, table: AtomTable):
        atoms = [FakeAtom(table, i) for i in range(len(table.atomic_nums))]
        for a in atoms:
            a.nb = [atoms[j] for j in table.neighbors[a.i]]
        self.OBMol = FakeOBMol(atoms)
        self.sssr = [FakeRing(r) for r in table.aromatic_rings]

    @property
    def clone(self):
        return self


def special_molecules():
    out = {}
    # guanidinium-like C(N)(N)N with one terminal N, sulfonate, phosphate, sulfonium, quaternary N, aryl halides
    out["guanidine"] = molblock([("C", 0, 0, 0), ("N", 1.3, 0, 0), ("N", -0.7, 1.1, 0), ("N", -0.7, -1.1, 0), ("C", 2.0, 1.2, 0)],
                                [(1, 2, 1), (1, 3, 2), (1, 4, 1), (2, 5, 1)])  # fmt: skip
    out["sulfonate"] = molblock([("S", 0, 0, 0), ("O", 1.4, 0, 0), ("O", -0.7, 1.2, 0), ("O", -0.7, -1.2, 0.3), ("C", 0, 0, 1.8)],
                                [(1, 2, 2), (1, 3, 2), (1, 4, 1), (1, 5, 1)], charges=[(4, -1)])  # fmt: skip
    out["sulfate"] = molblock([("S", 0, 0, 0), ("O", 1.4, 0, 0), ("O", -0.7, 1.2, 0), ("O", -0.7, -1.2, 0.3), ("O", 0, 0, 1.6), ("C", 0.5, 0.5, 2.8)],
                              [(1, 2, 2), (1, 3, 2), (1, 4, 1), (1, 5, 1), (5, 6, 1)], charges=[(4, -1)])  # fmt: skip
    out["phosphate"] = molblock([("P", 0, 0, 0), ("O", 1.5, 0, 0), ("O", -0.7, 1.3, 0), ("O", -0.7, -1.3, 0.3), ("O", 0, 0, 1.6), ("C", 0.5, 0.5, 2.8)],
                                [(1, 2, 2), (1, 3, 1), (1, 4, 1), (1, 5, 1), (5, 6, 1)], charges=[(3, -1), (4, -1)])  # fmt: skip
    out["sulfonium"] = molblock([("S", 0, 0, 0), ("C", 1.8, 0, 0), ("C", -0.9, 1.5, 0), ("C", -0.9, -1.5, 0.3)],
                                [(1, 2, 1), (1, 3, 1), (1, 4, 1)], charges=[(1, 1)])  # fmt: skip
    out["quat_n"] = molblock([("N", 0, 0, 0), ("C", 1, 1, 1), ("C", -1, -1, 1), ("C", -1, 1, -1), ("C", 1, -1, -1)],
                             [(1, k, 1) for k in range(2, 6)], charges=[(1, 1)])  # fmt: skip
    a, b = ring(6, "CCCCCC", [2, 1, 2, 1, 2, 1], [("Cl", 2.9, 0, 0), ("F", -2.7, 0.1, 0), ("Br", 0.1, 3.0, 0.1)], [(1, 7, 1), (4, 8, 1), (2, 9, 1)])
    out["aryl_halides"] = molblock(a, b)
    out["carboxylic_acid"] = molblock([("C", 0, 0, 0), ("C", 1.5, 0, 0), ("O", 2.2, 1.0, 0), ("O", 2.2, -1.0, 0), ("H", 3.1, -1.0, 0)],
                                      [(1, 2, 1), (2, 3, 2), (2, 4, 1), (4, 5, 1)])  # fmt: skip
    return out


def main():
    _, _, _, ligand_utils = ref_harness.import_reference()
    ligand_utils.ob.OBMolAtomIter = lambda obmol: iter(obmol.atoms)
    ligand_utils.ob.OBAtomAtomIter = lambda atom: iter(atom.nb)
    records = {}
    with tarfile.open(LIBRARY) as tar:
        members = sorted((m for m in tar.getmembers() if m.name.endswith(".sdf")), key=lambda m: m.name)[:N_FILES]
        for m in members:
            records[os.path.basename(m.name)] = tar.extractfile(m).read().decode()
    records.update(special_molecules())
    out = []
    counts = {}
    for name, text in records.items():
        table, _ = sdf.perceive(sdf.parse_sdf(text, 1)[0])
        nodes = ligand_utils.get_pharmacophore_nodes(FakeMol(table))
        ref = []
        for typ, lst in nodes.items():
            for n in lst:
                ref.append([typ, n.atom_indices, n.center_indices])
                counts[typ] = counts.get(typ, 0) + 1
        mine = [[t, a, c] for t, a, c in type_atoms(table)]
        norm = lambda x: json.loads(json.dumps(x))  # noqa: E731 - tuples -> lists
        assert norm(mine) == norm(ref), (name, mine, ref)
        out.append(dict(name=name, table={k: getattr(table, k) for k in AtomTable.__dataclass_fields__}, pharmacophores=ref))
    path = os.path.join(ROOT, "tests", "golden", "typing_examples.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"{len(out)} molecules, pharmacophore counts {counts} -> {path} ({os.path.getsize(path) / 1e3:.0f} KB)")


if __name__ == "__main__":
    main()
