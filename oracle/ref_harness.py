"""TEST INFRASTRUCTURE - import harness for the *real* reference (SeonghwanSeo/PharmacoNet) in the build container.

Used by `oracle/make_golden*.py` (build container, `/root/reference`) and by `oracle/ref_pool.py`, the CPU-baseline
arm of bench.py, which on the GPU box imports the unmodified copy `oracle/_ref/pmnet` made by `oracle/make_ref.py`.
The product package never imports anything from `oracle/`.

Recipe: SURVEY.md Appendix D. OpenBabel / molvoxel / omegaconf / biopython are absent, so they are stubbed; the
scoring core (`pmnet.scoring.graph_match`, `.tree`, `.match_utils_numba`, `pmnet.pharmacophore_model`,
`pmnet.utils.density_map`) needs only numpy + numba and runs unmodified.
"""

from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))
# build container: the reference where it lies; GPU box: the unmodified copy made by oracle/make_ref.py (git-ignored)
REFERENCE_SRC = os.environ.get("PMNET_REFERENCE_SRC") or (
    "/root/reference/src" if os.path.isdir("/root/reference/src/pmnet") else os.path.join(_HERE, "_ref")
)


def import_reference():
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "pmnet")):
        raise RuntimeError(f"reference sources not found at {REFERENCE_SRC} (run oracle/make_ref.py in the build container)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")  # reference tree is read-only, kernels use cache=True
    for name in ["openbabel", "openbabel.pybel", "molvoxel", "omegaconf", "Bio", "Bio.PDB", "Bio.PDB.PDBIO"]:
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import pmnet.pharmacophore_model as pm
    from pmnet.scoring import graph_match, ligand, ligand_utils

    # fake atoms expose .nb (neighbour atoms); the reference iterates neighbours through ob.OBAtomAtomIter
    ligand.ob.OBAtomAtomIter = lambda atom: iter(atom.nb)
    return pm, graph_match, ligand, ligand_utils


class _FakeAtom:
    def __init__(self, idx0: int, z: int):
        self._idx = idx0 + 1  # OpenBabel indices are 1-based
        self._z = z
        self.nb: list[_FakeAtom] = []

    def GetIdx(self):
        return self._idx

    def GetAtomicNum(self):
        return self._z


class RefLigand:
    """Duck-typed stand-in for the reference's `Ligand` (ligand.py:16-61) built from a TypedLigand, feeding the
    reference's own unmodified `LigandGraph`."""

    def __init__(self, typed):
        _, _, ligand, ligand_utils = import_reference()
        import numpy as np

        self.obatoms = [_FakeAtom(i, int(z)) for i, z in enumerate(typed.atomic_nums)]
        for i, nbrs in enumerate(typed.neighbors):
            self.obatoms[i].nb = [self.obatoms[j] for j in nbrs]
        self.num_atoms = len(self.obatoms)
        self.num_rotatable_bonds = 0
        self.atom_positions = np.asarray(typed.atom_positions, dtype=np.float32)
        self.num_conformers = self.atom_positions.shape[1]
        self.pharmacophore_list = [
            (typ, ligand_utils.PharmacophoreNode(atom_key, center_key)) for typ, atom_key, center_key in typed.pharmacophores
        ]
        self.graph = ligand.LigandGraph(self)


def ref_create_model(hotspot_infos, center=(0.0, 0.0, 0.0)):
    pm, *_ = import_reference()
    return pm.PharmacophoreModel.create("", center, hotspot_infos)


def ref_score(model, typed, weights=None) -> float:
    _, graph_match, _, _ = import_reference()
    return float(graph_match.GraphMatcher(model, RefLigand(typed), weights).run())
