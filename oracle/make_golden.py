"""TEST INFRASTRUCTURE - generate golden vectors by running the REAL reference (build container only).

    python oracle/make_golden.py            # writes tests/golden/*.pm, tests/golden/*.npz

The reference has no tests or expected outputs of its own, so parity is pinned on outputs of
the reference's unmodified `GraphMatcher.run` (src/pmnet/scoring/graph_match.py:94-101), imported from
/root/reference with the stub recipe of `ref_harness.py`, on seeded synthetic inputs:

* pharmacophore models are built by the reference's own `PharmacophoreModel.create`
  (src/pmnet/pharmacophore_model.py:108-149) and saved with its own `.save` (so the `.pm` files are also
  fixtures for this package's loader);
* ligand graphs are the reference's own `LigandGraph` (src/pmnet/scoring/ligand.py:110-259) fed with the typed
  atoms of `pharmaconet_b200.synthetic`; the script asserts that this package's host featuriser
  (`LigandBatch.from_typed`) packs them identically to packing the reference's graph objects.

Each case file holds the packed ligand batch, the packed model, the weights and `ref_scores` (fp64).
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

from pharmaconet_b200 import packing, synthetic  # noqa: E402
from pharmaconet_b200.constants import weights_vector  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> kwargs of synthetic.make_hotspot_infos
MODELS = {
    "syn0": dict(seed=0),  # the headline "6OIM-like" model: 35 nodes / 26 clusters
    "loose": dict(seed=5, n_hotspots=24, r_lo=2.0, r_hi=3.5),  # wide sigma: large DFS trees
    "sparse": dict(seed=7, n_hotspots=8, type_probs=[0.5, 0, 0, 0, 0, 0.25, 0.25, 0, 0, 0]),  # few types
    "xbond": dict(seed=11, n_hotspots=4, type_probs=[0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0]),  # most ligands match nothing
}

# name -> (model, make_ligands kwargs, weights)
CASES = {
    "syn0_c32": ("syn0", dict(n=192, num_conformers=32, seed=1), None),
    "syn0_c8": ("syn0", dict(n=128, num_conformers=8, seed=2), None),
    "syn0_c1": ("syn0", dict(n=48, num_conformers=1, seed=3), None),
    "syn0_c5_big": ("syn0", dict(n=24, num_conformers=5, seed=4, frag_range=(9, 15)), None),
    "syn0_c32_weights": (
        "syn0",
        dict(n=48, num_conformers=32, seed=6),
        dict(Cation=3.5, Anion=6.0, Aromatic=5.0, HBond_donor=2.0, HBond_acceptor=1.5, Halogen=7.0, Hydrophobic=0.5),
    ),
    "syn0_c48": ("syn0", dict(n=32, num_conformers=48, seed=12), None),  # two conformers per lane
    "syn0_c100": ("syn0", dict(n=16, num_conformers=100, seed=13), None),  # four conformers per lane
    "syn0_c4_deep": ("syn0", dict(n=16, num_conformers=4, seed=4, frag_range=(16, 24)), None),  # > 20 clusters
    "loose_c8": ("loose", dict(n=20, num_conformers=8, seed=8), None),
    "xbond_c4": ("xbond", dict(n=32, num_conformers=4, seed=10), None),
    "sparse_c16": ("sparse", dict(n=96, num_conformers=16, seed=9), None),
}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    pm, graph_match, _, _ = ref_harness.import_reference()
    models = {}
    for name, kw in MODELS.items():
        t = time.time()
        model = ref_harness.ref_create_model(synthetic.make_hotspot_infos(**kw))
        model.save(os.path.join(GOLDEN, f"model_{name}.pm"))
        models[name] = model
        print(f"model {name}: {len(model.nodes)} nodes, {len(model.node_clusters)} clusters ({time.time() - t:.1f}s)")
    for case, (mname, lkw, weights) in CASES.items():
        model = models[mname]
        ligs = synthetic.make_ligands(**lkw)
        t = time.time()
        graphs, ref = [], []
        for lig in ligs:
            rl = ref_harness.RefLigand(lig)
            graphs.append(rl.graph)
            ref.append(float(graph_match.GraphMatcher(model, rl, weights).run()))
        dt = time.time() - t
        b_ref = packing.LigandBatch.from_reference_graphs(graphs)
        b_own = packing.LigandBatch.from_typed(ligs)
        for k, v in b_ref.arrays().items():
            assert np.array_equal(v, b_own.arrays()[k]), f"{case}: host featuriser differs from reference graph in {k}"
        pmod = packing.PackedModel.from_model(model)
        out = {f"lig_{k}": v for k, v in b_ref.arrays().items()}
        out.update({f"model_{k}": v for k, v in pmod.arrays().items()})
        out["weights"] = np.asarray(weights_vector(weights), dtype=np.float32)
        out["ref_scores"] = np.asarray(ref, dtype=np.float64)
        out["gen_kwargs"] = np.asarray(repr(lkw))
        out["model_name"] = np.asarray(mname)
        np.savez_compressed(os.path.join(GOLDEN, f"{case}.npz"), **out)
        ref = np.asarray(ref)
        print(
            f"case {case}: {len(ligs)} ligands, reference {dt:.1f}s, score range [{ref.min():.3f}, {ref.max():.3f}], "
            f"zeros {(ref == 0).sum()}"
        )


if __name__ == "__main__":
    main()
