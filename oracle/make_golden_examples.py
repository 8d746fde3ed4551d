"""TEST INFRASTRUCTURE - golden vectors on REAL molecule geometries (build container only).

    python oracle/make_golden_examples.py      # writes tests/golden/examples_*.npz

Reads the reference's own example library (`/root/reference/examples/library.tar`: 1000 SD files, 8 conformers
each - the input of BASELINE configs[0]), types the ligands with this package's toolkit-free reader
(`pharmaconet_b200.sdf`, approximate perception - OpenBabel is absent), and scores them with the REAL reference's
unmodified `GraphMatcher.run` through `ref_harness` against the committed models. Both sides therefore see the
same typed ligands: the fixtures pin the scoring path on real conformer geometries, not the typing.
"""

from __future__ import annotations

import io
import os
import sys
import tarfile
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

from pharmaconet_b200 import packing, sdf  # noqa: E402
from pharmaconet_b200.constants import weights_vector  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
LIBRARY = "/root/reference/examples/library.tar"
CASES = {"examples_syn0": ("syn0", 0, 160), "examples_loose": ("loose", 160, 224)}  # model, first, last file


def main():
    pm, graph_match, _, _ = ref_harness.import_reference()
    with tarfile.open(LIBRARY) as tar:
        members = sorted((m for m in tar.getmembers() if m.name.endswith(".sdf")), key=lambda m: m.name)
        ligs = []
        with tempfile.TemporaryDirectory() as tmp:
            for m in members[: max(c[2] for c in CASES.values())]:
                path = os.path.join(tmp, os.path.basename(m.name))
                with open(path, "wb") as f:
                    f.write(tar.extractfile(m).read())
                lig = sdf.typed_ligand_from_sdf(path)
                lig.name = m.name
                ligs.append(lig)
    for case, (mname, a, b) in CASES.items():
        model = pm.PharmacophoreModel.load(os.path.join(GOLDEN, f"model_{mname}.pm"))
        sub = ligs[a:b]
        t = time.time()
        graphs, ref = [], []
        for lig in sub:
            rl = ref_harness.RefLigand(lig)
            graphs.append(rl.graph)
            ref.append(float(graph_match.GraphMatcher(model, rl, None).run()))
        dt = time.time() - t
        b_ref = packing.LigandBatch.from_reference_graphs(graphs)
        b_own = packing.LigandBatch.from_typed(sub)
        for k, v in b_ref.arrays().items():
            assert np.array_equal(v, b_own.arrays()[k]), f"{case}: host featuriser differs from reference graph in {k}"
        pmod = packing.PackedModel.from_model(model)
        out = {f"lig_{k}": v for k, v in b_ref.arrays().items()}
        out.update({f"model_{k}": v for k, v in pmod.arrays().items()})
        out["weights"] = np.asarray(weights_vector(None), dtype=np.float32)
        out["ref_scores"] = np.asarray(ref, dtype=np.float64)
        out["gen_kwargs"] = np.asarray(repr(dict(source="examples/library.tar", first=a, last=b)))
        out["model_name"] = np.asarray(mname)
        np.savez_compressed(os.path.join(GOLDEN, f"{case}.npz"), **out)
        ref = np.asarray(ref)
        print(f"case {case}: {len(sub)} ligands, reference {dt:.1f}s, score range [{ref.min():.3f}, {ref.max():.3f}], "
              f"zeros {(ref == 0).sum()}")


if __name__ == "__main__":
    main()
