"""TEST INFRASTRUCTURE - golden vectors of the reference's numpy FALLBACK scorer (build container only).

    python oracle/make_golden_fallback.py      # writes tests/golden/fallback_*.npz

`graph_match.py:12-15` selects `match_utils.py` (numpy, fp32 throughout, :9-122) when `match_utils_numba` cannot be
imported. This script blocks numba (`sys.modules["numba"] = None` makes the import raise), checks that the fallback
was really selected, and re-scores the ligands of existing golden cases with the unmodified `GraphMatcher.run`.
It is the reference's second statement of the same arithmetic (SURVEY.md section 8a, row `match_utils.py`): the C
oracle and the CUDA kernel are compared with it in tests/test_oracle_golden.py / tests/test_scoring_gpu.py.
Each output file holds `ref_scores` (fp64) and the name of the base case whose inputs it re-uses.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

sys.modules["numba"] = None  # any `import numba` now raises ImportError -> graph_match falls back to match_utils

import ref_harness  # noqa: E402

from pharmaconet_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BASE_CASES = ("syn0_c8", "syn0_c32_weights", "loose_c8")


def main():
    pm, graph_match, _, _ = ref_harness.import_reference()
    assert graph_match.scoring_matching_pair.__module__.endswith("match_utils"), "numba variant still selected"
    from make_golden import CASES  # noqa: PLC0415 - generation kwargs of the base cases

    for case in BASE_CASES:
        mname, lkw, weights = CASES[case]
        model = pm.PharmacophoreModel.load(os.path.join(GOLDEN, f"model_{mname}.pm"))
        ligs = synthetic.make_ligands(**lkw)
        ref = [float(graph_match.GraphMatcher(model, ref_harness.RefLigand(l), weights).run()) for l in ligs]
        base = np.load(os.path.join(GOLDEN, case + ".npz"))["ref_scores"]
        rel = np.abs(np.asarray(ref) - base) / np.maximum(np.abs(base), 1e-12)
        np.savez_compressed(
            os.path.join(GOLDEN, f"fallback_{case}.npz"), ref_scores=np.asarray(ref, dtype=np.float64), base_case=np.asarray(case)
        )
        print(f"fallback_{case}: {len(ref)} ligands, max rel diff numpy fallback vs numba variant {rel.max():.2e}")


if __name__ == "__main__":
    main()
