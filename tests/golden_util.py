"""Helpers to read tests/golden/*.npz (written by oracle/make_golden.py from the real reference)."""

import glob
import os

import numpy as np

from pharmaconet_b200.packing import LigandBatch, PackedModel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(
    os.path.basename(f)[:-4]
    for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
    if not os.path.basename(f).startswith(("cnn_", "fallback_", "typing_"))
)  # scoring cases; cnn_*.npz belong to tests/test_cnn_gpu.py
# scores of the same ligands by the reference's numpy fallback scorer (match_utils.py; oracle/make_golden_fallback.py)
FALLBACK_CASES = sorted(
    os.path.basename(f)[len("fallback_") : -4] for f in glob.glob(os.path.join(GOLDEN, "fallback_*.npz"))
)


def load_fallback(name):
    z = np.load(os.path.join(GOLDEN, f"fallback_{name}.npz"))
    assert str(z["base_case"]) == name
    return z["ref_scores"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = PackedModel.from_arrays({k[6:]: z[k] for k in z.files if k.startswith("model_") and k != "model_name"})
    batch = LigandBatch.from_arrays({k[4:]: z[k] for k in z.files if k.startswith("lig_")})
    return dict(
        model=model,
        batch=batch,
        weights=z["weights"],
        ref=z["ref_scores"],
        gen_kwargs=eval(str(z["gen_kwargs"])),  # noqa: S307 - our own fixture
        model_name=str(z["model_name"]),
    )


def weights_dict(vec):
    from pharmaconet_b200.constants import PHARMACOPHORE_TYPES

    return {t: float(v) for t, v in zip(PHARMACOPHORE_TYPES, vec)}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
